// Small in-register FFT building blocks for the hop-1 STFT kernel (K1).
// __host__ __device__ so the per-thread phases can be replayed on the CPU by tests/host_sim.
#pragma once
#include <cuda_runtime.h>

#ifndef HSSB_HD
#define HSSB_HD __host__ __device__ __forceinline__
#endif

namespace hssb {

HSSB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
HSSB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
HSSB_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
HSSB_HD float2 cmul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT, natural order in / out:  X[k] = sum_n a[n] exp(-2 pi i n k / 4)
HSSB_HD void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_neg_i(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// forward 8-point DFT, natural order.  n = 2a+b, k = c+4d.
HSSB_HD void fft8(float2 *a)
{
    const float h = 0.70710678118654752440f;
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    float2 o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    // o[c] *= W8^c
    o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));     // * (h, -h)
    o2 = cmul_neg_i(o2);                                        // * (0, -1)
    o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));    // * (-h, -h)
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

// forward 16-point DFT, natural order.  n = 4a+b, k = c+4d.
HSSB_HD void fft16(float2 *a)
{
    const float c1 = 0.92387953251128675613f;  // cos(pi/8)
    const float s1 = 0.38268343236508977173f;  // sin(pi/8)
    const float h  = 0.70710678118654752440f;
    // inner 4-point transforms over a (stride 4), one per residue b
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4(a[b], a[4 + b], a[8 + b], a[12 + b]);
    // after this a[4c + b] holds y[b][c]; twiddle y[b][c] *= W16^(b c)
    // b = 1: c = 1,2,3 -> W16^1, W16^2, W16^3
    a[4 * 1 + 1] = cmul(a[4 * 1 + 1], make_float2(c1, -s1));
    a[4 * 2 + 1] = make_float2(h * (a[4 * 2 + 1].x + a[4 * 2 + 1].y), h * (a[4 * 2 + 1].y - a[4 * 2 + 1].x));
    a[4 * 3 + 1] = cmul(a[4 * 3 + 1], make_float2(s1, -c1));
    // b = 2: W16^2, W16^4, W16^6
    a[4 * 1 + 2] = make_float2(h * (a[4 * 1 + 2].x + a[4 * 1 + 2].y), h * (a[4 * 1 + 2].y - a[4 * 1 + 2].x));
    a[4 * 2 + 2] = cmul_neg_i(a[4 * 2 + 2]);
    a[4 * 3 + 2] = make_float2(h * (a[4 * 3 + 2].y - a[4 * 3 + 2].x), -h * (a[4 * 3 + 2].x + a[4 * 3 + 2].y));
    // b = 3: W16^3, W16^6, W16^9
    a[4 * 1 + 3] = cmul(a[4 * 1 + 3], make_float2(s1, -c1));
    a[4 * 2 + 3] = make_float2(h * (a[4 * 2 + 3].y - a[4 * 2 + 3].x), -h * (a[4 * 2 + 3].x + a[4 * 2 + 3].y));
    a[4 * 3 + 3] = cmul(a[4 * 3 + 3], make_float2(-c1, s1));
    // outer 4-point transforms over b for each c: X[c + 4d]
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4(a[4 * c + 0], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
    // now a[4c + d] = X[c + 4d]  -> transpose 4x4 to natural order
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { float2 t = a[4 * c + d]; a[4 * c + d] = a[4 * d + c]; a[4 * d + c] = t; }
}

template <int R> HSSB_HD void fft_small(float2 *a);
template <> HSSB_HD void fft_small<8>(float2 *a) { fft8(a); }
template <> HSSB_HD void fft_small<16>(float2 *a) { fft16(a); }

}  // namespace hssb
