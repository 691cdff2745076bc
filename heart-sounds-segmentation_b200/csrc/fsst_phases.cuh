// Per-thread phases of the FSST kernels, written as __host__ __device__ functions over explicit
// (thread index, shared-memory pointers) so tests/host_sim can replay the exact index arithmetic
// on the CPU.  The __global__ wrappers live in fsst.cu.
//
// K1 (hop-1 STFT), NFFT = 16*R2, one time column per R2 threads, 16 points per thread:
//   n = R2*n2 + n1 (n1 = thread-in-column, n2 in registers), k = k1 + 16*k2.
//   phase 1: 16-point DFT over n2, twiddle W_NFFT^(n1*k1), scatter to the exchange buffer
//   phase 2: R2-point DFTs over n1 for PER = 16/R2 values of k1, stage Z[k][col]
//   phase 3: Hermitian split of Z = FFT(x*g + i*x*g') into S_g and S_dg, coalesced stores.
#pragma once
#include <cmath>
#include "fsst_fft.cuh"

namespace hssb {

template <int R2>
struct StftCfg {
    static constexpr int NFFT = 16 * R2;
    static constexpr int K = NFFT / 2 + 1;
    static constexpr int TT = 32;              // time columns per CTA
    static constexpr int NT = TT * R2;         // threads per CTA
    static constexpr int PER = 16 / R2;        // k1 values per thread in phase 2
    static constexpr int XS = NFFT + 8;        // exchange-buffer column stride (float2)
    static constexpr int ZS = TT + 2;          // stage row stride (float2)
    static constexpr int BUFN = (TT * XS > NFFT * ZS) ? TT * XS : NFFT * ZS;
    static constexpr int XSN = TT + NFFT;      // staged input samples (TT + NFFT - 1, padded)
    static constexpr size_t SMEM_BYTES = sizeof(float) * (XSN + 2 * NFFT) + sizeof(float2) * (NFFT + BUFN);
};

template <int R2>
HSSB_HD void stft_phase1(int col, int j, const float *xs, const float *gs, const float *dgs,
                         const float2 *tw, float2 *buf)
{
    using C = StftCfg<R2>;
    float2 a[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
        const int n = R2 * n2 + j;
        const float v = xs[col + n];
        a[n2] = make_float2(v * gs[n], v * dgs[n]);
    }
    fft16(a);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const float2 y = (k1 == 0) ? a[0] : cmul(a[k1], tw[j * k1]);
        buf[col * C::XS + k1 * R2 + (j ^ (k1 / C::PER))] = y;
    }
}

template <int R2>
HSSB_HD void stft_phase2_load(int col, int j, const float2 *buf, float2 (&y)[StftCfg<R2>::PER][R2])
{
    using C = StftCfg<R2>;
#pragma unroll
    for (int a = 0; a < C::PER; ++a) {
        const int k1 = C::PER * j + a;
#pragma unroll
        for (int n1 = 0; n1 < R2; ++n1) y[a][n1] = buf[col * C::XS + k1 * R2 + (n1 ^ j)];
        fft_small<R2>(y[a]);
    }
}

template <int R2>
HSSB_HD void stft_phase2_store(int col, int j, float2 *buf, const float2 (&y)[StftCfg<R2>::PER][R2])
{
    using C = StftCfg<R2>;
#pragma unroll
    for (int a = 0; a < C::PER; ++a) {
        const int k1 = C::PER * j + a;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) buf[(k1 + 16 * k2) * C::ZS + col] = y[a][k2];
    }
}

// Hermitian split for bin k, column c.  Z = U + iV with U = FFT(x g), V = FFT(x g') Hermitian.
template <int R2>
HSSB_HD void stft_phase3(int k, int c, const float2 *buf, float2 &sg, float2 &sdg)
{
    using C = StftCfg<R2>;
    const float2 zk = buf[k * C::ZS + c];
    const float2 zm = buf[((C::NFFT - k) & (C::NFFT - 1)) * C::ZS + c];
    sg = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    sdg = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
}

// ---------------------------------------------------------------------------------------------
// K2: instantaneous frequency + frequency reassignment of one time column.
//   sg/sdg: this column's bins, element k at sg[k*kstride]
//   acc   : this column's accumulators, band row r at acc[r*astride] (zeroed by the caller)
// Implements steps 4-7 of MATLAB fsst (the algorithm behind ssq.fsst, reference synchrosqueeze.py:48):
//   fcorr = -imag(Sdg/Sg) (non-finite -> 0); row = mod(round_half_away(k + fcorr*nfft/fs), nfft);
//   value = (-1)^k Sg[k] (phase shift exp(-i*pi*k), nwin even); the bins 1..nfft/2-1 also
//   contribute through their negative-frequency mirror (value conj, fcorr negated).
// ---------------------------------------------------------------------------------------------
HSSB_HD int wrap_row(float r, int nfft)
{
    // r is integer valued.  Below 2^31 the int conversion is exact; from 2^31 on every float is a multiple of 256,
    // i.e. of nfft (128 or 256): row 0.  Branch free (the IF of a bin with S_g ~ 0 is garbage, but it must not trap).
    return (fabsf(r) < 2147483648.0f) ? (((int)r) & (nfft - 1)) : 0;
}

// MATLAB round(): half away from zero.  x - trunc(x) is exact, so this is too.
HSSB_HD float round_half_away(float x)
{
    const float t = truncf(x);
    return (fabsf(x - t) >= 0.5f) ? t + copysignf(1.0f, x) : t;
}

HSSB_HD void reassign_one(int k, float2 a, float2 d, int nfft, float bins_per_hz, int k_lo, int k_hi,
                          float2 *acc, int astride)
{
    const float den = a.x * a.x + a.y * a.y;
    const float num = d.x * a.y - d.y * a.x;            // -imag(Sdg * conj(Sg))
#ifdef __CUDA_ARCH__
    float fc = __fdividef(num, den);                    // MUFU.RCP + FMUL: 2 ulp, no slow-path call
#else
    float fc = num / den;
#endif
    if (!(fabsf(fc) <= 3.0e38f)) fc = 0.0f;   // NaN or Inf -> 0
    const float off = fc * bins_per_hz;
    const float sgn = (k & 1) ? -1.0f : 1.0f;
    const float vx = sgn * a.x, vy = sgn * a.y;
    const int row = wrap_row(round_half_away((float)k + off), nfft);
    if (row >= k_lo && row <= k_hi) {
        float2 *p = acc + (row - k_lo) * astride;
        p->x += vx; p->y += vy;
    }
    if (k > 0 && k < nfft / 2) {
        // The mirror bin nfft - k lands on a kept row only after a correction of dozens of bins: between k_hi + 1 and
        // nfft + k_lo - 1 the (unwrapped) destination rounds to a row outside [k_lo, k_hi] whatever the wrap does, so the rounding,
        // wrapping and range test are skipped there -- the usual case (the test is conservative by a whole row: exact).
        const float vm = (float)(nfft - k) - off;
        if (!(vm > (float)(k_hi + 1) && vm < (float)(nfft + k_lo - 1))) {
            const int rowm = wrap_row(round_half_away(vm), nfft);
            if (rowm >= k_lo && rowm <= k_hi) {
                float2 *p = acc + (rowm - k_lo) * astride;
                p->x += vx; p->y -= vy;
            }
        }
    }
}

// Chan / Welford pairwise merge of (n, mean, M2): the parallel form of the streaming recurrences
// in reference hss/moments/__init__.py:16,35-36.
struct Moments { double n, mean, m2; };
HSSB_HD Moments merge_moments(Moments a, Moments b)
{
    const double n = a.n + b.n;
    if (n == 0.0) return Moments{0.0, 0.0, 0.0};
    const double delta = b.mean - a.mean;
    Moments r;
    r.n = n;
    if (a.n == b.n) {          // the common case inside a full tile: no division (x * 0.5 is exact)
        r.mean = a.mean + delta * 0.5;
        r.m2 = a.m2 + b.m2 + delta * delta * (a.n * 0.5);
    } else {
        r.mean = a.mean + delta * (b.n / n);
        r.m2 = a.m2 + b.m2 + delta * delta * (a.n * b.n / n);
    }
    return r;
}

}  // namespace hssb
