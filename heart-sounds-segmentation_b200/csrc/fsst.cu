// FSST kernels for sm_100a and their C-ABI entry points.
//
//   K1 stft_hop1_kernel      x[B,N] -> Sg,Sdg [B,K,N] complex64         fp32-FMA / smem bound
//   K2 if_reassign_kernel    Sg,Sdg -> T[B,Kt,N] + moment partials       HBM bound (1040+8*Kt B/col)
//   K3 stats_finalize_kernel partials -> (mean,std) per window
//      normalise_kernel      T -> out[B,N,2Kt] (or |T| [B,N,Kt])          HBM bound (16*Kt B/col)
//
// Replaces ssq.fsst + the wrapper post-processing of reference hss/transforms/synchrosqueeze.py:48-111.
#include "hssb_common.cuh"
#include "fsst_phases.cuh"
#include <mutex>
#include <cstdlib>
#include <algorithm>

namespace hssb {

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
template <int R2>
__global__ void __launch_bounds__(StftCfg<R2>::NT)
stft_hop1_kernel(const float *__restrict__ x, long long N, const float *__restrict__ g,
                 const float *__restrict__ dg, float2 *__restrict__ Sg, float2 *__restrict__ Sdg)
{
    using C = StftCfg<R2>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    float2 *tw = buf + C::BUFN;
    float *xs = reinterpret_cast<float *>(tw + C::NFFT);
    float *gs = xs + C::XSN;
    float *dgs = gs + C::NFFT;

    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * C::TT;

    for (int i = tid; i < C::NFFT; i += C::NT) {
        gs[i] = g[i];
        dgs[i] = dg[i];
        float s, c;
        sincospif(-2.0f * (float)i / (float)C::NFFT, &s, &c);
        tw[i] = make_float2(c, s);
    }
    // frame t = xp[t .. t+NFFT-1], xp = [zeros(NFFT/2); x; zeros(NFFT/2-1)]  =>  xs[i] = x[t0 - NFFT/2 + i]
    const float *xb = x + b * N;
    for (int i = tid; i < C::XSN; i += C::NT) {
        const long long src = t0 - C::NFFT / 2 + i;
        xs[i] = (src >= 0 && src < N) ? __ldg(xb + src) : 0.0f;
    }
    __syncthreads();

    const int col = tid / R2, j = tid % R2;
    stft_phase1<R2>(col, j, xs, gs, dgs, tw, buf);
    __syncthreads();
    float2 y[C::PER][R2];
    stft_phase2_load<R2>(col, j, buf, y);
    __syncthreads();
    stft_phase2_store<R2>(col, j, buf, y);
    __syncthreads();

    const long long ncols = (N - t0 < C::TT) ? (N - t0) : C::TT;
    float2 *sg_out = Sg + (size_t)b * C::K * N + t0;
    float2 *sdg_out = Sdg + (size_t)b * C::K * N + t0;
    for (int idx = tid; idx < C::K * C::TT; idx += C::NT) {
        const int k = idx / C::TT, c = idx % C::TT;
        if (c < ncols) {
            float2 sg, sdg;
            stft_phase3<R2>(k, c, buf, sg, sdg);
            __stcs(sg_out + (size_t)k * N + c, sg);
            __stcs(sdg_out + (size_t)k * N + c, sdg);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------
constexpr int RT = 128;        // time columns (= threads) per CTA of the reassignment kernel
// RU = bins loaded ahead per thread (memory-level parallelism); 8 by default, HSSB_RU=5|13 selects the other instantiations

__device__ __forceinline__ Moments shfl_xor_moments(Moments m, int lane_mask)
{
    Moments o;
    o.n = __shfl_xor_sync(0xffffffffu, m.n, lane_mask);
    o.mean = __shfl_xor_sync(0xffffffffu, m.mean, lane_mask);
    o.m2 = __shfl_xor_sync(0xffffffffu, m.m2, lane_mask);
    return o;
}

template <int RU>
__global__ void __launch_bounds__(RT)
if_reassign_kernel(const float2 *__restrict__ Sg, const float2 *__restrict__ Sdg, long long N, int nfft,
                   float bins_per_hz, int k_lo, int k_hi, float2 *__restrict__ T,
                   double *__restrict__ partials)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *acc = reinterpret_cast<float2 *>(smem_raw);     // [Kout][RT]
    __shared__ Moments warp_m[2][RT / 32];

    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    const long long t = (long long)blockIdx.x * RT + tid;
    const int K = nfft / 2 + 1;
    const int Kout = k_hi - k_lo + 1;
    const bool active = t < N;

    for (int r = 0; r < Kout; ++r) acc[r * RT + tid] = make_float2(0.f, 0.f);

    if (active) {
        const float2 *sg = Sg + (size_t)b * K * N + t;
        const float2 *sdg = Sdg + (size_t)b * K * N + t;
        for (int k0 = 0; k0 < K; k0 += RU) {
            float2 a[RU], d[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                if (k0 + u < K) {
                    a[u] = __ldcs(sg + (size_t)(k0 + u) * N);
                    d[u] = __ldcs(sdg + (size_t)(k0 + u) * N);
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u)
                if (k0 + u < K)
                    reassign_one(k0 + u, a[u], d[u], nfft, bins_per_hz, k_lo, k_hi, acc + tid, RT);
        }
        float2 *tout = T + (size_t)b * Kout * N + t;
        for (int r = 0; r < Kout; ++r) __stcs(tout + (size_t)r * N, acc[r * RT + tid]);
    }

    if (partials == nullptr) return;

    // per-thread two-pass moments of this column's Kout values (fp32: 22 values, two-pass), then Chan merges in
    // double (fixed tree order)
    Moments m[2];
    if (active) {
        float sx = 0.f, sy = 0.f;
        for (int r = 0; r < Kout; ++r) { const float2 v = acc[r * RT + tid]; sx += v.x; sy += v.y; }
        const float mx = sx / (float)Kout, my = sy / (float)Kout;
        float qx = 0.f, qy = 0.f, ex = 0.f, ey = 0.f;
        for (int r = 0; r < Kout; ++r) {
            const float2 v = acc[r * RT + tid];
            const float dx = v.x - mx, dy = v.y - my;
            qx = fmaf(dx, dx, qx); ex += dx;
            qy = fmaf(dy, dy, qy); ey += dy;
        }
        // corrected two-pass: the residual sums ex, ey absorb the rounding of the fp32 means
        const double cx = (double)ex / Kout, cy = (double)ey / Kout;
        m[0] = Moments{(double)Kout, (double)mx + cx, (double)qx - cx * cx * Kout};
        m[1] = Moments{(double)Kout, (double)my + cy, (double)qy - cy * cy * Kout};
    } else {
        m[0] = m[1] = Moments{0.0, 0.0, 0.0};
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const Moments o = shfl_xor_moments(m[c], s);
            // keep the merge order identical on both partners: lower lane's value first
            m[c] = ((tid & s) == 0) ? merge_moments(m[c], o) : merge_moments(o, m[c]);
        }
        if ((tid & 31) == 0) warp_m[c][tid >> 5] = m[c];
    }
    __syncthreads();
    if (tid < 2) {
        Moments r = warp_m[tid][0];
        for (int w = 1; w < RT / 32; ++w) r = merge_moments(r, warp_m[tid][w]);
        double *p = partials + (((size_t)b * gridDim.x + blockIdx.x) * 2 + tid) * 3;
        p[0] = r.n; p[1] = r.mean; p[2] = r.m2;
    }
}

// ------------------------------------------------------------------------------------------------
// Generic window lengths (4 <= nwin <= 1024, even or odd; nfft = nwin).  ssq.fsst takes any window (reference
// hss/transforms/synchrosqueeze.py:48); 128 and 256 run on the radix kernels above, every other length on these two: a direct
// DFT of every frame (double accumulators: the sums run over up to 1024 terms) and a reassignment kernel with modulo-nfft rows
// and the general phase shift exp(-2 pi i floor(nwin/2) k / nfft) (exactly (-1)^k only for even nwin).  Same Sg / Sdg / T /
// partial-moment layouts as K1 / K2, so K3 and the wrappers do not care which pair ran.  Correct, not tuned.
// ------------------------------------------------------------------------------------------------
constexpr int GT = 32;          // time columns per CTA of the generic STFT

__global__ void __launch_bounds__(256)
stft_dft_kernel(const float *__restrict__ x, long long N, const float *__restrict__ g, const float *__restrict__ dg, int nwin,
                float2 *__restrict__ Sg, float2 *__restrict__ Sdg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);       // [nwin]  exp(-2 pi i j / nwin)
    float *gs = reinterpret_cast<float *>(tw + nwin);        // [nwin]
    float *dgs = gs + nwin;                                   // [nwin]
    float *xs = dgs + nwin;                                   // [GT + nwin]
    const int tid = threadIdx.x;
    const long long b = blockIdx.y, t0 = (long long)blockIdx.x * GT;
    const int K = nwin / 2 + 1, left = nwin / 2;             // zero padding: nwin/2 (even) or (nwin-1)/2 (odd) samples on the left
    for (int i = tid; i < nwin; i += 256) {
        gs[i] = g[i];
        dgs[i] = dg[i];
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)nwin, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    const float *xb = x + b * N;
    for (int i = tid; i < GT + nwin; i += 256) {
        const long long src = t0 - left + i;
        xs[i] = (src >= 0 && src < N) ? __ldg(xb + src) : 0.0f;
    }
    __syncthreads();
    const long long ncols = (N - t0 < GT) ? (N - t0) : GT;
    for (int idx = tid; idx < K * GT; idx += 256) {
        const int k = idx / GT, c = idx % GT;                // a warp shares k: the twiddle and window reads are broadcasts
        if (c >= ncols) continue;
        double gr = 0.0, gi = 0.0, dr = 0.0, di = 0.0;
        int ph = 0;
        for (int n = 0; n < nwin; ++n) {
            const float v = xs[c + n];
            const float2 w = tw[ph];
            const float a = v * gs[n], d = v * dgs[n];
            gr += (double)(a * w.x); gi += (double)(a * w.y);
            dr += (double)(d * w.x); di += (double)(d * w.y);
            ph += k;
            if (ph >= nwin) ph -= nwin;
        }
        const size_t o = ((size_t)b * K + k) * N + t0 + c;
        Sg[o] = make_float2((float)gr, (float)gi);
        Sdg[o] = make_float2((float)dr, (float)di);
    }
}

__device__ __forceinline__ int wrap_row_generic(float r, int nfft)
{
    if (!(fabsf(r) < 2147483648.0f)) return 0;               // garbage IF of an empty bin: any row, but no trap
    int m = ((int)r) % nfft;
    return m < 0 ? m + nfft : m;
}

// One warp per CTA, 128 columns as four passes of 32 (the per-column accumulators of a pass: [Kout][32] in shared memory), one
// moment partial per 128-column tile like K2.
__global__ void __launch_bounds__(32)
if_reassign_generic_kernel(const float2 *__restrict__ Sg, const float2 *__restrict__ Sdg, long long N, int nfft, float bins_per_hz,
                           int k_lo, int k_hi, float2 *__restrict__ T, double *__restrict__ partials)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *acc = reinterpret_cast<float2 *>(smem_raw);      // [Kout][32]
    const int lane = threadIdx.x;
    const long long b = blockIdx.y;
    const int K = nfft / 2 + 1, Kout = k_hi - k_lo + 1, half = nfft / 2;
    Moments tile_m[2] = {Moments{0.0, 0.0, 0.0}, Moments{0.0, 0.0, 0.0}};
    for (int pass = 0; pass < RT / 32; ++pass) {
        const long long t = (long long)blockIdx.x * RT + pass * 32 + lane;
        const bool active = t < N;
        for (int r = 0; r < Kout; ++r) acc[r * 32 + lane] = make_float2(0.f, 0.f);
        Moments m[2] = {Moments{0.0, 0.0, 0.0}, Moments{0.0, 0.0, 0.0}};
        if (active) {
            const float2 *sg = Sg + (size_t)b * K * N + t;
            const float2 *sdg = Sdg + (size_t)b * K * N + t;
            for (int k = 0; k < K; ++k) {
                const float2 a = __ldcs(sg + (size_t)k * N), d = __ldcs(sdg + (size_t)k * N);
                const float den = a.x * a.x + a.y * a.y;
                const float num = d.x * a.y - d.y * a.x;     // -imag(Sdg * conj(Sg))
                float fc = __fdividef(num, den);
                if (!(fabsf(fc) <= 3.0e38f)) fc = 0.0f;
                const float off = fc * bins_per_hz;
                float sn, cs;                                 // exp(-2 pi i floor(nwin/2) k / nfft), argument reduced in integers
                sincospif(-2.0f * (float)((int)(((long long)half * k) % nfft)) / (float)nfft, &sn, &cs);
                if ((nfft & 1) == 0) { cs = (k & 1) ? -1.0f : 1.0f; sn = 0.0f; }
                const float vx = a.x * cs - a.y * sn, vy = a.x * sn + a.y * cs;
                const int row = wrap_row_generic(round_half_away((float)k + off), nfft);
                if (row >= k_lo && row <= k_hi) { float2 *p = acc + (row - k_lo) * 32 + lane; p->x += vx; p->y += vy; }
                if (k > 0 && 2 * k != nfft) {                 // the negative-frequency mirror of the bin: value conj, correction negated
                    const int rowm = wrap_row_generic(round_half_away((float)(nfft - k) - off), nfft);
                    if (rowm >= k_lo && rowm <= k_hi) { float2 *p = acc + (rowm - k_lo) * 32 + lane; p->x += vx; p->y -= vy; }
                }
            }
            float2 *tout = T + (size_t)b * Kout * N + t;
            double sx = 0.0, sy = 0.0;
            for (int r = 0; r < Kout; ++r) { const float2 v = acc[r * 32 + lane]; __stcs(tout + (size_t)r * N, v); sx += v.x; sy += v.y; }
            const double mx = sx / Kout, my = sy / Kout;
            double qx = 0.0, qy = 0.0;
            for (int r = 0; r < Kout; ++r) { const float2 v = acc[r * 32 + lane]; qx += (v.x - mx) * (v.x - mx); qy += (v.y - my) * (v.y - my); }
            m[0] = Moments{(double)Kout, mx, qx};
            m[1] = Moments{(double)Kout, my, qy};
        }
        if (partials) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int s2 = 1; s2 < 32; s2 <<= 1) {
                    const Moments o = shfl_xor_moments(m[c], s2);
                    m[c] = ((lane & s2) == 0) ? merge_moments(m[c], o) : merge_moments(o, m[c]);
                }
                tile_m[c] = merge_moments(tile_m[c], m[c]);
            }
        }
        __syncwarp();
    }
    if (partials && lane < 2) {
        double *p = partials + (((size_t)b * gridDim.x + blockIdx.x) * 2 + lane) * 3;
        const Moments r = lane ? tile_m[1] : tile_m[0];
        p[0] = r.n; p[1] = r.mean; p[2] = r.m2;
    }
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
// one CTA per window, warp c merges the tiles' partials of channel c -> final[b] = {mean_re, std_re, mean_im, std_im}
__global__ void __launch_bounds__(64)
stats_finalize_kernel(const double *__restrict__ partials, int ntiles, float *__restrict__ final_stats)
{
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = blockIdx.x;
    Moments m{0.0, 0.0, 0.0};
    for (int i = lane; i < ntiles; i += 32) {
        const double *p = partials + (((size_t)b * ntiles + i) * 2 + c) * 3;
        m = merge_moments(m, Moments{p[0], p[1], p[2]});
    }
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const Moments o = shfl_xor_moments(m, s);
        m = ((lane & s) == 0) ? merge_moments(m, o) : merge_moments(o, m);
    }
    if (lane == 0) {
        final_stats[b * 4 + 2 * c] = (float)m.mean;
        final_stats[b * 4 + 2 * c + 1] = (float)sqrt(m.m2 / (m.n - 1.0));   // unbiased (torch.std default)
    }
}

constexpr int FT = 128;  // time columns per CTA of the normalise kernel (64-column CTAs were launch / latency bound)

__global__ void __launch_bounds__(256)
normalise_kernel(const float2 *__restrict__ T, const float *__restrict__ final_stats, long long N, int Kt,
                 int mode, float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);     // [FT][W+1]
    const int W = (mode == HSSB_MODE_STACK) ? 2 * Kt : Kt;
    const int WS = W | 1;                                  // odd row stride: conflict-free transposition
    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * FT;
    const int ncols = (int)((N - t0 < FT) ? (N - t0) : FT);

    float mr = 0.f, sr = 1.f, mi = 0.f, si = 1.f;
    if (mode == HSSB_MODE_STACK) {
        mr = final_stats[b * 4 + 0]; sr = final_stats[b * 4 + 1];
        mi = final_stats[b * 4 + 2]; si = final_stats[b * 4 + 3];
    }
    // (v - mean) * (1 / std): one reciprocal per CTA instead of an IEEE division per element (<= 1 ulp from the quotient)
    const float ir = 1.0f / sr, ii = 1.0f / si;
    const float2 *tin = T + (size_t)b * Kt * N + t0;
    // 8 loads in flight per thread (one load per loop trip left the kernel latency bound at half the HBM rate)
    constexpr int MLP = 8;
    for (int base = tid; base < Kt * FT; base += 256 * MLP) {
        float2 v[MLP];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const int idx = base + 256 * u, r = idx / FT, c = idx % FT;
            if (idx < Kt * FT && c < ncols) v[u] = __ldcs(tin + (size_t)r * N + c);
        }
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const int idx = base + 256 * u, r = idx / FT, c = idx % FT;
            if (idx < Kt * FT && c < ncols) {
                if (mode == HSSB_MODE_STACK) {
                    tile[c * WS + r] = (v[u].x - mr) * ir;
                    tile[c * WS + Kt + r] = (v[u].y - mi) * ii;
                } else {
                    tile[c * WS + r] = hypotf(v[u].x, v[u].y);
                }
            }
        }
    }
    __syncthreads();
    float *o = out + ((size_t)b * N + t0) * W;
    const int total = ncols * W;
    if (((((size_t)b * N + t0) * W) & 3) == 0) {          // 16-byte stores over the contiguous [ncols][W] block
        for (int i4 = tid * 4; i4 < total; i4 += 256 * 4) {
            if (i4 + 3 < total) {
                float4 v;
                v.x = tile[(i4 / W) * WS + (i4 % W)];
                v.y = tile[((i4 + 1) / W) * WS + ((i4 + 1) % W)];
                v.z = tile[((i4 + 2) / W) * WS + ((i4 + 2) % W)];
                v.w = tile[((i4 + 3) / W) * WS + ((i4 + 3) % W)];
                __stcs(reinterpret_cast<float4 *>(o + i4), v);
            } else {
                for (int i = i4; i < total; ++i) __stcs(o + i, tile[(i / W) * WS + (i % W)]);
            }
        }
    } else {
        for (int i = tid; i < total; i += 256) __stcs(o + i, tile[(i / W) * WS + (i % W)]);
    }
}

}  // namespace hssb

using namespace hssb;

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
static int check_nwin(int nwin)
{
    if (nwin < 4 || nwin > 1024) return fail(HSSB_E_NWIN, "nwin=%d unsupported (4 .. 1024)", nwin);
    return 0;
}
static bool radix_nwin(int nwin) { return nwin == 128 || nwin == 256; }

static int ntiles_reassign(int64_t N) { return (int)((N + RT - 1) / RT); }

extern "C" int hssb_fsst_stft(const float *x, int64_t B, int64_t N, const float *g, const float *dg,
                              int nwin, hssb_c32 *Sg, hssb_c32 *Sdg, void *stream)
{
    if (!x || !g || !dg || !Sg || !Sdg) return fail(HSSB_E_NULL, "hssb_fsst_stft: null pointer");
    if (B < 0 || N < 0 || B > 65535) return fail(HSSB_E_SHAPE, "hssb_fsst_stft: B=%lld N=%lld", (long long)B, (long long)N);
    if (int rc = check_nwin(nwin)) return rc;
    if (B == 0 || N == 0) return 0;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    ProfScope prof("stft_hop1", st);
    if (nwin == 128) {
        using C = StftCfg<8>;
        HSSB_CUDA_OK(cudaFuncSetAttribute(stft_hop1_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));   // per device: set on every call
        dim3 grid((unsigned)((N + C::TT - 1) / C::TT), (unsigned)B);
        stft_hop1_kernel<8><<<grid, C::NT, C::SMEM_BYTES, st>>>(x, N, g, dg, (float2 *)Sg, (float2 *)Sdg);
    } else if (nwin != 256) {
        const size_t smem = sizeof(float2) * nwin + sizeof(float) * (2 * nwin + GT + nwin);
        dim3 grid((unsigned)((N + GT - 1) / GT), (unsigned)B);
        stft_dft_kernel<<<grid, 256, smem, st>>>(x, N, g, dg, nwin, (float2 *)Sg, (float2 *)Sdg);
    } else {
        using C = StftCfg<16>;
        HSSB_CUDA_OK(cudaFuncSetAttribute(stft_hop1_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        dim3 grid((unsigned)((N + C::TT - 1) / C::TT), (unsigned)B);
        stft_hop1_kernel<16><<<grid, C::NT, C::SMEM_BYTES, st>>>(x, N, g, dg, (float2 *)Sg, (float2 *)Sdg);
    }
    HSSB_LAUNCH_OK("stft_hop1_kernel");
    return 0;
}

extern "C" size_t hssb_fsst_stats_words(int64_t B, int64_t N)
{
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * ntiles_reassign(N) * 6 + (size_t)B * 2;   // partials + 4 floats per window
}

extern "C" int hssb_fsst_reassign(const hssb_c32 *Sg, const hssb_c32 *Sdg, int64_t B, int64_t N, int nwin,
                                  float fs, int k_lo, int k_hi, hssb_c32 *T, double *stats, void *stream)
{
    if (!Sg || !Sdg || !T) return fail(HSSB_E_NULL, "hssb_fsst_reassign: null pointer");
    if (B < 0 || N < 0 || B > 65535) return fail(HSSB_E_SHAPE, "hssb_fsst_reassign: B=%lld N=%lld", (long long)B, (long long)N);
    if (int rc = check_nwin(nwin)) return rc;
    if (k_lo < 0 || k_hi > nwin / 2 || k_hi < k_lo) return fail(HSSB_E_BAND, "band [%d,%d] outside [0,%d]", k_lo, k_hi, nwin / 2);
    if (!(fs > 0.f)) return fail(HSSB_E_SHAPE, "fs must be positive");
    if (B == 0 || N == 0) return 0;
    if (int rc = require_sm100()) return rc;
    const int Kout = k_hi - k_lo + 1;
    if (!radix_nwin(nwin)) {
        const size_t gsmem = sizeof(float2) * (size_t)Kout * 32;
        HSSB_CUDA_OK(cudaFuncSetAttribute(if_reassign_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float2) * 513 * 32)));
        dim3 ggrid((unsigned)ntiles_reassign(N), (unsigned)B);
        ProfScope prof("if_reassign", as_stream(stream));
        if_reassign_generic_kernel<<<ggrid, 32, gsmem, as_stream(stream)>>>((const float2 *)Sg, (const float2 *)Sdg, N, nwin,
                                                                             (float)((double)nwin / (double)fs), k_lo, k_hi, (float2 *)T, stats);
        HSSB_LAUNCH_OK("if_reassign_generic_kernel");
        return 0;
    }
    const size_t smem = sizeof(float2) * (size_t)Kout * RT;
    static PerDeviceInt attr_set;                 // the opt-in shared-memory size is a per-device attribute
    if (!attr_set.get()) {
        const int mx = (int)(sizeof(float2) * 129 * RT);
        HSSB_CUDA_OK(cudaFuncSetAttribute(if_reassign_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        HSSB_CUDA_OK(cudaFuncSetAttribute(if_reassign_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        HSSB_CUDA_OK(cudaFuncSetAttribute(if_reassign_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr_set.set(1);
    }
    dim3 grid((unsigned)ntiles_reassign(N), (unsigned)B);
    const float bins_per_hz = (float)((double)nwin / (double)fs);
    const char *ru_env = getenv("HSSB_RU");
    const int ru = ru_env ? atoi(ru_env) : 8;
    ProfScope prof("if_reassign", as_stream(stream));
    if (ru == 5)
        if_reassign_kernel<5><<<grid, RT, smem, as_stream(stream)>>>((const float2 *)Sg, (const float2 *)Sdg, N, nwin, bins_per_hz, k_lo, k_hi, (float2 *)T, stats);
    else if (ru == 13)
        if_reassign_kernel<13><<<grid, RT, smem, as_stream(stream)>>>((const float2 *)Sg, (const float2 *)Sdg, N, nwin, bins_per_hz, k_lo, k_hi, (float2 *)T, stats);
    else
        if_reassign_kernel<8><<<grid, RT, smem, as_stream(stream)>>>((const float2 *)Sg, (const float2 *)Sdg, N, nwin, bins_per_hz, k_lo, k_hi, (float2 *)T, stats);
    HSSB_LAUNCH_OK("if_reassign_kernel");
    return 0;
}

extern "C" int hssb_fsst_finish(const hssb_c32 *T, const double *stats, int64_t B, int64_t N, int Kt, int mode,
                                float *out, void *stream)
{
    if (!T || !out) return fail(HSSB_E_NULL, "hssb_fsst_finish: null pointer");
    if (mode != HSSB_MODE_ABS && mode != HSSB_MODE_STACK) return fail(HSSB_E_MODE, "hssb_fsst_finish: mode %d", mode);
    if (mode == HSSB_MODE_STACK && !stats) return fail(HSSB_E_NULL, "hssb_fsst_finish: STACK needs stats");
    if (B < 0 || N < 0 || Kt < 1 || Kt > 513 || B > 65535) return fail(HSSB_E_SHAPE, "hssb_fsst_finish: bad shape");
    if (B == 0 || N == 0) return 0;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    const int ntiles = ntiles_reassign(N);
    float *final_stats = nullptr;
    if (mode == HSSB_MODE_STACK) {
        final_stats = reinterpret_cast<float *>(const_cast<double *>(stats) + (size_t)B * ntiles * 6);
        ProfScope prof("stats_finalize", st);
        stats_finalize_kernel<<<(unsigned)B, 64, 0, st>>>(stats, ntiles, final_stats);
        HSSB_LAUNCH_OK("stats_finalize_kernel");
    }
    const int W = (mode == HSSB_MODE_STACK) ? 2 * Kt : Kt;
    const size_t smem = sizeof(float) * FT * (size_t)(W | 1);
    if (smem > 227 * 1024) return fail(HSSB_E_BAND, "hssb_fsst_finish: %d rows do not fit the normalise tile (<= 220 in stack mode, <= 440 otherwise)", Kt);
    HSSB_CUDA_OK(cudaFuncSetAttribute(normalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, sizeof(float) * FT * 259)));
    dim3 grid((unsigned)((N + FT - 1) / FT), (unsigned)B);
    ProfScope prof("normalise", st);
    normalise_kernel<<<grid, 256, smem, st>>>((const float2 *)T, final_stats, N, Kt, mode, out);
    HSSB_LAUNCH_OK("normalise_kernel");
    return 0;
}

namespace {
struct FsstWs { size_t sg, sdg, t, stats, total; };
FsstWs fsst_ws_layout(int64_t B, int64_t N, int nwin, int k_lo, int k_hi, int mode)
{
    const size_t K = nwin / 2 + 1, Kt = k_hi - k_lo + 1;
    FsstWs w{};
    size_t off = 0;
    w.sg = off;  off += align_up(sizeof(float2) * (size_t)B * K * N, 256);
    w.sdg = off; off += align_up(sizeof(float2) * (size_t)B * K * N, 256);
    w.t = off;   if (mode != HSSB_MODE_RAW) off += align_up(sizeof(float2) * (size_t)B * Kt * N, 256);
    w.stats = off; if (mode == HSSB_MODE_STACK) off += align_up(sizeof(double) * hssb_fsst_stats_words(B, N), 256);
    w.total = off;
    return w;
}
}  // namespace

extern "C" size_t hssb_fsst_workspace_bytes(int64_t B, int64_t N, int nwin, int k_lo, int k_hi, int mode)
{
    if (B <= 0 || N <= 0 || k_hi < k_lo) return 0;
    return fsst_ws_layout(B, N, nwin, k_lo, k_hi, mode).total;
}

extern "C" int hssb_fsst_forward(const float *x, int64_t B, int64_t N, const float *g, const float *dg, int nwin,
                                 float fs, int k_lo, int k_hi, int mode, void *out, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    if (!x || !g || !dg || !out) return fail(HSSB_E_NULL, "hssb_fsst_forward: null pointer");
    if (mode < HSSB_MODE_RAW || mode > HSSB_MODE_STACK) return fail(HSSB_E_MODE, "hssb_fsst_forward: mode %d", mode);
    if (int rc = check_nwin(nwin)) return rc;
    if (k_lo < 0 || k_hi > nwin / 2 || k_hi < k_lo) return fail(HSSB_E_BAND, "band [%d,%d] outside [0,%d]", k_lo, k_hi, nwin / 2);
    if (B < 0 || N < 0) return fail(HSSB_E_SHAPE, "hssb_fsst_forward: bad shape");
    if (B == 0 || N == 0) return 0;
    const FsstWs w = fsst_ws_layout(B, N, nwin, k_lo, k_hi, mode);
    if (!workspace || workspace_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(HSSB_E_WORKSPACE, "hssb_fsst_forward: workspace %zu < %zu or misaligned", workspace_bytes, w.total);
    char *ws = static_cast<char *>(workspace);
    hssb_c32 *Sg = reinterpret_cast<hssb_c32 *>(ws + w.sg), *Sdg = reinterpret_cast<hssb_c32 *>(ws + w.sdg);
    hssb_c32 *T = (mode == HSSB_MODE_RAW) ? static_cast<hssb_c32 *>(out) : reinterpret_cast<hssb_c32 *>(ws + w.t);
    double *stats = (mode == HSSB_MODE_STACK) ? reinterpret_cast<double *>(ws + w.stats) : nullptr;
    if (int rc = hssb_fsst_stft(x, B, N, g, dg, nwin, Sg, Sdg, stream)) return rc;
    if (int rc = hssb_fsst_reassign(Sg, Sdg, B, N, nwin, fs, k_lo, k_hi, T, stats, stream)) return rc;
    if (mode != HSSB_MODE_RAW)
        if (int rc = hssb_fsst_finish(T, stats, B, N, k_hi - k_lo + 1, mode, static_cast<float *>(out), stream)) return rc;
    return 0;
}

// Host entry point: grow-only device scratch shared by the calling process (guarded by a mutex), tied to the device it
// was allocated on -- a call with another current device gets a fresh buffer.
namespace {
std::mutex g_host_mu;
void *g_host_buf = nullptr;
size_t g_host_cap = 0;
int g_host_dev = -1;
}  // namespace

extern "C" int hssb_fsst_host(const float *x, int64_t B, int64_t N, double fs, const double *window,
                              const double *dwindow, int nwin, int k_lo, int k_hi, int mode, void *out)
{
    if (!x || !window || !dwindow || !out) return fail(HSSB_E_NULL, "hssb_fsst_host: null pointer");
    if (mode < HSSB_MODE_RAW || mode > HSSB_MODE_STACK) return fail(HSSB_E_MODE, "hssb_fsst_host: mode %d", mode);
    if (int rc = check_nwin(nwin)) return rc;
    if (k_lo < 0 || k_hi > nwin / 2 || k_hi < k_lo) return fail(HSSB_E_BAND, "band [%d,%d] outside [0,%d]", k_lo, k_hi, nwin / 2);
    if (B < 0 || N < 0) return fail(HSSB_E_SHAPE, "hssb_fsst_host: bad shape");
    if (B == 0 || N == 0) return 0;
    if (int rc = require_sm100()) return rc;
    const size_t Kt = k_hi - k_lo + 1;
    const size_t out_bytes = (mode == HSSB_MODE_RAW) ? sizeof(float2) * B * Kt * N
                             : sizeof(float) * B * N * (mode == HSSB_MODE_STACK ? 2 * Kt : Kt);
    const size_t x_bytes = align_up(sizeof(float) * (size_t)B * N, 256);
    const size_t win_bytes = align_up(sizeof(float) * 2 * nwin, 256);
    const size_t ws_bytes = hssb_fsst_workspace_bytes(B, N, nwin, k_lo, k_hi, mode);
    const size_t total = x_bytes + win_bytes + align_up(out_bytes, 256) + ws_bytes;

    std::lock_guard<std::mutex> lock(g_host_mu);
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (total > g_host_cap || dev != g_host_dev) {
        if (g_host_buf) cudaFree(g_host_buf);
        g_host_buf = nullptr; g_host_cap = 0; g_host_dev = -1;
        HSSB_CUDA_OK(cudaMalloc(&g_host_buf, total));
        g_host_cap = total;
        g_host_dev = dev;
    }
    char *base = static_cast<char *>(g_host_buf);
    float *dx = reinterpret_cast<float *>(base);
    float *dwin = reinterpret_cast<float *>(base + x_bytes);
    void *dout = base + x_bytes + win_bytes;
    void *dws = base + x_bytes + win_bytes + align_up(out_bytes, 256);
    float hwin[2048];
    for (int i = 0; i < nwin; ++i) { hwin[i] = (float)window[i]; hwin[nwin + i] = (float)dwindow[i]; }
    cudaStream_t st = nullptr;   // legacy default stream: ordered with the synchronous copies below
    HSSB_CUDA_OK(cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)B * N, cudaMemcpyHostToDevice, st));
    HSSB_CUDA_OK(cudaMemcpyAsync(dwin, hwin, sizeof(float) * 2 * nwin, cudaMemcpyHostToDevice, st));
    if (int rc = hssb_fsst_forward(dx, B, N, dwin, dwin + nwin, nwin, (float)fs, k_lo, k_hi, mode, dout, dws, ws_bytes, st)) return rc;
    HSSB_CUDA_OK(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, st));
    HSSB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
