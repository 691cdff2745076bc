// CPU replay of the per-thread phases of the FSST kernels (K1, K2) -- test infrastructure.
// Runs the exact __host__ __device__ phase functions of csrc/fsst_phases.cuh in plain loops, with
// the kernels' __syncthreads() boundaries as loop boundaries, so the index arithmetic of the CUDA
// kernels is checked against the float64 oracle without a GPU (tests/test_host_sim.py).
#include <vector>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../../heart-sounds-segmentation_b200/csrc/fsst_phases.cuh"

using namespace hssb;

template <int R2>
static void sim_stft(const float *x, long long N, const float *g, const float *dg, float2 *Sg, float2 *Sdg)
{
    using C = StftCfg<R2>;
    std::vector<float2> buf(C::BUFN), tw(C::NFFT);
    std::vector<float> xs(C::XSN);
    for (int i = 0; i < C::NFFT; ++i) {
        double a = -2.0 * M_PI * i / C::NFFT;
        tw[i] = make_float2((float)cos(a), (float)sin(a));
    }
    const long long ntiles = (N + C::TT - 1) / C::TT;
    for (long long tile = 0; tile < ntiles; ++tile) {
        const long long t0 = tile * C::TT;
        for (int i = 0; i < C::XSN; ++i) {
            long long src = t0 - C::NFFT / 2 + i;
            xs[i] = (src >= 0 && src < N) ? x[src] : 0.0f;
        }
        for (int tid = 0; tid < C::NT; ++tid) stft_phase1<R2>(tid / R2, tid % R2, xs.data(), g, dg, tw.data(), buf.data());
        std::vector<float2> ys((size_t)C::NT * 16);
        for (int tid = 0; tid < C::NT; ++tid) {
            float2 y[C::PER][R2];
            stft_phase2_load<R2>(tid / R2, tid % R2, buf.data(), y);
            memcpy(&ys[(size_t)tid * 16], y, sizeof(y));
        }
        for (int tid = 0; tid < C::NT; ++tid) {
            float2 y[C::PER][R2];
            memcpy(y, &ys[(size_t)tid * 16], sizeof(y));
            stft_phase2_store<R2>(tid / R2, tid % R2, buf.data(), y);
        }
        const long long ncols = (N - t0 < C::TT) ? (N - t0) : C::TT;
        for (int idx = 0; idx < C::K * C::TT; ++idx) {
            const int k = idx / C::TT, c = idx % C::TT;
            if (c < ncols) stft_phase3<R2>(k, c, buf.data(), Sg[(size_t)k * N + t0 + c], Sdg[(size_t)k * N + t0 + c]);
        }
    }
}

extern "C" int hssb_sim_fsst(const float *x, long long N, const float *g, const float *dg, int nwin, float fs,
                             int k_lo, int k_hi, float2 *Sg, float2 *Sdg, float2 *T)
{
    if (nwin == 128) sim_stft<8>(x, N, g, dg, Sg, Sdg);
    else if (nwin == 256) sim_stft<16>(x, N, g, dg, Sg, Sdg);
    else return -3;
    const int K = nwin / 2 + 1, Kout = k_hi - k_lo + 1;
    const float bins_per_hz = (float)((double)nwin / (double)fs);
    for (long long i = 0; i < (long long)Kout * N; ++i) T[i] = make_float2(0.f, 0.f);
    for (long long t = 0; t < N; ++t)
        for (int k = 0; k < K; ++k)
            reassign_one(k, Sg[(size_t)k * N + t], Sdg[(size_t)k * N + t], nwin, bins_per_hz, k_lo, k_hi, T + t, (int)N);
    return 0;
}

extern "C" void hssb_sim_merge(const double *a, const double *b, double *out)
{
    Moments r = merge_moments(Moments{a[0], a[1], a[2]}, Moments{b[0], b[1], b[2]});
    out[0] = r.n; out[1] = r.mean; out[2] = r.m2;
}
