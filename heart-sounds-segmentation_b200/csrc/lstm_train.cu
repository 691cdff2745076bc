// Training recurrences of one bidirectional LSTM layer: the forward that keeps what back-propagation
// through time needs, and the backward recurrence.  fp32 SIMT -- a first, correct version; the plain GEMMs either side
// (x W_ih^T, dG^T x, dG^T h_prev, dG W_ih) are library GEMMs issued by the host wrapper (hss/model/_train.py).
// Replaces what autograd does for nn.LSTM in the reference's training step (main.py:67-82 over hss/model/segmenter.py:80-83).
//
// Layouts (all fp32): gates [dir][B*T][4H] (row b*T + t, torch gate order i,f,g,o), cells [dir][B*T][H], out [B][T][2H]
// (dir 0 = forward in columns 0..H-1, dir 1 = reverse in H..2H-1), states [dir][B][H].
#include "model.cuh"
#include <cstdlib>

namespace hssb {

namespace {

constexpr int TR = 4;          // batch rows of one direction per CTA
constexpr int TQ = 4;          // thread groups that split the 4H-long contraction of the backward step
constexpr int TRAIN_THREADS = 256;

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// gates: in = x W_ih^T + b_ih + b_hh (pre-activations without the recurrent term), out = activated gates i,f,g,o.
__global__ void __launch_bounds__(TRAIN_THREADS *TQ)
lstm_train_fwd_kernel(float *__restrict__ gates, const float *__restrict__ w0T, const float *__restrict__ w1T,
                      const float *__restrict__ h0, const float *__restrict__ c0, long long B, long long T, int H,
                      float *__restrict__ out, float *__restrict__ cells, float *__restrict__ hn, float *__restrict__ cn)
{
    extern __shared__ float sm[];
    float *h_s = sm;                    // [TR][H]
    float *c_s = h_s + TR * H;          // [TR][H]
    float *g_s = c_s + TR * H;          // [TR][4H]
    const int tid = threadIdx.x, nthreads = TRAIN_THREADS * TQ;   // one gate column per thread for H <= 256
    const int dir = blockIdx.y;
    const long long b0 = (long long)blockIdx.x * TR;
    const int G = 4 * H;
    const float *wT = dir ? w1T : w0T;                         // [H][4H] = W_hh^T
    float *gd = gates + (size_t)dir * B * T * G;
    float *cd = cells + (size_t)dir * B * T * H;

    for (int i = tid; i < TR * H; i += nthreads) {
        const int r = i / H, u = i % H;
        const bool ok = b0 + r < B;
        h_s[i] = ok ? h0[((size_t)dir * B + b0 + r) * H + u] : 0.f;
        c_s[i] = ok ? c0[((size_t)dir * B + b0 + r) * H + u] : 0.f;
    }
    __syncthreads();

    for (long long step = 0; step < T; ++step) {
        const long long t = dir ? (T - 1 - step) : step;
        for (int n = tid; n < G; n += nthreads) {
            float acc[TR];
#pragma unroll
            for (int r = 0; r < TR; ++r) acc[r] = (b0 + r < B) ? gd[((size_t)(b0 + r) * T + t) * G + n] : 0.f;
#pragma unroll 8
            for (int k = 0; k < H; ++k) {
                const float w = __ldg(wT + (size_t)k * G + n);
#pragma unroll
                for (int r = 0; r < TR; ++r) acc[r] = fmaf(h_s[r * H + k], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < TR; ++r) g_s[r * G + n] = acc[r];
        }
        __syncthreads();
        for (int i = tid; i < TR * H; i += nthreads) {
            const int r = i / H, u = i % H;
            const float ig = sigmoid_f(g_s[r * G + u]);
            const float fg = sigmoid_f(g_s[r * G + H + u]);
            const float gg = tanhf(g_s[r * G + 2 * H + u]);
            const float og = sigmoid_f(g_s[r * G + 3 * H + u]);
            const float c = fg * c_s[i] + ig * gg;
            const float h = og * tanhf(c);
            c_s[i] = c;
            h_s[i] = h;
            if (b0 + r < B) {
                const size_t row = (size_t)(b0 + r) * T + t;
                gd[row * G + u] = ig;
                gd[row * G + H + u] = fg;
                gd[row * G + 2 * H + u] = gg;
                gd[row * G + 3 * H + u] = og;
                cd[row * H + u] = c;
                out[row * (2 * H) + dir * H + u] = h;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < TR * H; i += nthreads) {
        const int r = i / H, u = i % H;
        if (b0 + r < B) {
            hn[((size_t)dir * B + b0 + r) * H + u] = h_s[i];
            cn[((size_t)dir * B + b0 + r) * H + u] = c_s[i];
        }
    }
}

// gates: in = activated gates of the forward, out = gradient of the loss w.r.t. the gate pre-activations (dG).
// Visits the steps in the reverse of the forward order; carries dL/dh (recurrent part) and dL/dc in shared memory.
__global__ void __launch_bounds__(TRAIN_THREADS *TQ)
lstm_train_bwd_kernel(float *__restrict__ gates, const float *__restrict__ cells, const float *__restrict__ w0,
                      const float *__restrict__ w1, const float *__restrict__ c0, const float *__restrict__ d_out,
                      const float *__restrict__ d_hn, const float *__restrict__ d_cn, long long B, long long T, int H,
                      float *__restrict__ dh0, float *__restrict__ dc0)
{
    extern __shared__ float sm[];
    float *dh_s = sm;                       // [TR][H]   recurrent part of dL/dh_t
    float *dc_s = dh_s + TR * H;            // [TR][H]   dL/dc_t carried from the later step
    float *dg_s = dc_s + TR * H;            // [TR][4H]  this step's pre-activation gradients
    float *part_s = dg_s + TR * 4 * H;      // [TQ][TR][H] partial sums of dG W_hh
    const int tid = threadIdx.x, nthreads = TRAIN_THREADS * TQ;
    const int dir = blockIdx.y;
    const long long b0 = (long long)blockIdx.x * TR;
    const int G = 4 * H;
    const float *w = dir ? w1 : w0;                            // [4H][H] torch layout
    float *gd = gates + (size_t)dir * B * T * G;
    const float *cd = cells + (size_t)dir * B * T * H;

    for (int i = tid; i < TR * H; i += nthreads) {
        const int r = i / H, u = i % H;
        const bool ok = b0 + r < B;
        dh_s[i] = (ok && d_hn) ? d_hn[((size_t)dir * B + b0 + r) * H + u] : 0.f;
        dc_s[i] = (ok && d_cn) ? d_cn[((size_t)dir * B + b0 + r) * H + u] : 0.f;
    }
    __syncthreads();

    for (long long step = T - 1; step >= 0; --step) {         // forward-order index of the step being differentiated
        const long long t = dir ? (T - 1 - step) : step;
        const long long t_prev = dir ? t + 1 : t - 1;          // time of the forward's previous step (step > 0)
        for (int i = tid; i < TR * H; i += nthreads) {
            const int r = i / H, u = i % H;
            float dai = 0.f, daf = 0.f, dag = 0.f, dao = 0.f;
            if (b0 + r < B) {
                const size_t row = (size_t)(b0 + r) * T + t;
                const float ig = gd[row * G + u], fg = gd[row * G + H + u], gg = gd[row * G + 2 * H + u], og = gd[row * G + 3 * H + u];
                const float c = cd[row * H + u];
                const float c_prev = step > 0 ? cd[((size_t)(b0 + r) * T + t_prev) * H + u] : c0[((size_t)dir * B + b0 + r) * H + u];
                const float tc = tanhf(c);
                const float dh = d_out[row * (2 * H) + dir * H + u] + dh_s[i];
                const float dc = dc_s[i] + dh * og * (1.f - tc * tc);
                dao = dh * tc * og * (1.f - og);
                dai = dc * gg * ig * (1.f - ig);
                dag = dc * ig * (1.f - gg * gg);
                daf = dc * c_prev * fg * (1.f - fg);
                dc_s[i] = dc * fg;
                gd[row * G + u] = dai;
                gd[row * G + H + u] = daf;
                gd[row * G + 2 * H + u] = dag;
                gd[row * G + 3 * H + u] = dao;
            }
            dg_s[r * G + u] = dai;
            dg_s[r * G + H + u] = daf;
            dg_s[r * G + 2 * H + u] = dag;
            dg_s[r * G + 3 * H + u] = dao;
        }
        __syncthreads();
        // dL/dh_{prev}[r][j] = sum_n dG[r][n] W_hh[n][j]; thread group q takes the n = q (mod TQ) rows, lanes run over j
        {
            const int q = tid / TRAIN_THREADS, j = tid % TRAIN_THREADS;
            if (j < H) {
                float acc[TR] = {};
#pragma unroll 8
                for (int n = q; n < G; n += TQ) {
                    const float wv = __ldg(w + (size_t)n * H + j);
#pragma unroll
                    for (int r = 0; r < TR; ++r) acc[r] = fmaf(dg_s[r * G + n], wv, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < TR; ++r) part_s[(q * TR + r) * H + j] = acc[r];
            }
        }
        __syncthreads();
        for (int i = tid; i < TR * H; i += nthreads) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < TQ; ++q) s += part_s[q * TR * H + i];
            dh_s[i] = s;
        }
        __syncthreads();
    }
    for (int i = tid; i < TR * H; i += nthreads) {
        const int r = i / H, u = i % H;
        if (b0 + r < B) {
            dh0[((size_t)dir * B + b0 + r) * H + u] = dh_s[i];
            dc0[((size_t)dir * B + b0 + r) * H + u] = dc_s[i];
        }
    }
}

// HSSB_TRAIN_IMPL=stream forces the generic kernels below (weights re-read from L2 every step) where the cluster-resident
// ones of lstm_train_cluster.cu would run, =gather the all-gather form of the cluster backward: the cross-implementation
// test uses both.
bool use_cluster_kernels(int H)
{
    const char *e = getenv("HSSB_TRAIN_IMPL");
    return train_cluster_supported(H) && !(e && e[0] == 's');
}

int check_train_args(const char *what, int64_t B, int64_t T, int H)
{
    if (B < 0 || T < 0 || H < 1) return fail(HSSB_E_SHAPE, "%s: B=%lld T=%lld H=%d", what, (long long)B, (long long)T, H);
    if (H > TRAIN_THREADS) return fail(HSSB_E_MODEL, "%s: hidden_size %d > %d unsupported", what, H, TRAIN_THREADS);
    if ((B + TR - 1) / TR > 0x7fffffffLL) return fail(HSSB_E_SHAPE, "%s: B=%lld too large", what, (long long)B);
    return 0;
}

}  // namespace
}  // namespace hssb

using namespace hssb;

extern "C" int hssb_lstm_train_forward(float *gates, const float *w_hhT_fwd, const float *w_hhT_rev, const float *h0, const float *c0,
                                       int64_t B, int64_t T, int H, float *out, float *cells, float *hn, float *cn, void *stream)
{
    if (!gates || !w_hhT_fwd || !w_hhT_rev || !h0 || !c0 || !out || !cells || !hn || !cn)
        return fail(HSSB_E_NULL, "hssb_lstm_train_forward: null pointer");
    if (int rc = check_train_args("hssb_lstm_train_forward", B, T, H)) return rc;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    if (B == 0) return 0;
    if (T == 0) {       // no steps: the final state is the initial state
        HSSB_CUDA_OK(cudaMemcpyAsync(hn, h0, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToDevice, st));
        HSSB_CUDA_OK(cudaMemcpyAsync(cn, c0, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    if (use_cluster_kernels(H)) return train_fwd_cluster_launch(gates, w_hhT_fwd, w_hhT_rev, h0, c0, B, T, out, cells, hn, cn, st);
    const size_t smem = sizeof(float) * (size_t)TR * H * 6;
    HSSB_CUDA_OK(cudaFuncSetAttribute(lstm_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((B + TR - 1) / TR), 2);
    ProfScope prof("lstm_train_fwd", st);
    lstm_train_fwd_kernel<<<grid, TRAIN_THREADS * TQ, smem, st>>>(gates, w_hhT_fwd, w_hhT_rev, h0, c0, B, T, H, out, cells, hn, cn);
    HSSB_LAUNCH_OK("lstm_train_fwd_kernel");
    return 0;
}

extern "C" int hssb_lstm_train_backward(float *gates, const float *cells, const float *w_hh_fwd, const float *w_hh_rev, const float *c0,
                                        const float *d_out, const float *d_hn, const float *d_cn, int64_t B, int64_t T, int H,
                                        float *dh0, float *dc0, void *stream)
{
    if (!gates || !cells || !w_hh_fwd || !w_hh_rev || !c0 || !d_out || !dh0 || !dc0)
        return fail(HSSB_E_NULL, "hssb_lstm_train_backward: null pointer");
    if (int rc = check_train_args("hssb_lstm_train_backward", B, T, H)) return rc;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    if (B == 0) return 0;
    if (T == 0) {       // identity between the initial and the final state
        if (d_hn) HSSB_CUDA_OK(cudaMemcpyAsync(dh0, d_hn, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToDevice, st));
        else HSSB_CUDA_OK(cudaMemsetAsync(dh0, 0, sizeof(float) * 2 * B * H, st));
        if (d_cn) HSSB_CUDA_OK(cudaMemcpyAsync(dc0, d_cn, sizeof(float) * 2 * B * H, cudaMemcpyDeviceToDevice, st));
        else HSSB_CUDA_OK(cudaMemsetAsync(dc0, 0, sizeof(float) * 2 * B * H, st));
        return 0;
    }
    if (use_cluster_kernels(H)) {
        const char *e = getenv("HSSB_TRAIN_IMPL");       // "gather": the all-gather backward (3x slower; kept as a cross-check)
        return train_bwd_cluster_launch(gates, cells, w_hh_fwd, w_hh_rev, c0, d_out, d_hn, d_cn, B, T, dh0, dc0, e && e[0] == 'g', st);
    }
    const size_t smem = sizeof(float) * (size_t)TR * H * (6 + TQ);
    HSSB_CUDA_OK(cudaFuncSetAttribute(lstm_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((B + TR - 1) / TR), 2);
    ProfScope prof("lstm_train_bwd", st);
    lstm_train_bwd_kernel<<<grid, TRAIN_THREADS * TQ, smem, st>>>(gates, cells, w_hh_fwd, w_hh_rev, c0, d_out, d_hn, d_cn, B, T, H, dh0, dc0);
    HSSB_LAUNCH_OK("lstm_train_bwd_kernel");
    return 0;
}

extern "C" size_t hssb_lstm_train_backward_tc_workspace_bytes(void) { return bptt_tc_workspace_bytes(); }

extern "C" int hssb_lstm_train_backward_tc(const float *gates, float *dG, float *dG_hi, float *dG_lo, float *db, const float *cells, const float *w_hh_fwd, const float *w_hh_rev, const float *c0,
                                           const float *d_out, const float *d_hn, const float *d_cn, int64_t B, int64_t T,
                                           float *dh0, float *dc0, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!gates || !cells || !w_hh_fwd || !w_hh_rev || !c0 || !d_out || !dh0 || !dc0)
        return fail(HSSB_E_NULL, "hssb_lstm_train_backward_tc: null pointer");
    if ((!dG && !dG_hi) || (!dG_hi != !dG_lo)) return fail(HSSB_E_NULL, "hssb_lstm_train_backward_tc: need dG and / or the (dG_hi, dG_lo) pair");
    if (int rc = check_train_args("hssb_lstm_train_backward_tc", B, T, 240)) return rc;
    if (B == 0 || T == 0) {     // nothing for the tensor cores to do: the generic entry point handles the degenerate shapes (no dG rows)
        if (db) HSSB_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * 2 * 960, as_stream(stream)));
        return hssb_lstm_train_backward(dG ? dG : dG_hi, cells, w_hh_fwd, w_hh_rev, c0, d_out, d_hn, d_cn, B, T, 240, dh0, dc0, stream);
    }
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(HSSB_E_WORKSPACE, "hssb_lstm_train_backward_tc: workspace must be 256-byte aligned");
    if (int rc = require_sm100()) return rc;
    return bptt_tc_backward(gates, dG, dG_hi, dG_lo, db, cells, w_hh_fwd, w_hh_rev, c0, d_out, d_hn, d_cn, B, T, dh0, dc0, workspace, workspace_bytes, as_stream(stream));
}
