// fp32 SIMT kernels of the BiLSTM segmenter: plain-FMA input projection, recurrence and head.
// The recurrence/in-projection here are the on-device VALIDATION path (impl = 1) for the tcgen05
// kernels of lstm_tc.cu; the head kernel (K6) is shared by both paths.
// Replaces reference hss/model/segmenter.py:80-87 (nn.LSTM x2, ReLU, Linear, LogSoftmax), eval mode.
#include "model.cuh"

namespace hssb {

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * Wt[K,N] + bias[N]      (64x64 tile, 256 threads, 4x4 per thread)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
simt_inproj_kernel(const float *__restrict__ A, long long M, int K, const float *__restrict__ Wt,
                   const float *__restrict__ bias, int N, float *__restrict__ C)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            const int r = i / 16, kk = i % 16;
            const long long m = m0 + r;
            As[kk][r] = (m < M && k0 + kk < K) ? A[m * K + k0 + kk] : 0.f;
        }
        for (int i = tid; i < 16 * 64; i += 256) {
            const int kk = i / 64, c = i % 64;
            Bs[kk][c] = (k0 + kk < K && n0 + c < N) ? Wt[(size_t)(k0 + kk) * N + n0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[m * N + n] = acc[i][j] + bias[n];
        }
    }
}

int simt_inproj(const float *A, int64_t M, int K, const float *Wt, const float *bias, int N, float *C, cudaStream_t st)
{
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    ProfScope prof("simt_inproj", st);
    simt_inproj_kernel<<<grid, 256, 0, st>>>(A, M, K, Wt, bias, N, C);
    HSSB_LAUNCH_OK("simt_inproj_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Recurrence: one CTA = SR batch rows of one direction, T sequential steps.
// ------------------------------------------------------------------------------------------------
constexpr int SR = 4;

__device__ __forceinline__ float sigmoid_accurate(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(256)
simt_recurrent_kernel(const float *__restrict__ xproj, const float *__restrict__ w0, const float *__restrict__ w1,
                      const float *__restrict__ h0, const float *__restrict__ c0, long long B, long long T, int H,
                      float *__restrict__ out, float *__restrict__ hn, float *__restrict__ cn)
{
    extern __shared__ float sm[];
    float *h_s = sm;                    // [SR][H]
    float *c_s = h_s + SR * H;          // [SR][H]
    float *g_s = c_s + SR * H;          // [SR][4H]
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const long long b0 = (long long)blockIdx.x * SR;
    const int G = 4 * H;
    const float *wT = dir ? w1 : w0;                     // [H][4H]
    const float *xp = xproj + (size_t)dir * B * T * G;  // [B*T][4H]

    for (int i = tid; i < SR * H; i += 256) {
        const int r = i / H, u = i % H;
        const bool ok = b0 + r < B;
        h_s[i] = ok ? h0[((size_t)dir * B + b0 + r) * H + u] : 0.f;
        c_s[i] = ok ? c0[((size_t)dir * B + b0 + r) * H + u] : 0.f;
    }
    __syncthreads();

    for (long long step = 0; step < T; ++step) {
        const long long t = dir ? (T - 1 - step) : step;
        for (int n = tid; n < G; n += 256) {
            float acc[SR];
#pragma unroll
            for (int r = 0; r < SR; ++r) acc[r] = (b0 + r < B) ? xp[((size_t)(b0 + r) * T + t) * G + n] : 0.f;
            for (int k = 0; k < H; ++k) {
                const float w = __ldg(wT + (size_t)k * G + n);
#pragma unroll
                for (int r = 0; r < SR; ++r) acc[r] = fmaf(h_s[r * H + k], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < SR; ++r) g_s[r * G + n] = acc[r];
        }
        __syncthreads();
        for (int i = tid; i < SR * H; i += 256) {
            const int r = i / H, u = i % H;
            const float ig = sigmoid_accurate(g_s[r * G + u]);
            const float fg = sigmoid_accurate(g_s[r * G + H + u]);
            const float gg = tanhf(g_s[r * G + 2 * H + u]);
            const float og = sigmoid_accurate(g_s[r * G + 3 * H + u]);
            const float c = fg * c_s[i] + ig * gg;
            const float h = og * tanhf(c);
            c_s[i] = c;
            h_s[i] = h;
            if (b0 + r < B) out[((size_t)(b0 + r) * T + t) * (2 * H) + dir * H + u] = fmaxf(h, 0.f);
        }
        __syncthreads();
    }
    for (int i = tid; i < SR * H; i += 256) {
        const int r = i / H, u = i % H;
        if (b0 + r < B) {
            hn[((size_t)dir * B + b0 + r) * H + u] = h_s[i];
            cn[((size_t)dir * B + b0 + r) * H + u] = c_s[i];
        }
    }
}

int simt_recurrent(const float *xproj, const float *const w_hhT[2], const float *h0, const float *c0, int64_t B,
                   int64_t T, int H, float *out, float *hn, float *cn, cudaStream_t st)
{
    const size_t smem = sizeof(float) * (size_t)SR * H * 6;
    if (smem > 200 * 1024) return fail(HSSB_E_MODEL, "hidden_size %d too large for the SIMT recurrence", H);
    cudaError_t e = cudaFuncSetAttribute(simt_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(simt_recurrent_kernel)");
    dim3 grid((unsigned)((B + SR - 1) / SR), 2);
    ProfScope prof("simt_recurrent", st);
    simt_recurrent_kernel<<<grid, 256, smem, st>>>(xproj, w_hhT[0], w_hhT[1], h0, c0, B, T, H, out, hn, cn);
    HSSB_LAUNCH_OK("simt_recurrent_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K6 head: logits = act[M,2H] * lin_w[4,2H]^T + lin_b ; log_softmax over 4 ; argmax.
// One warp per row; memory bound (reads 2H floats, writes 4 floats + 1 label per row).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_kernel(const float *__restrict__ act, long long M, int H2, const float *__restrict__ lin_w,
            const float *__restrict__ lin_b, float *__restrict__ logp, int32_t *__restrict__ labels, const int *__restrict__ poison,
            long long T, long long Tp)
{
    // T != Tp: the activations are [B][Tp][H2] with Tp >= T time rows per window (act_pitch), the outputs dense [B][T]
    const bool small = M < 0x7fffffffLL;       // 32-bit division (a 64-bit one costs more than the row's dot products)
    auto src_row = [&](long long r) -> long long {
        if (T == Tp) return r;
        if (small) { const unsigned b = (unsigned)r / (unsigned)T; return (long long)b * Tp + ((unsigned)r - b * (unsigned)T); }
        return (r / T) * Tp + r % T;
    };
    extern __shared__ float w_s[];   // [4][H2]
    // `poison` (nullable): a producer / consumer wait upstream gave up (see POLL_TIMEOUT_NS) -- the activations are not the
    // forward's result, so the outputs are NaN / -1 rather than plausible numbers
    const bool bad = poison && *poison != 0;
    for (int i = threadIdx.x; i < 4 * H2; i += blockDim.x) w_s[i] = lin_w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const bool vec = (H2 & 127) == 0;        // 16-byte loads, two rows in flight per warp (the tcgen05 path: H2 = 512)
    for (long long row0 = 2 * ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); row0 < M; row0 += 2 * warps_total) {
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const int nrow = (row0 + 1 < M) ? 2 : 1;
        if (vec) {
            const float4 *a0 = reinterpret_cast<const float4 *>(act + (size_t)src_row(row0) * H2);
            const float4 *a1 = reinterpret_cast<const float4 *>(act + (size_t)src_row(row0 + nrow - 1) * H2);
            for (int k4 = lane; k4 < H2 / 4; k4 += 32) {
                const float4 v0 = __ldcs(a0 + k4), v1 = __ldcs(a1 + k4);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 w = *reinterpret_cast<const float4 *>(w_s + c * H2 + 4 * k4);
                    s[0][c] = fmaf(v0.w, w.w, fmaf(v0.z, w.z, fmaf(v0.y, w.y, fmaf(v0.x, w.x, s[0][c]))));
                    s[1][c] = fmaf(v1.w, w.w, fmaf(v1.z, w.z, fmaf(v1.y, w.y, fmaf(v1.x, w.x, s[1][c]))));
                }
            }
        } else {
            for (int r = 0; r < nrow; ++r) {
                const float *a = act + (size_t)src_row(row0 + r) * H2;
                for (int k = lane; k < H2; k += 32) {
                    const float v = __ldcs(a + k);
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[r][c] = fmaf(v, w_s[c * H2 + k], s[r][c]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[r][c] += __shfl_xor_sync(0xffffffffu, s[r][c], o);
        if (lane < nrow) {
            const int r = lane;
            const long long row = row0 + r;
            const float z[4] = {(r ? s[1][0] : s[0][0]) + lin_b[0], (r ? s[1][1] : s[0][1]) + lin_b[1], (r ? s[1][2] : s[0][2]) + lin_b[2],
                                (r ? s[1][3] : s[0][3]) + lin_b[3]};
            const float mx = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
            const float lse = mx + logf(expf(z[0] - mx) + expf(z[1] - mx) + expf(z[2] - mx) + expf(z[3] - mx));
            const float o[4] = {z[0] - lse, z[1] - lse, z[2] - lse, z[3] - lse};
            // labels = argmax of the log-probabilities (first maximum, as torch.argmax), which is what
            // the reference's callers take (main.py:69-72 feeds logp to the metrics)
            int arg = 0;
            float best = o[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) if (o[c] > best) { best = o[c]; arg = c; }
            const float nan = __int_as_float(0x7fc00000);
            if (logp) *reinterpret_cast<float4 *>(logp + (size_t)row * 4) = bad ? make_float4(nan, nan, nan, nan) : make_float4(o[0], o[1], o[2], o[3]);
            if (labels) labels[row] = bad ? -1 : arg;
        }
    }
}

int head_forward(const float *act, int64_t M, int H2, const float *lin_w, const float *lin_b, float *logp,
                 int32_t *labels, cudaStream_t st, const int *poison, int64_t T, int64_t Tp)
{
    if (T <= 0 || Tp <= 0) T = Tp = 1;             // dense rows
    if (M == 0) return 0;
    long long blocks = (M + 15) / 16;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = sizeof(float) * 4 * H2;
    if (smem > 48 * 1024) {          // hidden sizes above 1536: opt in to the larger dynamic shared memory (per device, every call)
        cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(head_kernel)");
    }
    ProfScope prof("head", st);
    head_kernel<<<(unsigned)blocks, 256, smem, st>>>(act, M, H2, lin_w, lin_b, logp, labels, poison, T, Tp);
    HSSB_LAUNCH_OK("head_kernel");
    return 0;
}

}  // namespace hssb
