// Fused training-step kernels around the recurrences (reference main.py:67-82,130-135,222-230):
//
//   hssb_ce_head_forward / _backward   linear(480 -> 4) + log_softmax + nn.CrossEntropyLoss on the permuted output
//                                      (segmenter.py:86-87 + main.py:69-70) and its gradient w.r.t. the activations, the linear
//                                      weight and bias -- one pass each instead of linear / log_softmax / permute / cross_entropy
//                                      and their five backward kernels
//   hssb_clip_adam_step                gradient clipping by global norm (pl.Trainer(gradient_clip_val=1), main.py:226 ==
//                                      clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))) fused with torch.optim.Adam
//                                      (main.py:130, default betas / eps, no weight decay, no amsgrad) over ALL parameter tensors in
//                                      two launches; the learning rate of the step (LambdaLR 0.9^epoch, main.py:133-134) is an argument
//   hssb_split_tf32                    a = hi + lo with hi exactly representable in TF32: the operands of the weight-gradient /
//                                      input-gradient GEMMs of back-propagation (what autograd's fp32 mm kernels compute for nn.LSTM),
//                                      which then run as three TF32 tensor-core GEMMs hi.hi + hi.lo + lo.hi with fp32 accumulation
//                                      (product error 2^-21) instead of SIMT fp32 ones
// fp32, same arithmetic order as the torch ops they replace up to the reduction order of the sums.
#include "hssb_common.cuh"
#include <algorithm>
#include <cmath>

namespace hssb {

constexpr int TRAIN_MAX_TENSORS = 32;
struct AdamTable {
    float *p[TRAIN_MAX_TENSORS];
    const float *g[TRAIN_MAX_TENSORS];
    float *m[TRAIN_MAX_TENSORS];
    float *v[TRAIN_MAX_TENSORS];
    long long start[TRAIN_MAX_TENSORS + 1];      // prefix sums of the element counts: tensor i owns [start[i], start[i+1]) of the flat index
    int n;
};

__device__ __forceinline__ int tensor_of(const AdamTable &t, long long i)
{
    int lo = 0, hi = t.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.start[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// sum of squares of all gradients -> norm2[0] (double, zeroed by the host)
__global__ void __launch_bounds__(256) grad_norm_kernel(const __grid_constant__ AdamTable t, double *__restrict__ norm2)
{
    const long long total = t.start[t.n];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = tensor_of(t, i);
        const float g = t.g[k][i - t.start[k]];
        acc += (double)g * (double)g;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ double warp_acc[8];
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += warp_acc[w];
        atomicAdd(norm2, a);
    }
}

// g <- g * clip;  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;  p <- p - step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) clip_adam_kernel(const __grid_constant__ AdamTable t, const double *__restrict__ norm2, float max_norm,
                                                         float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float *__restrict__ norm_out)
{
    const long long total = t.start[t.n];
    const float norm = (float)sqrt(*norm2);
    float clip = 1.0f;
    if (max_norm > 0.f) clip = fminf(max_norm / (norm + 1e-6f), 1.0f);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    const float step_size = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = tensor_of(t, i);
        const long long j = i - t.start[k];
        const float g = t.g[k][j] * clip;
        const float m = beta1 * t.m[k][j] + (1.0f - beta1) * g;
        const float v = beta2 * t.v[k][j] + (1.0f - beta2) * g * g;
        t.m[k][j] = m;
        t.v[k][j] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[k][j] -= step_size * (m / denom);
    }
}

// ------------------------------------------------------------------------------------------------
// head + loss.  act [M, K] (K = 2H, after ReLU / dropout), w [4, K], b [4], target [M] int64.
// forward: logp [M, 4] (kept for the backward and the metrics), loss_sum += sum_rows (lse(logp) - logp[target]); rows with a
// target outside 0..3 do not contribute (none in the reference's data).  One warp per row, 16-byte loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_head_forward_kernel(const float *__restrict__ act, long long M, int K, const float *__restrict__ w,
                                                               const float *__restrict__ b, const int64_t *__restrict__ target,
                                                               float *__restrict__ logp, double *__restrict__ loss_sum)
{
    extern __shared__ float w_s[];                      // [4][K]
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) w_s[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    double loss = 0.0;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        const float *a = act + (size_t)row * K;
        for (int k = lane; k < K; k += 32) {
            const float v = __ldg(a + k);
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] = fmaf(v, w_s[c * K + k], s[c]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
        if (lane == 0) {
            const float z[4] = {s[0] + b[0], s[1] + b[1], s[2] + b[2], s[3] + b[3]};
            const float mx = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
            const float lse = mx + logf(expf(z[0] - mx) + expf(z[1] - mx) + expf(z[2] - mx) + expf(z[3] - mx));
            const float lp[4] = {z[0] - lse, z[1] - lse, z[2] - lse, z[3] - lse};
            *reinterpret_cast<float4 *>(logp + (size_t)row * 4) = make_float4(lp[0], lp[1], lp[2], lp[3]);
            const long long t = target[row];
            if (t >= 0 && t < 4) {
                // CrossEntropyLoss applies log_softmax again to its input (the log-probabilities)
                const float m2 = fmaxf(fmaxf(lp[0], lp[1]), fmaxf(lp[2], lp[3]));
                const float lse2 = m2 + logf(expf(lp[0] - m2) + expf(lp[1] - m2) + expf(lp[2] - m2) + expf(lp[3] - m2));
                loss += (double)(lse2 - lp[(int)t]);
            }
        }
    }
    if (lane == 0 && loss != 0.0) atomicAdd(loss_sum, loss);
}

// backward of mean-reduced CE(log_softmax(log_softmax(z))) w.r.t. z: (softmax(z) - onehot) * scale  (log_softmax is idempotent),
// then d_act = dz . W, dW += dz^T . act, db += sum dz.  grid-stride over rows; dW / db accumulate per block in shared memory.
__global__ void __launch_bounds__(256) ce_head_backward_kernel(const float *__restrict__ act, const float *__restrict__ logp, long long M, int K,
                                                                const float *__restrict__ w, const int64_t *__restrict__ target, float scale,
                                                                const float *__restrict__ upstream, float *__restrict__ d_act,
                                                                float *__restrict__ d_w, float *__restrict__ d_b)
{
    if (upstream) scale *= __ldg(upstream);             // dL/d(loss) left on the device: no host round trip in the backward pass
    extern __shared__ float sm[];                       // w_s [4][K], dw_s [4][K], db_s [4]
    float *w_s = sm, *dw_s = sm + 4 * K, *db_s = sm + 8 * K;
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) { w_s[i] = w[i]; dw_s[i] = 0.f; }
    if (threadIdx.x < 4) db_s[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    // every warp keeps its share of dW in registers: columns lane, lane + 32, ... (K <= 32 * 16)
    float dw[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) dw[c][i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
        const float4 lp = *reinterpret_cast<const float4 *>(logp + (size_t)row * 4);
        const long long t = target[row];
        float dz[4] = {0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < 4) {
            dz[0] = (expf(lp.x) - (t == 0 ? 1.f : 0.f)) * scale;
            dz[1] = (expf(lp.y) - (t == 1 ? 1.f : 0.f)) * scale;
            dz[2] = (expf(lp.z) - (t == 2 ? 1.f : 0.f)) * scale;
            dz[3] = (expf(lp.w) - (t == 3 ? 1.f : 0.f)) * scale;
        }
        const float *a = act + (size_t)row * K;
        float *da = d_act + (size_t)row * K;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = lane + 32 * i;
            if (k < K) {
                const float v = __ldg(a + k);
                da[k] = dz[0] * w_s[k] + dz[1] * w_s[K + k] + dz[2] * w_s[2 * K + k] + dz[3] * w_s[3 * K + k];
#pragma unroll
                for (int c = 0; c < 4; ++c) dw[c][i] = fmaf(dz[c], v, dw[c][i]);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) db[c] += dz[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = lane + 32 * i;
            if (k < K) atomicAdd(&dw_s[c * K + k], dw[c][i]);
        }
        if (lane == 0) atomicAdd(&db_s[c], db[c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) atomicAdd(d_w + i, dw_s[i]);
    if (threadIdx.x < 4) atomicAdd(d_b + threadIdx.x, db_s[threadIdx.x]);
}

}  // namespace hssb

using namespace hssb;

extern "C" int hssb_clip_adam_step(int n_tensors, float *const *params, const float *const *grads, float *const *exp_avg,
                                   float *const *exp_avg_sq, const int64_t *numel, float lr, float beta1, float beta2, float eps,
                                   int64_t step, float max_norm, double *norm2_scratch, float *grad_norm_out, void *stream)
{
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !norm2_scratch) return fail(HSSB_E_NULL, "hssb_clip_adam_step: null pointer");
    if (n_tensors < 1 || n_tensors > TRAIN_MAX_TENSORS) return fail(HSSB_E_SHAPE, "hssb_clip_adam_step: %d tensors (1..%d)", n_tensors, TRAIN_MAX_TENSORS);
    if (step < 1) return fail(HSSB_E_SHAPE, "hssb_clip_adam_step: step counts from 1");
    AdamTable t;
    t.n = n_tensors;
    t.start[0] = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || numel[i] < 0) return fail(HSSB_E_NULL, "hssb_clip_adam_step: tensor %d", i);
        t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i];
        t.start[i + 1] = t.start[i] + numel[i];
    }
    if (t.start[n_tensors] == 0) return 0;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    HSSB_CUDA_OK(cudaMemsetAsync(norm2_scratch, 0, sizeof(double), st));
    const long long total = t.start[n_tensors];
    int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 4);
    {
        ProfScope prof("grad_norm", st);
        grad_norm_kernel<<<blocks, 256, 0, st>>>(t, norm2_scratch);
        HSSB_LAUNCH_OK("grad_norm_kernel");
    }
    const float bc1 = 1.0f - (float)std::pow((double)beta1, (double)step);
    const float bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)beta2, (double)step));
    ProfScope prof("clip_adam", st);
    clip_adam_kernel<<<blocks, 256, 0, st>>>(t, norm2_scratch, max_norm, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_norm_out);
    HSSB_LAUNCH_OK("clip_adam_kernel");
    return 0;
}

extern "C" int hssb_ce_head_forward(const float *act, int64_t M, int K, const float *w, const float *b, const int64_t *target,
                                    float *logp, double *loss_sum, void *stream)
{
    if (!act || !w || !b || !target || !logp || !loss_sum) return fail(HSSB_E_NULL, "hssb_ce_head_forward: null pointer");
    if (M < 0 || K < 1 || K > 512) return fail(HSSB_E_SHAPE, "hssb_ce_head_forward: M=%lld K=%d (K <= 512)", (long long)M, K);
    if (M == 0) return 0;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    const int blocks = (int)std::min<long long>((M + 7) / 8, 148 * 8);
    ProfScope prof("ce_head_fwd", st);
    ce_head_forward_kernel<<<blocks, 256, sizeof(float) * 4 * K, st>>>(act, M, K, w, b, target, logp, loss_sum);
    HSSB_LAUNCH_OK("ce_head_forward_kernel");
    return 0;
}

extern "C" int hssb_ce_head_backward(const float *act, const float *logp, int64_t M, int K, const float *w, const int64_t *target,
                                     float scale, const float *upstream, float *d_act, float *d_w, float *d_b, void *stream)
{
    if (!act || !logp || !w || !target || !d_act || !d_w || !d_b) return fail(HSSB_E_NULL, "hssb_ce_head_backward: null pointer");
    if (M < 0 || K < 1 || K > 512) return fail(HSSB_E_SHAPE, "hssb_ce_head_backward: M=%lld K=%d (K <= 512)", (long long)M, K);
    if (M == 0) return 0;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    const int blocks = (int)std::min<long long>((M + 7) / 8, 148 * 2);
    ProfScope prof("ce_head_bwd", st);
    ce_head_backward_kernel<<<blocks, 256, sizeof(float) * (8 * K + 4), st>>>(act, logp, M, K, w, target, scale, upstream, d_act, d_w, d_b);
    HSSB_LAUNCH_OK("ce_head_backward_kernel");
    return 0;
}

namespace hssb {
// hi = a rounded to TF32 (10 explicit mantissa bits, round half away from zero; inf / NaN pass through), lo = a - hi (exact in fp32; 0 for inf)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4 *__restrict__ a, long long n4, const float *__restrict__ tail_src,
                                                         int tail, float4 *__restrict__ hi, float4 *__restrict__ lo, float *__restrict__ tail_hi,
                                                         float *__restrict__ tail_lo)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(a + i);
        float4 h, l;
        tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
        hi[i] = h;
        lo[i] = l;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
        float h, l;
        tf32_split(tail_src[threadIdx.x], h, l);
        tail_hi[threadIdx.x] = h;
        tail_lo[threadIdx.x] = l;
    }
}
}  // namespace hssb

extern "C" int hssb_split_tf32(const float *a, int64_t n, float *hi, float *lo, void *stream)
{
    using namespace hssb;
    if (n < 0) return fail(HSSB_E_SHAPE, "hssb_split_tf32: n=%lld", (long long)n);
    if (n == 0) return 0;
    if (!a || !hi || !lo) return fail(HSSB_E_NULL, "hssb_split_tf32: null pointer");
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15)
        return fail(HSSB_E_SHAPE, "hssb_split_tf32: pointers must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);
    const long long n4 = n / 4;
    const int tail = (int)(n % 4);
    ProfScope prof("split_tf32", st);
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, 148 * 16));
    split_tf32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(a), n4, a + 4 * n4, tail, reinterpret_cast<float4 *>(hi),
                                              reinterpret_cast<float4 *>(lo), hi + 4 * n4, lo + 4 * n4);
    HSSB_LAUNCH_OK("split_tf32_kernel");
    return 0;
}
