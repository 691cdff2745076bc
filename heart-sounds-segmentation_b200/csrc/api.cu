// libhssb.so: error reporting, device check, metric counters.
#include "hssb_common.cuh"
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace hssb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int require_sm100()
{
    static thread_local int checked_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HSSB_E_DEVICE, "no CUDA device: %s", cudaGetErrorString(e)); }
    if (dev == checked_dev) return 0;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HSSB_E_DEVICE, "cannot query device %d: %s", dev, cudaGetErrorString(e)); }
    if (major != 10) return fail(HSSB_E_DEVICE, "device %d is sm_%dx; libhssb is built for sm_100a only (no fallback)", dev, major);
    checked_dev = dev;
    return 0;
}

// ---- per-launch timing -------------------------------------------------------------------------
namespace {
struct ProfRec { const char *name; cudaEvent_t start, stop; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<bool> g_prof_on{false};
}  // namespace

ProfScope::ProfScope(const char *name, cudaStream_t s) : slot(-1), st(s)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, st);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].stop, st);
}

// 4x4 confusion counts (replaces the torchmetrics state of reference main.py:36-62).
__global__ void confusion_kernel(const int32_t *__restrict__ pred, const int64_t *__restrict__ target,
                                 long long n, unsigned long long *__restrict__ cm)
{
    __shared__ unsigned int local[16];
    if (threadIdx.x < 16) local[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = pred[i];
        const long long t = target[i];
        if (p >= 0 && p < 4 && t >= 0 && t < 4) atomicAdd(&local[(int)t * 4 + p], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 16 && local[threadIdx.x]) atomicAdd(&cm[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

// K7 complete ("confusion matrix (+ NLL sum, count) ... <= 32 scalars"): the whole metric state of one evaluation
// step from the log-probabilities and targets -- 16 confusion counts, the summed loss and the element count -- as 18 doubles
// (counts are exact below 2^53), so ONE all-reduce(SUM) merges ranks.  The loss term restates what the reference feeds its
// logger (main.py:69-70,91-92,112-113): nn.CrossEntropyLoss on the permuted log-probabilities, i.e. a second log-softmax
// over the four log-probabilities, lse(logp) - logp[target]; labels = first maximum of logp unless `pred` is given.
__global__ void __launch_bounds__(256)
metrics_kernel(const float4 *__restrict__ logp, const int32_t *__restrict__ pred, const int64_t *__restrict__ target, long long n,
               double *__restrict__ state)
{
    __shared__ unsigned int local[16];
    __shared__ double warp_nll[8];
    __shared__ unsigned int warp_cnt[8];
    if (threadIdx.x < 16) local[threadIdx.x] = 0;
    __syncthreads();
    double nll = 0.0;
    unsigned int cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long t = target[i];
        const float4 v = __ldcs(&logp[i]);
        const float lp[4] = {v.x, v.y, v.z, v.w};
        int p = 0;
        if (pred) p = pred[i];
        else {
            float best = lp[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) if (lp[c] > best) { best = lp[c]; p = c; }
        }
        if (t >= 0 && t < 4) {
            if (p >= 0 && p < 4) atomicAdd(&local[(int)t * 4 + p], 1u);
            const float mx = fmaxf(fmaxf(lp[0], lp[1]), fmaxf(lp[2], lp[3]));
            const float lse = mx + logf(expf(lp[0] - mx) + expf(lp[1] - mx) + expf(lp[2] - mx) + expf(lp[3] - mx));
            nll += (double)(lse - lp[(int)t]);
            ++cnt;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        nll += __shfl_xor_sync(0xffffffffu, nll, s);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    }
    if ((threadIdx.x & 31) == 0) { warp_nll[threadIdx.x >> 5] = nll; warp_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 16 && local[threadIdx.x]) atomicAdd(&state[threadIdx.x], (double)local[threadIdx.x]);
    if (threadIdx.x == 0) {
        double a = 0.0;
        unsigned int c = 0;
        for (int w = 0; w < 8; ++w) { a += warp_nll[w]; c += warp_cnt[w]; }
        if (c) { atomicAdd(&state[16], a); atomicAdd(&state[17], (double)c); }
    }
}

// One-vs-rest score histograms for the binned multiclass AUROC (reference main.py:48,60: torchmetrics AUROC on the
// class probabilities).  hist[c][target == c][bin(exp(logp[c]))] += 1 -- plain counters, so shards and ranks add up
// (all-reduce) exactly like the confusion counts.  Block-private shared histograms, one warp-aggregated shared atomic
// per distinct bin of a warp (trained models put most scores into the two end bins), one global atomic per non-empty bin.
__global__ void __launch_bounds__(512)
auroc_hist_kernel(const float4 *__restrict__ logp, const int64_t *__restrict__ target, long long n, int nbins,
                  unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned int sh[];   // [4][2][nbins]
    const int words = 8 * nbins;
    for (int i = threadIdx.x; i < words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const float scale = (float)nbins;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n + stride - 1) / stride;   // whole warps stay converged for __match_any_sync
    for (long long r = 0; r < rounds; ++r) {
        const long long i = first + r * stride;
        const bool live = i < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        long long t = -1;
        if (live) {
            v = __ldcs(&logp[i]);
            t = target[i];
        }
        const float lp[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float p = expf(lp[c]);
            int b = (int)(p * scale);
            b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
            if (!(p == p)) b = 0;              // NaN score: lowest bin (never ranks above anything)
            const int key = live && t >= 0 && t < 4 ? (c * 2 + (t == c ? 1 : 0)) * nbins + b : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[key], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < words; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

}  // namespace hssb

extern "C" int hssb_version(void) { return HSSB_VERSION; }
extern "C" const char *hssb_last_error(void) { return hssb::g_err; }

extern "C" int hssb_prof_enable(int on)
{
    std::lock_guard<std::mutex> lock(hssb::g_prof_mu);
    hssb::g_prof_on = on != 0;
    return 0;
}

// Writes one line per kernel name: "<name> <launches> <total_ms>\n"; clears the records.
extern "C" int hssb_prof_read(char *buf, size_t n)
{
    using namespace hssb;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> acc;
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            auto &a = acc[r.name];
            a.first += 1;
            a.second += ms;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof.clear();
    std::string out;
    for (auto &kv : acc) {
        char line[160];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (!buf || n == 0) return (int)out.size();
    snprintf(buf, n, "%s", out.c_str());
    return (int)out.size();
}

extern "C" int hssb_confusion(const int32_t *pred, const int64_t *target, int64_t n, int64_t *cm16, void *stream)
{
    using namespace hssb;
    if (!pred || !target || !cm16) return fail(HSSB_E_NULL, "hssb_confusion: null pointer");
    if (n < 0) return fail(HSSB_E_SHAPE, "hssb_confusion: n=%lld", (long long)n);
    if (n == 0) return 0;
    if (int rc = require_sm100()) return rc;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ProfScope prof("confusion", as_stream(stream));
    confusion_kernel<<<blocks, 256, 0, as_stream(stream)>>>(pred, target, n, reinterpret_cast<unsigned long long *>(cm16));
    HSSB_LAUNCH_OK("confusion_kernel");
    return 0;
}

extern "C" int hssb_metrics_update(const float *logp, const int32_t *pred, const int64_t *target, int64_t n, double *state18, void *stream)
{
    using namespace hssb;
    if (!logp || !target || !state18) return fail(HSSB_E_NULL, "hssb_metrics_update: null pointer");
    if (n < 0) return fail(HSSB_E_SHAPE, "hssb_metrics_update: n=%lld", (long long)n);
    if ((reinterpret_cast<uintptr_t>(logp) & 15) != 0) return fail(HSSB_E_SHAPE, "hssb_metrics_update: logp must be 16-byte aligned");
    if (n == 0) return 0;
    if (int rc = require_sm100()) return rc;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ProfScope prof("metrics", as_stream(stream));
    metrics_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(logp), pred, target, n, state18);
    HSSB_LAUNCH_OK("metrics_kernel");
    return 0;
}

extern "C" int hssb_auroc_hist(const float *logp, const int64_t *target, int64_t n, int nbins, int64_t *hist, void *stream)
{
    using namespace hssb;
    if (!logp || !target || !hist) return fail(HSSB_E_NULL, "hssb_auroc_hist: null pointer");
    if (n < 0) return fail(HSSB_E_SHAPE, "hssb_auroc_hist: n=%lld", (long long)n);
    if (nbins < 2 || nbins > 4096) return fail(HSSB_E_SHAPE, "hssb_auroc_hist: nbins=%d outside [2, 4096]", nbins);
    if ((reinterpret_cast<uintptr_t>(logp) & 15) != 0) return fail(HSSB_E_SHAPE, "hssb_auroc_hist: logp must be 16-byte aligned");
    if (n == 0) return 0;
    if (int rc = require_sm100()) return rc;
    const size_t smem = (size_t)8 * nbins * sizeof(unsigned int);
    HSSB_CUDA_OK(cudaFuncSetAttribute(auroc_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 * 4));
    long long blocks = (n + 511) / 512;
    if (blocks > 148) blocks = 148;          // one 128 KB histogram per SM at nbins = 4096
    ProfScope prof("auroc_hist", as_stream(stream));
    auroc_hist_kernel<<<(int)blocks, 512, smem, as_stream(stream)>>>(reinterpret_cast<const float4 *>(logp), target, n, nbins,
                                                                      reinterpret_cast<unsigned long long *>(hist));
    HSSB_LAUNCH_OK("auroc_hist_kernel");
    return 0;
}
