// Shared helpers for libhssb.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include "../../include/hssb.h"

namespace hssb {

// thread-local message of the last failure (returned by hssb_last_error()).
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);          // sets the message, returns code
int cuda_fail(cudaError_t e, const char *what);    // sets the message, returns (int)e

#define HSSB_CUDA_OK(expr)                                                    \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::hssb::cuda_fail(_e, #expr);           \
    } while (0)

#define HSSB_LAUNCH_OK(name)                                                  \
    do {                                                                      \
        cudaError_t _e = cudaGetLastError();                                  \
        if (_e != cudaSuccess) return ::hssb::cuda_fail(_e, name);            \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// fails with HSSB_E_DEVICE unless the current device is compute capability 10.x
int require_sm100();

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One cached int per CUDA device (0 = not yet known).  Function attributes (opt-in shared memory) and occupancy figures
// (co-resident clusters) belong to a device / context, not to the process: a second GPU used by the same process needs its own.
constexpr int HSSB_MAX_DEVICES = 64;
struct PerDeviceInt {
    std::atomic<int> v[HSSB_MAX_DEVICES];
    static int slot() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < HSSB_MAX_DEVICES) ? d : 0; }
    int get() const { return v[slot()].load(std::memory_order_acquire); }
    void set(int x) { v[slot()].store(x, std::memory_order_release); }
};

#ifdef __CUDACC__
// v = hi + lo with hi exactly representable in TF32 (10 explicit mantissa bits, round half away from zero; inf / NaN pass through
// with lo = 0 / NaN): the operand split of the three-pass TF32 GEMMs of back-propagation (hssb_split_tf32, K5b)
__device__ __forceinline__ void tf32_split(float v, float &hi, float &lo)
{
    const unsigned u = __float_as_uint(v);
    hi = ((u & 0x7f800000u) == 0x7f800000u) ? v : __uint_as_float((u + 0x1000u) & 0xffffe000u);
    lo = v == hi ? 0.0f : v - hi;
}
#endif

// Optional per-launch timing (hssb_prof_*): a pair of CUDA events recorded on the launching stream
// around one kernel launch.  Costs nothing when disabled.
struct ProfScope {
    ProfScope(const char *name, cudaStream_t st);
    ~ProfScope();
    int slot;
    cudaStream_t st;
};

}  // namespace hssb
