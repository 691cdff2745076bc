// Shared declarations of the tcgen05 BiLSTM kernels (lstm_tc.cu: operand packing, projection GEMM K4 and the forward;
// lstm_rc_mc.cu: the L2-multicast recurrence K5m; lstm_rc_dsmem.cu: the DSMEM recurrences K5 / K5p).
#pragma once
#include "model.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>
#include <algorithm>

namespace hssb {

using namespace ptx;

constexpr int TC_H = 240;
constexpr int TC_G = 960;          // gate rows per direction
constexpr int TC_NG = 2 * TC_G;    // both directions
// "slot layout" of the hidden state handed from one layer to the next: column = dir*256 + rank*32 + slot
// (rank = recurrence CTA 0..7, slot = unit within the rank 0..29; slots 30, 31 are zero).  Every CTA's 8-unit
// k-chunk is then a 16-byte aligned, non-overlapping run, which is what lets the recurrence write its outputs
// with TMA stores straight from the shared-memory image.
constexpr int TC_OP = 512;
// xproj is [dir][t][Bp][960] with an odd number of batch rows per time step: with B = 512 the t stride would be 15 * 2^17 bytes and
// every one of the 128 rows a projection tile writes would fall on the same HBM channel / L2 slice.
static inline long long xproj_pitch(long long B) { return B | 1; }
// The [B][T][512] activations between the layers (relu(h1) planes, relu(h2)) are pitched to an ODD number of time rows per
// window for the same reason: with T = 2000 consecutive windows would be 2 000 KB apart, and the CTAs of the projection GEMM --
// which all work on the same time tile of different windows -- would keep hitting the same HBM channels / L2 slices.
static inline long long act_pitch(long long T) { return T | 1; }
constexpr size_t TC_GATHER_BYTES = (size_t)16 * 8 * 3 * 2 * 4096;   // L2 scratch of the multicast all-gather: [cluster][rank][S][parity][4 KB]


__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// Range of the split: hi = fp16(v) is finite for |v| <= 65504.  Hidden states are in (-1, 1) and weights are checked when they are
// packed, so only the model INPUT can leave that range (un-normalised features): inputs whose magnitude exceeds TC_SPLIT_SAFE are
// pre-scaled by a power of two taken from the tensor's max-abs (exact in fp32) and the projection result is scaled back in the
// fp32 epilogue of K4 -- see tc_forward.  range word 0: "some |x| > TC_SPLIT_SAFE (or non-finite)", word 1: bits of max |x|.
constexpr float TC_SPLIT_SAFE = 32768.0f;
// exponent e >= 0 such that max|x| * 2^-e < 2^15 (0 while the input is in range); 2^-e and 2^e as floats
__device__ __forceinline__ int range_exponent(unsigned amax_bits)
{
    const int ex = (int)((amax_bits >> 23) & 255u) - 127;       // floor(log2(max|x|)); 128 for inf / nan
    const int e = ex - 14;
    return e < 0 ? 0 : (e > 110 ? 110 : e);
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((unsigned)(127 + e) << 23); }


__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long POLL_TIMEOUT_NS = 4000000000ull;     // 4 s: a producer that never shows up must not hang the GPU

constexpr int TC_TT = 128;          // time tile: the projection GEMM's M tile, and the granularity of the producer / consumer flags

// recurrence geometry shared by all variants: 8 CTAs per cluster, 30 hidden units (120 gate rows) per CTA
constexpr int RC_CL = 8;            // CTAs per cluster
constexpr int RC_U = 30;            // real units per CTA
constexpr int RC_KP = 256;          // padded K (8 ranks x 32 slots)
constexpr int RC_XW = 4 * RC_U;     // xproj floats per (t, b) owned by one CTA (120)

struct RecurParams {
    const float *xproj;         // [dir][t][b][g'(960)] fp32 (cluster gate order)
    const __half *whh;          // [dir][rank][plane][128][256] fp16 (cluster gate order, zero padded)
    const float *h0, *c0;       // [2][B][240]
    float *hn, *cn;             // [2][B][240]  raw final state
    __half *out_hi, *out_lo;    // layer 1: relu(h) planes [B*T][480]  (nullptr for layer 2)
    float *out_f32;             // layer 2: relu(h) [B*T][480]         (nullptr for layer 1)
    long long B, T;
    long long Tp;               // time rows per window of the [B][Tp][512] outputs (act_pitch(T))
    long long Bp;               // row pitch of xproj in batch rows (xproj_pitch(B))
    int b_base;                 // first batch column handled by this launch
    int groups;                 // groups of S*NB columns per direction in this launch
    int stagger_ns;             // initial phase offset between the sub-tiles of a cluster
    unsigned long long *trace;  // diagnostic (hssb_debug_trace): clock64 stamps of cluster 0 / rank 0, or nullptr
    int trace_steps;
    // pair kernel: TMA stores of relu(h) into the slot-layout outputs [B][T][512] (fp16 hi, lo planes or one fp32 tensor)
    alignas(64) CUtensorMap out_map[2];
    alignas(64) CUtensorMap out_map16[2];   // the same with boxes of 16 batch columns (two epilogue warps per quadrant)
    // fused layer-1 input projection (multicast kernel): x planes [t][32-column tile][chunk 8][32 cols][8] fp16 (hi, lo) -- the
    // operand of one sub-tile and step is one contiguous 3 KB run --, W_ih slices [dir][rank][plane][128 rows in fragment
    // order][64] fp16 and b_ih + b_hh [dir][rank][128] in the same row order
    const __half *x_hi, *x_lo;
    long long x_tiles;          // 32-column tiles per time step = ceil(B / 32)
    const __half *wih0;
    const float *bias0;
    unsigned char *gather;      // multicast kernel: L2 scratch [cluster][rank][S][2][4 KB] of the all-gather
    int debug;                  // HSSB_RC_DEBUG knock-out switches for timing experiments (results are wrong when set)
    int layer;
    // input-range guard of the fused layer-1 path: the launch is a no-op when (*skip_flag != 0) == skip_when (nullptr: always runs)
    const int *skip_flag;
    int skip_when;
    // producer / consumer flags of the overlapped layer-2 projection (tc_forward).  chunk_done (nullable): the xproj of time tile tt
    // and this direction may be read once chunk_done[q] reached chunk_need (q = 2 tt forward, 2 (t_tiles - 1 - tt) + 1 reverse);
    // tile_done (nullable): bumped per epilogue warp when its relu(h) stores of a time tile are complete, [dir][t_tiles];
    // resident (nullable): [0] bumped once per CTA, [1] set when all CTAs of a launch are on the machine; timeout_flag: raised instead of hanging
    // training forward (TRAIN variants of K5m, hssb_lstm_train_forward_tc): instead of the relu'd outputs the kernel keeps what
    // back-propagation needs, in the layouts of hssb_lstm_train_backward -- activated gates [2][B*T][960] (row b*T + t, gate
    // order i,f,g,o), cell states [2][B*T][240], raw h [B][T][480]
    float *tr_gates, *tr_cells, *tr_out;
    const unsigned *chunk_done;
    unsigned chunk_need;
    unsigned *tile_done;
    unsigned *resident;
    int *timeout_flag;
    int t_tiles;
};

// trace events (per step, per sub-tile): see scripts/trace_recurrent.py
enum { TR_MMA_HFULL = 0, TR_MMA_ISSUED, TR_EPI_DFULL, TR_EPI_ACT, TR_EPI_CELL, TR_EPI_IMAGE, TR_EPI_COPIES, TR_EVENTS = 16 };
#define HSSB_TRACE(ev, step, sub)                                                                         \
    do {                                                                                                  \
        if (p.trace && blockIdx.x == 0 && (step) >= 0 && (step) < p.trace_steps)                          \
            p.trace[(((step) * 4 + (sub)) * TR_EVENTS) + (ev)] = clock64();                               \
    } while (0)

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// B-operand geometry of the 32-column sub-tiles (pair and multicast kernels)
constexpr int RP_NB = 64;                          // columns of one sub-tile (pair MMA N)
constexpr int RP_NBH = 32;                         // columns held (and produced per epilogue warp) per CTA half
constexpr int RP_G = 4;                            // arrival groups per buffer (= source pairs)
constexpr int RP_PIECE = RP_NBH * 8 * 2 * 2;       // [plane][32 cols][8 units] fp16 = 1 KB: one epilogue warp's output
constexpr int RP_SLICE = 4 * RP_PIECE;             // one source rank: 4 k-chunks
constexpr int RP_HBUF = RC_CL * RP_SLICE;          // 32 KB


__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void sts_b16(uint32_t addr, __half v)
{
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }


// global -> own shared memory, completing `bytes` on the mbarrier
__device__ __forceinline__ void bulk_load_global(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void bulk_store_global(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// global -> the same smem offset in every CTA of `mask`, completing `bytes` on each one's mbarrier
__device__ __forceinline__ void bulk_load_multicast(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}


constexpr int RX_KSTEPS = 3;                       // fused layer-1 input projection: K16 steps of the x operand (input_size <= 48)

// ---- host side, one definition each --------------------------------------------------------------------------
extern unsigned long long *g_trace_buf;      // hssb_debug_trace: device buffer of clock64 stamps (nullptr = off)
extern int g_trace_steps;

// Launch one recurrence kernel over as many batch columns of [b_base, B) as are co-resident; *cols_done = columns covered.
int rc_dsmem_launch(int nb, int s, int pair, const RecurParams &prm, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st);
int rc_dsmem_max_clusters(int *out);                 // co-resident 8-CTA clusters of the 32 x 3 DSMEM geometry
int rc_pair_launch(int s, const RecurParams &prm, const __half *whh_frag, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st);
// variant 2: one publisher per sub-tile, 3: per-warp publishing, 4: per-warp + two epilogue warps per TMEM quadrant
int rc_mc_prepare();                                  // load / configure every default K5m variant on the current device
// info (nullable): CTAs of the launch and the number of epilogue warps per direction that run the step loop (tile_done signals)
struct RecurLaunchInfo { int ctas; unsigned signals_per_dir; };
int rc_mc_launch(int s, int variant, bool fused, const RecurParams &prm, const __half *whh_frag, int64_t rem, int *cols_done,
                 const float *xproj, cudaStream_t st, RecurLaunchInfo *info = nullptr, bool train = false);

}  // namespace hssb
