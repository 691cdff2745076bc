// tcgen05 / TMEM / TMA kernels of the BiLSTM segmenter (hidden_size = 240).
//
// Precision: every gate contraction runs as a split-fp16 "3-pass" product on the 5th-gen tensor
// cores with fp32 accumulation in TMEM:  x = hi + lo (hi = fp16(x), lo = fp16(x - hi), 22 mantissa
// bits together), and  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (the dropped lo*lo term is 2^-22
// relative).  That keeps log-probabilities within ~1e-6 of the fp32 reference, which is what the
// bit-identical-labels requirement needs, at 3 fp16 MMAs per product.
//
// "Cluster gate order": the 960 gate rows of one direction are permuted to g' = r*120 + 4*u + q
// (r = CTA rank in the 8-CTA recurrence cluster, u = unit 0..29 of that rank, q = gate i/f/g/o;
// torch row = q*240 + 30*r + u), so that every recurrence CTA owns one contiguous 120-wide slice and
// the four gates of a unit sit in four adjacent TMEM lanes (= four adjacent threads of one warp).
//
// This file: operand packing, K4 tc_inproj_kernel (xproj[dir][t][b][g'] = A[b,t,:] . W_ih[g',:] + (b_ih + b_hh)[g']), the
// per-layer recurrence dispatch and the forward.  The recurrence kernels (T sequential steps, 8-CTA clusters, weights
// resident in TMEM) live in lstm_rc_mc.cu (K5m, the default) and lstm_rc_dsmem.cu (K5 / K5p); shared declarations in
// lstm_tc_common.cuh.
#include "lstm_tc_common.cuh"
#include <mutex>
#include <cstring>
#include <climits>

namespace hssb {

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
__global__ void pack_whh_kernel(const float *__restrict__ w, int dir, int frag, __half *__restrict__ dst);   // defined below
__global__ void pack_linw_kernel(const float *__restrict__ w, float *__restrict__ dst);
__global__ void pack_wih0_frag_kernel(const float *__restrict__ w, const float *__restrict__ b_ih, const float *__restrict__ b_hh, int F,
                                      int dir, __half *__restrict__ dst, float *__restrict__ bias);
static PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

static int make_tmap(CUtensorMap *m, CUtensorMapDataType dt, int rank, const void *base, const uint64_t *dims,
                     const uint64_t *strides_bytes, const uint32_t *box, CUtensorMapSwizzle sw)
{
    auto enc = get_encode();
    if (!enc) return fail(HSSB_E_DEVICE, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HSSB_E_SHAPE, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
// x[M,F] fp32 -> hi/lo fp16 planes [M,Kp] (zero padded columns)
// range (nullable): {flag, bits of max|x|} -- the values are pre-scaled by 2^-range_exponent; run_flag (nullable): no-op while *run_flag == 0
__global__ void split_planes_kernel(const float *__restrict__ x, long long M, int F, int Kp, __half *__restrict__ hi,
                                    __half *__restrict__ lo, const unsigned *__restrict__ range, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const float down = range ? pow2f(-range_exponent(range[1])) : 1.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M * Kp; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / Kp;
        const int k = (int)(i % Kp);
        __half h = __float2half_rn(0.f), l = h;
        if (k < F) split_f16(x[m * F + k] * down, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// act[B][T][480] fp32 (torch layout: column dir*240 + unit) -> hi/lo fp16 planes [B][Tp][512] in slot layout (column dir*256 +
// rank*32 + u, unit = rank*30 + u; slots u >= 30 are zero): the layer-2 projection operand of the training forward, whose
// layer-1 output passes through torch's dropout between the two layers.
__global__ void split_slots_kernel(const float *__restrict__ act, long long B, long long T, long long Tp, __half *__restrict__ hi,
                                   __half *__restrict__ lo)
{
    const long long n = B * T * (TC_OP / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (TC_OP / 2);
        const int slot = 2 * (int)(i % (TC_OP / 2));             // two slots per thread: 30 and 240 are even, so a pair never straddles
        const int dir = slot >> 8, r = (slot & 255) >> 5, u = slot & 31;
        __half h0 = __float2half_rn(0.f), l0 = h0, h1 = h0, l1 = h0;
        if (u < RC_U) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(act + row * (2 * TC_H) + dir * TC_H + r * RC_U + u));
            split_f16(v.x, h0, l0);
            split_f16(v.y, h1, l1);
        }
        const long long b = row / T, t = row % T;
        const size_t o = ((size_t)b * Tp + t) * TC_OP + slot;
        *reinterpret_cast<__half2 *>(hi + o) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2 *>(lo + o) = __halves2half2(l0, l1);
    }
}

// max |x| of the model input as float bits (word 1 of `range`; non-negative floats order like unsigned integers, NaN sorts above inf)
__global__ void input_amax_kernel(const float *__restrict__ x, long long n, unsigned *__restrict__ range, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    unsigned m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(__ldg(x + i))));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(range + 1, m);
}

// torch W_ih[960][Kin] (rows q*240 + unit) -> planes [dir*960 + g'][Kp]; bias[dir*960 + g'] = b_ih + b_hh
// slots != 0: the K index is in slot layout (column = dir*256 + rank*32 + slot) and maps to torch column dir*240 + rank*30 + slot
__global__ void pack_wih_kernel(const float *__restrict__ w, const float *__restrict__ b_ih, const float *__restrict__ b_hh, int Kin,
                                int Kp, int dir, int slots, __half *__restrict__ hi, __half *__restrict__ lo, float *__restrict__ bias)
{
    const int gp = blockIdx.x;                     // g' = r*120 + 4*u + q
    const int r = gp / 120, u = (gp % 120) / 4, q = gp % 4;
    const int row = q * TC_H + 30 * r + u;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        __half h = __float2half_rn(0.f), l = h;
        if (slots) {
            const int slot = k & 31;
            if (slot < 30) split_f16(w[(size_t)row * Kin + (k >> 8) * TC_H + ((k >> 5) & 7) * 30 + slot], h, l);
        } else if (k < Kin) split_f16(w[(size_t)row * Kin + k], h, l);
        hi[((size_t)dir * TC_G + gp) * Kp + k] = h;
        lo[((size_t)dir * TC_G + gp) * Kp + k] = l;
    }
    if (threadIdx.x == 0) bias[dir * TC_G + gp] = b_ih[row] + b_hh[row];
}

// ------------------------------------------------------------------------------------------------
// K4: input projection GEMM, persistent + warp specialised, clusters of CM (along M) x CN (along N) CTAs.
//
//   xproj[dir][t][b][g'] = A[b,t,:] . W_ih[g',:] + (b_ih + b_hh)[g']          M = B*T, N = 1920, K = 48 / 512
//
// Per CTA one 128(t) x 240(g') output tile at a time; three fp16 MMAs per product (hi*hi, hi*lo with the A operand kept in the
// collector, lo*hi).  SS-mode MMAs are bound by shared-memory bandwidth -- every instruction reads its A (4 KB) and B (N x 32 B)
// operands -- so N is as wide as a direction's 960 gate rows allow (4 n-tiles) and the hi*lo product reuses the A read.
// The operands come out of L2, whose bandwidth is what bounds a lone CTA (it re-reads (128 + 240) x K x 4 bytes per tile).
// The CTAs of a cluster work on CM consecutive batch rows x CN consecutive n-tiles of one time tile and share their loads by
// TMA multicast: the A stage (128 rows) is fetched in CN parts by the CTAs of a cluster column and multicast to all of them,
// the W stage (240 rows) in CM parts by the CTAs of a cluster row.
//
// Work order.  The consumer of layer 2's projection is the layer-2 recurrence, whose forward direction walks t = 0.. and whose
// reverse direction walks t = T-1..; the producer of its A operand is the layer-1 recurrence, which finishes the MIDDLE time
// steps first.  Work is therefore scheduled in units of one time tile (128 steps), in one of three orders per launch:
//   mode 0  whole tiles (the n-tiles of both directions: the A tile is fetched once), outside-in: 0, last, 1, last-1, ...
//   mode 1  whole tiles, middle-out
//   mode 2  one direction of a tile ("chunk" q = 2k: (forward, tile k), q = 2k + 1: (reverse, tile t_tiles - 1 - k)) in the
//           order in which the recurrence needs them, over one or two ranges of q
// Every finished tile bumps chunk_done[q]; the recurrence polls it before it reads a chunk's xproj (lstm_rc_mc.cu), which is
// what lets parts of this GEMM run on the SMs the latency-bound recurrences leave idle (tc_forward).
// Items (CM batch rows x CN n-tiles of one chunk) are handed out by an atomic counter, not by blockIdx: a launch that shares
// the GPU with the recurrence has fewer clusters resident than launched, and a statically assigned item of a cluster that is
// not resident would never be produced while the recurrence waits for it.
//   warp 0          : TMA producer (ring of STAGES stages, BK = 32, SW64)
//   warp 1          : tcgen05.mma issuer; accumulators double-buffered in TMEM (columns 0.. and 256..) so that the
//                     epilogue of tile i overlaps the main loop of tile i+1; smem stages are released to all CTAs
//                     that write into this one with a multicast tcgen05.commit
//   warps 2..       : epilogue TMEM -> registers (+bias) -> swizzled smem tile -> coalesced 16-byte global stores
//   last warp       : item scheduler (cluster rank 0 fetches the next item and posts it into every CTA's item ring)
// ------------------------------------------------------------------------------------------------
#ifndef HSSB_IP_BN
#define HSSB_IP_BN 240
#endif
constexpr int IP_BM = 128, IP_BN = HSSB_IP_BN, IP_BK = 32;      // BK = 32 fp16 = 64-byte rows: SWIZZLE_64B
// the large cluster shape (launches that have the machine to themselves) and the stage count of the layer-2 configuration
constexpr int IP_BIG_CM = (IP_BN == 240) ? 2 : 4, IP_BIG_CN = (IP_BN == 240) ? 4 : 2, IP_L2_STAGES = (IP_BN == 240) ? 4 : 5;
constexpr int IP_NT_DIR = TC_G / IP_BN;                  // n-tiles per direction (4)
static_assert(IP_BM == TC_TT, "the M tile of the projection is the time tile of the producer / consumer flags");
static_assert(TC_G % IP_BN == 0 && IP_BN % 16 == 0 && IP_BN <= 256, "n-tiles must not straddle the two directions");
constexpr int IP_A_BYTES = IP_BM * IP_BK * 2;            // one fp16 plane of the A stage (8 KB)
constexpr int IP_B_BYTES = IP_BN * IP_BK * 2;            // one fp16 plane of the W stage (15 KB)
constexpr int IP_STAGE_BYTES = 2 * IP_A_BYTES + 2 * IP_B_BYTES;   // 46 KB
constexpr int IP_OUT_TILE = 32 * 32 * 4;                 // epilogue transposition tile of one warp: 32 rows x 32 fp32 (4 KB)
constexpr int IP_BIAS_BYTES = (TC_NG + 32) * 4;          // all 1920 folded biases, staged once per CTA (+ slack for the last half chunk)
constexpr int IP_RING = 8;                               // item ring depth: the producer runs up to this many items ahead of the slowest epilogue warp of the cluster
constexpr int IP_NCHUNK = (IP_BN + 31) / 32;             // 32-column chunks of the accumulator (7 + one half chunk)
// STAGES / EPI_WARPS: layer 2's projection (K = 512) is main-loop bound: 4 smem stages, 4 epilogue warps.  Layer 1's (K = 48)
// is epilogue bound -- one warp per SMSP cannot hide the TMEM-load / shared-memory latencies of the drain: 3 stages, 8 warps.
template <int STAGES, int EPI_WARPS, int CM, int CN>
struct IpCfg {
    static constexpr int CL = CM * CN;
    static constexpr int OUT_BYTES = EPI_WARPS * IP_OUT_TILE;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int SMEM_BYTES = STAGES * IP_STAGE_BYTES + OUT_BYTES + IP_BIAS_BYTES + 1024 /*align*/ + BAR_BYTES;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;
    static constexpr int CONSUMERS = 2 + EPI_WARPS;      // roles of one CTA that read an item slot
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "one or two warps per TMEM lane quadrant");
    static_assert((2 * IP_NT_DIR) % CN == 0 && IP_BM % CN == 0 && IP_BN % CM == 0 && (IP_BN / CM) % 8 == 0, "cluster shape must split the tiles");
    static_assert((2 * STAGES + 4 + 2 * IP_RING) * 8 + 16 + 16 + 32 * IP_RING <= BAR_BYTES, "barrier area too small");
};
constexpr int IP_TMEM_COLS = 512;                        // 2 accumulators of 240 columns (at 0 and 256)

struct InprojParams {
    CUtensorMap a_hi, a_lo;   // [k, t, b] fp16, box (32, 128 / CN, 1), SW64   (my part of the A stage)
    CUtensorMap w_hi, w_lo;   // [k, g'(1920)] fp16, box (32, 240 / CM), SW64  (my part of the W stage)
    float *out;               // xproj [dir][t][Bp][960] fp32
    const float *bias;        // [1920]
    long long B, Bp;          // batch, and the row pitch of xproj in batch rows (see xproj_pitch)
    int debug;                // HSSB_IP_DEBUG bit 0: skip the global stores (timing experiment; results are wrong when set)
    const unsigned *range;    // nullable: {flag, bits of max|x|} of a pre-scaled A operand -- the accumulator is scaled back by 2^e here
    const int *run_flag;      // nullable: the launch is a no-op while *run_flag == 0 (stand-in path of the input-range guard)
    const int *skip_flag;     // nullable: the launch is a no-op when *skip_flag != 0
    int k_real;               // true K rounded up to 16 (48 / 512)
    int T;
    int t_tiles;              // ceil(T/128)
    int b_groups;             // ceil(B / CM)
    int unit_mode;            // 0 whole tiles outside-in, 1 whole tiles middle-out, 2 single-direction chunks (see above)
    int u_lo, n_units;        // modes 0 / 1: units [u_lo, u_lo + n_units) of the order; mode 2: n_units chunks ...
    int q_lo1, len1, q_lo2;   // ... q_lo1 + u for u < len1, q_lo2 + (u - len1) beyond
    unsigned *next_item;      // item counter of this launch (zeroed by the host)
    unsigned *chunk_done;     // nullable: [2 * t_tiles] finished (tile, epilogue warp) pairs per chunk
    const unsigned *src_done; // nullable: [2][t_tiles] counters of the layer-1 recurrence (its relu(h1) tile is in memory) ...
    unsigned src_need;        // ... a time tile may be loaded once both directions reached this count
    int *timeout_flag;        // raised instead of hanging when src_done never arrives
};

template <int IP_STAGES, int EPI_WARPS, int CM, int CN>
__global__ void __launch_bounds__(IpCfg<IP_STAGES, EPI_WARPS, CM, CN>::THREADS, 1) tc_inproj_kernel(const __grid_constant__ InprojParams p)
{
    using C = IpCfg<IP_STAGES, EPI_WARPS, CM, CN>;
    constexpr int IP_OUT_BYTES = C::OUT_BYTES;
    constexpr int CL = C::CL;
    if (p.run_flag && *p.run_flag == 0) return;        // uniform over the grid; before any barrier / TMEM allocation
    if (p.skip_flag && *p.skip_flag != 0) return;
    const float up = p.range ? pow2f(range_exponent(p.range[1])) : 1.0f;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    unsigned char *out_base = smem + IP_STAGES * IP_STAGE_BYTES;
    float *bias_s = reinterpret_cast<float *>(out_base + IP_OUT_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_base + IP_OUT_BYTES + IP_BIAS_BYTES);
    uint64_t *full = bars, *empty = bars + IP_STAGES, *tmem_full = bars + 2 * IP_STAGES, *tmem_empty = bars + 2 * IP_STAGES + 2;
    uint64_t *item_full = bars + 2 * IP_STAGES + 4, *item_empty = item_full + IP_RING;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(item_empty + IP_RING);
    // item ring: 16-byte slots (the granularity of a bulk copy); item_src is the scheduler's staging copy in cluster rank 0
    int4 *item_ring = reinterpret_cast<int4 *>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~(uintptr_t)15);
    int4 *item_src = item_ring + IP_RING;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cx = rank % CM, cy = rank / CM;                      // position in the cluster: m / n
    uint16_t mask_a = 0, mask_w = 0;
#pragma unroll
    for (int j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (cx + CM * j));        // CTAs sharing my A tile (same batch row)
#pragma unroll
    for (int i = 0; i < CM; ++i) mask_w |= (uint16_t)(1u << (i + CM * cy));        // CTAs sharing my W tile (same n-tile)
    const int kblocks = (p.k_real + IP_BK - 1) / IP_BK;
    const int n_pairs = (p.unit_mode == 2 ? IP_NT_DIR : 2 * IP_NT_DIR) / CN;      // cluster columns per unit
    const int items_per_unit = p.b_groups * n_pairs;
    const int n_items = p.n_units * items_per_unit;

    // item -> tile coordinates of this CTA
    struct Tile { int q, dir, t0, b, n0; };
    auto tile_of = [&](int item) {
        Tile t;
        const int u = item / items_per_unit, r = item % items_per_unit;
        int tt, ntile;
        if (p.unit_mode == 2) {
            const int q = u < p.len1 ? p.q_lo1 + u : p.q_lo2 + (u - p.len1);
            const int dir = q & 1, k = q >> 1;
            tt = dir ? p.t_tiles - 1 - k : k;
            ntile = dir * IP_NT_DIR + (r % n_pairs) * CN + cy;
        } else {
            const int j = p.u_lo + u;
            if (p.unit_mode == 0) tt = (j & 1) ? p.t_tiles - 1 - (j >> 1) : (j >> 1);
            else if (p.t_tiles & 1) tt = (j & 1) ? (p.t_tiles - 1) / 2 - ((j + 1) >> 1) : (p.t_tiles - 1) / 2 + (j >> 1);
            else tt = (j & 1) ? p.t_tiles / 2 + (j >> 1) : p.t_tiles / 2 - 1 - (j >> 1);
            ntile = (r % n_pairs) * CN + cy;
        }
        t.dir = ntile / IP_NT_DIR;
        t.q = t.dir ? 2 * (p.t_tiles - 1 - tt) + 1 : 2 * tt;
        t.t0 = tt * IP_BM;
        t.b = (r / n_pairs) * CM + cx;
        t.n0 = ntile * IP_BN;
        return t;
    };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.a_hi); prefetch_tmap(&p.a_lo); prefetch_tmap(&p.w_hi); prefetch_tmap(&p.w_lo);
        for (int s = 0; s < IP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CM + CN - 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
        for (int i = 0; i < IP_RING; ++i) { mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], CL * C::CONSUMERS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<IP_TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < TC_NG + 32; i += C::THREADS) bias_s[i] = i < TC_NG ? __ldg(p.bias + i) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();     // barriers of every CTA exist before any multicast / remote arrive can target them

    // every consumer role (the producer thread, the MMA thread, each epilogue warp) reads the ring in order and hands a slot back
    // to the scheduler with ONE arrival as soon as the value is in its registers; an item >= n_items is the end marker
    uint32_t ring_it = 0;
    auto next_item = [&]() {                   // called by a single thread
        const int slot = ring_it % IP_RING;
        mbar_wait_cluster(&item_full[slot], (ring_it / IP_RING) & 1);
        const int item = *reinterpret_cast<volatile int *>(&item_ring[slot].x);
        // relaxed: nothing but the value just read is ordered by this hand-back (a release would wait for the epilogue's
        // outstanding global stores); the comparison makes the arrive depend on the load having returned
        if (item != INT_MIN) mbar_arrive_remote_relaxed(&item_empty[slot], 0);
        ++ring_it;
        return item;
    };
    auto next_item_warp = [&]() {              // called by a converged warp
        const int slot = ring_it % IP_RING;
        mbar_wait_cluster(&item_full[slot], (ring_it / IP_RING) & 1);
        const int item = *reinterpret_cast<volatile int *>(&item_ring[slot].x);
        __syncwarp();
        if (lane == 0 && item != INT_MIN) mbar_arrive_remote_relaxed(&item_empty[slot], 0);
        ++ring_it;
        return item;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            unsigned long long t_start = 0;
            for (int item = next_item(); item < n_items; item = next_item()) {
                const Tile tl = tile_of(item);
                if (p.src_done) {
                    // layer 1 still running: both of its directions must have stored this time tile (generic-proxy flag, then a
                    // proxy fence so that the TMA reads below are ordered behind it)
                    const int tt = tl.t0 / IP_BM;
                    while (ld_acquire_u32(p.src_done + tt) < p.src_need || ld_acquire_u32(p.src_done + p.t_tiles + tt) < p.src_need) {
                        __nanosleep(500);
                        if (!t_start) t_start = globaltimer_ns();
                        else if (globaltimer_ns() - t_start > POLL_TIMEOUT_NS) { *p.timeout_flag = 1; break; }
                    }
                    t_start = 0;
                    fence_proxy_async_all();
                }
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % IP_STAGES;
                    mbar_wait_cluster(&empty[s], ((it / IP_STAGES) & 1) ^ 1);
                    unsigned char *st = stage_base + s * IP_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], IP_STAGE_BYTES);
                    // my part of the A tile (rows cy * 128/CN ..) -> the CTAs of my cluster column
                    tma_load_3d_mc(st + cy * (IP_A_BYTES / CN), &p.a_hi, &full[s], kb * IP_BK, tl.t0 + cy * (IP_BM / CN), tl.b, mask_a);
                    tma_load_3d_mc(st + IP_A_BYTES + cy * (IP_A_BYTES / CN), &p.a_lo, &full[s], kb * IP_BK, tl.t0 + cy * (IP_BM / CN), tl.b, mask_a);
                    // my part of the W tile (rows cx * 240/CM ..) -> the CTAs of my cluster row
                    tma_load_2d_mc(st + 2 * IP_A_BYTES + cx * (IP_B_BYTES / CM), &p.w_hi, &full[s], kb * IP_BK, tl.n0 + cx * (IP_BN / CM), mask_w);
                    tma_load_2d_mc(st + 2 * IP_A_BYTES + IP_B_BYTES + cx * (IP_B_BYTES / CM), &p.w_lo, &full[s], kb * IP_BK, tl.n0 + cx * (IP_BN / CM), mask_w);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(IP_BM, IP_BN);
            const uint16_t mask_rel = mask_a | mask_w;     // every CTA that writes into my stages
            uint32_t it = 0, tile = 0;
            for (int item = next_item(); item < n_items; item = next_item(), ++tile) {
                const uint32_t acc = tile & 1;
                mbar_wait(&tmem_empty[acc], ((tile >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % IP_STAGES;
                    mbar_wait_cluster(&full[s], (it / IP_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stage_base + s * IP_STAGE_BYTES);
                    const uint32_t a_lo = a_hi + IP_A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * IP_A_BYTES;
                    const uint32_t b_lo = b_hi + IP_B_BYTES;
                    const int ksteps = min(IP_BK, p.k_real - kb * IP_BK) / 16;   // skip all-padding K16 steps
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t off = ks * 32;   // 16 fp16 = 32 bytes inside the 64-byte swizzle row
                        const uint64_t da_hi = make_smem_desc(a_hi + off, 16, 512, LAYOUT_SW64);
                        const uint64_t da_lo = make_smem_desc(a_lo + off, 16, 512, LAYOUT_SW64);
                        const uint64_t db_hi = make_smem_desc(b_hi + off, 16, 512, LAYOUT_SW64);
                        const uint64_t db_lo = make_smem_desc(b_lo + off, 16, 512, LAYOUT_SW64);
                        mma_f16_ss_keep_a(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);      // a_hi stays in the collector ...
                        mma_f16_ss_reuse_a(d_tmem, da_hi, db_lo, idesc, 1);                  // ... for the hi * lo product
                        mma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1);
                    }
                    mma_commit_mc(&empty[s], mask_rel);    // frees this stage in every CTA that fills it
                }
                mma_commit(&tmem_full[acc]);               // accumulator complete
            }
        }
    } else if (warp < 2 + EPI_WARPS) {
        // ===== epilogue: TMEM lane quadrant = warp % 4; with 8 warps the two warps of a quadrant take alternate 32-column chunks =====
        // A warp first pulls ALL of its chunks of the accumulator into registers (three at a time) and hands the accumulator back to the
        // MMA warp as soon as the last one is loaded (the stores below are slow and must not hold TMEM), then per chunk:
        // registers (one row of 32 columns per thread, + bias) -> xor-swizzled smem tile -> 16-byte global stores in which 8
        // consecutive lanes cover one 128-byte row segment (every store instruction writes four full lines).
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CSTEP = EPI_WARPS / 4;
        unsigned char *ob = out_base + (warp - 2) * IP_OUT_TILE;
        uint32_t tile = 0;
        // The completion of a tile is published one tile late, right after the NEXT accumulator has been pulled into registers:
        // the fence in front of the flag then only waits for stores issued a whole tile ago (long since written) instead of
        // stalling the drain of the tile that follows.
        int pending_q = -1;
        auto publish_pending = [&]() {
            if (pending_q >= 0) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(p.chunk_done + pending_q, 1u);
                }
                pending_q = -1;
            }
        };
        for (int item = next_item_warp(); item < n_items; item = next_item_warp(), ++tile) {
            const Tile tl = tile_of(item);
            const int nl0 = tl.n0 - tl.dir * TC_G;
            const uint32_t acc = tile & 1;
            mbar_wait(&tmem_full[acc], (tile >> 1) & 1);
            tc_fence_after();
            for (int c0 = half; c0 < IP_NCHUNK; c0 += 3 * CSTEP) {
                uint32_t v[3][32];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (c0 + i * CSTEP < IP_NCHUNK)
                        tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (c0 + i * CSTEP) * 32, v[i]);
                tmem_ld_wait();
                if (c0 + 3 * CSTEP >= IP_NCHUNK) {          // my part of the accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    publish_pending();                      // (the previous tile's stores: see above)
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int c = c0 + i * CSTEP;
                    if (c >= IP_NCHUNK) break;
                    const float4 *bias = reinterpret_cast<const float4 *>(bias_s + tl.n0 + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bj = bias[j];                 // shared-memory broadcast
                        float4 o;
                        o.x = fmaf(__uint_as_float(v[i][4 * j + 0]), up, bj.x);      // up = 1 unless the A operand was pre-scaled
                        o.y = fmaf(__uint_as_float(v[i][4 * j + 1]), up, bj.y);
                        o.z = fmaf(__uint_as_float(v[i][4 * j + 2]), up, bj.z);
                        o.w = fmaf(__uint_as_float(v[i][4 * j + 3]), up, bj.w);
                        *reinterpret_cast<float4 *>(ob + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;   // 128B xor swizzle: conflict free both ways
                    }
                    __syncwarp();
                    const int rsub = lane >> 3, c16 = lane & 7;
                    float *gout = p.out + (((size_t)tl.dir * p.T + tl.t0 + q * 32) * p.Bp + tl.b) * TC_G + nl0 + c * 32 + c16 * 4;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = 4 * j + rsub;
                        const float4 o = *reinterpret_cast<const float4 *>(ob + row * 128 + ((c16 ^ (row & 7)) << 4));
                        if (tl.b < p.B && tl.t0 + q * 32 + row < p.T && c * 32 + c16 * 4 < IP_BN && !(p.debug & 1))   // (b >= B: padding tiles of the last batch group)
                            __stcs(reinterpret_cast<float4 *>(gout + (size_t)row * p.Bp * TC_G), o);
                    }
                    __syncwarp();
                }
            }
            if (p.chunk_done && tl.b < p.B) pending_q = tl.q;      // this warp's part of a real tile is stored (the recurrence counts tiles x epilogue warps)
        }
        publish_pending();
    } else {
        // ===== item scheduler (cluster rank 0): the next item of the launch -> every CTA's ring =====
        // Fully asynchronous broadcast: the item goes into a staging slot, then per PEER one remote arrive.expect_tx arms its
        // item_full barrier and one 16-byte bulk copy (shared -> shared::cluster, complete_tx on that barrier) delivers the slot.
        // Nothing here waits for a round trip (release-arrives behind remote stores cost one per CTA: about as long as a tile
        // takes); the next item is fetched from the global counter right after the previous one was posted.
        // A shared::cta -> shared::cluster bulk copy must target ANOTHER CTA, so rank 0's own ring is filled by rank 1's scheduler
        // warp, which forwards every slot it receives (one more hop, hidden by the ring depth); a single-CTA "cluster" has no peer
        // and takes the item with a plain store + arrive.
        if (rank == 0 && elect_one()) {
            int item = (int)atomicAdd(p.next_item, 1u);
            for (uint32_t it = 0;; ++it) {
                const int slot = it % IP_RING;
                mbar_wait_cluster(&item_empty[slot], ((it / IP_RING) & 1) ^ 1);      // every consumer of the cluster has read the slot's previous item
                if (CL > 1) {
                    *reinterpret_cast<volatile int *>(&item_src[slot].x) = item;    // (the copies of the previous use of this staging slot completed
                    fence_proxy_async_smem();                                        //  before its consumers could read, i.e. before item_empty)
                    mbar_arrive_expect_tx(&item_full[slot], 16);                     // my own slot: the bytes come back from rank 1
#pragma unroll
                    for (int r = 1; r < CL; ++r) {
                        mbar_arrive_expect_tx_remote(&item_full[slot], 16, r);
                        bulk_copy_to_cta(&item_ring[slot], &item_src[slot], 16, &item_full[slot], r);
                    }
                } else {
                    *reinterpret_cast<volatile int *>(&item_ring[slot].x) = item;
                    mbar_arrive(&item_full[slot]);
                }
                if (item >= n_items) break;
                item = (int)atomicAdd(p.next_item, 1u);
            }
        } else if (CL > 1 && rank == 1 && elect_one()) {
            // relay: my ring slot (filled through the async proxy, so no fence) -> the same slot of rank 0.  The slot is not
            // reposted before rank 0's consumers have handed it back, i.e. not before this copy has been read and delivered.
            for (uint32_t it = 0;; ++it) {
                const int slot = it % IP_RING;
                mbar_wait_cluster(&item_full[slot], (it / IP_RING) & 1);
                const int item = *reinterpret_cast<volatile int *>(&item_ring[slot].x);
                bulk_copy_to_cta(&item_ring[slot], &item_ring[slot], 16, &item_full[slot], 0);
                if (item >= n_items) break;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();     // nobody leaves while peers may still multicast into / arrive on this CTA
    if (warp == 1) tmem_dealloc<IP_TMEM_COLS>(tmem_base);
}

// xproj[dir][t][b][g'] -> canonical [dir][b*T + t][q*240 + unit]   (debug / validation only)
__global__ void unpermute_xproj_kernel(const float *__restrict__ src, long long B, long long Bp, long long T, float *__restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 2 * B * T * TC_G;
    if (i >= total) return;
    const int gp = (int)(i % TC_G);
    const long long rest = i / TC_G;
    const long long b = rest % B, t = (rest / B) % T, dir = rest / (B * T);
    const int r = gp / 120, u = (gp % 120) / 4, q = gp % 4;
    dst[((size_t)dir * B * T + b * T + t) * TC_G + q * TC_H + 30 * r + u] = src[((size_t)(dir * T + t) * Bp + b) * TC_G + gp];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int kp_of_layer(int layer, int F) { return layer == 0 ? 64 : 512; }
static int kreal_of_layer(int layer, int F) { return layer == 0 ? ((F + 15) / 16) * 16 : TC_OP; }

static int tc_prepare();

size_t tc_pack_bytes(int F, int H)
{
    if (H != TC_H || F > 64) return 0;
    size_t n = 0;
    for (int l = 0; l < 2; ++l) {
        n += align_up(sizeof(__half) * 2 * TC_NG * kp_of_layer(l, F), 256);   // wih hi+lo
        n += align_up(sizeof(float) * TC_NG, 256);                             // bias
        n += 2 * align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 256, 256);        // whh planes [dir][rank][plane][128][256], two row orders
    }
    n += align_up(sizeof(float) * 4 * TC_OP, 256);                             // linear weights in slot layout
    n += align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 64, 256);                 // layer-1 W_ih slices for the fused projection
    n += align_up(sizeof(float) * 2 * 8 * 128, 256);                           // ... and their folded biases
    return n;
}

int tc_pack(hssb_model *m, const hssb_model_params *p, void *dst, cudaStream_t st)
{
    m->tc_ready = false;
    if (m->H != TC_H || m->F > 64) return 0;
    char *base = static_cast<char *>(dst);
    size_t off = 0;
    const int kin[2] = {m->F, 2 * TC_H};
    // the raw torch tensors may be host pointers: stage them through the model's staging buffer (hssb_model_create sized it for
    // the largest tensor group + the two range words; no per-call allocation: this runs after every optimiser step in training)
    const size_t tmp_floats = (size_t)TC_G * (2 * TC_H) + 2 * TC_G;
    float *tmp = m->stage;
    cudaError_t e;
    // max |w| over the LSTM weights (word 1): outside the fp16-split range the model keeps to the generic fp32 kernels
    unsigned *w_range = reinterpret_cast<unsigned *>(tmp + tmp_floats);
    if ((e = cudaMemsetAsync(w_range, 0, 8, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(tc_pack)");
    int rc = 0;
    m->tc_wih0_frag = reinterpret_cast<__half *>(base + off);
    off += align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 64, 256);
    m->tc_bias0_frag = reinterpret_cast<float *>(base + off);
    off += align_up(sizeof(float) * 2 * 8 * 128, 256);
    for (int l = 0; l < 2 && !rc; ++l) {
        const int Kp = kp_of_layer(l, m->F);
        m->tc_wih[l] = reinterpret_cast<__half *>(base + off);
        off += align_up(sizeof(__half) * 2 * TC_NG * Kp, 256);
        m->tc_bias[l] = reinterpret_cast<float *>(base + off);
        off += align_up(sizeof(float) * TC_NG, 256);
        m->tc_whh[l] = reinterpret_cast<__half *>(base + off);
        off += align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 256, 256);
        m->tc_whh_frag[l] = reinterpret_cast<__half *>(base + off);
        off += align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 256, 256);
        __half *hi = m->tc_wih[l], *lo = hi + (size_t)TC_NG * Kp;
        for (int d = 0; d < 2 && !rc; ++d) {
            float *w = tmp, *bi = tmp + (size_t)TC_G * kin[l], *bh = bi + TC_G;
            if ((e = cudaMemcpyAsync(w, p->w_ih[l][d], sizeof(float) * TC_G * kin[l], cudaMemcpyDefault, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(bi, p->b_ih[l][d], sizeof(float) * TC_G, cudaMemcpyDefault, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(bh, p->b_hh[l][d], sizeof(float) * TC_G, cudaMemcpyDefault, st)) != cudaSuccess) {
                rc = cuda_fail(e, "cudaMemcpyAsync(tc_pack)");
                break;
            }
            input_amax_kernel<<<64, 256, 0, st>>>(w, (long long)TC_G * kin[l], w_range, nullptr);
            pack_wih_kernel<<<TC_G, 128, 0, st>>>(w, bi, bh, kin[l], Kp, d, l, hi, lo, m->tc_bias[l]);
            if (l == 0) pack_wih0_frag_kernel<<<8 * 128, 64, 0, st>>>(w, bi, bh, kin[0], d, m->tc_wih0_frag, m->tc_bias0_frag);
            if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "pack_wih_kernel"); break; }
            if ((e = cudaMemcpyAsync(w, p->w_hh[l][d], sizeof(float) * TC_G * TC_H, cudaMemcpyDefault, st)) != cudaSuccess) {
                rc = cuda_fail(e, "cudaMemcpyAsync(tc_pack w_hh)");
                break;
            }
            input_amax_kernel<<<64, 256, 0, st>>>(w, (long long)TC_G * TC_H, w_range, nullptr);
            pack_whh_kernel<<<8 * 128, 128, 0, st>>>(w, d, 0, m->tc_whh[l]);
            pack_whh_kernel<<<8 * 128, 128, 0, st>>>(w, d, 1, m->tc_whh_frag[l]);
            if ((e = cudaGetLastError()) != cudaSuccess) rc = cuda_fail(e, "pack_whh_kernel");
        }
    }
    if (!rc) {
        m->tc_lin_w = reinterpret_cast<float *>(base + off);
        off += align_up(sizeof(float) * 4 * TC_OP, 256);
        cudaError_t e2 = cudaMemcpyAsync(tmp, p->lin_w, sizeof(float) * 4 * 2 * TC_H, cudaMemcpyDefault, st);
        if (e2 != cudaSuccess) rc = cuda_fail(e2, "cudaMemcpyAsync(tc_pack lin_w)");
        else {
            pack_linw_kernel<<<4, TC_OP, 0, st>>>(tmp, m->tc_lin_w);
            if ((e2 = cudaGetLastError()) != cudaSuccess) rc = cuda_fail(e2, "pack_linw_kernel");
        }
    }
    unsigned w_bits[2] = {0, 0};
    if (!rc && ((e = cudaMemcpyAsync(w_bits, w_range, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess))
        rc = cuda_fail(e, "tc_pack: weight range");
    if (rc) return rc;
    float w_max;
    memcpy(&w_max, &w_bits[1], 4);
    m->tc_ready = (w_max <= TC_SPLIT_SAFE);      // false for larger / non-finite weights: hssb_model_forward then runs the fp32 SIMT kernels
    if (m->tc_ready) return tc_prepare();
    return 0;
}

// What one launch of K4 covers and how it is synchronised with the recurrences either side of it (see tc_forward).
struct InprojJob {
    int unit_mode = 0;                    // 0 whole tiles outside-in, 1 whole tiles middle-out, 2 single-direction chunks
    int u_lo = 0, n_units = -1;           // modes 0 / 1: units of the order (n_units < 0: all t_tiles tiles)
    int q_lo1 = 0, len1 = 0, q_lo2 = 0;   // mode 2: chunks q_lo1 .. (len1 of them), then q_lo2 .. up to n_units in total
    int shape = 0;                        // cluster shape CM x CN: 0 = 2x4, 1 = 2x2, 2 = 2x1, 3 = 1x1
    unsigned *next_item = nullptr;        // zeroed device word: the item counter of this launch (required)
    unsigned *chunk_done = nullptr;       // [2 * t_tiles] zeroed device counters bumped per finished (tile, epilogue warp)
    const unsigned *src_done = nullptr;   // [2][t_tiles] progress of the layer-1 recurrence (middle-out launch only)
    unsigned src_need = 0;
    int *timeout_flag = nullptr;
    const int *skip_flag = nullptr;       // the launch is a no-op when *skip_flag != 0
    const char *name = nullptr;
};

// Per device, once: opt in to the kernel's shared memory and find its co-resident clusters (see prepare_recurrent_mc for why this
// also runs when a model is created).
template <int STAGES, int EPI_WARPS, int CM, int CN>
static int prepare_inproj(int *max_clusters_out)
{
    using C = IpCfg<STAGES, EPI_WARPS, CM, CN>;
    static PerDeviceInt cached_clusters;
    int max_clusters = cached_clusters.get();
    if (!max_clusters) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaFuncSetAttribute(tc_inproj_kernel<STAGES, EPI_WARPS, CM, CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_inproj_kernel)");
        cfg.gridDim = dim3(C::CL * 16);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_inproj_kernel<STAGES, EPI_WARPS, CM, CN>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_inproj_kernel)");
        if (n < 1) return fail(HSSB_E_DEVICE, "device cannot host an input-projection cluster");
        max_clusters = n;
        cached_clusters.set(n);
    }
    *max_clusters_out = max_clusters;
    return 0;
}

template <int STAGES, int EPI_WARPS, int CM, int CN>
static int launch_inproj(const InprojParams &prm, int n_items, const char *name, cudaStream_t st)
{
    using C = IpCfg<STAGES, EPI_WARPS, CM, CN>;
    int max_clusters = 0;
    if (int rc = prepare_inproj<STAGES, EPI_WARPS, CM, CN>(&max_clusters)) return rc;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(C::CL * std::max(1, std::min(max_clusters, n_items))));
    ProfScope prof(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_inproj_kernel<STAGES, EPI_WARPS, CM, CN>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_inproj_kernel)");
    return 0;
}

__global__ void resident_gate_kernel(const unsigned *__restrict__ resident);
__global__ void tile_gate_kernel(const unsigned *__restrict__ fwd, const unsigned *__restrict__ rev, unsigned need);

// Loads and configures every kernel that may be launched while another one is polling for it (see prepare_recurrent_mc).
static int tc_prepare()
{
    int n = 0;
    if (int rc = prepare_inproj<3, 8, IP_BIG_CM, IP_BIG_CN>(&n)) return rc;
    if (int rc = prepare_inproj<IP_L2_STAGES, 4, IP_BIG_CM, IP_BIG_CN>(&n)) return rc;
    if (int rc = prepare_inproj<IP_L2_STAGES, 4, 2, 2>(&n)) return rc;
    if (int rc = prepare_inproj<IP_L2_STAGES, 4, 2, 1>(&n)) return rc;
    if (int rc = prepare_inproj<IP_L2_STAGES, 4, 1, 1>(&n)) return rc;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, resident_gate_kernel);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncGetAttributes(resident_gate_kernel)");
    if ((e = cudaFuncGetAttributes(&fa, tile_gate_kernel)) != cudaSuccess) return cuda_fail(e, "cudaFuncGetAttributes(tile_gate_kernel)");
    return rc_mc_prepare();
}

// One layer's input projection on the tensor cores.  a_hi/a_lo: [B*T][pitch] fp16 planes.
int tc_inproj(const hssb_model *m, int layer, const __half *a_hi, const __half *a_lo, int pitch_elems, int64_t B, int64_t T,
              float *xproj /*[2][T][B][960]*/, cudaStream_t st, const InprojJob &job, const unsigned *range = nullptr,
              const int *run_flag = nullptr, int64_t t_pitch = 0)
{
    if (t_pitch <= 0) t_pitch = T;          // time rows per window of the A planes
    static const int SHAPES[4][2] = {{IP_BIG_CM, IP_BIG_CN}, {2, 2}, {2, 1}, {1, 1}};
    if (job.shape < 0 || job.shape > 3 || !job.next_item) return fail(HSSB_E_MODE, "tc_inproj: bad job");
    const int CM = SHAPES[job.shape][0], CN = SHAPES[job.shape][1];
    InprojParams prm = {};
    prm.range = range;
    prm.run_flag = run_flag;
    const int Kp = kp_of_layer(layer, m->F);
    const int kreal = kreal_of_layer(layer, m->F);
    {
        const uint64_t dims[3] = {(uint64_t)(layer == 0 ? Kp : kreal), (uint64_t)T, (uint64_t)B};
        const uint64_t strides[2] = {(uint64_t)pitch_elems * 2, (uint64_t)t_pitch * pitch_elems * 2};
        const uint32_t box[3] = {IP_BK, (uint32_t)(IP_BM / CN), 1};
        if (int rc = make_tmap(&prm.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        if (int rc = make_tmap(&prm.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)TC_NG};
        const uint64_t strides[1] = {(uint64_t)Kp * 2};
        const uint32_t box[2] = {IP_BK, (uint32_t)(IP_BN / CM)};
        const __half *hi = m->tc_wih[layer], *lo = hi + (size_t)TC_NG * Kp;
        if (int rc = make_tmap(&prm.w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        if (int rc = make_tmap(&prm.w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    prm.out = xproj;
    prm.B = B;
    prm.Bp = xproj_pitch(B);
    prm.debug = 0;
#ifdef HSSB_KNOCKOUTS   // timing experiment that skips the global stores (results are WRONG): only in -DHSSB_KNOCKOUTS builds
    if (const char *e = getenv("HSSB_IP_DEBUG")) prm.debug = atoi(e);
#endif
    prm.bias = m->tc_bias[layer];
    prm.k_real = kreal;
    prm.T = (int)T;
    prm.t_tiles = (int)((T + IP_BM - 1) / IP_BM);
    prm.b_groups = (int)((B + CM - 1) / CM);
    prm.unit_mode = job.unit_mode;
    prm.u_lo = job.u_lo;
    prm.n_units = job.n_units < 0 ? prm.t_tiles : job.n_units;
    prm.q_lo1 = job.q_lo1; prm.len1 = job.len1; prm.q_lo2 = job.q_lo2;
    prm.next_item = job.next_item;
    prm.chunk_done = job.chunk_done;
    prm.src_done = job.src_done;
    prm.src_need = job.src_need;
    prm.timeout_flag = job.timeout_flag;
    prm.skip_flag = job.skip_flag;
    const int nt_unit = (job.unit_mode == 2 ? IP_NT_DIR : 2 * IP_NT_DIR);
    if (job.unit_mode < 0 || job.unit_mode > 2 || nt_unit % CN != 0) return fail(HSSB_E_MODE, "tc_inproj: cluster shape %dx%d does not fit unit mode %d", CM, CN, job.unit_mode);
    if (job.unit_mode == 2) {
        if (prm.len1 < 0 || prm.len1 > prm.n_units || prm.q_lo1 < 0 || prm.q_lo1 + prm.len1 > 2 * prm.t_tiles ||
            (prm.n_units > prm.len1 && (prm.q_lo2 < 0 || prm.q_lo2 + prm.n_units - prm.len1 > 2 * prm.t_tiles)))
            return fail(HSSB_E_SHAPE, "tc_inproj: bad chunk ranges");
    } else if (prm.u_lo < 0 || prm.u_lo + prm.n_units > prm.t_tiles) return fail(HSSB_E_SHAPE, "tc_inproj: bad tile range");
    if (prm.n_units == 0) return 0;
    if (prm.src_done && !prm.timeout_flag) return fail(HSSB_E_NULL, "tc_inproj: src_done needs a timeout flag");

    const long long items = (long long)prm.n_units * prm.b_groups * (nt_unit / CN);
    if (items > 0x7fffff00ll) return fail(HSSB_E_SHAPE, "tc_inproj: %lld work items do not fit the item counter", items);
    const int n_items = (int)items;
    const char *name = job.name ? job.name : (layer == 0 ? "tc_inproj_l0" : "tc_inproj_l1");
    if (layer == 0) {                      // K = 48: epilogue bound -> 3 stages, 8 epilogue warps
        if (job.shape != 0) return fail(HSSB_E_MODE, "tc_inproj: layer 1 runs on the large clusters");
        return launch_inproj<3, 8, IP_BIG_CM, IP_BIG_CN>(prm, n_items, name, st);
    }
    switch (job.shape) {
    case 0: return launch_inproj<IP_L2_STAGES, 4, IP_BIG_CM, IP_BIG_CN>(prm, n_items, name, st);
    case 1: return launch_inproj<IP_L2_STAGES, 4, 2, 2>(prm, n_items, name, st);
    case 2: return launch_inproj<IP_L2_STAGES, 4, 2, 1>(prm, n_items, name, st);
    default: return launch_inproj<IP_L2_STAGES, 4, 1, 1>(prm, n_items, name, st);
    }
}

// torch W_hh[960][240] -> planes [dir][rank][plane][128 rows][256 k' = 32 r' + u'];  row (TMEM lane) order:
//   frag == 0: lane = 4*u + gate                          (tc_recurrent_kernel: 32x32b loads + quad shuffles)
//   frag == 1: lane = 32*(u/8) + 8*gate + u%8             (tc_recurrent_pair_kernel: 16x256b fragment loads)
__global__ void pack_whh_kernel(const float *__restrict__ w, int dir, int frag, __half *__restrict__ dst)
{
    const int rank = blockIdx.x / 128, row = blockIdx.x % 128;
    const int u = frag ? (row / 32) * 8 + row % 8 : row / 4;
    const int q = frag ? (row % 32) / 8 : row % 4;
    __half *hi = dst + ((((size_t)dir * RC_CL + rank) * 2 + 0) * 128 + row) * RC_KP;
    __half *lo = dst + ((((size_t)dir * RC_CL + rank) * 2 + 1) * 128 + row) * RC_KP;
    for (int kp = threadIdx.x; kp < RC_KP; kp += blockDim.x) {
        const int r2 = kp / 32, u2 = kp % 32;
        __half h = __float2half_rn(0.f), l = h;
        if (u < RC_U && u2 < RC_U) split_f16(w[(size_t)(q * TC_H + RC_U * rank + u) * TC_H + RC_U * r2 + u2], h, l);
        hi[kp] = h;
        lo[kp] = l;
    }
}

// Fused layer-1 projection: torch W_ih[960][F] -> planes [dir][rank][plane][128 rows, lane = 32*(u/8) + 8*gate + u%8][64] fp16 and
// the folded bias b_ih + b_hh [dir][rank][128] in the same row order (zero rows for the padding units 30, 31)
__global__ void pack_wih0_frag_kernel(const float *__restrict__ w, const float *__restrict__ b_ih, const float *__restrict__ b_hh, int F,
                                      int dir, __half *__restrict__ dst, float *__restrict__ bias)
{
    const int rank = blockIdx.x / 128, row = blockIdx.x % 128;
    const int u = (row / 32) * 8 + row % 8, q = (row % 32) / 8;
    const int trow = q * TC_H + RC_U * rank + u;
    __half *hi = dst + ((((size_t)dir * RC_CL + rank) * 2 + 0) * 128 + row) * 64;
    __half *lo = dst + ((((size_t)dir * RC_CL + rank) * 2 + 1) * 128 + row) * 64;
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        __half h = __float2half_rn(0.f), l = h;
        if (u < RC_U && k < F) split_f16(w[(size_t)trow * F + k], h, l);
        hi[k] = h;
        lo[k] = l;
    }
    if (threadIdx.x == 0) bias[((size_t)dir * RC_CL + rank) * 128 + row] = (u < RC_U) ? b_ih[trow] + b_hh[trow] : 0.f;
}

// x[B][T][F] fp32 -> hi/lo fp16 planes [t][32-column tile][chunk 8][32 cols][8 features] (zero padded features and columns): the
// K-major operand of the fused projection for one (step, sub-tile) is then one contiguous run, fetched by a single bulk copy
// *range_flag (nullable) is raised when some |x| > TC_SPLIT_SAFE or x is not finite: the fused recurrence then stands down for the
// pre-scaled projection path (tc_forward)
__global__ void split_planes_tiled_kernel(const float *__restrict__ x, long long B, long long T, int F, __half *__restrict__ hi,
                                          __half *__restrict__ lo, int *__restrict__ range_flag)
{
    const long long tiles = (B + RP_NBH - 1) / RP_NBH;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((t * tiles + tile) * 8 + chunk) * 32 + col
    if (i >= T * tiles * 8 * RP_NBH) return;
    const int col = (int)(i % RP_NBH), c = (int)((i / RP_NBH) % 8);
    const long long tile = (i / (8 * RP_NBH)) % tiles, t = i / (8 * RP_NBH * tiles);
    const long long b = tile * RP_NBH + col;
    __half h[8], l[8];
    bool out_of_range = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = 8 * c + e;
        h[e] = l[e] = __float2half_rn(0.f);
        if (k < F && b < B) {
            const float v = __ldg(x + (b * T + t) * F + k);
            out_of_range |= !(fabsf(v) <= TC_SPLIT_SAFE);
            split_f16(v, h[e], l[e]);
        }
    }
    *reinterpret_cast<uint4 *>(hi + i * 8) = *reinterpret_cast<const uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo + i * 8) = *reinterpret_cast<const uint4 *>(l);
    if (out_of_range && range_flag) *range_flag = 1;          // benign race: every writer stores the same value
}

// linear.weight[4][480] -> slot layout [4][512]
__global__ void pack_linw_kernel(const float *__restrict__ w, float *__restrict__ dst)
{
    const int c = blockIdx.x, k = threadIdx.x, slot = k & 31;
    dst[c * TC_OP + k] = slot < 30 ? w[c * 2 * TC_H + (k >> 8) * TC_H + ((k >> 5) & 7) * 30 + slot] : 0.f;
}

// Holds a stream back until every CTA of a recurrence launch is on the machine (resident[1] != 0): the projection launch behind it
// either waits on that recurrence's progress flags or must leave it the SMs it needs, so its persistent CTAs must not be placed
// first.  Gives up quietly after 20 ms: under a profiler that serialises kernels (ncu) the recurrence is queued BEHIND this
// kernel and can never become resident -- the projection launch then simply runs before it, alone, which is correct.
constexpr unsigned long long GATE_TIMEOUT_NS = 20000000ull;
__global__ void resident_gate_kernel(const unsigned *__restrict__ resident)
{
    if (threadIdx.x != 0) return;
    const unsigned long long t_start = globaltimer_ns();
    while (ld_acquire_u32(resident + 1) == 0) {
        __nanosleep(1000);
        if (globaltimer_ns() - t_start > GATE_TIMEOUT_NS) break;
    }
}

// The same for the middle-out projection launch M: held back until the layer-1 recurrence has stored the first tile M will work on
// (both directions), i.e. until about half of that recurrence is done.  Until then its persistent CTAs would only sit on the idle
// SMs polling; gated like this those SMs stay free for whatever else the caller has queued (hss.pipeline: the next batch's FSST).
__global__ void tile_gate_kernel(const unsigned *__restrict__ fwd, const unsigned *__restrict__ rev, unsigned need)
{
    if (threadIdx.x != 0) return;
    const unsigned long long t_start = globaltimer_ns();
    while (ld_acquire_u32(fwd) < need || ld_acquire_u32(rev) < need) {
        __nanosleep(2000);
        if (globaltimer_ns() - t_start > GATE_TIMEOUT_NS) break;
    }
}

unsigned long long *g_trace_buf = nullptr;
int g_trace_steps = 0;

// Optional hooks of a recurrence (input-range guard, flags of the overlapped projection) and what its launches looked like.
struct RecurSync {
    const int *skip_flag = nullptr;       // the launches are no-ops when (*skip_flag != 0) == skip_when
    int skip_when = 0;
    const unsigned *chunk_done = nullptr; // consumer side: wait for the projection chunk before reading its xproj
    unsigned chunk_need = 0;
    unsigned *tile_done = nullptr;        // producer side: relu(h) of a time tile is in memory
    unsigned *resident = nullptr;         // bumped by every CTA once it holds its SM
    int *timeout_flag = nullptr;
    // training forward: the activated gates [2][B*T][960], cell states [2][B*T][240] and raw h [B][T][480] replace the relu'd outputs
    float *tr_gates = nullptr, *tr_cells = nullptr, *tr_out = nullptr;
    // out
    int launches = 0, ctas_first = 0;
    unsigned signals_per_dir = 0;
    bool multicast = true;                // every launch was a K5m launch (the only kernel that implements the hooks)
};

// One layer's recurrence for batch columns [0, B): picks the sub-tile geometry from B.
static int tc_recurrent(const hssb_model *m, int layer, float *xproj, const float *h0, const float *c0, float *hn, float *cn,
                        __half *out_hi, __half *out_lo, float *out_f32, unsigned char *gather, int64_t B, int64_t T, cudaStream_t st,
                        const __half *x_hi = nullptr, const __half *x_lo = nullptr, RecurSync *sync = nullptr)
{
    // x_hi / x_lo != nullptr: layer 1 with the input projection fused into the recurrence (tile-major x planes, no xproj)
    const bool fused = x_hi != nullptr;
    RecurSync no_sync;
    if (!sync) sync = &no_sync;
    const int *skip_flag = sync->skip_flag;
    RecurParams prm = {};
    prm.gather = gather;
    prm.layer = layer;
    prm.skip_flag = sync->skip_flag;
    prm.skip_when = sync->skip_when;
    prm.chunk_done = sync->chunk_done;
    prm.chunk_need = sync->chunk_need;
    prm.tile_done = sync->tile_done;
    prm.resident = sync->resident;
    prm.timeout_flag = sync->timeout_flag;
    prm.t_tiles = (int)((T + TC_TT - 1) / TC_TT);
    sync->launches = 0; sync->ctas_first = 0; sync->signals_per_dir = 0; sync->multicast = true;
    if (fused) {
        prm.x_hi = x_hi;
        prm.x_lo = x_lo;
        prm.x_tiles = (B + RP_NBH - 1) / RP_NBH;
        prm.wih0 = m->tc_wih0_frag;
        prm.bias0 = m->tc_bias0_frag;
    }
    prm.whh = m->tc_whh[layer];
    prm.h0 = h0; prm.c0 = c0; prm.hn = hn; prm.cn = cn;
    prm.out_hi = out_hi; prm.out_lo = out_lo; prm.out_f32 = out_f32;
    prm.B = B; prm.T = T;
    prm.Tp = act_pitch(T);
    prm.Bp = xproj_pitch(B);
    const bool train = sync->tr_gates != nullptr;
    prm.tr_gates = sync->tr_gates; prm.tr_cells = sync->tr_cells; prm.tr_out = sync->tr_out;
    if (train && (fused || !sync->tr_cells || !sync->tr_out)) return fail(HSSB_E_MODE, "training recurrence: separate projection and all three outputs needed");
    if (!train) {
        const bool f32 = out_f32 != nullptr;
        const uint64_t es = f32 ? 4 : 2;
        const uint64_t dims[3] = {(uint64_t)TC_OP, (uint64_t)T, (uint64_t)B};
        const uint64_t strides[2] = {(uint64_t)TC_OP * es, (uint64_t)prm.Tp * TC_OP * es};
        const uint32_t box[3] = {8, 1, RP_NBH};
        const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        if (int rc = make_tmap(&prm.out_map[0], dt, 3, f32 ? (const void *)out_f32 : (const void *)out_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        if (int rc = make_tmap(&prm.out_map[1], dt, 3, f32 ? (const void *)out_f32 : (const void *)out_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        const uint32_t box16[3] = {8, 1, RP_NBH / 2};
        if (int rc = make_tmap(&prm.out_map16[0], dt, 3, f32 ? (const void *)out_f32 : (const void *)out_hi, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        if (int rc = make_tmap(&prm.out_map16[1], dt, 3, f32 ? (const void *)out_f32 : (const void *)out_lo, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    }
    // at most 8 groups per direction are co-resident (16 clusters of 8 CTAs on 148 SMs); larger batches
    // run as successive launches over blocks of batch columns
    // Geometry: as few batch columns per cluster as the co-resident clusters allow (the DSMEM all-gather
    // volume per CTA and step is 1 KB per column), S sub-tiles of NB columns each.
    // HSSB_RC_GEOM="NB,S[,pair]" forces one geometry (experiments).
    int force_nb = 0, force_s = 0, force_pair = -1;
    if (const char *e = getenv("HSSB_RC_GEOM")) sscanf(e, "%d,%d,%d", &force_nb, &force_s, &force_pair);
    int max_clusters = 0;
    if (int rc = rc_dsmem_max_clusters(&max_clusters)) return rc;
    const int max_groups = max_clusters / 2;
    for (int64_t base = 0; base < B;) {
        const int64_t rem = B - base;
        prm.b_base = (int)base;
        const int64_t per_group = (rem + max_groups - 1) / max_groups;
        int nb, s, pair;
        if (force_nb) { nb = force_nb; s = force_s; pair = force_pair < 0 ? (nb % 32 == 0) : force_pair; }
        // the L2-multicast kernel (K5m) is the fastest at every batch size measured (scripts/sweep_recurrent.py); variants
        // pair = 2: one publisher per sub-tile, 3: per-warp publishing, 4: per-warp + two epilogue warps per TMEM quadrant
        else if (per_group <= 32) { nb = 32; s = 1; pair = 4; }
        else if (per_group <= 64) { nb = 32; s = 2; pair = 4; }
        else { nb = 32; s = 3; pair = 3; }
        if (fused && force_nb) return fail(HSSB_E_MODE, "HSSB_RC_GEOM cannot be combined with the fused projection");
        if ((skip_flag || train) && !(nb == 32 && pair >= 2)) return fail(HSSB_E_MODE, "the input-range guard / training forward need the multicast recurrence");
        // pair: 0 = DSMEM all-gather (K5), 1 = its cta_group::2 mode or, with nb = 64, the CTA-pair kernel (K5p); 2..4 = K5m variants
        int rc, done = 0;
        RecurLaunchInfo info = {0, 0};
        if (nb == 32 && pair >= 2) rc = rc_mc_launch(s, pair, fused, prm, m->tc_whh_frag[layer], rem, &done, xproj, st, &info, train);
        else {
            sync->multicast = false;
            if (sync->chunk_done || sync->tile_done) return fail(HSSB_E_MODE, "the overlapped projection needs the multicast recurrence");
            if (nb == 64) rc = rc_pair_launch(s, prm, m->tc_whh_frag[layer], rem, &done, xproj, st);
            else rc = rc_dsmem_launch(nb, s, pair, prm, rem, &done, xproj, st);
        }
        if (rc) return rc;
        if (sync->launches++ == 0) sync->ctas_first = info.ctas;
        sync->signals_per_dir += info.signals_per_dir;
        base += done;
    }
    return 0;
}

namespace {
constexpr int SYNC_MAX_CHUNKS = 8192;                                   // chunk / tile counters of the overlapped projection
constexpr size_t SYNC_HEAD = 256;                                        // next_item[8], timeout flag, resident counter
constexpr size_t SYNC_BYTES = SYNC_HEAD + 2 * sizeof(unsigned) * SYNC_MAX_CHUNKS;
struct TcWs { size_t xhi, xlo, xproj, o1hi, o1lo, out2, hn, cn, gather, range, sync, total; };
TcWs tc_ws_layout(int64_t B, int64_t T)
{
    const size_t M = (size_t)B * T;
    TcWs w{};
    size_t off = 0;
    const size_t Mp = (size_t)T * ((B + 31) / 32) * 32;          // batch padded to whole 32-column tiles (fused projection operand)
    w.xhi = off;   off += align_up(sizeof(__half) * Mp * 64, 1024);
    w.xlo = off;   off += align_up(sizeof(__half) * Mp * 64, 1024);
    w.xproj = off; off += align_up(sizeof(float) * 2 * (size_t)T * xproj_pitch(B) * TC_G, 1024);
    const size_t Ma = (size_t)B * act_pitch(T);                  // rows of the pitched [B][Tp][512] activations
    w.o1hi = off;  off += align_up(sizeof(__half) * Ma * TC_OP, 1024);
    w.o1lo = off;  off += align_up(sizeof(__half) * Ma * TC_OP, 1024);
    w.out2 = off;  off += align_up(sizeof(float) * Ma * TC_OP, 1024);
    w.hn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
    w.cn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
    w.gather = off; off += TC_GATHER_BYTES;
    w.range = off; off += 256;                                     // input-range guard: {flag, bits of max|x|}
    w.sync = off; off += align_up(SYNC_BYTES, 256);                // flags of the overlapped layer-2 projection
    w.total = off;
    return w;
}
}  // namespace

size_t tc_workspace_bytes(const hssb_model *, int64_t B, int64_t T) { return tc_ws_layout(B, T).total; }

// Holds `side` back until the layer-1 recurrence of the forward most recently enqueued on (m, ws) holds its SMs: work queued on
// `side` behind this (the next batch's FSST, hss.pipeline) then runs on the SMs the two latency-bound recurrences leave idle
// instead of delaying the placement of their clusters.
int tc_side_gate(const hssb_model *m, int64_t B, int64_t T, void *ws, cudaStream_t side)
{
    const TcWs w = tc_ws_layout(B, T);
    std::lock_guard<std::mutex> enqueue_lock(*m->enqueue_mu);
    unsigned char *sync_base = reinterpret_cast<unsigned char *>(static_cast<char *>(ws) + w.sync);
    m->pipelined->store(1, std::memory_order_relaxed);
    HSSB_CUDA_OK(cudaStreamWaitEvent(side, m->ev[4], 0));
    resident_gate_kernel<<<1, 32, 0, side>>>(reinterpret_cast<const unsigned *>(sync_base + 36));
    HSSB_LAUNCH_OK("resident_gate_kernel");
    return 0;
}

// Training forward of ONE layer (the caller applies ReLU / dropout between the layers, like reference segmenter.py:80-85):
// x[B][T][Kin] fp32 (Kin = input_size for layer 0, 480 for layer 1) -> split planes -> K4 -> K5m (TRAIN variant).  Layer 0 goes
// through the range-scaled projection, so any finite input is exact; layer 1's input is a dropout-scaled relu(h) and needs none.
int tc_train_forward(const hssb_model *m, int layer, const float *x, int64_t B, int64_t T, const float *h0, const float *c0,
                     float *gates, float *out, float *cells, float *hn, float *cn, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const TcWs w = tc_ws_layout(B, T);
    if (!ws || ws_bytes < w.total) return fail(HSSB_E_WORKSPACE, "model workspace %zu < %zu", ws_bytes, w.total);
    std::lock_guard<std::mutex> enqueue_lock(*m->enqueue_mu);
    char *base = static_cast<char *>(ws);
    float *xproj = reinterpret_cast<float *>(base + w.xproj);
    unsigned char *gather = reinterpret_cast<unsigned char *>(base + w.gather);
    unsigned *range = reinterpret_cast<unsigned *>(base + w.range);
    unsigned char *sync_base = reinterpret_cast<unsigned char *>(base + w.sync);
    HSSB_CUDA_OK(cudaMemsetAsync(range, 0, 8, st));
    HSSB_CUDA_OK(cudaMemsetAsync(sync_base, 0, SYNC_HEAD, st));
    const int64_t M = B * T;
    InprojJob job;
    job.next_item = reinterpret_cast<unsigned *>(sync_base);
    if (layer == 0) {
        __half *xhi = reinterpret_cast<__half *>(base + w.xhi), *xlo = reinterpret_cast<__half *>(base + w.xlo);
        {
            ProfScope prof("split_planes", st);
            input_amax_kernel<<<148 * 4, 256, 0, st>>>(x, M * m->F, range, nullptr);
            split_planes_kernel<<<(unsigned)std::min<long long>((M * 64 + 255) / 256, 148 * 16), 256, 0, st>>>(x, M, m->F, 64, xhi, xlo, range, nullptr);
            HSSB_LAUNCH_OK("split_planes_kernel");
        }
        if (int rc = tc_inproj(m, 0, xhi, xlo, 64, B, T, xproj, st, job, range, nullptr)) return rc;
    } else {
        __half *ahi = reinterpret_cast<__half *>(base + w.o1hi), *alo = reinterpret_cast<__half *>(base + w.o1lo);
        {
            ProfScope prof("split_planes", st);
            split_slots_kernel<<<148 * 16, 256, 0, st>>>(x, B, T, act_pitch(T), ahi, alo);
            HSSB_LAUNCH_OK("split_slots_kernel");
        }
        if (int rc = tc_inproj(m, 1, ahi, alo, TC_OP, B, T, xproj, st, job, nullptr, nullptr, act_pitch(T))) return rc;
    }
    RecurSync rs;
    rs.timeout_flag = reinterpret_cast<int *>(sync_base + 32);
    rs.tr_gates = gates; rs.tr_cells = cells; rs.tr_out = out;
    return tc_recurrent(m, layer, xproj, h0, c0, hn, cn, nullptr, nullptr, nullptr, gather, B, T, st, nullptr, nullptr, &rs);
}

namespace {
// Caller-owned pre-split input (hssb_model_split_input): [x hi planes][x lo planes][range words], the tile-major planes of tc_forward
struct SplitLayout { size_t hi, lo, range, total; };
SplitLayout split_layout(int64_t B, int64_t T)
{
    const size_t Mp = (size_t)T * ((B + 31) / 32) * 32;
    SplitLayout l{};
    size_t off = 0;
    l.hi = off; off += align_up(sizeof(__half) * Mp * 64, 1024);
    l.lo = off; off += align_up(sizeof(__half) * Mp * 64, 1024);
    l.range = off; off += 256;
    l.total = off;
    return l;
}
bool fused_input_path(const hssb_model *m)
{
    const char *fx = getenv("HSSB_FUSE_X");
    return m->F <= 16 * RX_KSTEPS && !getenv("HSSB_RC_GEOM") && !(fx && fx[0] == '0');
}
}  // namespace

size_t tc_split_bytes(const hssb_model *, int64_t B, int64_t T) { return split_layout(B, T).total; }

// The first kernel of tc_forward, ahead of time and into a caller-owned buffer: a pipelined caller runs it behind the next batch's
// FSST on its side stream, so that the forward starts with the layer-1 recurrence.
int tc_split_input(const hssb_model *m, const float *x, int64_t B, int64_t T, void *planes, size_t planes_bytes, cudaStream_t st)
{
    const SplitLayout l = split_layout(B, T);
    if (!planes || planes_bytes < l.total) return fail(HSSB_E_WORKSPACE, "split buffer %zu < %zu", planes_bytes, l.total);
    char *base = static_cast<char *>(planes);
    int *range_flag = reinterpret_cast<int *>(base + l.range);
    HSSB_CUDA_OK(cudaMemsetAsync(range_flag, 0, 8, st));
    if (!fused_input_path(m)) return 0;            // the separate-projection path splits (and scales) inside the forward
    ProfScope prof("split_planes", st);
    const long long n = T * ((B + RP_NBH - 1) / RP_NBH) * 8 * RP_NBH;
    split_planes_tiled_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, B, T, m->F, reinterpret_cast<__half *>(base + l.hi),
                                                                             reinterpret_cast<__half *>(base + l.lo), range_flag);
    HSSB_LAUNCH_OK("split_planes_tiled_kernel");
    return 0;
}

int tc_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0, const float *c0, float *logp,
               int32_t *labels, void *ws, size_t ws_bytes, cudaStream_t st, void *presplit)
{
    const TcWs w = tc_ws_layout(B, T);
    if (!ws || ws_bytes < w.total) return fail(HSSB_E_WORKSPACE, "model workspace %zu < %zu", ws_bytes, w.total);
    // Calls on different caller streams (each with its own workspace) share the model's internal streams and events: their
    // launches are enqueued one call at a time, the device work of different calls still overlaps where the streams allow.
    std::lock_guard<std::mutex> enqueue_lock(*m->enqueue_mu);
    char *base = static_cast<char *>(ws);
    __half *xhi = reinterpret_cast<__half *>(base + w.xhi), *xlo = reinterpret_cast<__half *>(base + w.xlo);
    float *xproj = reinterpret_cast<float *>(base + w.xproj);
    __half *o1hi = reinterpret_cast<__half *>(base + w.o1hi), *o1lo = reinterpret_cast<__half *>(base + w.o1lo);
    float *out2 = reinterpret_cast<float *>(base + w.out2);
    float *hn = reinterpret_cast<float *>(base + w.hn), *cn = reinterpret_cast<float *>(base + w.cn);
    unsigned char *gather = reinterpret_cast<unsigned char *>(base + w.gather);
    const int64_t M = B * T;
    // Layer 1's input projection (K = input_size <= 48) rides in the recurrence kernel: W_ih . x_t is issued into the accumulator
    // while the step's h_{t-1} is still in flight, so no xproj tensor (7.7 KB per sample written and read back) exists for it.
    // HSSB_FUSE_X=0 or a forced recurrence geometry selects the separate projection kernel instead.
    const bool fused = fused_input_path(m);
    // Input-range guard.  The fp16 hi/lo split is exact (22 bits) for |x| <= 65504; beyond that hi would be inf where the reference
    // (plain fp32, segmenter.py:80) is finite.  The separate-projection path therefore pre-scales x by 2^-e (e from the tensor's
    // max-abs, 0 while |x| < 2^15) and K4 scales the fp32 accumulator back by 2^e: exact for every finite input.  The fused
    // recurrence cannot scale one operand of its mixed accumulator, so its split kernel only raises a flag when an input leaves
    // the safe range; the fused launch is then a no-op and the pre-scaled path, launched behind it (a no-op otherwise), runs.
    int *range_flag = reinterpret_cast<int *>(base + w.range);
    const bool have_split = presplit && fused;                     // planes and range flag come from hssb_model_split_input
    if (have_split) {
        const SplitLayout l = split_layout(B, T);
        char *pb = static_cast<char *>(presplit);
        xhi = reinterpret_cast<__half *>(pb + l.hi); xlo = reinterpret_cast<__half *>(pb + l.lo);
        range_flag = reinterpret_cast<int *>(pb + l.range);
    } else {
        HSSB_CUDA_OK(cudaMemsetAsync(range_flag, 0, 8, st));
    }
    unsigned *range = reinterpret_cast<unsigned *>(range_flag);
    const int *standin = fused ? range_flag : nullptr;             // stand-in kernels run only when the flag was raised

    // ---- flags of the overlapped layer-2 projection ----------------------------------------------------------------------
    // Layer 2's projection GEMM (K4, throughput bound, 4.8 ms alone at 512 x 2000) sits between two latency-bound recurrences
    // that leave ~50 of the 148 SMs idle.  It runs as up to three launches (work orders: see K4):
    //   M  the middle time tiles, middle-out, on small clusters on the side stream UNDER the layer-1 recurrence, each tile as soon
    //      as both layer-1 directions have stored it (tile_done) -- gated until every layer-1 CTA is resident;
    //   A  the outer time tiles (both directions of a tile together: its A operand is read once) on all SMs, alone, until the
    //      layer-2 recurrence has enough of a head start;
    //   B  the tiles in between, one direction at a time in the order the layer-2 recurrence needs them, on small clusters UNDER
    //      that recurrence, which polls chunk_done before it reads a chunk -- gated until its clusters are on the machine.
    // The layer-2 recurrence runs on a high-priority internal stream; everything is joined into the caller's stream before the head.
    unsigned char *sync_base = reinterpret_cast<unsigned char *>(base + w.sync);
    unsigned *next_item = reinterpret_cast<unsigned *>(sync_base);                   // [8]
    int *timeout_flag = reinterpret_cast<int *>(sync_base + 32);
    unsigned *resident = reinterpret_cast<unsigned *>(sync_base + 36);               // [2][2]: {CTA counter, all-resident flag} of the layer-1 / layer-2 recurrence
    unsigned *chunk_done = reinterpret_cast<unsigned *>(sync_base + SYNC_HEAD);
    unsigned *tile_done = chunk_done + SYNC_MAX_CHUNKS;
    const int t_tiles = (int)((T + TC_TT - 1) / TC_TT), Q = 2 * t_tiles;
    auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    const bool overlap = env_int("HSSB_OVERLAP", 1) != 0 && Q <= SYNC_MAX_CHUNKS && t_tiles >= 8 && !getenv("HSSB_RC_GEOM") &&
                         m->sm_count >= 128 && m->hi_stream && m->side_stream;
    const int shape_small = std::min(3, std::max(1, env_int("HSSB_K4_SHAPE", 2)));
    // tiles [0, kA) and [t_tiles - kA, t_tiles): launch A (whole tiles, all SMs, alone); the tM middle tiles: launch M (whole tiles,
    // under layer 1); what lies between: launch B (single-direction chunks in the order the layer-2 recurrence needs them, under it)
    int kA = t_tiles / 2, tM = 0;
    if (overlap) {
        tM = t_tiles * std::min(40, std::max(0, env_int("HSSB_K4_MID", 30))) / 100;
        if (tM && ((tM ^ t_tiles) & 1)) --tM;                        // the middle tiles must be symmetric about the centre
        kA = std::min((t_tiles - tM) / 2, std::max(1, (t_tiles * std::min(100, std::max(5, env_int("HSSB_K4_SPLIT", 40))) / 100 + 1) / 2));
        HSSB_CUDA_OK(cudaMemsetAsync(sync_base, 0, SYNC_HEAD + sizeof(unsigned) * Q, st));
        if (tM) {
            HSSB_CUDA_OK(cudaMemsetAsync(tile_done, 0, sizeof(unsigned) * Q, st));
            HSSB_CUDA_OK(cudaEventRecord(m->ev[0], st));           // the side stream starts behind the cleared flags (and behind the previous forward)
        }
    } else {
        HSSB_CUDA_OK(cudaMemsetAsync(sync_base, 0, SYNC_HEAD, st));
    }
    HSSB_CUDA_OK(cudaEventRecord(m->ev[4], st));        // hssb_model_side_gate: the flags of THIS forward are cleared from here on

    // ---- layer 1 -----------------------------------------------------------------------------------------------------------
    RecurSync l1;
    l1.timeout_flag = timeout_flag;
    l1.resident = resident;                              // (also read by hssb_model_side_gate)
    if (tM) l1.tile_done = tile_done;
    int l1_ctas = 0;
    unsigned l1_signals = 0;
    bool l1_single = true;
    if (fused) {
        if (!have_split) {
            ProfScope prof("split_planes", st);
            const long long n = T * ((B + RP_NBH - 1) / RP_NBH) * 8 * RP_NBH;
            split_planes_tiled_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, B, T, m->F, xhi, xlo, range_flag);
            HSSB_LAUNCH_OK("split_planes_tiled_kernel");
        }
        l1.skip_flag = range_flag; l1.skip_when = 1;
        if (int rc = tc_recurrent(m, 0, nullptr, h0, c0, hn, cn, o1hi, o1lo, nullptr, gather, B, T, st, xhi, xlo, &l1)) return rc;
        l1_ctas = l1.ctas_first; l1_signals = l1.signals_per_dir; l1_single = l1.launches == 1;
    }
    {
        ProfScope prof(fused ? "range_standin" : "split_planes", st);
        input_amax_kernel<<<148 * 4, 256, 0, st>>>(x, M * m->F, range, standin);
        split_planes_kernel<<<(unsigned)std::min<long long>((M * 64 + 255) / 256, 148 * 16), 256, 0, st>>>(x, M, m->F, 64, xhi, xlo, range, standin);
        HSSB_LAUNCH_OK("split_planes_kernel");
    }
    {
        InprojJob job;
        job.next_item = next_item + 0;
        if (fused) job.name = "range_standin";
        if (int rc = tc_inproj(m, 0, xhi, xlo, 64, B, T, xproj, st, job, range, standin)) return rc;
    }
    l1.skip_flag = standin; l1.skip_when = 0;
    if (int rc = tc_recurrent(m, 0, xproj, h0, c0, hn, cn, o1hi, o1lo, nullptr, gather, B, T, st, nullptr, nullptr, &l1)) return rc;
    if (!fused) { l1_ctas = l1.ctas_first; l1_signals = l1.signals_per_dir; l1_single = l1.launches == 1; }
    else if (l1.ctas_first != l1_ctas || l1.signals_per_dir != l1_signals) l1_single = false;    // (the stand-in must look like the launch it replaces)
    // a forced geometry (HSSB_RC_GEOM) runs kernels that do not report residency: raise the flag behind them, so that a
    // hssb_model_side_gate waits for the end of layer 1 instead of its 20 ms time-out
    if (!l1.multicast) HSSB_CUDA_OK(cudaMemsetAsync(resident + 1, 1, sizeof(unsigned), st));

    // ---- layer 2 -----------------------------------------------------------------------------------------------------------
    const unsigned chunk_need = (unsigned)(B * IP_NT_DIR * 4);        // finished (tile, epilogue warp) pairs of a chunk: 4 epilogue warps
    if (!overlap) {
        InprojJob job;
        job.next_item = next_item + 1;
        if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, st, job, nullptr, nullptr, act_pitch(T))) return rc;
        if (int rc = tc_recurrent(m, 1, xproj, hn, cn, hn, cn, nullptr, nullptr, out2, gather, B, T, st)) return rc;
    } else {
        bool mid_running = false;
        if (tM && l1_single && l1.multicast && l1_ctas > 0 && m->sm_count - l1_ctas >= 16) {
            // launch M: behind the gate on the side stream, concurrent with the layer-1 launches enqueued above
            HSSB_CUDA_OK(cudaStreamWaitEvent(m->side_stream, m->ev[0], 0));
            // Pipelined callers (hssb_model_side_gate seen): hold the launch back until its first tile is stored, so that the idle SMs
            // of the first half of layer 1 are really free (costs 0.1 ms of the step on its own, wins 0.5 ms with the next batch's
            // FSST running there).  Otherwise: as soon as the recurrence is resident (its CTAs then poll in place).  HSSB_M_GATE forces.
            if (env_int("HSSB_M_GATE", m->pipelined->load(std::memory_order_relaxed))) {
                const int tt0 = (t_tiles & 1) ? (t_tiles - 1) / 2 : t_tiles / 2 - 1;       // the first tile of the middle-out order
                tile_gate_kernel<<<1, 32, 0, m->side_stream>>>(tile_done + tt0, tile_done + t_tiles + tt0, l1_signals);
            } else {
                resident_gate_kernel<<<1, 32, 0, m->side_stream>>>(resident);
            }
            HSSB_LAUNCH_OK("gate kernel");
            InprojJob mid;
            mid.unit_mode = 1; mid.u_lo = 0; mid.n_units = tM; mid.shape = shape_small;
            mid.next_item = next_item + 3; mid.chunk_done = chunk_done;
            mid.src_done = tile_done; mid.src_need = l1_signals; mid.timeout_flag = timeout_flag;
            // (input-range guard fired: layer 1 runs as the slower stand-in chain, which this launch must not crowd out -- it
            //  stands down and its tiles are produced by a stand-in of launch B below)
            mid.skip_flag = standin;
            mid.name = "tc_inproj_l1_mid";
            if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, m->side_stream, mid, nullptr, nullptr, act_pitch(T))) return rc;
            HSSB_CUDA_OK(cudaEventRecord(m->ev[3], m->side_stream));
            mid_running = true;
        } else {
            tM = 0;
        }
        const int m_lo = (t_tiles - tM) / 2, m_hi = m_lo + tM - 1;      // middle tiles [m_lo, m_hi] (empty when tM == 0)
        InprojJob a;
        a.unit_mode = 0; a.u_lo = 0; a.n_units = 2 * kA; a.shape = 0; a.next_item = next_item + 1; a.chunk_done = chunk_done;
        if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, st, a, nullptr, nullptr, act_pitch(T))) return rc;
        HSSB_CUDA_OK(cudaEventRecord(m->ev[1], st));
        // launch B: chunks k in [kA, m_lo) and (m_hi, t_tiles - kA) of both directions, i.e. q in [2 kA, 2 m_lo) and [2 m_hi + 2, Q - 2 kA).
        // It is ENQUEUED before the recurrence (a profiler that serialises kernels then runs it first and the recurrence finds every
        // flag set) but held back by the gate until the recurrence's clusters are on the machine: they are placed first, B takes
        // the SMs that are left.
        InprojJob b;
        b.unit_mode = 2; b.shape = shape_small; b.next_item = next_item + 2; b.chunk_done = chunk_done;
        b.q_lo1 = 2 * kA; b.len1 = tM ? 2 * (m_lo - kA) : Q - 4 * kA;
        b.q_lo2 = 2 * m_hi + 2; b.n_units = tM ? b.len1 + (Q - 2 * kA) - (2 * m_hi + 2) : b.len1;
        b.name = "tc_inproj_l1_tail";
        if (b.n_units > 0) {
            resident_gate_kernel<<<1, 32, 0, st>>>(resident + 2);
            HSSB_LAUNCH_OK("resident_gate_kernel");
            if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, st, b, nullptr, nullptr, act_pitch(T))) return rc;
        }
        if (mid_running && standin) {
            InprojJob bm;                  // the middle tiles, when launch M stood down (no-op otherwise)
            bm.unit_mode = 2; bm.shape = shape_small; bm.next_item = next_item + 4; bm.chunk_done = chunk_done;
            bm.q_lo1 = 2 * m_lo; bm.len1 = 2 * tM; bm.n_units = bm.len1;
            bm.name = "range_standin";
            if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, st, bm, nullptr, standin, act_pitch(T))) return rc;
        }
        HSSB_CUDA_OK(cudaStreamWaitEvent(m->hi_stream, m->ev[1], 0));
        RecurSync l2;
        l2.chunk_done = chunk_done; l2.chunk_need = chunk_need; l2.timeout_flag = timeout_flag; l2.resident = resident + 2;
        if (int rc = tc_recurrent(m, 1, xproj, hn, cn, hn, cn, nullptr, nullptr, out2, gather, B, T, m->hi_stream, nullptr, nullptr, &l2)) return rc;
        HSSB_CUDA_OK(cudaEventRecord(m->ev[2], m->hi_stream));
        HSSB_CUDA_OK(cudaStreamWaitEvent(st, m->ev[2], 0));
        if (mid_running) HSSB_CUDA_OK(cudaStreamWaitEvent(st, m->ev[3], 0));
    }
    return head_forward(out2, M, TC_OP, m->tc_lin_w, m->lin_b, logp, labels, st, timeout_flag, T, act_pitch(T));
}

}  // namespace hssb

// Diagnostic: clock64 stamps of the recurrence roles (cluster 0, rank 0) for the first `steps` steps of every
// following recurrence launch; buf = steps*4*16 uint64 on the device, nullptr disables.
extern "C" int hssb_debug_max_clusters(void)
{
    int n = 0;
    if (hssb::rc_dsmem_max_clusters(&n)) return -1;
    return n;
}

// Diagnostic: byte offset of the producer / consumer flag area inside the model workspace ({next_item[8], timeout flag, resident
// counter} in the first 256 bytes, then chunk_done[8192] and tile_done[8192]); read by scripts/dump_sync.py.
extern "C" long long hssb_debug_sync_offset(long long B, long long T)
{
    if (B <= 0 || T <= 0) return -1;
    return (long long)hssb::tc_ws_layout(B, T).sync;
}

extern "C" int hssb_debug_trace(unsigned long long *buf, int steps)
{
    hssb::g_trace_buf = buf;
    hssb::g_trace_steps = buf ? steps : 0;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Diagnostic entry point: layer-1 input projection only, canonical layout, for kernel-level parity
// tests (impl 0 = tcgen05 kernel, 1 = SIMT kernel).  xproj: [2][B*T][960] fp32 (torch gate order).
// workspace: 2*T*(B|1)*960*4 + 2*B*T*64*2*2 bytes (xproj rows are pitched to an odd batch count).
// ------------------------------------------------------------------------------------------------
extern "C" int hssb_debug_inproj(const hssb_model *m, const float *x, int64_t B, int64_t T, int impl, float *xproj, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    using namespace hssb;
    if (!m || !x || !xproj) return fail(HSSB_E_NULL, "hssb_debug_inproj: null pointer");
    if (B <= 0 || T <= 0) return fail(HSSB_E_SHAPE, "hssb_debug_inproj: bad shape");
    cudaStream_t st = as_stream(stream);
    const int64_t M = B * T;
    if (impl == 1) {
        for (int d = 0; d < 2; ++d)
            if (int rc = simt_inproj(x, M, m->F, m->w_ihT[0][d], m->bias[0][d], 4 * m->H, xproj + (size_t)d * M * 4 * m->H, st)) return rc;
        return 0;
    }
    if (m->H != TC_H || m->F > 64 || !m->tc_wih[0]) return fail(HSSB_E_MODEL, "tcgen05 kernels need hidden_size 240");
    const size_t raw_floats = (size_t)2 * T * xproj_pitch(B) * TC_G;
    const size_t need = sizeof(float) * raw_floats + sizeof(__half) * 2 * M * 64 + 256;
    if (!workspace || workspace_bytes < need) return fail(HSSB_E_WORKSPACE, "hssb_debug_inproj: workspace %zu < %zu", workspace_bytes, need);
    float *raw = static_cast<float *>(workspace);
    __half *hi = reinterpret_cast<__half *>(raw + raw_floats), *lo = hi + M * 64;
    split_planes_kernel<<<(unsigned)std::min<long long>((M * 64 + 255) / 256, 148 * 16), 256, 0, st>>>(x, M, m->F, 64, hi, lo, nullptr, nullptr);
    HSSB_LAUNCH_OK("split_planes_kernel");
    InprojJob job;
    job.next_item = reinterpret_cast<unsigned *>(hi + 2 * M * 64);
    HSSB_CUDA_OK(cudaMemsetAsync(job.next_item, 0, 256, st));
    if (int rc = tc_inproj(m, 0, hi, lo, 64, B, T, raw, st, job)) return rc;
    unpermute_xproj_kernel<<<(unsigned)((2 * M * TC_G + 255) / 256), 256, 0, st>>>(raw, B, xproj_pitch(B), T, xproj);
    HSSB_LAUNCH_OK("unpermute_xproj_kernel");
    return 0;
}

