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
// K4  tc_inproj_kernel : xproj[dir][t][b][g'] = A[b,t,:] . W_ih[g',:] + (b_ih + b_hh)[g']   (see the kernel)
// K5  tc_recurrent_kernel : the T sequential steps, 8-CTA clusters, weights resident in TMEM (see the kernel)
#include "model.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <mutex>
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
constexpr size_t TC_GATHER_BYTES = (size_t)16 * 8 * 3 * 2 * 4096;   // L2 scratch of the multicast all-gather: [cluster][rank][S][parity][4 KB]

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
__global__ void pack_whh_kernel(const float *__restrict__ w, int dir, int frag, __half *__restrict__ dst);   // defined with K5
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
__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// x[M,F] fp32 -> hi/lo fp16 planes [M,Kp] (zero padded columns)
__global__ void split_planes_kernel(const float *__restrict__ x, long long M, int F, int Kp, __half *__restrict__ hi,
                                    __half *__restrict__ lo)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * Kp) return;
    const long long m = i / Kp;
    const int k = (int)(i % Kp);
    __half h = __float2half_rn(0.f), l = h;
    if (k < F) split_f16(x[m * F + k], h, l);
    hi[i] = h;
    lo[i] = l;
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
// K4: input projection GEMM, persistent + warp specialised, clusters of 8 CTAs (4 along M x 2 along N).
//
//   xproj[dir][t][b][g'] = A[b,t,:] . W_ih[g',:] + (b_ih + b_hh)[g']          M = B*T, N = 1920, K = 48 / 480
//
// Per CTA one 128(t) x 192(g') output tile at a time; three fp16 MMAs per product (hi*hi, lo*hi, hi*lo).
// The operands come out of L2, whose bandwidth (~7 TB/s, the same order as HBM) is what bounded the first
// version of this kernel: a lone CTA re-reads (128 + 192) x K x 4 bytes per tile.  Here the eight CTAs of a
// cluster work on 4 consecutive m-tiles x 2 consecutive n-tiles and share their loads by TMA multicast:
// the A stage (128 rows) is fetched in two halves by the two CTAs of a column and multicast to both, the
// W stage (192 rows) in four quarters by the four CTAs of a row -> (128/2 + 192/4) rows per CTA and stage.
//   warp 0   : TMA producer (4-stage ring of 40 KB stages, BK = 32, SW64: the refill of a stage waits for the
//              slowest of the 5 CTAs that read it, so depth matters more than stage size)
//   warp 1   : tcgen05.mma issuer; accumulators double-buffered in TMEM (2 x 192 columns) so that the
//              epilogue of tile i overlaps the main loop of tile i+1; smem stages are released to all
//              CTAs that write into this one with a multicast tcgen05.commit
//   warps 2-5: epilogue TMEM -> registers (+bias) -> swizzled smem -> TMA store
// ------------------------------------------------------------------------------------------------
constexpr int IP_BM = 128, IP_BN = 192, IP_BK = 32;   // BK = 32 fp16 = 64-byte rows: SWIZZLE_64B
constexpr int IP_CM = 4, IP_CN = 2, IP_CL = IP_CM * IP_CN;   // cluster shape
constexpr int IP_A_BYTES = IP_BM * IP_BK * 2;            // one fp16 plane of the A stage (8 KB)
constexpr int IP_B_BYTES = IP_BN * IP_BK * 2;            // one fp16 plane of the B stage (12 KB)
constexpr int IP_STAGE_BYTES = 2 * IP_A_BYTES + 2 * IP_B_BYTES;   // 40 KB
constexpr int IP_OUT_TILE = 32 * 32 * 4;                 // epilogue transposition tile of one warp: 32 rows x 32 fp32 (4 KB)
constexpr int IP_OUT_RING = 1;                           // tiles per warp
constexpr int IP_BIAS_BYTES = TC_NG * 4;                 // all 1920 folded biases, staged once per CTA
// STAGES / EPI_WARPS: layer 2's projection (K = 512) is main-loop bound: 5 smem stages, 4 epilogue warps.  Layer 1's (K = 48)
// is epilogue bound -- one warp per SMSP cannot hide the TMEM-load / shared-memory latencies of the drain: 3 stages, 8 warps.
template <int STAGES, int EPI_WARPS>
struct IpCfg {
    static constexpr int OUT_BYTES = EPI_WARPS * IP_OUT_RING * IP_OUT_TILE;
    static constexpr int SMEM_BYTES = STAGES * IP_STAGE_BYTES + OUT_BYTES + IP_BIAS_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "one or two warps per TMEM lane quadrant");
};
constexpr int IP_TMEM_COLS = 512;                        // 2 accumulators of 192 columns (at 0 and 256)
constexpr int IP_N_TILES = TC_NG / IP_BN;                // 10
static_assert(IP_N_TILES % IP_CN == 0, "n-tiles must split over the cluster");

struct InprojParams {
    CUtensorMap a_hi, a_lo;   // [k, t, b] fp16, box (32, 64, 1), SW64   (half of the A stage)
    CUtensorMap w_hi, w_lo;   // [k, g'(1920)] fp16, box (32, 48), SW64  (quarter of the W stage)
    float *out;               // xproj [dir][t][Bp][960] fp32
    const float *bias;        // [1920]
    long long B, Bp;          // batch, and the row pitch of xproj in batch rows (see xproj_pitch)
    int debug;                // HSSB_IP_DEBUG bit 0: skip the global stores (timing experiment; results are wrong when set)
    int k_real;               // true K rounded up to 16 (48 / 512)
    int T;
    int t_tiles;              // ceil(T/128)
    int m_groups;             // ceil(B * t_tiles / 4)
};

template <int IP_STAGES, int EPI_WARPS>
__global__ void __launch_bounds__(IpCfg<IP_STAGES, EPI_WARPS>::THREADS, 1) tc_inproj_kernel(const __grid_constant__ InprojParams p)
{
    using C = IpCfg<IP_STAGES, EPI_WARPS>;
    constexpr int IP_OUT_BYTES = C::OUT_BYTES;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    unsigned char *out_base = smem + IP_STAGES * IP_STAGE_BYTES;
    float *bias_s = reinterpret_cast<float *>(out_base + IP_OUT_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_base + IP_OUT_BYTES + IP_BIAS_BYTES);
    uint64_t *full = bars, *empty = bars + IP_STAGES, *tmem_full = bars + 2 * IP_STAGES, *tmem_empty = bars + 2 * IP_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * IP_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cx = rank % IP_CM, cy = rank / IP_CM;                // position in the cluster: m / n
    const uint16_t mask_a = (uint16_t)((1u << cx) | (1u << (cx + IP_CM)));          // CTAs sharing my A tile
    const uint16_t mask_w = (uint16_t)(((1u << IP_CM) - 1u) << (IP_CM * cy));       // CTAs sharing my W tile
    const int kblocks = (p.k_real + IP_BK - 1) / IP_BK;
    const int cluster_id = blockIdx.x / IP_CL, n_clusters = gridDim.x / IP_CL;
    const int n_items = p.m_groups * (IP_N_TILES / IP_CN);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.a_hi); prefetch_tmap(&p.a_lo); prefetch_tmap(&p.w_hi); prefetch_tmap(&p.w_lo);
        for (int s = 0; s < IP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], IP_CM + IP_CN - 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<IP_TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < TC_NG; i += C::THREADS) bias_s[i] = __ldg(p.bias + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();     // barriers of every CTA exist before any multicast can target them

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int m_tile = (item / (IP_N_TILES / IP_CN)) * IP_CM + cx;
                const int b = m_tile / p.t_tiles, t0 = (m_tile % p.t_tiles) * IP_BM;
                const int n0 = ((item % (IP_N_TILES / IP_CN)) * IP_CN + cy) * IP_BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % IP_STAGES;
                    mbar_wait_cluster(&empty[s], ((it / IP_STAGES) & 1) ^ 1);
                    unsigned char *st = stage_base + s * IP_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], IP_STAGE_BYTES);
                    // my half of the A tile (rows cy*64 ..) -> both CTAs of my cluster column
                    tma_load_3d_mc(st + cy * (IP_A_BYTES / 2), &p.a_hi, &full[s], kb * IP_BK, t0 + cy * (IP_BM / 2), b, mask_a);
                    tma_load_3d_mc(st + IP_A_BYTES + cy * (IP_A_BYTES / 2), &p.a_lo, &full[s], kb * IP_BK, t0 + cy * (IP_BM / 2), b, mask_a);
                    // my quarter of the W tile (rows cx*48 ..) -> the four CTAs of my cluster row
                    tma_load_2d_mc(st + 2 * IP_A_BYTES + cx * (IP_B_BYTES / 4), &p.w_hi, &full[s], kb * IP_BK, n0 + cx * (IP_BN / 4), mask_w);
                    tma_load_2d_mc(st + 2 * IP_A_BYTES + IP_B_BYTES + cx * (IP_B_BYTES / 4), &p.w_lo, &full[s], kb * IP_BK, n0 + cx * (IP_BN / 4), mask_w);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(IP_BM, IP_BN);
            const uint16_t mask_rel = mask_a | mask_w;     // every CTA that writes into my stages
            uint32_t it = 0, tile = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, ++tile) {
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
                        mma_f16_ss(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);
                        mma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1);
                        mma_f16_ss(d_tmem, da_hi, db_lo, idesc, 1);
                    }
                    mma_commit_mc(&empty[s], mask_rel);    // frees this stage in every CTA that fills it
                }
                mma_commit(&tmem_full[acc]);               // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2.., TMEM lane quadrant = warp % 4; with 8 warps the two warps of a quadrant take alternate chunks =====
        // A warp first pulls ALL of its 32-column chunks of the accumulator into registers and hands the accumulator back to the
        // MMA warp at once (the stores below are slow -- 1.5 ms of layer 2's projection -- and must not hold TMEM), then per chunk:
        // registers (one row of 32 columns per thread, + bias) -> xor-swizzled smem tile -> 16-byte global stores in which 8
        // consecutive lanes cover one 128-byte row segment (every store instruction writes four full lines).
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CSTEP = EPI_WARPS / 4;
        constexpr int NCH = (IP_BN / 32) / CSTEP;           // chunks per warp: 6 (4 warps) or 3 (8 warps)
        constexpr int NPASS = (NCH > 3) ? 2 : 1;            // registers hold 3 chunks at a time
        constexpr int PCH = NCH / NPASS;
        unsigned char *ob = out_base + (warp - 2) * IP_OUT_TILE;
        uint32_t tile = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++tile) {
            const int m_tile = (item / (IP_N_TILES / IP_CN)) * IP_CM + cx;
            const int b = m_tile / p.t_tiles, t0 = (m_tile % p.t_tiles) * IP_BM;
            const int n0 = ((item % (IP_N_TILES / IP_CN)) * IP_CN + cy) * IP_BN;
            const int dir = n0 / TC_G, nl0 = n0 % TC_G;
            const uint32_t acc = tile & 1;
            mbar_wait(&tmem_full[acc], (tile >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
                uint32_t v[PCH][32];
#pragma unroll
                for (int i = 0; i < PCH; ++i)
                    tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (half + (pass * PCH + i) * CSTEP) * 32, v[i]);
                tmem_ld_wait();
                if (pass == NPASS - 1) {                    // my part of the accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
#pragma unroll
                for (int i = 0; i < PCH; ++i) {
                    const int c = half + (pass * PCH + i) * CSTEP;
                    const float4 *bias = reinterpret_cast<const float4 *>(bias_s + n0 + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bj = bias[j];                 // shared-memory broadcast
                        float4 o;
                        o.x = __uint_as_float(v[i][4 * j + 0]) + bj.x;
                        o.y = __uint_as_float(v[i][4 * j + 1]) + bj.y;
                        o.z = __uint_as_float(v[i][4 * j + 2]) + bj.z;
                        o.w = __uint_as_float(v[i][4 * j + 3]) + bj.w;
                        *reinterpret_cast<float4 *>(ob + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;   // 128B xor swizzle: conflict free both ways
                    }
                    __syncwarp();
                    const int rsub = lane >> 3, c16 = lane & 7;
                    float *gout = p.out + (((size_t)dir * p.T + t0 + q * 32) * p.Bp + b) * TC_G + nl0 + c * 32 + c16 * 4;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = 4 * j + rsub;
                        const float4 o = *reinterpret_cast<const float4 *>(ob + row * 128 + ((c16 ^ (row & 7)) << 4));
                        if (b < p.B && t0 + q * 32 + row < p.T && !(p.debug & 1))                    // (b >= B: padding tiles of the last m-group)
                            __stcs(reinterpret_cast<float4 *>(gout + (size_t)row * p.Bp * TC_G), o);
                    }
                    __syncwarp();
                }
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
    // the raw torch tensors may be host pointers: stage them through a temporary device buffer
    float *tmp = nullptr;
    const size_t tmp_floats = (size_t)TC_G * (2 * TC_H) + 2 * TC_G;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&tmp), sizeof(float) * tmp_floats, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(tc_pack)");
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
            pack_wih_kernel<<<TC_G, 128, 0, st>>>(w, bi, bh, kin[l], Kp, d, l, hi, lo, m->tc_bias[l]);
            if (l == 0) pack_wih0_frag_kernel<<<8 * 128, 64, 0, st>>>(w, bi, bh, kin[0], d, m->tc_wih0_frag, m->tc_bias0_frag);
            if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "pack_wih_kernel"); break; }
            if ((e = cudaMemcpyAsync(w, p->w_hh[l][d], sizeof(float) * TC_G * TC_H, cudaMemcpyDefault, st)) != cudaSuccess) {
                rc = cuda_fail(e, "cudaMemcpyAsync(tc_pack w_hh)");
                break;
            }
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
    cudaFreeAsync(tmp, st);
    if (rc) return rc;
    m->tc_ready = true;
    return 0;
}

template <int STAGES, int EPI_WARPS>
static int launch_inproj(const InprojParams &prm, int n_items, const char *name, cudaStream_t st)
{
    using C = IpCfg<STAGES, EPI_WARPS>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = IP_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_clusters = 0;
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(tc_inproj_kernel<STAGES, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_inproj_kernel)");
        cfg.gridDim = dim3(IP_CL * 16);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_inproj_kernel<STAGES, EPI_WARPS>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_inproj_kernel)");
        if (n < 1) return fail(HSSB_E_DEVICE, "device cannot host an input-projection cluster");
        max_clusters = n;
    }
    cfg.gridDim = dim3((unsigned)(IP_CL * std::min(max_clusters, n_items)));
    ProfScope prof(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_inproj_kernel<STAGES, EPI_WARPS>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_inproj_kernel)");
    return 0;
}

// One layer's input projection on the tensor cores.  a_hi/a_lo: [B*T][pitch] fp16 planes.
int tc_inproj(const hssb_model *m, int layer, const __half *a_hi, const __half *a_lo, int pitch_elems, int64_t B, int64_t T,
              float *xproj /*[2][T][B][960]*/, cudaStream_t st)
{
    InprojParams prm;
    const int Kp = kp_of_layer(layer, m->F);
    const int kreal = kreal_of_layer(layer, m->F);
    {
        const uint64_t dims[3] = {(uint64_t)(layer == 0 ? Kp : kreal), (uint64_t)T, (uint64_t)B};
        const uint64_t strides[2] = {(uint64_t)pitch_elems * 2, (uint64_t)T * pitch_elems * 2};
        const uint32_t box[3] = {IP_BK, IP_BM / IP_CN, 1};
        if (int rc = make_tmap(&prm.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        if (int rc = make_tmap(&prm.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)TC_NG};
        const uint64_t strides[1] = {(uint64_t)Kp * 2};
        const uint32_t box[2] = {IP_BK, IP_BN / IP_CM};
        const __half *hi = m->tc_wih[layer], *lo = hi + (size_t)TC_NG * Kp;
        if (int rc = make_tmap(&prm.w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        if (int rc = make_tmap(&prm.w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    prm.out = xproj;
    prm.B = B;
    prm.Bp = xproj_pitch(B);
    prm.debug = 0;
    if (const char *e = getenv("HSSB_IP_DEBUG")) prm.debug = atoi(e);
    prm.bias = m->tc_bias[layer];
    prm.k_real = kreal;
    prm.T = (int)T;
    prm.t_tiles = (int)((T + IP_BM - 1) / IP_BM);
    prm.m_groups = (int)((B * prm.t_tiles + IP_CM - 1) / IP_CM);

    const int n_items = prm.m_groups * (IP_N_TILES / IP_CN);
    if (layer == 0) return launch_inproj<3, 8>(prm, n_items, "tc_inproj_l0", st);
    if (const char *e = getenv("HSSB_IP_L1")) {
        if (atoi(e) == 48) return launch_inproj<4, 8>(prm, n_items, "tc_inproj_l1", st);
        if (atoi(e) == 44) return launch_inproj<4, 4>(prm, n_items, "tc_inproj_l1", st);
    }
    return launch_inproj<5, 4>(prm, n_items, "tc_inproj_l1", st);
}

// ------------------------------------------------------------------------------------------------
// K5: recurrence.  One 8-CTA cluster per (direction, group of S*NB batch columns).
//
// Orientation: gates are the MMA M dimension and stay put, the batch is N:
//     G^T[g' (128 lanes), b (NB cols)] = W_hh,slice[g', k] . h_{t-1}^T[k, b]  (+ xproj^T added in the epilogue)
// CTA rank r owns units 30r..30r+29 -> gate rows (TMEM lanes) 4*u + q (q = i,f,g,o; lanes 120..127 are zero
// padding), so the four gates of a unit are four adjacent lanes of one warp.
//   * W_hh slice (hi and lo fp16 planes, K padded 240 -> 8*32) is loaded ONCE into TMEM columns
//     [0,256) and is the A operand of every MMA (tcgen05.mma with A in TMEM) -- weights never move.
//   * h_{t-1}^T lives in shared memory as the B operand (K-major, no swizzle, [rank][plane][k-chunk][b][8]).
//     After its epilogue each CTA owns 30 fresh h values per batch column; it writes them as an fp16
//     hi/lo "image" and one elected thread pushes that image into the B buffer of all 8 CTAs with
//     cp.async.bulk shared::cta -> shared::cluster, completing on the receiver's mbarrier
//     (the all-gather of the recurrence, no global memory round trip, no cluster barrier).
//   * Epilogue per step (4 warps per sub-tile, one TMEM lane quadrant each): tcgen05.ld the 32 x NB
//     accumulator slice, add xproj (plain coalesced 128-byte loads, prefetched one step ahead into
//     registers), branch-free sigmoid / tanh (tanh x = 2 sigmoid 2x - 1; MUFU.EX2 + MUFU.RCP), 4x4
//     lane transposes (shfl.xor 1, 2) that hand thread (u, j) the four gates of unit u for the batch
//     columns b = j (mod 4), then the c/h update with the cell state in registers.  No shared-memory
//     round trip and no block barrier between the gate activations and the cell update.
// Sub-tiles: S independent groups of NB batch columns are interleaved per cluster so that the tensor
// pipe (one sub-tile's MMAs) overlaps the MUFU work of another's epilogue and the DSMEM all-gather of
// the third.
// ------------------------------------------------------------------------------------------------
constexpr int RC_CL = 8;            // CTAs per cluster
constexpr int RC_U = 30;            // real units per CTA
constexpr int RC_KP = 256;          // padded K (8 ranks x 32 slots)
constexpr int RC_XW = 4 * RC_U;     // xproj floats per (t, b) owned by one CTA (120)

// PAIR: the two CTAs of a TPC issue one tcgen05.mma.cta_group::2 (M = 256 gate rows, N = NB columns) whose B
// operand is split between them (NB/2 columns each), so each CTA receives only half of the all-gather.
template <int NB, int S, bool PAIR>
struct RcCfg {
    static constexpr int NBH = PAIR ? NB / 2 : NB;              // batch columns of the B operand held by one CTA
    static constexpr int SLICE_BYTES = NBH * 32 * 2 * 2;        // one rank's slot: [plane][4 chunks][NBH][8] fp16
    static constexpr int HBUF_BYTES = RC_CL * SLICE_BYTES;      // one B-operand buffer (hi+lo planes)
    static constexpr int IMG_BYTES = NB * 32 * 2 * 2;           // this CTA's h_t of all NB columns ([half] x slot layout)
    static constexpr int PER_SUB = 2 * HBUF_BYTES + 2 * IMG_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = S * PER_SUB + BAR_BYTES + 1024;
    static constexpr int THREADS = 32 * S + 128 * S;         // S MMA-issuer warps + S epilogue groups of 4 warps
    static_assert(NB % 16 == 0 && NB <= 64, "NB must be 16, 32, 48 or 64");
    static_assert(S * NB <= 256, "accumulators must fit in the TMEM columns left of the weights");
    static_assert(5 * S * 8 <= BAR_BYTES - 8, "barrier area too small");
    static_assert(!PAIR || NB % 32 == 0, "pair mode splits NB in two halves of a multiple of 16 columns");
};

struct RecurParams {
    const float *xproj;         // [dir][t][b][g'(960)] fp32 (cluster gate order)
    const __half *whh;          // [dir][rank][plane][128][256] fp16 (cluster gate order, zero padded)
    const float *h0, *c0;       // [2][B][240]
    float *hn, *cn;             // [2][B][240]  raw final state
    __half *out_hi, *out_lo;    // layer 1: relu(h) planes [B*T][480]  (nullptr for layer 2)
    float *out_f32;             // layer 2: relu(h) [B*T][480]         (nullptr for layer 1)
    long long B, T;
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
// 4x4 transpose across the 4 lanes of a quad: in r[c] = a[lane j][c]  ->  out r[g] = a[lane g][j]
__device__ __forceinline__ void quad_transpose(float (&r)[4], int j)
{
    const bool o1 = (j & 1) != 0, o2 = (j & 2) != 0;
    float s0 = o1 ? r[0] : r[1], s1 = o1 ? r[2] : r[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    if (o1) { r[0] = s0; r[2] = s1; } else { r[1] = s0; r[3] = s1; }
    s0 = o2 ? r[0] : r[2];
    s1 = o2 ? r[1] : r[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (o2) { r[0] = s0; r[1] = s1; } else { r[2] = s0; r[3] = s1; }
}

template <int NB, int S, bool PAIR>
__global__ void __launch_bounds__(RcCfg<NB, S, PAIR>::THREADS, 1) tc_recurrent_kernel(const __grid_constant__ RecurParams p)
{
    using C = RcCfg<NB, S, PAIR>;
    constexpr int NBH = C::NBH;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    auto hbuf = [&](int s, int par) { return smem + s * C::PER_SUB + par * C::HBUF_BYTES; };
    auto image = [&](int s, int par) { return smem + s * C::PER_SUB + 2 * C::HBUF_BYTES + par * C::IMG_BYTES; };
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * C::PER_SUB);
    uint64_t *h_full = bars;                 // [S][2]  my B-operand buffer is complete
    uint64_t *d_full = bars + 2 * S;         // [S]     accumulator complete
    uint64_t *peer_full = bars + 3 * S;      // [S][2]  (pair leader) the odd CTA's buffer is complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5 * S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x / RC_CL;
    const int dir = cid & 1;
    const int group = cid >> 1;
    const long long T = p.T, B = p.B;
    auto sub_b0 = [&](int s) { return (long long)p.b_base + ((long long)group * S + s) * NB; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&h_full[2 * s], 1); mbar_init(&h_full[2 * s + 1], 1); mbar_init(&d_full[s], 1);
            mbar_init(&peer_full[2 * s], 1); mbar_init(&peer_full[2 * s + 1], 1);
        }
        fence_barrier_init();
    }
    if (PAIR) cluster_sync();                // both CTAs of a pair are resident before the paired TMEM allocation
    if (warp == 0) { if (PAIR) tmem_alloc2<512>(tmem_slot); else tmem_alloc<512>(tmem_slot); }
    // zero the buffers (padding slots u = 30, 31 of every image must be finite zeros forever)
    for (int i = threadIdx.x; i < S * C::PER_SUB / 16; i += C::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();     // every CTA's barriers are initialised before any remote copy can target them

    if (warp < S) {
        // ================= MMA issuer of sub-tile s = warp (one elected thread) =================
        const int s = warp;
        named_barrier(S + 1, 32 * S + 128);      // weights are in TMEM (loaded by epilogue group 0)
        tc_fence_after();
        if (sub_b0(s) < B && elect_one()) {
            if (PAIR && (rank & 1)) {
                // odd CTA of a pair: tell the leader when my half of the B operand has landed
                for (long long t = 0; t < T; ++t) {
                    const int par = (int)(t & 1);
                    mbar_wait_cluster(&h_full[2 * s + par], (uint32_t)((t >> 1) & 1));
                    mbar_arrive_remote(&peer_full[2 * s + par], rank ^ 1u);
                }
            } else {
                constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, NB);
                const uint32_t d_tmem = tmem_base + 256 + s * NB;
                const uint16_t pair_mask = (uint16_t)(3u << (rank & ~1u));
                for (long long t = 0; t < T; ++t) {
                    const int par = (int)(t & 1);
                    mbar_wait_cluster(&h_full[2 * s + par], (uint32_t)((t >> 1) & 1));
                    if (PAIR) mbar_wait_cluster(&peer_full[2 * s + par], (uint32_t)((t >> 1) & 1));
                    tc_fence_after();
                    HSSB_TRACE(TR_MMA_HFULL, t, s);
                    const uint32_t hb = smem_u32(hbuf(s, par));
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t blk = hb + (j >> 1) * (NBH * 128) + (j & 1) * (NBH * 32);
                        const uint64_t b_hi = make_smem_desc(blk, NBH * 16, 128, LAYOUT_NONE);
                        const uint64_t b_lo = make_smem_desc(blk + NBH * 64, NBH * 16, 128, LAYOUT_NONE);
                        const uint32_t a_hi = tmem_base + j * 8, a_lo = tmem_base + 128 + j * 8;
                        if (PAIR) {
                            mma_f16_ts2(d_tmem, a_hi, b_hi, idesc, j != 0);
                            mma_f16_ts2(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_f16_ts2(d_tmem, a_hi, b_lo, idesc, 1);
                        } else {
                            mma_f16_ts(d_tmem, a_hi, b_hi, idesc, j != 0);
                            mma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    }
                    if (PAIR) mma_commit2_mc(&d_full[s], pair_mask); else mma_commit(&d_full[s]);
                    HSSB_TRACE(TR_MMA_ISSUED, t, s);
                }
            }
        }
    } else {
        // ================= epilogue group s: warps S+4s .. S+4s+3 =================
        const int s = (warp - S) >> 2;
        const int q = warp & 3;                  // TMEM lane quadrant of this warp
        const int row = q * 32 + lane;           // TMEM lane = gate row 4*u + j of this CTA
        const int u = row >> 2, j = lane & 3;    // unit 0..31 (30, 31 padding), gate / column residue
        const bool unit_ok = u < RC_U;
        const long long b0 = sub_b0(s);
        const int hcol = dir * (TC_OP / 2) + (int)rank * 32 + u;    // column in the [.., 512] slot-layout outputs
        const bool leader = (warp == S + 4 * s);

        if (s == 0) {
            // one-time: W_hh slice -> TMEM.  This thread owns lane `row`; column c holds k' = 2c, 2c+1.
            const __half *wrow = p.whh + ((((size_t)dir * RC_CL + rank) * 2) * 128 + row) * RC_KP;
#pragma unroll 1
            for (int plane = 0; plane < 2; ++plane) {
                const uint4 *src = reinterpret_cast<const uint4 *>(wrow + (size_t)plane * 128 * RC_KP);
#pragma unroll 4
                for (int c8 = 0; c8 < 16; ++c8) {
                    const uint4 v0 = __ldg(src + 2 * c8), v1 = __ldg(src + 2 * c8 + 1);
                    const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    tmem_st_x8(tmem_base + ((uint32_t)(q * 32) << 16) + plane * 128 + c8 * 8, r);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            named_barrier(S + 1, 32 * S + 128);
        }

        if (b0 < B) {
            constexpr int NI = NB / 4;
            constexpr float LOG2E = 1.4426950408889634f;
            // sigmoid for i, f, o; tanh x = 2 sigmoid(2x) - 1 for g: act = ksc * rcp(1 + 2^(nsc * x)) + kof
            const float ksc = (j == 2) ? 2.0f : 1.0f, nsc = -ksc * LOG2E, kof = 1.0f - ksc;
            const int ncols = (int)((B - b0 < NB) ? (B - b0) : NB);
            const bool full = ncols == NB;
            const int ni_valid = (ncols - j + 3) / 4;                   // columns 4i + j < ncols  <=>  i < ni_valid
            // xproj of this thread's gate row (padding lanes re-read row 119, result unused):
            // element (t, b) at xp_base + (t*B + b)*960
            const int xrow = unit_ok ? row : RC_XW - 1;
            const float *xp_base = p.xproj + (size_t)dir * T * p.Bp * TC_G + (size_t)b0 * TC_G + rank * RC_XW + xrow;
            const long long xstep = (dir ? -1 : 1) * p.Bp * TC_G;       // one time step
            const float *xp_next = xp_base + (dir ? (size_t)(T - 1) * p.Bp * TC_G : 0);
            float c_state[NI], hv[NI], xnext[NB];
            auto load_x = [&]() {                                        // xproj of the next step -> registers
                if (full) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) xnext[b] = __ldcs(xp_next + b * TC_G);
                } else {
#pragma unroll
                    for (int b = 0; b < NB; ++b) xnext[b] = (b < ncols) ? __ldcs(xp_next + b * TC_G) : 0.0f;
                }
                xp_next += xstep;
            };
            load_x();
            // outputs of (unit u, column 4i + j): element offset of step tt = o_base + i*o_stride + tt*480
            const size_t o_stride = (size_t)4 * T * TC_OP;
            size_t o_next = ((size_t)(b0 + j) * T + (dir ? T - 1 : 0)) * TC_OP + hcol;
            const long long o_step = (dir ? -1 : 1) * TC_OP;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + s * NB;
            if (p.stagger_ns) __nanosleep((unsigned)(s * p.stagger_ns));   // de-phase the sub-tiles of a cluster
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const long long bg = b0 + 4 * i + j;
                const bool ok = unit_ok && i < ni_valid;
                hv[i] = ok ? __ldg(p.h0 + ((size_t)dir * B + bg) * TC_H + rank * RC_U + u) : 0.f;
                c_state[i] = ok ? __ldg(p.c0 + ((size_t)dir * B + bg) * TC_H + rank * RC_U + u) : 0.f;
            }
            for (long long t = -1; t < T; ++t) {
                if (t >= 0) {
                    mbar_wait(&d_full[s], (uint32_t)(t & 1));
                    tc_fence_after();
                    if (leader && lane == 0) HSSB_TRACE(TR_EPI_DFULL, t, s);
                    uint32_t v[NB];
#pragma unroll
                    for (int c16 = 0; c16 < NB / 16; ++c16) tmem_ld_x16(taddr + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
                    tmem_ld_wait();
                    tc_fence_before();
                    float act[NB];
#pragma unroll
                    for (int b = 0; b < NB; ++b) act[b] = (__uint_as_float(v[b]) + xnext[b]) * nsc;
                    if (t + 1 < T) load_x();               // lands during this step's math and the all-gather
#pragma unroll
                    for (int b = 0; b < NB; ++b) act[b] = fmaf(rcp_approx(1.0f + ex2_approx(act[b])), ksc, kof);
                    if (leader && lane == 0) HSSB_TRACE(TR_EPI_ACT, t, s);
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        float g4[4] = {act[4 * i], act[4 * i + 1], act[4 * i + 2], act[4 * i + 3]};
                        quad_transpose(g4, j);                         // -> i, f, g, o of (unit u, column 4i + j)
                        const float c = fmaf(g4[1], c_state[i], g4[0] * g4[2]);
                        c_state[i] = c;
                        const float th = fmaf(rcp_approx(1.0f + ex2_approx(c * (-2.0f * LOG2E))), 2.0f, -1.0f);
                        hv[i] = unit_ok ? g4[3] * th : 0.0f;
                    }
                    if (leader && lane == 0) HSSB_TRACE(TR_EPI_CELL, t, s);
                }
                // h_t of (unit u, columns 4i + j) -> fp16 hi/lo image [plane][k-chunk q][b][8 units]
                if (t + 1 < T) {
                    // (pair mode: columns [0, NB/2) form the half sent to the even CTAs, the rest goes to the odd ones)
                    __half *img_hi = reinterpret_cast<__half *>(image(s, (int)(t & 1))) + q * (NBH * 8) + j * 8 + (lane >> 2);
                    __half *img_lo = img_hi + NBH * 32;
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        __half hh, hl;
                        split_f16(hv[i], hh, hl);
                        const int off = (4 * i >= NBH) ? (NBH * 64 + (4 * i - NBH) * 8) : 4 * i * 8;   // [half][plane][chunk][col][8]
                        img_hi[off] = hh;
                        img_lo[off] = hl;
                    }
                    fence_proxy_async_smem();
                    named_barrier(1 + s, 128);
                    if (leader && lane == 0) HSSB_TRACE(TR_EPI_IMAGE, t, s);
                    if (leader) {
                        // all-gather: this CTA's image -> slot `rank` of every CTA's B buffer for step t+1;
                        // lane r pushes to CTA r
                        const int par = (int)((t + 1) & 1);
                        if (lane == 0) mbar_arrive_expect_tx(&h_full[2 * s + par], C::HBUF_BYTES);
                        __syncwarp();
                        if (lane < RC_CL)
                            bulk_copy_to_cta(hbuf(s, par) + rank * C::SLICE_BYTES, image(s, (int)(t & 1)) + (PAIR ? (lane & 1) * C::SLICE_BYTES : 0),
                                             C::SLICE_BYTES, &h_full[2 * s + par], lane);
                    }
                    if (leader && lane == 0) HSSB_TRACE(TR_EPI_COPIES, t, s);
                }
                // ---- off the critical path: this step's outputs to global memory ----
                if (t >= 0) {
                    {                                   // padding slots 30, 31 are written too (zeros)
                        size_t o = o_next;
                        if (p.out_f32) {
#pragma unroll
                            for (int i = 0; i < NI; ++i, o += o_stride)
                                if (i < ni_valid) p.out_f32[o] = fmaxf(hv[i], 0.f);
                        } else {
#pragma unroll
                            for (int i = 0; i < NI; ++i, o += o_stride)
                                if (i < ni_valid) {
                                    __half hh, hl;
                                    split_f16(fmaxf(hv[i], 0.f), hh, hl);
                                    p.out_hi[o] = hh;
                                    p.out_lo[o] = hl;
                                }
                        }
                    }
                    o_next += o_step;
                }
            }
            if (unit_ok) {
#pragma unroll
                for (int i = 0; i < NI; ++i)
                    if (i < ni_valid) {
                        const size_t o = ((size_t)dir * B + b0 + 4 * i + j) * TC_H + rank * RC_U + u;
                        p.hn[o] = hv[i];
                        p.cn[o] = c_state[i];
                    }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) { if (PAIR) tmem_dealloc2<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
}

// ------------------------------------------------------------------------------------------------
// K5p: recurrence for large batches -- CTA pairs (cta_group::2), 64 batch columns per MMA.
//
// What bounds the kernel above is the per-step chain (MMA -> TMEM load -> activations -> all-gather) and, at 96
// columns per cluster, the all-gather itself: every CTA pushes 1 KB per column to 7 peers through DSMEM
// (~17 B/cycle/SM measured, scripts/microbench/ub_cluster.cu), about twice the tensor time of the same columns.
// Here the two CTAs of a TPC issue ONE tcgen05.mma.cta_group::2 (M = 256 gate rows, N = 64 columns; 33 cycles,
// the same as a cta_group::1 MMA of that N) whose B operand is split between them: the even CTA holds columns
// [0, 32) of h_{t-1}, the odd CTA columns [32, 64), so every CTA receives -- and sends -- half as much.
//   * TMEM lanes in "fragment order" (lane = 32*(u/8) + 8*gate + u%8): two tcgen05.ld.16x256b.x4 hand thread
//     (ul = lane/4, cp = lane%4) the four gates of unit 8q+ul for the 8 columns 8k + 2cp + {0,1} -- no shuffles,
//     and the matching xproj values are 8 coalesced 16-byte loads (xproj keeps the 4*u + gate order of K4).
//   * 8 epilogue warps per sub-tile = 4 TMEM lane quadrants x 2 column halves; the warps of half `hf` produce
//     exactly the part of the image that goes to the CTAs of parity `hf` (4 bulk copies of 4 KB per half).
//   * every B buffer has one mbarrier per SOURCE PAIR, so the MMA issuer starts on the K range of a pair as soon
//     as that pair's slices landed (the group holding this pair's own slices first: its arrival also proves that
//     all 16 epilogue warps of the pair have read the previous accumulator).  The odd CTA relays its arrivals to
//     the even (issuing) CTA.
//   * activations with 8 instead of 10 MUFU ops per (unit, column): the reciprocals of i.g and o.tanh(c) are
//     shared, i*g = (1 - e_g) / ((1 + e_i)(1 + e_g)) with e_x = exp(-x) (exp(-2x) for g and c).
//   * outputs (relu(h) of the step) leave through a per-warp shared-memory tile and one TMA tensor store per
//     plane (box 8 units x 32 columns of the slot-layout [B][T][512] tensors; ragged batches are clipped by the
//     TMA unit) instead of 16 scattered 2-byte global stores per thread.
// Layout of one B buffer: [source rank 8][k-chunk 4][plane 2][column 32][8 units] fp16 (K-major core matrices:
// LBO = 1 KB between k-chunks, SBO = 128 B between 8-column groups).
// ------------------------------------------------------------------------------------------------
constexpr int RP_NB = 64;                          // columns of one sub-tile (pair MMA N)
constexpr int RP_NBH = 32;                         // columns held (and produced per epilogue warp) per CTA half
constexpr int RP_G = 4;                            // arrival groups per buffer (= source pairs)
constexpr int RP_PIECE = RP_NBH * 8 * 2 * 2;       // [plane][32 cols][8 units] fp16 = 1 KB: one epilogue warp's output
constexpr int RP_SLICE = 4 * RP_PIECE;             // one source rank: 4 k-chunks
constexpr int RP_HBUF = RC_CL * RP_SLICE;          // 32 KB

template <int S>
struct RpCfg {
    static constexpr int PER_SUB = 2 * RP_HBUF + 4 * RP_SLICE;       // 2 B buffers + images [parity][half]
    static constexpr int OUT_BYTES = S * 8 * 1024;                   // per epilogue warp: relu(h) tile for the TMA store
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = S * PER_SUB + OUT_BYTES + BAR_BYTES + 1024;
    static constexpr int THREADS = 32 * S + 256 * S;                 // S issuer / relay warps + S x 8 epilogue warps
    static_assert(S * RP_NB <= 256, "accumulators must fit in the TMEM columns left of the weights");
    static_assert((4 * S * RP_G + S) * 8 + 8 <= BAR_BYTES, "barrier area too small");
};

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

template <int S>
__global__ void __launch_bounds__(RpCfg<S>::THREADS, 1) tc_recurrent_pair_kernel(const __grid_constant__ RecurParams p)
{
    using C = RpCfg<S>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    auto hbuf = [&](int s, int par) { return smem + s * C::PER_SUB + par * RP_HBUF; };
    auto image = [&](int s, int par, int hf) { return smem + s * C::PER_SUB + 2 * RP_HBUF + (par * 2 + hf) * RP_SLICE; };
    unsigned char *out_tiles = smem + S * C::PER_SUB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_tiles + C::OUT_BYTES);
    uint64_t *own_full = bars;                       // [S][2][G]  slices of source pair g have landed in my buffer
    uint64_t *peer_full = bars + 2 * S * RP_G;       // [S][2][G]  (even CTA) ... and in the odd CTA's buffer
    uint64_t *d_full = bars + 4 * S * RP_G;          // [S]        accumulator complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x / RC_CL;
    const int dir = cid & 1;
    const int group = cid >> 1;
    const long long T = p.T, B = p.B;
    auto sub_b0 = [&](int s) { return (long long)p.b_base + ((long long)group * S + s) * RP_NB; };
    unsigned long long *const tr_buf = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
#define RP_TRACE(ev, step, sub)                                                                                       \
    do {                                                                                                              \
        if (tr_buf && (step) >= 0 && (step) < p.trace_steps) tr_buf[(((step) * 4 + (sub)) * TR_EVENTS) + (ev)] = clock64(); \
    } while (0)

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * S * RP_G + S; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
        prefetch_tmap(&p.out_map[0]);
        if (!p.out_f32) prefetch_tmap(&p.out_map[1]);
    }
    cluster_sync();                          // both CTAs of a pair are resident before the paired TMEM allocation
    if (warp == 0) tmem_alloc2<512>(tmem_slot);
    for (int i = threadIdx.x; i < (S * C::PER_SUB + C::OUT_BYTES) / 16; i += C::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();                          // every CTA's barriers are initialised before any remote copy can target them

    if (warp < S) {
        // ================= MMA issuer (even CTA) / arrival relay (odd CTA) of sub-tile s = warp =================
        const int s = warp;
        named_barrier(9, 32 * S + 128);          // weights are in TMEM
        tc_fence_after();
        if (sub_b0(s) < B && elect_one()) {
            const int g0 = (int)(rank >> 1);
            for (int i = 0; i < 2 * RP_G; ++i) mbar_arrive_expect_tx(&own_full[s * 2 * RP_G + i], 2 * RP_SLICE);
            if (rank & 1) {
                for (long long t = 0; t < T; ++t) {
                    const int par = (int)(t & 1);
                    const uint32_t ph = (uint32_t)((t >> 1) & 1);
#pragma unroll
                    for (int gi = 0; gi < RP_G; ++gi) {
                        const int bi = (s * 2 + par) * RP_G + ((g0 + gi) & (RP_G - 1));
                        mbar_wait_cluster(&own_full[bi], ph);
                        mbar_arrive_remote(&peer_full[bi], rank ^ 1u);
                        if (t + 2 < T) mbar_arrive_expect_tx(&own_full[bi], 2 * RP_SLICE);
                    }
                }
            } else {
                constexpr uint32_t idesc = make_idesc_f16(256, RP_NB);
                const uint32_t d_tmem = tmem_base + 256 + s * RP_NB;
                const uint16_t pair_mask = (uint16_t)(3u << rank);
                for (long long t = 0; t < T; ++t) {
                    const int par = (int)(t & 1);
                    const uint32_t ph = (uint32_t)((t >> 1) & 1);
                    const uint32_t hb = smem_u32(hbuf(s, par));
#pragma unroll
                    for (int gi = 0; gi < RP_G; ++gi) {
                        const int g = (g0 + gi) & (RP_G - 1);
                        const int bi = (s * 2 + par) * RP_G + g;
                        mbar_wait_cluster(&own_full[bi], ph);
                        mbar_wait_cluster(&peer_full[bi], ph);
                        if (t + 2 < T) mbar_arrive_expect_tx(&own_full[bi], 2 * RP_SLICE);
                        tc_fence_after();
                        if (gi == 0) RP_TRACE(TR_MMA_HFULL, t, s);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = 4 * g + jj;                      // K16 step: source rank j >> 1, k-chunks 2(j&1), 2(j&1)+1
                            const uint32_t blk = hb + (j >> 1) * RP_SLICE + (j & 1) * (2 * RP_PIECE);
                            const uint64_t b_hi = make_smem_desc(blk, RP_PIECE, 128, LAYOUT_NONE);
                            const uint64_t b_lo = make_smem_desc(blk + RP_PIECE / 2, RP_PIECE, 128, LAYOUT_NONE);
                            const uint32_t a_hi = tmem_base + j * 8, a_lo = tmem_base + 128 + j * 8;
                            mma_f16_ts2(d_tmem, a_hi, b_hi, idesc, (gi | jj) != 0);
                            mma_f16_ts2(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_f16_ts2(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    }
                    mma_commit2_mc(&d_full[s], pair_mask);
                    RP_TRACE(TR_MMA_ISSUED, t, s);
                }
            }
        }
    } else {
        // ================= epilogue warp: sub-tile s, column half hf, TMEM lane quadrant q =================
        const int k = (warp - S) >> 2;
        const int s = k >> 1, hf = k & 1;
        const int q = warp & 3;
        const int ul = lane >> 2, cp = lane & 3;     // unit within the k-chunk q; column pair
        const int u = 8 * q + ul;                    // unit 0..31 of this CTA (30, 31 padding)
        const bool unit_ok = u < RC_U;
        const long long b0 = sub_b0(s) + hf * RP_NBH;
        const bool tracer = (hf == 0 && q == 0 && lane == 0);
        unsigned char *out_tile = out_tiles + (warp - S) * 1024;

        if (k == 0) {
            // one-time: W_hh slice -> TMEM.  This thread owns lane 32q + lane; column c holds k' = 2c, 2c+1.
            const __half *wrow = p.whh + ((((size_t)dir * RC_CL + rank) * 2) * 128 + q * 32 + lane) * RC_KP;
#pragma unroll 1
            for (int plane = 0; plane < 2; ++plane) {
                const uint4 *src = reinterpret_cast<const uint4 *>(wrow + (size_t)plane * 128 * RC_KP);
#pragma unroll 4
                for (int c8 = 0; c8 < 16; ++c8) {
                    const uint4 v0 = __ldg(src + 2 * c8), v1 = __ldg(src + 2 * c8 + 1);
                    const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    tmem_st_x8(tmem_base + ((uint32_t)(q * 32) << 16) + plane * 128 + c8 * 8, r);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            named_barrier(9, 32 * S + 128);
        }

        if (sub_b0(s) < B) {
            constexpr int NI = RP_NBH / 4;              // 8 (unit, column) cells per thread: columns 8*(i/2) + 2*cp + (i&1)
            constexpr float LOG2E = 1.4426950408889634f;
            constexpr float EMAX = 60.0f;               // exponent clamp: (1 + 2^60)^2 is finite, sigmoid(-41) = 0 in fp32 anyway
            const long long left = B - b0;
            const int ncols = (int)(left < 0 ? 0 : (left < RP_NBH ? left : RP_NBH));    // valid columns of this half (may be 0)
            auto col_of = [&](int i) { return 8 * (i >> 1) + 2 * cp + (i & 1); };
            // xproj of (unit u, column c): 4 consecutive floats i, f, g, o at xp + c*960 (16-byte aligned)
            const int ux = unit_ok ? u : RC_U - 1;
            const float *xp_next = p.xproj + ((size_t)dir * T + (dir ? T - 1 : 0)) * p.Bp * TC_G + (size_t)b0 * TC_G + rank * RC_XW + 4 * ux;
            const long long xstep = (dir ? -1 : 1) * p.Bp * TC_G;
            float4 xnext[NI];
            float c_state[NI];
            // xproj of the next step -> registers.  Issued as the LAST thing of a step: every later long-scoreboard
            // wait of the warp (spill reloads, TMA issue, ...) would otherwise sit behind these HBM loads.
            auto load_x = [&]() {
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int c = col_of(i);
                    xnext[i] = (c < ncols) ? __ldcs(reinterpret_cast<const float4 *>(xp_next + (size_t)c * TC_G)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                xp_next += xstep;
            };
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + s * RP_NB + hf * RP_NBH;
            const uint32_t my_group = rank >> 1;          // my slices complete barrier `my_group` of every destination
            const int out_c0 = dir * (TC_OP / 2) + (int)rank * 32 + 8 * q;      // first column of this warp's 8 units
            const size_t state_o = ((size_t)dir * B + b0) * TC_H + rank * RC_U + u;      // + column * TC_H
            // h_t of (unit u, 8 columns) -> this warp's piece [plane][col][8 units] of the fp16 hi/lo image, then
            // (after the 4 warps of this half have written theirs) 4 bulk copies of the 4 KB half-image
            auto publish = [&](const float (&hv)[NI], int t) {
                const uint32_t img = smem_u32(image(s, (int)(t & 1), hf)) + q * RP_PIECE + ul * 2;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    __half hh, hl;
                    split_f16(hv[i], hh, hl);
                    sts_b16(img + col_of(i) * 16, hh);
                    sts_b16(img + RP_NBH * 16 + col_of(i) * 16, hl);
                }
                fence_proxy_async_smem();
                named_barrier(1 + k, 128);
                if (tracer) RP_TRACE(TR_EPI_IMAGE, t, s);
                if (q == 0 && elect_one()) {
                    const int par = (int)((t + 1) & 1);
                    uint64_t *bar = &own_full[(s * 2 + par) * RP_G + my_group];
#pragma unroll
                    for (int d = 0; d < 4; ++d)
                        bulk_copy_to_cta(hbuf(s, par) + rank * RP_SLICE, image(s, (int)(t & 1), hf), RP_SLICE, bar, (uint32_t)(2 * d + hf));
                }
                if (tracer) RP_TRACE(TR_EPI_COPIES, t, s);
            };
            {
                float h_init[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const bool ok = unit_ok && col_of(i) < ncols;
                    h_init[i] = ok ? __ldg(p.h0 + state_o + (size_t)col_of(i) * TC_H) : 0.f;
                    c_state[i] = ok ? __ldg(p.c0 + state_o + (size_t)col_of(i) * TC_H) : 0.f;
                }
                if (p.stagger_ns) __nanosleep((unsigned)(s * p.stagger_ns));   // de-phase the sub-tiles of a cluster
                publish(h_init, -1);
            }
            load_x();
            const int Ti = (int)T;
            int t_idx = dir ? Ti - 1 : 0;
            for (int t = 0; t < Ti; ++t) {
                mbar_wait(&d_full[s], (uint32_t)(t & 1));
                tc_fence_after();
                if (tracer) RP_TRACE(TR_EPI_DFULL, t, s);
                float hv[NI];
                {
                    // two passes of 16 columns keep the register peak (xproj prefetch + accumulators + exponentials) under 96
                    float ei[NI], ef[NI], eg[NI], eo[NI];
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        uint32_t a[8], b[8];    // a: gates i (lane ul), f (lane ul+8);  b: gates g, o;  [4k + 2*gate + c] = column 16*pass + 8k + 2cp + c
                        tmem_ld_16x256b_x2(taddr + 16 * pass, a);
                        tmem_ld_16x256b_x2(taddr + (16u << 16) + 16 * pass, b);
                        tmem_ld_wait();
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const int i = 4 * pass + ii, r = 4 * (ii >> 1) + (ii & 1);
                            ei[i] = ex2_approx(fminf((__uint_as_float(a[r]) + xnext[i].x) * -LOG2E, EMAX));
                            ef[i] = ex2_approx(fminf((__uint_as_float(a[r + 2]) + xnext[i].y) * -LOG2E, EMAX));
                            eg[i] = ex2_approx(fminf((__uint_as_float(b[r]) + xnext[i].z) * (-2.0f * LOG2E), EMAX));
                            eo[i] = ex2_approx(fminf((__uint_as_float(b[r + 2]) + xnext[i].w) * -LOG2E, EMAX));
                        }
                    }
                    tc_fence_before();
                    if (tracer) RP_TRACE(TR_EPI_ACT, t, s);
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const float ig = (1.0f - eg[i]) * rcp_approx((1.0f + ei[i]) * (1.0f + eg[i]));      // sigmoid(i) tanh(g)
                        const float c = fmaf(rcp_approx(1.0f + ef[i]), c_state[i], ig);
                        c_state[i] = c;
                        const float ec = ex2_approx(fminf(c * (-2.0f * LOG2E), EMAX));
                        const float h = (1.0f - ec) * rcp_approx((1.0f + eo[i]) * (1.0f + ec));             // sigmoid(o) tanh(c)
                        hv[i] = unit_ok ? h : 0.0f;
                    }
                    if (tracer) RP_TRACE(TR_EPI_CELL, t, s);
                }
                if (t + 1 < Ti) publish(hv, t);
                // ---- off the critical path: relu(h_t) -> global memory by TMA from this warp's tile ----
                if (elect_one()) tma_store_wait_read<0>();        // the previous step's store has read the tile
                __syncwarp();
                const uint32_t tile = smem_u32(out_tile);
                if (p.out_f32) {
#pragma unroll
                    for (int i = 0; i < NI; ++i) sts_b32(tile + col_of(i) * 32 + ul * 4, fmaxf(hv[i], 0.f));
                } else {
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        __half hh, hl;
                        split_f16(fmaxf(hv[i], 0.f), hh, hl);
                        sts_b16(tile + col_of(i) * 16 + ul * 2, hh);
                        sts_b16(tile + 512 + col_of(i) * 16 + ul * 2, hl);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (ncols > 0 && elect_one()) {
                    tma_store_3d(&p.out_map[0], out_tile, out_c0, t_idx, (int)b0);
                    if (!p.out_f32) tma_store_3d(&p.out_map[1], out_tile + 512, out_c0, t_idx, (int)b0);
                    tma_store_commit();
                }
                t_idx += dir ? -1 : 1;
                if (t + 1 < Ti) {
                    load_x();
                } else if (unit_ok) {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        if (col_of(i) < ncols) {
                            p.hn[state_o + (size_t)col_of(i) * TC_H] = hv[i];
                            p.cn[state_o + (size_t)col_of(i) * TC_H] = c_state[i];
                        }
                }
            }
            if (elect_one()) tma_store_wait<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2<512>(tmem_base);
#undef RP_TRACE
}

// ------------------------------------------------------------------------------------------------
// K5m: recurrence with the all-gather through L2 (bulk store + TMA multicast) -- the large-batch kernel.
//
// Measured on B200 (scripts/microbench/ub_cluster.cu): a CTA can push ~17 B/cycle into DSMEM, so the 8-way
// all-gather of the kernels above costs 5 750 cycles per step at 96 columns per cluster -- twice the tensor time.
// The same exchange through L2 -- every CTA bulk-stores its 4 KB image once and then issues ONE multicast bulk
// load that delivers it to all 8 CTAs of the cluster -- moves 50-60 B/cycle into every SM and takes ~1 100
// cycles end to end, with two bulk operations per sub-tile and step instead of eight.
//   * one CTA per 30 units as before (cta_group::1, M = 128, N = 32, W_hh hi/lo resident in TMEM), S = 1..3
//     independent sub-tiles of 32 batch columns interleaved per cluster;
//   * TMEM lanes in fragment order, xproj as 16-byte loads issued at the END of a step, outputs by TMA store:
//     see the pair kernel above (same epilogue);
//   * every B buffer has one mbarrier per pair of source ranks; the MMA issuer starts on a pair's K range as
//     soon as its two slices landed (own pair first: its arrival proves that the four epilogue warps have
//     read the previous accumulator);
//   * the image is single-buffered: the issuing thread waits for its bulk store (cp.async.bulk.wait_group)
//     before it issues the multicast load, and nobody rewrites the image before the next accumulator, which
//     depends on that load; the L2 scratch slot is double-buffered by step parity.
// ------------------------------------------------------------------------------------------------
constexpr int RX_KSTEPS = 3;                       // fused layer-1 input projection: K16 steps of the x operand (input_size <= 48)
constexpr int RX_CHUNKS = 2 * RX_KSTEPS;           // 8-feature k-chunks
constexpr int RX_PLANE = RX_CHUNKS * RP_NBH * 16;  // [chunk][32 cols][8 features] fp16 = 3 KB
constexpr int RX_TMEM = 256 + 96;                  // TMEM column of the W_ih slice (hi plane; lo plane 8*RX_KSTEPS columns further)

template <int S, int EW, bool FUSE_X = false>
struct RmCfg {
    static constexpr int NW = RP_NBH / EW;                           // batch columns per epilogue warp
    static constexpr int PER_SUB = 2 * RP_HBUF + RP_SLICE;           // 2 B buffers + one image
    static constexpr int TILE_BYTES = 1024 / EW;                     // per epilogue warp: relu(h) tile for the TMA store
    // fused: the relu(h) tile of a warp reuses its piece of the image (free again once the bulk store of the publish has
    // completed), which makes room for the x operand buffers
    static constexpr int OUT_BYTES = FUSE_X ? 0 : S * 4 * EW * TILE_BYTES;
    static constexpr int X_BYTES = FUSE_X ? S * 2 * RX_PLANE : 0;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = S * PER_SUB + OUT_BYTES + X_BYTES + BAR_BYTES + 1024;
    static constexpr int THREADS = 32 * S + 128 * S * EW;            // S issuer warps + S x 4 x EW epilogue warps
    static_assert(EW == 1 || EW == 2, "one or two epilogue warps per TMEM lane quadrant and sub-tile");
    static_assert(S * RP_NBH <= 96, "accumulators sit in TMEM columns [256, 352)");
    static_assert((2 * S * RP_G + 4 * S) * 8 + 8 <= BAR_BYTES, "barrier area too small");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

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

template <int S, bool WARP_PUBLISH, int EW, bool FUSE_X>
__global__ void __launch_bounds__(RmCfg<S, EW, FUSE_X>::THREADS, 1) tc_recurrent_mc_kernel(const __grid_constant__ RecurParams p)
{
    using C = RmCfg<S, EW, FUSE_X>;
    static_assert(EW == 1 || WARP_PUBLISH, "two warps per quadrant publish per warp");
    static_assert(!FUSE_X || WARP_PUBLISH, "the fused kernel reuses each warp's image piece as its output tile");
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    auto hbuf = [&](int s, int par) { return smem + s * C::PER_SUB + par * RP_HBUF; };
    auto image = [&](int s) { return smem + s * C::PER_SUB + 2 * RP_HBUF; };
    unsigned char *out_tiles = smem + S * C::PER_SUB;
    unsigned char *xbufs = out_tiles + C::OUT_BYTES;             // fused: [S][plane][chunk][32 cols][8 features] fp16
    uint64_t *bars = reinterpret_cast<uint64_t *>(xbufs + C::X_BYTES);
    uint64_t *h_full = bars;                         // [S][2][G]  slices of source pair g have landed in my buffer
    uint64_t *d_full = bars + 2 * S * RP_G;          // [S]        accumulator complete
    uint64_t *d_empty = d_full + S;                  // [S]        (fused) every epilogue warp has read the accumulator
    uint64_t *x_full = d_full + 2 * S;               // [S]        (fused) x_t operand landed
    uint64_t *x_empty = d_full + 3 * S;              // [S]        (fused) the MMAs reading the x operand are complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 4 * S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x / RC_CL;
    const int dir = cid & 1;
    const int group = cid >> 1;
    const long long T = p.T, B = p.B;
    auto sub_b0 = [&](int s) { return (long long)p.b_base + ((long long)group * S + s) * RP_NBH; };
    unsigned long long *const tr_buf = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
#define RM_TRACE(ev, step, sub)                                                                                       \
    do {                                                                                                              \
        if (tr_buf && (step) >= 0 && (step) < p.trace_steps) tr_buf[(((step) * 4 + (sub)) * TR_EVENTS) + (ev)] = clock64(); \
    } while (0)

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * S * RP_G + 4 * S; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < S; ++i) mbar_init(&d_empty[i], 4 * EW);
        fence_barrier_init();
        const CUtensorMap *om = (EW == 1) ? p.out_map : p.out_map16;
        prefetch_tmap(&om[0]);
        if (!p.out_f32) prefetch_tmap(&om[1]);
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < (S * C::PER_SUB + C::OUT_BYTES + C::X_BYTES) / 16; i += C::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();                          // every CTA's barriers are initialised before any multicast can target them

    if (warp < S) {
        // ================= MMA issuer of sub-tile s = warp (one elected thread) =================
        const int s = warp;
        named_barrier(9, 32 * S + 128);          // weights are in TMEM
        tc_fence_after();
        if (sub_b0(s) < B && elect_one()) {
            const int g0 = (int)(rank >> 1);
            for (int i = 0; i < 2 * RP_G; ++i) mbar_arrive_expect_tx(&h_full[s * 2 * RP_G + i], 2 * RP_SLICE);
            constexpr uint32_t idesc = make_idesc_f16(128, RP_NBH);
            const uint32_t d_tmem = tmem_base + 256 + s * RP_NBH;
            const int Ti = (int)T;
            unsigned char *xb = xbufs + s * 2 * RX_PLANE;
            auto load_x_operand = [&](int t) {           // x_t of this sub-tile's 32 columns: both fp16 planes, [chunk][col][8], 3 KB each
                const int t_idx = dir ? Ti - 1 - t : t;
                const size_t off = ((size_t)t_idx * p.x_tiles + (size_t)(sub_b0(s) / RP_NBH)) * (8 * RP_NBH * 8);     // halves
                mbar_arrive_expect_tx(&x_full[s], 2 * RX_PLANE);
                bulk_load_global(xb, p.x_hi + off, RX_PLANE, &x_full[s]);
                bulk_load_global(xb + RX_PLANE, p.x_lo + off, RX_PLANE, &x_full[s]);
            };
            if (FUSE_X) load_x_operand(0);
            for (int t = 0; t < Ti; ++t) {
                const int par = t & 1;
                const uint32_t ph = (uint32_t)((t >> 1) & 1);
                const uint32_t hb = smem_u32(hbuf(s, par));
                if (FUSE_X) {
                    // W_ih . x_t first: it does not depend on h_{t-1}, only on the epilogue having read the previous accumulator
                    if (t > 0) mbar_wait(&d_empty[s], (uint32_t)((t - 1) & 1));
                    mbar_wait(&x_full[s], (uint32_t)(t & 1));
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < RX_KSTEPS; ++j) {
                        const uint32_t blk = smem_u32(xb) + j * (2 * RP_NBH * 16);
                        const uint64_t x_hi = make_smem_desc(blk, RP_NBH * 16, 128, LAYOUT_NONE);
                        const uint64_t x_lo = make_smem_desc(blk + RX_PLANE, RP_NBH * 16, 128, LAYOUT_NONE);
                        const uint32_t w_hi = tmem_base + RX_TMEM + j * 8, w_lo = w_hi + 8 * RX_KSTEPS;
                        mma_f16_ts(d_tmem, w_hi, x_hi, idesc, j != 0);
                        mma_f16_ts(d_tmem, w_lo, x_hi, idesc, 1);
                        mma_f16_ts(d_tmem, w_hi, x_lo, idesc, 1);
                    }
                    mma_commit(&x_empty[s]);
                }
#pragma unroll
                for (int gi = 0; gi < RP_G; ++gi) {
                    const int g = (g0 + gi) & (RP_G - 1);
                    uint64_t *bar = &h_full[(s * 2 + par) * RP_G + g];
                    mbar_wait_cluster(bar, ph);
                    if (t + 2 < Ti) mbar_arrive_expect_tx(bar, 2 * RP_SLICE);
                    tc_fence_after();
                    if (gi == 0) RM_TRACE(TR_MMA_HFULL, t, s);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * g + jj;                      // K16 step: source rank j >> 1, k-chunks 2(j&1), 2(j&1)+1
                        const uint32_t blk = hb + (j >> 1) * RP_SLICE + (j & 1) * (2 * RP_PIECE);
                        const uint64_t b_hi = make_smem_desc(blk, RP_PIECE, 128, LAYOUT_NONE);
                        const uint64_t b_lo = make_smem_desc(blk + RP_PIECE / 2, RP_PIECE, 128, LAYOUT_NONE);
                        const uint32_t a_hi = tmem_base + j * 8, a_lo = tmem_base + 128 + j * 8;
                        mma_f16_ts(d_tmem, a_hi, b_hi, idesc, FUSE_X || (gi | jj) != 0);
                        mma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1);
                        mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                }
                mma_commit(&d_full[s]);
                RM_TRACE(TR_MMA_ISSUED, t, s);
                if (FUSE_X && t + 1 < Ti) {
                    mbar_wait(&x_empty[s], (uint32_t)(t & 1));      // long since complete: the h part was issued behind it
                    load_x_operand(t + 1);
                }
            }
        }
    } else {
        // ================= epilogue warp: sub-tile s, TMEM lane quadrant q =================
        const int s = (warp - S) / (4 * EW);
        const int half = ((warp - S) >> 2) % EW;     // which NW-column part of the sub-tile this warp drains
        const int cbase = half * C::NW;
        const int q = warp & 3;
        const int ul = lane >> 2, cp = lane & 3;     // unit within the k-chunk q; column pair
        const int u = 8 * q + ul;                    // unit 0..31 of this CTA (30, 31 padding)
        const bool unit_ok = u < RC_U;
        const long long b0 = sub_b0(s);
        const bool tracer = (q == 0 && lane == 0 && half == 0);
        // fused: the relu(h) tile (fp16 hi / lo) lives in this warp's own two runs of its image piece, free once its publish completed
        unsigned char *out_tile = FUSE_X ? image(s) + q * RP_PIECE + cbase * 16 : out_tiles + (warp - S) * C::TILE_BYTES;
        constexpr int LO_OFF = FUSE_X ? RP_PIECE / 2 : C::NW * 16;      // lo plane of the tile

        if (s == 0 && half == 0) {
            // one-time: W_hh slice -> TMEM.  This thread owns lane 32q + lane; column c holds k' = 2c, 2c+1.
            const __half *wrow = p.whh + ((((size_t)dir * RC_CL + rank) * 2) * 128 + q * 32 + lane) * RC_KP;
#pragma unroll 1
            for (int plane = 0; plane < 2; ++plane) {
                const uint4 *src = reinterpret_cast<const uint4 *>(wrow + (size_t)plane * 128 * RC_KP);
#pragma unroll 4
                for (int c8 = 0; c8 < 16; ++c8) {
                    const uint4 v0 = __ldg(src + 2 * c8), v1 = __ldg(src + 2 * c8 + 1);
                    const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    tmem_st_x8(tmem_base + ((uint32_t)(q * 32) << 16) + plane * 128 + c8 * 8, r);
                }
            }
            if (FUSE_X) {
                // W_ih slice (rows in the same fragment order, K = 16*RX_KSTEPS features): hi plane, then lo plane
                const __half *xrow = p.wih0 + ((((size_t)dir * RC_CL + rank) * 2) * 128 + q * 32 + lane) * 64;
#pragma unroll 1
                for (int plane = 0; plane < 2; ++plane) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(xrow + (size_t)plane * 128 * 64);
#pragma unroll
                    for (int c8 = 0; c8 < RX_KSTEPS; ++c8) {
                        const uint4 v0 = __ldg(src + 2 * c8), v1 = __ldg(src + 2 * c8 + 1);
                        const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        tmem_st_x8(tmem_base + ((uint32_t)(q * 32) << 16) + RX_TMEM + plane * 8 * RX_KSTEPS + c8 * 8, r);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            named_barrier(9, 32 * S + 128);
        }

        if (b0 < B) {
            constexpr int NW = C::NW;
            constexpr int NI = NW / 4;                  // (unit, column) cells per thread: columns cbase + 8*(i/2) + 2*cp + (i&1)
            constexpr float LOG2E = 1.4426950408889634f;
            constexpr float EMAX = 60.0f;               // exponent clamp: (1 + 2^60)^2 is finite, sigmoid(-41) = 0 in fp32 anyway
            const long long left = B - b0;
            const int ncols = (int)(left < RP_NBH ? left : RP_NBH);
            auto col_of = [&](int i) { return cbase + 8 * (i >> 1) + 2 * cp + (i & 1); };
            const int ux = unit_ok ? u : RC_U - 1;
            const float *xp_next = p.xproj + ((size_t)dir * T + (dir ? T - 1 : 0)) * p.Bp * TC_G + (size_t)b0 * TC_G + rank * RC_XW + 4 * ux;
            const long long xstep = (dir ? -1 : 1) * p.Bp * TC_G;
            float4 xnext[NI];
            float c_state[NI];
            // xproj of the next step -> registers.  Issued as the LAST thing of a step: every later long-scoreboard
            // wait of the warp (TMA issue, spill reloads, ...) would otherwise sit behind these HBM loads.
            auto load_x = [&]() {
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int c = col_of(i);
                    if (FUSE_X) continue;               // fused: xnext holds the (constant) biases of this unit's four gates
                    xnext[i] = (c < ncols && !(p.debug & 1)) ? __ldcs(reinterpret_cast<const float4 *>(xp_next + (size_t)c * TC_G)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                xp_next += xstep;
            };
            if (FUSE_X) {
                const float *bz = p.bias0 + ((size_t)dir * RC_CL + rank) * 128 + q * 32 + ul;      // rows 32q + 8*gate + ul
                const float4 b4 = make_float4(__ldg(bz), __ldg(bz + 8), __ldg(bz + 16), __ldg(bz + 24));
#pragma unroll
                for (int i = 0; i < NI; ++i) xnext[i] = b4;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + s * RP_NBH + cbase;
            const int out_c0 = dir * (TC_OP / 2) + (int)rank * 32 + 8 * q;      // first column of this warp's 8 units
            const size_t state_o = ((size_t)dir * B + b0) * TC_H + rank * RC_U + u;      // + column * TC_H
            unsigned char *gslot = p.gather + ((((size_t)cid * RC_CL + rank) * S + s) * 2) * RP_SLICE;     // [parity][4 KB]
            // h_t -> this warp's piece [plane][col][8 units] of the fp16 hi/lo image; then one thread stores the 4 KB image
            // to its L2 slot and multicasts it into slot `rank` of every CTA's B buffer for step t + 1
            auto publish = [&](const float (&hv)[NI], int t) {
                if (FUSE_X) {                       // the previous step's output store has read the tile that shares this piece
                    if (elect_one()) tma_store_wait_read<0>();
                    __syncwarp();
                }
                const uint32_t img = smem_u32(image(s)) + q * RP_PIECE + ul * 2;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    __half hh, hl;
                    split_f16(hv[i], hh, hl);
                    sts_b16(img + col_of(i) * 16, hh);
                    sts_b16(img + RP_NBH * 16 + col_of(i) * 16, hl);
                }
                fence_proxy_async_smem();
                const int par = (t + 1) & 1;
                uint64_t *bar = &h_full[(s * 2 + par) * RP_G + (rank >> 1)];
                if (WARP_PUBLISH) {
                    // every warp publishes its own piece: no block barrier, the exchange starts with the first warp done
                    __syncwarp();
                    if (tracer) RM_TRACE(TR_EPI_IMAGE, t, s);
                    if (elect_one()) {
                        if (EW == 1) {
                            unsigned char *g = gslot + par * RP_SLICE + q * RP_PIECE;
                            bulk_store_global(g, image(s) + q * RP_PIECE, RP_PIECE);
                            tma_store_commit();
                            tma_store_wait<0>();
                            bulk_load_multicast(hbuf(s, par) + rank * RP_SLICE + q * RP_PIECE, g, RP_PIECE, bar, (uint16_t)0xFF);
                        } else {
                            // my NW columns are one run of NW*16 bytes in each plane of the piece
                            const int o0 = q * RP_PIECE + cbase * 16, o1 = o0 + RP_PIECE / 2;
                            unsigned char *g = gslot + par * RP_SLICE;
                            bulk_store_global(g + o0, image(s) + o0, NW * 16);
                            bulk_store_global(g + o1, image(s) + o1, NW * 16);
                            tma_store_commit();
                            tma_store_wait<0>();
                            bulk_load_multicast(hbuf(s, par) + rank * RP_SLICE + o0, g + o0, NW * 16, bar, (uint16_t)0xFF);
                            bulk_load_multicast(hbuf(s, par) + rank * RP_SLICE + o1, g + o1, NW * 16, bar, (uint16_t)0xFF);
                        }
                    }
                } else {
                    named_barrier(1 + s, 128);
                    if (tracer) RM_TRACE(TR_EPI_IMAGE, t, s);
                    if (q == 0 && elect_one()) {
                        unsigned char *g = gslot + par * RP_SLICE;
                        bulk_store_global(g, image(s), RP_SLICE);
                        tma_store_commit();
                        tma_store_wait<0>();
                        bulk_load_multicast(hbuf(s, par) + rank * RP_SLICE, g, RP_SLICE, bar, (uint16_t)0xFF);
                    }
                }
                if (tracer) RM_TRACE(TR_EPI_COPIES, t, s);
            };
            {
                float h_init[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const bool ok = unit_ok && col_of(i) < ncols;
                    h_init[i] = ok ? __ldg(p.h0 + state_o + (size_t)col_of(i) * TC_H) : 0.f;
                    c_state[i] = ok ? __ldg(p.c0 + state_o + (size_t)col_of(i) * TC_H) : 0.f;
                }
                if (p.stagger_ns) __nanosleep((unsigned)(s * p.stagger_ns));   // de-phase the sub-tiles of a cluster
                publish(h_init, -1);
            }
            load_x();
            const int Ti = (int)T;
            int t_idx = dir ? Ti - 1 : 0;
            for (int t = 0; t < Ti; ++t) {
                mbar_wait(&d_full[s], (uint32_t)(t & 1));
                tc_fence_after();
                if (tracer) RM_TRACE(TR_EPI_DFULL, t, s);
                float hv[NI];
                {
                    uint32_t a[2 * NI], b[2 * NI];      // a: gates i (lane ul), f (lane ul+8);  b: gates g, o;  [4k + 2*gate + c] = column cbase + 8k + 2cp + c
                    if (EW == 1) {
                        tmem_ld_16x256b_x4(taddr, *reinterpret_cast<uint32_t(*)[16]>(&a[0]));
                        tmem_ld_16x256b_x4(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[16]>(&b[0]));
                    } else {
                        tmem_ld_16x256b_x2(taddr, *reinterpret_cast<uint32_t(*)[8]>(&a[0]));
                        tmem_ld_16x256b_x2(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[8]>(&b[0]));
                    }
                    tmem_ld_wait();
                    tc_fence_before();
                    if (FUSE_X) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_empty[s]);       // the issuer may start W_ih . x_{t+1} into this accumulator
                    }
                    float ei[NI], ef[NI], eg[NI], eo[NI];
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const int r = 4 * (i >> 1) + (i & 1);
                        // e_i, e_f, e_o may overflow to +inf (1/inf = 0 is the right limit); e_g and e_c are clamped because
                        // (1 - e) * 0 must not become inf * 0
                        ei[i] = ex2_approx((__uint_as_float(a[r]) + xnext[i].x) * -LOG2E);
                        ef[i] = ex2_approx((__uint_as_float(a[r + 2]) + xnext[i].y) * -LOG2E);
                        eg[i] = ex2_approx(fminf((__uint_as_float(b[r]) + xnext[i].z) * (-2.0f * LOG2E), EMAX));
                        eo[i] = ex2_approx((__uint_as_float(b[r + 2]) + xnext[i].w) * -LOG2E);
                    }
                    if (tracer) RM_TRACE(TR_EPI_ACT, t, s);
                    if (p.debug & 4) {                                  // (timing experiment: no cell math)
#pragma unroll
                        for (int i = 0; i < NI; ++i) hv[i] = unit_ok ? 0.25f * (ei[i] + ef[i]) * 1e-3f + 1e-3f * (eg[i] + eo[i]) : 0.0f;
                    } else
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const float ig = (1.0f - eg[i]) * rcp_approx((1.0f + ei[i]) * (1.0f + eg[i]));      // sigmoid(i) tanh(g)
                        const float c = fmaf(rcp_approx(1.0f + ef[i]), c_state[i], ig);
                        c_state[i] = c;
                        const float ec = ex2_approx(fminf(c * (-2.0f * LOG2E), EMAX));
                        const float h = (1.0f - ec) * rcp_approx((1.0f + eo[i]) * (1.0f + ec));             // sigmoid(o) tanh(c)
                        hv[i] = unit_ok ? h : 0.0f;
                    }
                    if (tracer) RM_TRACE(TR_EPI_CELL, t, s);
                }
                if (t + 1 < Ti) publish(hv, t);
                // ---- off the critical path: relu(h_t) -> global memory by TMA from this warp's tile ----
                if (elect_one()) tma_store_wait_read<0>();        // the previous step's store has read the tile
                __syncwarp();
                const uint32_t tile = smem_u32(out_tile);
                if (p.out_f32) {
#pragma unroll
                    for (int i = 0; i < NI; ++i) sts_b32(tile + (col_of(i) - cbase) * 32 + ul * 4, fmaxf(hv[i], 0.f));
                } else {
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        __half hh, hl;
                        split_f16(fmaxf(hv[i], 0.f), hh, hl);
                        sts_b16(tile + (col_of(i) - cbase) * 16 + ul * 2, hh);
                        sts_b16(tile + LO_OFF + (col_of(i) - cbase) * 16 + ul * 2, hl);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (!(p.debug & 2) && elect_one()) {
                    const CUtensorMap *om = (EW == 1) ? p.out_map : p.out_map16;       // box of 32 / 16 batch columns
                    tma_store_3d(&om[0], out_tile, out_c0, t_idx, (int)b0 + cbase);
                    if (!p.out_f32) tma_store_3d(&om[1], out_tile + LO_OFF, out_c0, t_idx, (int)b0 + cbase);
                    tma_store_commit();
                }
                t_idx += dir ? -1 : 1;
                if (t + 1 < Ti) {
                    load_x();
                } else if (unit_ok) {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        if (col_of(i) < ncols) {
                            p.hn[state_o + (size_t)col_of(i) * TC_H] = hv[i];
                            p.cn[state_o + (size_t)col_of(i) * TC_H] = c_state[i];
                        }
                }
            }
            if (elect_one()) tma_store_wait<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
#undef RM_TRACE
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
__global__ void split_planes_tiled_kernel(const float *__restrict__ x, long long B, long long T, int F, __half *__restrict__ hi,
                                          __half *__restrict__ lo)
{
    const long long tiles = (B + RP_NBH - 1) / RP_NBH;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // ((t * tiles + tile) * 8 + chunk) * 32 + col
    if (i >= T * tiles * 8 * RP_NBH) return;
    const int col = (int)(i % RP_NBH), c = (int)((i / RP_NBH) % 8);
    const long long tile = (i / (8 * RP_NBH)) % tiles, t = i / (8 * RP_NBH * tiles);
    const long long b = tile * RP_NBH + col;
    __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = 8 * c + e;
        h[e] = l[e] = __float2half_rn(0.f);
        if (k < F && b < B) split_f16(__ldg(x + (b * T + t) * F + k), h[e], l[e]);
    }
    *reinterpret_cast<uint4 *>(hi + i * 8) = *reinterpret_cast<const uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo + i * 8) = *reinterpret_cast<const uint4 *>(l);
}

// linear.weight[4][480] -> slot layout [4][512]
__global__ void pack_linw_kernel(const float *__restrict__ w, float *__restrict__ dst)
{
    const int c = blockIdx.x, k = threadIdx.x, slot = k & 31;
    dst[c * TC_OP + k] = slot < 30 ? w[c * 2 * TC_H + (k >> 8) * TC_H + ((k >> 5) & 7) * 30 + slot] : 0.f;
}

static unsigned long long *g_trace_buf = nullptr;
static int g_trace_steps = 0;

template <int NB, int S, bool PAIR>
static cudaLaunchConfig_t recurrent_config(int clusters, cudaStream_t st, cudaLaunchAttribute *attr)
{
    using C = RcCfg<NB, S, PAIR>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * RC_CL));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

// How many 8-CTA clusters of this geometry are co-resident on the current device (cached per geometry).
template <int NB, int S, bool PAIR>
static int max_resident_clusters(int *out)
{
    using C = RcCfg<NB, S, PAIR>;
    static int cached = 0;
    if (!cached) {
        cudaError_t e = cudaFuncSetAttribute(tc_recurrent_kernel<NB, S, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_recurrent_kernel)");
        cudaLaunchAttribute attr[1];
        cudaLaunchConfig_t cfg = recurrent_config<NB, S, PAIR>(16, nullptr, attr);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_recurrent_kernel<NB, S, PAIR>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_recurrent_kernel)");
        if (n < 2) return fail(HSSB_E_DEVICE, "device fits only %d recurrence clusters", n);
        cached = n;
    }
    *out = cached;
    return 0;
}

template <int NB, int S, bool PAIR>
static int launch_recurrent(const RecurParams &prm_in, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st)
{
    RecurParams prm = prm_in;
    prm.trace = g_trace_buf;
    prm.trace_steps = g_trace_steps;
    int max_clusters = 0;
    if (int rc = max_resident_clusters<NB, S, PAIR>(&max_clusters)) return rc;
    // one cluster per (direction, group): never launch more groups than are co-resident, a second wave
    // of clusters would double the latency of the whole launch
    const int per = NB * S;
    const int groups = (int)std::min<int64_t>(max_clusters / 2, (rem + per - 1) / per);
    *cols_done = groups * per;
    prm.xproj = xproj;
    prm.groups = groups;
    prm.stagger_ns = 800;
    if (const char *e = getenv("HSSB_RC_STAGGER")) prm.stagger_ns = atoi(e);
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = recurrent_config<NB, S, PAIR>(2 * groups, st, attr);
    ProfScope prof("tc_recurrent", st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_recurrent_kernel<NB, S, PAIR>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_recurrent_kernel)");
    return 0;
}

template <int S>
static int launch_recurrent_pair(const RecurParams &prm_in, const __half *whh_frag, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st)
{
    using C = RpCfg<S>;
    RecurParams prm = prm_in;
    prm.whh = whh_frag;
    prm.trace = g_trace_buf;
    prm.trace_steps = g_trace_steps;
    static int max_clusters = 0;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(tc_recurrent_pair_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_recurrent_pair_kernel)");
        cfg.gridDim = dim3(16 * RC_CL);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_recurrent_pair_kernel<S>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_recurrent_pair_kernel)");
        if (n < 2) return fail(HSSB_E_DEVICE, "device fits only %d recurrence clusters", n);
        max_clusters = n;
    }
    const int per = RP_NB * S;
    const int groups = (int)std::min<int64_t>(max_clusters / 2, (rem + per - 1) / per);
    *cols_done = groups * per;
    prm.xproj = xproj;
    prm.groups = groups;
    prm.stagger_ns = 1500;
    if (const char *e = getenv("HSSB_RC_STAGGER")) prm.stagger_ns = atoi(e);
    cfg.gridDim = dim3((unsigned)(2 * groups * RC_CL));
    ProfScope prof("tc_recurrent", st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_recurrent_pair_kernel<S>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_recurrent_pair_kernel)");
    return 0;
}

template <int S, bool WARP_PUBLISH, int EW, bool FUSE_X = false>
static int launch_recurrent_mc(const RecurParams &prm_in, const __half *whh_frag, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st)
{
    using C = RmCfg<S, EW, FUSE_X>;
    RecurParams prm = prm_in;
    prm.whh = whh_frag;
    prm.trace = g_trace_buf;
    prm.trace_steps = g_trace_steps;
    if (const char *e = getenv("HSSB_TRACE_LAYER")) if (atoi(e) != prm.layer) prm.trace = nullptr;
    static int max_clusters = 0;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_recurrent_mc_kernel)");
        cfg.gridDim = dim3(16 * RC_CL);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_recurrent_mc_kernel)");
        if (n < 2) return fail(HSSB_E_DEVICE, "device fits only %d recurrence clusters", n);
        max_clusters = std::min(n, 16);
    }
    const int per = RP_NBH * S;
    const int groups = (int)std::min<int64_t>(max_clusters / 2, (rem + per - 1) / per);
    *cols_done = groups * per;
    prm.xproj = xproj;
    prm.groups = groups;
    prm.stagger_ns = 600;
    if (const char *e = getenv("HSSB_RC_STAGGER")) prm.stagger_ns = atoi(e);
    if (const char *e = getenv("HSSB_RC_DEBUG")) prm.debug = atoi(e);
    cfg.gridDim = dim3((unsigned)(2 * groups * RC_CL));
    ProfScope prof("tc_recurrent", st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_recurrent_mc_kernel)");
    return 0;
}

// One layer's recurrence for batch columns [0, B): picks the sub-tile geometry from B.
static int tc_recurrent(const hssb_model *m, int layer, float *xproj, const float *h0, const float *c0, float *hn, float *cn,
                        __half *out_hi, __half *out_lo, float *out_f32, unsigned char *gather, int64_t B, int64_t T, cudaStream_t st,
                        const __half *x_hi = nullptr, const __half *x_lo = nullptr)
{
    // x_hi / x_lo != nullptr: layer 1 with the input projection fused into the recurrence (chunk-major x planes, no xproj)
    const bool fused = x_hi != nullptr;
    RecurParams prm = {};
    prm.gather = gather;
    prm.layer = layer;
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
    prm.Bp = xproj_pitch(B);
    {
        const bool f32 = out_f32 != nullptr;
        const uint64_t es = f32 ? 4 : 2;
        const uint64_t dims[3] = {(uint64_t)TC_OP, (uint64_t)T, (uint64_t)B};
        const uint64_t strides[2] = {(uint64_t)TC_OP * es, (uint64_t)T * TC_OP * es};
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
    if (int rc = max_resident_clusters<32, 3, false>(&max_clusters)) return rc;
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
        if (fused) {
            if (force_nb) return fail(HSSB_E_MODE, "HSSB_RC_GEOM cannot be combined with the fused projection");
            int rcf, donef = 0;
            if (s == 1) rcf = launch_recurrent_mc<1, true, 2, true>(prm, m->tc_whh_frag[layer], rem, &donef, xproj, st);
            else if (s == 2) rcf = launch_recurrent_mc<2, true, 2, true>(prm, m->tc_whh_frag[layer], rem, &donef, xproj, st);
            else rcf = launch_recurrent_mc<3, true, 1, true>(prm, m->tc_whh_frag[layer], rem, &donef, xproj, st);
            if (rcf) return rcf;
            base += donef;
            continue;
        }
        // (the CTA-pair variants -- cta_group::2, half the all-gather volume -- are validated but measured slower on
        //  B200: at N = 32 the paired MMA is issue-overhead bound, ~30 cycles each against ~18 for cta_group::1)
        int rc, done = 0;
        const int key = nb * 100 + s * 10 + pair;
        switch (key) {
        case 1610: rc = launch_recurrent<16, 1, false>(prm, rem, &done, xproj, st); break;
        case 1620: rc = launch_recurrent<16, 2, false>(prm, rem, &done, xproj, st); break;
        case 1630: rc = launch_recurrent<16, 3, false>(prm, rem, &done, xproj, st); break;
        case 3220: rc = launch_recurrent<32, 2, false>(prm, rem, &done, xproj, st); break;
        case 3230: rc = launch_recurrent<32, 3, false>(prm, rem, &done, xproj, st); break;
        case 3211: rc = launch_recurrent<32, 1, true>(prm, rem, &done, xproj, st); break;
        case 3221: rc = launch_recurrent<32, 2, true>(prm, rem, &done, xproj, st); break;
        case 3231: rc = launch_recurrent<32, 3, true>(prm, rem, &done, xproj, st); break;
        case 3241: rc = launch_recurrent<32, 4, true>(prm, rem, &done, xproj, st); break;
        case 3212: rc = launch_recurrent_mc<1, false, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3222: rc = launch_recurrent_mc<2, false, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3232: rc = launch_recurrent_mc<3, false, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3213: rc = launch_recurrent_mc<1, true, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3223: rc = launch_recurrent_mc<2, true, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3233: rc = launch_recurrent_mc<3, true, 1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3214: rc = launch_recurrent_mc<1, true, 2>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3224: rc = launch_recurrent_mc<2, true, 2>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 3234: rc = launch_recurrent_mc<3, true, 2>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 6411: rc = launch_recurrent_pair<1>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        case 6421: rc = launch_recurrent_pair<2>(prm, m->tc_whh_frag[layer], rem, &done, xproj, st); break;
        default: return fail(HSSB_E_MODE, "HSSB_RC_GEOM=%d,%d,%d unsupported", nb, s, pair);
        }
        if (rc) return rc;
        base += done;
    }
    return 0;
}

namespace {
struct TcWs { size_t xhi, xlo, xproj, o1hi, o1lo, out2, hn, cn, gather, total; };
TcWs tc_ws_layout(int64_t B, int64_t T)
{
    const size_t M = (size_t)B * T;
    TcWs w{};
    size_t off = 0;
    const size_t Mp = (size_t)T * ((B + 31) / 32) * 32;          // batch padded to whole 32-column tiles (fused projection operand)
    w.xhi = off;   off += align_up(sizeof(__half) * Mp * 64, 1024);
    w.xlo = off;   off += align_up(sizeof(__half) * Mp * 64, 1024);
    w.xproj = off; off += align_up(sizeof(float) * 2 * (size_t)T * xproj_pitch(B) * TC_G, 1024);
    w.o1hi = off;  off += align_up(sizeof(__half) * M * TC_OP, 1024);
    w.o1lo = off;  off += align_up(sizeof(__half) * M * TC_OP, 1024);
    w.out2 = off;  off += align_up(sizeof(float) * M * TC_OP, 1024);
    w.hn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
    w.cn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
    w.gather = off; off += TC_GATHER_BYTES;
    w.total = off;
    return w;
}
}  // namespace

size_t tc_workspace_bytes(const hssb_model *, int64_t B, int64_t T) { return tc_ws_layout(B, T).total; }

int tc_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0, const float *c0, float *logp,
               int32_t *labels, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const TcWs w = tc_ws_layout(B, T);
    if (!ws || ws_bytes < w.total) return fail(HSSB_E_WORKSPACE, "model workspace %zu < %zu", ws_bytes, w.total);
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
    const char *fx = getenv("HSSB_FUSE_X");
    const bool fused = m->F <= 16 * RX_KSTEPS && !getenv("HSSB_RC_GEOM") && !(fx && fx[0] == '0');
    if (fused) {
        {
            ProfScope prof("split_planes", st);
            const long long n = T * ((B + RP_NBH - 1) / RP_NBH) * 8 * RP_NBH;
            split_planes_tiled_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, B, T, m->F, xhi, xlo);
            HSSB_LAUNCH_OK("split_planes_tiled_kernel");
        }
        if (int rc = tc_recurrent(m, 0, nullptr, h0, c0, hn, cn, o1hi, o1lo, nullptr, gather, B, T, st, xhi, xlo)) return rc;
    } else {
        {
            ProfScope prof("split_planes", st);
            split_planes_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(x, M, m->F, 64, xhi, xlo);
            HSSB_LAUNCH_OK("split_planes_kernel");
        }
        if (int rc = tc_inproj(m, 0, xhi, xlo, 64, B, T, xproj, st)) return rc;
        if (int rc = tc_recurrent(m, 0, xproj, h0, c0, hn, cn, o1hi, o1lo, nullptr, gather, B, T, st)) return rc;
    }
    if (int rc = tc_inproj(m, 1, o1hi, o1lo, TC_OP, B, T, xproj, st)) return rc;
    if (int rc = tc_recurrent(m, 1, xproj, hn, cn, hn, cn, nullptr, nullptr, out2, gather, B, T, st)) return rc;
    return head_forward(out2, M, TC_OP, m->tc_lin_w, m->lin_b, logp, labels, st);
}

}  // namespace hssb

// Diagnostic: clock64 stamps of the recurrence roles (cluster 0, rank 0) for the first `steps` steps of every
// following recurrence launch; buf = steps*4*16 uint64 on the device, nullptr disables.
extern "C" int hssb_debug_max_clusters(void)
{
    int n = 0;
    if (hssb::max_resident_clusters<32, 3, false>(&n)) return -1;
    return n;
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
    const size_t need = sizeof(float) * raw_floats + sizeof(__half) * 2 * M * 64;
    if (!workspace || workspace_bytes < need) return fail(HSSB_E_WORKSPACE, "hssb_debug_inproj: workspace %zu < %zu", workspace_bytes, need);
    float *raw = static_cast<float *>(workspace);
    __half *hi = reinterpret_cast<__half *>(raw + raw_floats), *lo = hi + M * 64;
    split_planes_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(x, M, m->F, 64, hi, lo);
    HSSB_LAUNCH_OK("split_planes_kernel");
    if (int rc = tc_inproj(m, 0, hi, lo, 64, B, T, raw, st)) return rc;
    unpermute_xproj_kernel<<<(unsigned)((2 * M * TC_G + 255) / 256), 256, 0, st>>>(raw, B, xproj_pitch(B), T, xproj);
    HSSB_LAUNCH_OK("unpermute_xproj_kernel");
    return 0;
}
