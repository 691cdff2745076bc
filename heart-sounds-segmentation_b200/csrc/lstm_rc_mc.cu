// K5m: the recurrence kernel of the BiLSTM (tc_recurrent_mc_kernel) -- all-gather through L2 with TMA multicast, optional fused
// layer-1 input projection -- and its launcher.
#include "lstm_tc_common.cuh"

namespace hssb {

// ------------------------------------------------------------------------------------------------
// K5m: the T sequential steps of one BiLSTM layer.  One 8-CTA cluster per (direction, S <= 3 sub-tiles of 32 batch columns).
//
// Orientation: gates are the MMA M dimension and stay put, the batch is N:
//     G^T[gate row (128 TMEM lanes), b (32 cols)] = W_hh,slice . h_{t-1}^T  (+ W_ih,slice . x_t^T when fused, else + xproj^T)
// CTA rank r owns units 30r..30r+29.  Its W_hh slice (fp16 hi and lo planes, K padded 240 -> 8 x 32) is loaded ONCE into TMEM
// columns [0, 256) and is the A operand of every MMA (tcgen05.mma with A in TMEM): weights never move.  h_{t-1}^T lives in
// shared memory as the K-major B operand [source rank 8][k-chunk 4][plane 2][column 32][8 units] (no swizzle: LBO = 1 KB
// between k-chunks, SBO = 128 B between 8-column groups), double-buffered by step parity.
//
// All-gather through L2.  Measured on B200 (scripts/microbench/ub_cluster.cu): a CTA pushes ~17 B/cycle into DSMEM, so an
// 8-way DSMEM all-gather of 96 columns costs 5 750 cycles per step -- twice the tensor time (that is lstm_rc_dsmem.cu).  Here
// every epilogue warp bulk-stores its 1 KB piece of h_t (fp16 hi / lo image) to an L2-resident scratch slot, waits for the
// store (cp.async.bulk.wait_group) and issues ONE multicast bulk load that lands the piece in slot `rank` of all 8 CTAs' B
// buffers and completes bytes on their mbarriers: 50-60 B/cycle into every SM, ~1 100-1 600 cycles end to end.
//   * every B buffer has one mbarrier per pair of source ranks; the MMA issuer starts on a pair's K range as soon as its two
//     slices landed (own pair first: its arrival proves that this CTA's epilogue warps have read the previous accumulator);
//   * the image is single-buffered (the publishing thread has waited for its bulk store, and nobody rewrites the image before
//     the next accumulator, which depends on that publish); the L2 scratch slot is double-buffered by step parity.
// Epilogue (4 warps per sub-tile, or 8 with EW = 2: one or two per TMEM lane quadrant).  TMEM lanes are in "fragment order"
// (lane = 32*(u/8) + 8*gate + u%8), so two tcgen05.ld.16x256b hand thread (ul = lane/4, cp = lane%4) the four gates of unit
// 8q+ul for the columns 8k + 2cp + {0,1}: no shuffles, and the matching xproj values are coalesced 16-byte loads, issued as
// the LAST thing of a step (an LDG in flight stalls every later long-scoreboard wait of the warp).  Activations with 8
// instead of 10 MUFU ops per cell: i*g = (1 - e_g) / ((1 + e_i)(1 + e_g)) with e_x = exp(-x) (exp(-2x) for g and c).
// relu(h_t) leaves through a per-warp shared-memory tile and one TMA tensor store per plane (slot-layout [B][T][512]
// outputs; ragged batches are clipped by the TMA unit).
// FUSE_X (layer 1): W_ih's slice sits in TMEM too; x_t (fp16 hi / lo planes, tile-major: one contiguous 3 KB run per step and
// sub-tile) is bulk-loaded one step ahead and W_ih . x_t is issued into the accumulator as soon as the epilogue has read the
// previous one (d_empty), i.e. while h_{t-1} is still in flight; the relu tile then reuses the warp's image piece.
// ------------------------------------------------------------------------------------------------
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
    static constexpr int USED_BYTES = S * PER_SUB + OUT_BYTES + X_BYTES + BAR_BYTES + 1024;
    // One CTA per SM, always: every CTA holds all 512 TMEM columns of its SM for the whole launch, so a second CTA on the same SM
    // would block in tcgen05.alloc until the first exits -- and two clusters holding each other's SMs that way never finish.  With
    // S = 1 the real need (~78 KB) would let two CTAs share an SM; asking for more than half of the SM's shared memory forbids it.
    static constexpr int SMEM_BYTES = USED_BYTES > 118 * 1024 ? USED_BYTES : 118 * 1024;
    static constexpr int THREADS = 32 * S + 128 * S * EW;            // S issuer warps + S x 4 x EW epilogue warps
    static_assert(EW == 1 || EW == 2, "one or two epilogue warps per TMEM lane quadrant and sub-tile");
    static_assert(S * RP_NBH <= 96, "accumulators sit in TMEM columns [256, 352)");
    static_assert((2 * S * RP_G + 4 * S) * 8 + 8 <= BAR_BYTES, "barrier area too small");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int S, bool WARP_PUBLISH, int EW, bool FUSE_X, bool TRAIN = false>
__global__ void __launch_bounds__(RmCfg<S, EW, FUSE_X>::THREADS, 1) tc_recurrent_mc_kernel(const __grid_constant__ RecurParams p)
{
    static_assert(!TRAIN || !FUSE_X, "the training forward takes its input projection from K4");
    using C = RmCfg<S, EW, FUSE_X>;
    // range guard (tc_forward): the fused launch is skipped when the input left the fp16-split range, its stand-in when it did not.
    // The flag is final before the launch, so every CTA of the grid takes the same branch (before any barrier / TMEM allocation).
    if (p.skip_flag && ((*p.skip_flag != 0) == (p.skip_when != 0))) return;
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
    // clock64 stamps of the roles (scripts/trace_recurrent.py) and the knock-out switches exist only in diagnostic builds
    // (-DHSSB_TRACE / -DHSSB_KNOCKOUTS): every instruction of the epilogue warps is on the dependent chain of a step
#ifdef HSSB_TRACE
#define RM_TRACE(ev, step, sub)                                                                                       \
    do {                                                                                                              \
        if (tr_buf && (step) >= 0 && (step) < p.trace_steps) tr_buf[(((step) * 4 + (sub)) * TR_EVENTS) + (ev)] = clock64(); \
    } while (0)
#else
#define RM_TRACE(ev, step, sub) do { (void)tr_buf; } while (0)
#endif
#ifdef HSSB_KNOCKOUTS
    const int knock = p.debug;
#else
    constexpr int knock = 0;
#endif

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * S * RP_G + 4 * S; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < S; ++i) mbar_init(&d_empty[i], 4 * EW);
        fence_barrier_init();
        if (!TRAIN) {
            const CUtensorMap *om = (EW == 1) ? p.out_map : p.out_map16;
            prefetch_tmap(&om[0]);
            if (!p.out_f32) prefetch_tmap(&om[1]);
        }
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < (S * C::PER_SUB + C::OUT_BYTES + C::X_BYTES) / 16; i += C::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();                          // every CTA's barriers are initialised before any multicast can target them
    if (p.resident && threadIdx.x == 0 && atomicAdd(p.resident, 1u) + 1 == gridDim.x)
        atomicExch(p.resident + 1, 1u);      // every CTA of this launch holds its SM: the gated projection launch may start

    // one-time: the W_hh slice (and, fused, the W_ih slice) -> TMEM, by the first four epilogue warps (one TMEM lane quadrant each)
    if (warp >= S && warp < S + 4) {
        const int q = warp & 3;
        // this thread owns lane 32q + lane; column c holds k' = 2c, 2c+1
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
    }
    tc_fence_before();
    __syncthreads();                         // weights are in TMEM (a block-wide barrier: every warp passes here exactly once)
    tc_fence_after();

    if (warp < S) {
        // ================= MMA issuer of sub-tile s = warp (one elected thread) =================
        const int s = warp;
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

        if (b0 < B) {
            constexpr int NW = C::NW;
            constexpr int NI = NW / 4;                  // (unit, column) cells per thread: columns cbase + 8*(i/2) + 2*cp + (i&1)
            constexpr bool EARLY_X = !FUSE_X && !TRAIN && S <= 2;
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
            int x_t = dir ? (int)T - 1 : 0, x_tile = -1;        // time index of the next xproj load and the time tile already waited for
            // xproj of the next step -> registers.  With three sub-tiles issued as the LAST thing of a step: every later long-scoreboard
            // wait of the warp (TMA issue, spill reloads, ...) would otherwise sit behind these HBM loads; with one or two sub-tiles
            // (and in the training variant) right after the publish: there the HBM round trip itself is what must be hidden.
            auto load_into = [&](float4 *dst) {
                if (!FUSE_X && p.chunk_done && (x_t >> 7) != x_tile) {
                    // the projection GEMM runs concurrently: wait until this direction's chunk of the time tile is in memory
                    x_tile = x_t >> 7;
                    if (lane == 0) {
                        const unsigned *flag = p.chunk_done + (dir ? 2 * (p.t_tiles - 1 - x_tile) + 1 : 2 * x_tile);
                        unsigned long long t_start = 0;
                        while (ld_acquire_u32(flag) < p.chunk_need) {
                            __nanosleep(256);
                            if (!t_start) t_start = globaltimer_ns();
                            else if (globaltimer_ns() - t_start > POLL_TIMEOUT_NS) { *p.timeout_flag = 1; break; }
                        }
                    }
                    __syncwarp();
                }
                x_t += dir ? -1 : 1;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int c = col_of(i);
                    if (FUSE_X) continue;               // fused: xnext holds the (constant) biases of this unit's four gates
                    dst[i] = (c < ncols && !(knock & 1)) ? __ldcs(reinterpret_cast<const float4 *>(xp_next + (size_t)c * TC_G)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                xp_next += xstep;
            };
            auto load_x = [&]() { load_into(xnext); };
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
                float kept[TRAIN ? 4 * NI : 1];          // TRAIN: the activated gates i, f, g, o of this thread's cells
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
                    if (knock & 4) {                                  // (timing experiment: no cell math)
#pragma unroll
                        for (int i = 0; i < NI; ++i) hv[i] = unit_ok ? 0.25f * (ei[i] + ef[i]) * 1e-3f + 1e-3f * (eg[i] + eo[i]) : 0.0f;
                    } else if (TRAIN) {
                        // back-propagation needs the four activated gates themselves: no shared reciprocals here
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            const float si = rcp_approx(1.0f + ei[i]), sf = rcp_approx(1.0f + ef[i]), so = rcp_approx(1.0f + eo[i]);
                            const float tg = (1.0f - eg[i]) * rcp_approx(1.0f + eg[i]);
                            const float c = fmaf(sf, c_state[i], si * tg);
                            c_state[i] = c;
                            const float ec = ex2_approx(fminf(c * (-2.0f * LOG2E), EMAX));
                            hv[i] = unit_ok ? so * ((1.0f - ec) * rcp_approx(1.0f + ec)) : 0.0f;
                            kept[(TRAIN ? 4 : 0) * i + 0] = si; kept[(TRAIN ? 4 * i + 1 : 0)] = sf;
                            kept[(TRAIN ? 4 * i + 2 : 0)] = tg; kept[(TRAIN ? 4 * i + 3 : 0)] = so;
                        }
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
                if (TRAIN) {
                    // the next step's xproj loads go out FIRST: they have one step of exchange latency to come back, and the 24 stores
                    // below would push them ~200 cycles later (the stores wait on no scoreboard, so the loads in flight cost them nothing).
                    // Measured: 7.97 -> 7.03 ms for the two layers at 50 x 2000; loading two steps ahead gains nothing more (7.2 ms).
                    if (t + 1 < Ti) load_x();
                    // ---- off the critical path: what back-propagation needs, straight to global memory (8 lanes = 8 consecutive units) ----
                    if (unit_ok) {
                        const int U = (int)rank * RC_U + u;
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            const int c = col_of(i);
                            if (c < ncols) {
                                const size_t row = (size_t)(b0 + c) * T + t_idx;
                                float *gp = p.tr_gates + ((size_t)dir * B * T + row) * TC_G + U;
                                gp[0] = kept[(TRAIN ? 4 * i : 0)]; gp[TC_H] = kept[(TRAIN ? 4 * i + 1 : 0)];
                                gp[2 * TC_H] = kept[(TRAIN ? 4 * i + 2 : 0)]; gp[3 * TC_H] = kept[(TRAIN ? 4 * i + 3 : 0)];
                                p.tr_cells[((size_t)dir * B * T + row) * TC_H + U] = c_state[i];
                                p.tr_out[row * (2 * TC_H) + dir * TC_H + U] = hv[i];
                            }
                        }
                    }
                } else {
                // With one or two sub-tiles per cluster a step is about as long as an HBM round trip: the next step's xproj loads go out
                // before the output store (measured at T = 2000: 6.41 -> 6.18 ms at 50 windows, 9.8 -> 9.55 ms at 300; with three
                // sub-tiles the other two sub-tiles' steps hide the latency and the late issue below is as fast or faster)
                if (EARLY_X && t + 1 < Ti) load_x();
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
                if (!(knock & 2) && elect_one()) {
                    const CUtensorMap *om = (EW == 1) ? p.out_map : p.out_map16;       // box of 32 / 16 batch columns
                    tma_store_3d(&om[0], out_tile, out_c0, t_idx, (int)b0 + cbase);
                    if (!p.out_f32) tma_store_3d(&om[1], out_tile + LO_OFF, out_c0, t_idx, (int)b0 + cbase);
                    tma_store_commit();
                }
                }
                if (p.tile_done && ((dir ? (t_idx & (TC_TT - 1)) == 0 : (t_idx & (TC_TT - 1)) == TC_TT - 1) || t + 1 == Ti)) {
                    // this warp's relu(h) of a whole time tile is on its way: once the bulk stores have completed, tell the
                    // projection GEMM of the next layer (generic-proxy release behind the async-proxy writes)
                    if (elect_one()) {
                        tma_store_wait<0>();
                        fence_proxy_async_all();
                        __threadfence();
                        atomicAdd(p.tile_done + dir * p.t_tiles + (t_idx >> 7), 1u);
                    }
                    __syncwarp();
                }
                t_idx += dir ? -1 : 1;
                if (t + 1 < Ti) {
                    if (!TRAIN && !EARLY_X) load_x();
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

// Per device, once: opt in to the kernel's shared memory and find out how many of its clusters are co-resident.  Also called for
// every default variant when a model is created (rc_mc_prepare): the first use of a kernel loads its code, which may synchronise
// the device -- that must not happen while a recurrence is polling for a projection launch that the host has not issued yet.
template <int S, bool WARP_PUBLISH, int EW, bool FUSE_X, bool TRAIN = false>
static int prepare_recurrent_mc(int *max_clusters_out)
{
    using C = RmCfg<S, EW, FUSE_X>;
    static PerDeviceInt cached_clusters;
    int max_clusters = cached_clusters.get();
    if (!max_clusters) {
        cudaLaunchAttribute attr[1];
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaFuncSetAttribute(tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_recurrent_mc_kernel)");
        cfg.gridDim = dim3(16 * RC_CL);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X, TRAIN>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_recurrent_mc_kernel)");
        if (n < 2) return fail(HSSB_E_DEVICE, "device fits only %d recurrence clusters", n);
        max_clusters = std::min(n, 16);
        cached_clusters.set(max_clusters);
    }
    *max_clusters_out = max_clusters;
    return 0;
}

int rc_mc_prepare()
{
    int n = 0;
    if (int rc = prepare_recurrent_mc<1, true, 2, true>(&n)) return rc;
    if (int rc = prepare_recurrent_mc<2, true, 2, true>(&n)) return rc;
    if (int rc = prepare_recurrent_mc<3, true, 1, true>(&n)) return rc;
    if (int rc = prepare_recurrent_mc<1, true, 2, false>(&n)) return rc;
    if (int rc = prepare_recurrent_mc<2, true, 2, false>(&n)) return rc;
    return prepare_recurrent_mc<3, true, 1, false>(&n);
}

template <int S, bool WARP_PUBLISH, int EW, bool FUSE_X = false, bool TRAIN = false>
static int launch_recurrent_mc(const RecurParams &prm_in, const __half *whh_frag, int64_t rem, int *cols_done, const float *xproj, cudaStream_t st,
                               RecurLaunchInfo *info)
{
    using C = RmCfg<S, EW, FUSE_X>;
    RecurParams prm = prm_in;
    prm.whh = whh_frag;
    prm.trace = g_trace_buf;
    prm.trace_steps = g_trace_steps;
    if (const char *e = getenv("HSSB_TRACE_LAYER")) if (atoi(e) != prm.layer) prm.trace = nullptr;
    int max_clusters = 0;
    if (int rc = prepare_recurrent_mc<S, WARP_PUBLISH, EW, FUSE_X, TRAIN>(&max_clusters)) return rc;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int per = RP_NBH * S;
    const int groups = (int)std::min<int64_t>(max_clusters / 2, (rem + per - 1) / per);
    *cols_done = groups * per;
    prm.xproj = xproj;
    prm.groups = groups;
    prm.stagger_ns = 600;
    if (const char *e = getenv("HSSB_RC_STAGGER")) prm.stagger_ns = atoi(e);
#ifdef HSSB_KNOCKOUTS   // timing experiments that skip parts of the step (results are WRONG): only in -DHSSB_KNOCKOUTS builds
    if (const char *e = getenv("HSSB_RC_DEBUG")) prm.debug = atoi(e);
#endif
    cfg.gridDim = dim3((unsigned)(2 * groups * RC_CL));
    if (info) {
        info->ctas = 2 * groups * RC_CL;
        unsigned live = 0;                          // sub-tiles of one direction that hold batch columns
        for (int g = 0; g < groups; ++g)
            for (int i = 0; i < S; ++i) live += (prm.b_base + ((long long)g * S + i) * RP_NBH < prm.B) ? 1u : 0u;
        info->signals_per_dir = live * RC_CL * 4 * EW;
    }
    // (the stand-in launch of the input-range guard is a no-op unless the guard fired: timed under its own name)
    ProfScope prof(TRAIN ? "tc_recurrent_train" : (prm.skip_flag && !prm.skip_when) ? "range_standin" : (prm.layer ? "tc_recurrent_l2" : "tc_recurrent_l1"), st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_recurrent_mc_kernel<S, WARP_PUBLISH, EW, FUSE_X, TRAIN>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_recurrent_mc_kernel)");
    return 0;
}


int rc_mc_launch(int s, int variant, bool fused, const RecurParams &prm, const __half *whh_frag, int64_t rem, int *done,
                 const float *xproj, cudaStream_t st, RecurLaunchInfo *info, bool train)
{
    if (train) {        // training forward: keeps activated gates / cell states / raw h instead of the relu'd outputs
        if (s == 1) return launch_recurrent_mc<1, true, 2, false, true>(prm, whh_frag, rem, done, xproj, st, info);
        if (s == 2) return launch_recurrent_mc<2, true, 2, false, true>(prm, whh_frag, rem, done, xproj, st, info);
        return launch_recurrent_mc<3, true, 1, false, true>(prm, whh_frag, rem, done, xproj, st, info);
    }
    if (fused) {        // the default variants only: two epilogue warps per quadrant up to 64 columns per cluster, one beyond
        if (s == 1) return launch_recurrent_mc<1, true, 2, true>(prm, whh_frag, rem, done, xproj, st, info);
        if (s == 2) return launch_recurrent_mc<2, true, 2, true>(prm, whh_frag, rem, done, xproj, st, info);
        if (s == 3) return launch_recurrent_mc<3, true, 1, true>(prm, whh_frag, rem, done, xproj, st, info);
    }
    switch (s * 10 + variant) {
    case 12: return launch_recurrent_mc<1, false, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 22: return launch_recurrent_mc<2, false, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 32: return launch_recurrent_mc<3, false, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 13: return launch_recurrent_mc<1, true, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 23: return launch_recurrent_mc<2, true, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 33: return launch_recurrent_mc<3, true, 1>(prm, whh_frag, rem, done, xproj, st, info);
    case 14: return launch_recurrent_mc<1, true, 2>(prm, whh_frag, rem, done, xproj, st, info);
    case 24: return launch_recurrent_mc<2, true, 2>(prm, whh_frag, rem, done, xproj, st, info);
    case 34: return launch_recurrent_mc<3, true, 2>(prm, whh_frag, rem, done, xproj, st, info);
    default: return fail(HSSB_E_MODE, "multicast recurrence: %d sub-tiles, variant %d unsupported", s, variant);
    }
}

}  // namespace hssb
