// K5 / K5p: the recurrence kernels whose all-gather runs over distributed shared memory (tc_recurrent_kernel, the kernel of
// the first sessions, and tc_recurrent_pair_kernel, its cta_group::2 variant).  Validated, selectable with HSSB_RC_GEOM, not
// used by default: the L2-multicast kernel of lstm_rc_mc.cu is faster at every batch size measured.
#include "lstm_tc_common.cuh"

namespace hssb {

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
    static constexpr int USED_BYTES = S * PER_SUB + BAR_BYTES + 1024;
    static constexpr int SMEM_BYTES = USED_BYTES > 118 * 1024 ? USED_BYTES : 118 * 1024;     // one CTA per SM: each holds all of the SM's TMEM (see K5m)
    static constexpr int THREADS = 32 * S + 128 * S;         // S MMA-issuer warps + S epilogue groups of 4 warps
    static_assert(NB % 16 == 0 && NB <= 64, "NB must be 16, 32, 48 or 64");
    static_assert(S * NB <= 256, "accumulators must fit in the TMEM columns left of the weights");
    static_assert(5 * S * 8 <= BAR_BYTES - 8, "barrier area too small");
    static_assert(!PAIR || NB % 32 == 0, "pair mode splits NB in two halves of a multiple of 16 columns");
};

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
            const size_t o_stride = (size_t)4 * p.Tp * TC_OP;
            size_t o_next = ((size_t)(b0 + j) * p.Tp + (dir ? T - 1 : 0)) * TC_OP + hcol;
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
template <int S>
struct RpCfg {
    static constexpr int PER_SUB = 2 * RP_HBUF + 4 * RP_SLICE;       // 2 B buffers + images [parity][half]
    static constexpr int OUT_BYTES = S * 8 * 1024;                   // per epilogue warp: relu(h) tile for the TMA store
    static constexpr int BAR_BYTES = 512;
    static constexpr int USED_BYTES = S * PER_SUB + OUT_BYTES + BAR_BYTES + 1024;
    static constexpr int SMEM_BYTES = USED_BYTES > 118 * 1024 ? USED_BYTES : 118 * 1024;     // one CTA per SM: each holds all of the SM's TMEM (see K5m)
    static constexpr int THREADS = 32 * S + 256 * S;                 // S issuer / relay warps + S x 8 epilogue warps
    static_assert(S * RP_NB <= 256, "accumulators must fit in the TMEM columns left of the weights");
    static_assert((4 * S * RP_G + S) * 8 + 8 <= BAR_BYTES, "barrier area too small");
};

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
    static PerDeviceInt cache;
    int cached = cache.get();
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
        cache.set(n);
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
    static PerDeviceInt cached_clusters;
    int max_clusters = cached_clusters.get();
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
        cached_clusters.set(n);
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


int rc_dsmem_max_clusters(int *out) { return max_resident_clusters<32, 3, false>(out); }

int rc_dsmem_launch(int nb, int s, int pair, const RecurParams &prm, int64_t rem, int *done, const float *xproj, cudaStream_t st)
{
    switch (nb * 100 + s * 10 + pair) {
    case 1610: return launch_recurrent<16, 1, false>(prm, rem, done, xproj, st);
    case 1620: return launch_recurrent<16, 2, false>(prm, rem, done, xproj, st);
    case 1630: return launch_recurrent<16, 3, false>(prm, rem, done, xproj, st);
    case 3220: return launch_recurrent<32, 2, false>(prm, rem, done, xproj, st);
    case 3230: return launch_recurrent<32, 3, false>(prm, rem, done, xproj, st);
    case 3211: return launch_recurrent<32, 1, true>(prm, rem, done, xproj, st);
    case 3221: return launch_recurrent<32, 2, true>(prm, rem, done, xproj, st);
    case 3231: return launch_recurrent<32, 3, true>(prm, rem, done, xproj, st);
    case 3241: return launch_recurrent<32, 4, true>(prm, rem, done, xproj, st);
    default: return fail(HSSB_E_MODE, "HSSB_RC_GEOM=%d,%d,%d unsupported", nb, s, pair);
    }
}

int rc_pair_launch(int s, const RecurParams &prm, const __half *whh_frag, int64_t rem, int *done, const float *xproj, cudaStream_t st)
{
    if (s == 1) return launch_recurrent_pair<1>(prm, whh_frag, rem, done, xproj, st);
    if (s == 2) return launch_recurrent_pair<2>(prm, whh_frag, rem, done, xproj, st);
    return fail(HSSB_E_MODE, "pair recurrence: %d sub-tiles unsupported", s);
}

}  // namespace hssb
