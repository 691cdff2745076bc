// K5b: back-propagation through time of one bidirectional LSTM layer on the tcgen05 tensor cores (training) --
// what autograd runs under nn.LSTM for loss.backward() in the reference (main.py:72 over segmenter.py:80-83).
//
// Same cluster geometry as the forward recurrence K5m: 8 CTAs per (direction, group of NB batch columns), CTA `rank` owns units
// 30r .. 30r+29.  Per step of the reverse recurrence a CTA
//   1. sums the eight partial dL/dh blocks its peers (and itself) sent for its 30 units           (reduce-scatter, fp32)
//   2. turns dL/dh + dL/dc into the gate gradients dG of its units -- every factor that depends only on the saved forward
//      values (activated gates, c_t, c_{t-1}, d_out) is computed BEFORE the partials arrive, so 8 adds and 6 FMAs per cell sit
//      on the dependent chain -- writes dG (fp32, in place over the gates) for the weight-gradient GEMMs, and its fp16 hi / lo
//      split, scaled by a power of two, into the K-major B operand [k = gate*32 + unit (128)][column]
//   3. issues  P[slot (256 = 8 ranks x 32), column] = W_hh,slice^T . dG   (two M = 128 tiles x K = 128 x three split products):
//      the TRANSPOSED W_hh slice (the 120 gate rows of its own units) sits in TMEM columns [0, 256) for the whole launch and
//      is the A operand (tcgen05.mma with A in TMEM); no exchange is needed BEFORE the MMA because the contraction runs over the
//      CTA's own gate rows
//   4. drains the accumulator: TMEM lane quadrant q of tile j holds exactly the 32 slots of peer 4j + q, so each of the eight
//      epilogue warps stages one [column][32 units] block and sends it with ONE bulk copy (shared::cta -> shared::cluster,
//      complete_tx on the receiver's mbarrier): NB x 128 bytes per peer and step.  (The block for the CTA itself is stored
//      straight into its receive buffer; a named barrier of the eight epilogue warps after the drain orders it.)
// Hazards (no explicit "buffer free" signalling is needed): a peer can send step k+2 only after it received step k+1 from every
// CTA, i.e. after all eight of my warps finished the math of step k+1 -- which read the receive buffer of step k (the one k+2
// overwrites), drained the accumulator of step k and, one step earlier still, had their staged block of step k delivered.
// Receive and staging buffers are therefore double-buffered by step parity and nothing else.
//
// fp16 range: gradients are tiny (1 / (B T) from the mean loss), far below fp16's normal range.  dG is scaled by 2^e, e from the
// max-abs of the incoming gradients (d_out, d_hn) such that it lands in [4, 8): 2^13 of head room for growth through the
// recurrence (the split saturates at +-60000 instead of producing inf), absolute resolution 2^-25 of that maximum -- the
// rounding noise of an fp32 sum of 960 products.  The dG written to memory is the unscaled fp32 value.
#include "lstm_tc_common.cuh"

namespace hssb {

struct BpttParams {
    const float *gates;      // [2][B*T][960]  activated gates i, f, g, o of the forward
    float *dg;               // [2][B*T][960]  out: dG (may alias gates: a cell's gates are in registers before its dG is written); nullable
    float *dg_hi, *dg_lo;    // the same values as a TF32-exact head and its fp32 residue (operands of the gradient GEMMs); nullable pair,
                             // laid out [B*T][2][960]: both directions of a row side by side, so that dG^T x and dG W_ih are ONE GEMM each
    float *db;               // [2][960]  out: sum of dG over batch and time = the gradient of b_ih and of b_hh (zeroed by the launcher); nullable
    const float *cells;      // [2][B*T][240]
    const float *c0;         // [2][B][240]
    const float *d_out;      // [B][T][480]
    const float *d_hn, *d_cn;   // [2][B][240], nullable
    float *dh0, *dc0;        // [2][B][240]
    const __half *whhT;      // [dir][rank][plane][256 slots][128 k]: W_hh[gate*240 + 30 rank + ku][30 (slot >> 5) + (slot & 31)], k = 32 gate + ku
    const unsigned *range;   // word 1: bits of max |d_out|, |d_hn|
    long long B, T;
    int b_base;
};

template <int NB>
struct BpCfg {
    static constexpr int MMA_N = NB < 16 ? 16 : NB;            // M = 128 needs N % 16 == 0; with NB = 8 the upper 8 columns stay zero
    static constexpr int NC = NB / 8;                          // (unit, column) cells per thread: columns w8 + 8 i
    static constexpr int CH_STRIDE = 2 * MMA_N * 16 + 16;      // one k-chunk [plane][column][8 k] fp16, +16 B: the four chunks a warp
                                                               // writes with one store land in different banks
    static constexpr int B_BYTES = (16 * CH_STRIDE + 127) / 128 * 128;
    static constexpr int BLOCK = NB * 128;                     // one partial block [column][32 units] fp32
    static constexpr int RECV_BYTES = 2 * RC_CL * BLOCK;       // [parity][source rank]
    static constexpr int STAGE_BYTES = 2 * RC_CL * BLOCK;      // [parity][epilogue warp]
    static constexpr int BAR_BYTES = 128;
    static constexpr int USED_BYTES = RECV_BYTES + STAGE_BYTES + B_BYTES + BAR_BYTES + 1024;
    // Every CTA allocates all 512 TMEM columns of its SM for the whole launch: a second CTA on the same SM would block in
    // tcgen05.alloc until the first one exits, and two clusters that hold each other's SMs that way never finish.  Asking for more
    // than half of the SM's shared memory keeps it to one CTA per SM (the larger kernels K4 / K5m get that from their real needs).
    static constexpr int SMEM_BYTES = USED_BYTES > 118 * 1024 ? USED_BYTES : 118 * 1024;
    static constexpr int THREADS = 32 + 8 * 32;                // issuer warp + 8 epilogue warps (two per TMEM lane quadrant: one per M tile)
    static_assert(NB == 8 || NB == 16 || NB == 32, "batch columns per cluster");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld_x8(taddr, v); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_x16(taddr, v); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_x32(taddr, v); }

template <int NB>
__global__ void __cluster_dims__(RC_CL, 1, 1) __launch_bounds__(BpCfg<NB>::THREADS, 1) tc_bptt_kernel(const __grid_constant__ BpttParams p)
{
    using C = BpCfg<NB>;
    constexpr int NC = C::NC;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    float *recv = reinterpret_cast<float *>(smem);                                  // [parity][source][column][32 units]
    float *stage = reinterpret_cast<float *>(smem + C::RECV_BYTES);                 // [parity][warp][column][32 units]
    unsigned char *bop = smem + C::RECV_BYTES + C::STAGE_BYTES;                     // B operand: 16 k-chunks
    uint64_t *bars = reinterpret_cast<uint64_t *>(bop + C::B_BYTES);
    uint64_t *r_full = bars;          // [2]  the eight partial blocks of a step have landed
    uint64_t *b_full = bars + 2;      //      every epilogue warp has written its part of the B operand
    uint64_t *d_full = bars + 3;      //      accumulator complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x / RC_CL;
    const int dir = cid & 1, group = cid >> 1;
    const long long T = p.T, B = p.B;
    const long long b0 = (long long)p.b_base + (long long)group * NB;
    const int Ti = (int)T;

    if (threadIdx.x == 0) {
        mbar_init(&r_full[0], 1);        // armed with the bytes of the seven peer blocks (my own block does not travel)
        mbar_init(&r_full[1], 1);
        mbar_init(b_full, 8);
        mbar_init(d_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < C::B_BYTES / 16; i += C::THREADS) reinterpret_cast<uint4 *>(bop)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // one-time: W_hh,slice^T -> TMEM.  Tile j, plane pl: columns (2j + pl) * 64 .. + 63, lane = slot within the tile, column c = k 2c, 2c+1
    if (warp >= 1 && warp <= 4) {
        const int q = warp & 3;
#pragma unroll 1
        for (int jp = 0; jp < 4; ++jp) {
            const int j = jp >> 1, plane = jp & 1;
            const __half *row = p.whhT + (((((size_t)dir * RC_CL + rank) * 2 + plane) * 256) + 128 * j + 32 * q + lane) * 128;
            const uint4 *src = reinterpret_cast<const uint4 *>(row);
#pragma unroll 4
            for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 v0 = __ldg(src + 2 * c8), v1 = __ldg(src + 2 * c8 + 1);
                const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                tmem_st_x8(tmem_base + ((uint32_t)(q * 32) << 16) + jp * 64 + c8 * 8, r);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync();                          // every CTA's barriers are initialised before a peer's bulk copy can complete on them

    if (warp == 0) {
        // ================= MMA issuer (one elected thread); also arms the receive barriers =================
        if (elect_one()) {
            mbar_arrive_expect_tx(&r_full[0], (RC_CL - 1) * C::BLOCK);
            if (Ti > 1) mbar_arrive_expect_tx(&r_full[1], (RC_CL - 1) * C::BLOCK);
            constexpr uint32_t idesc = make_idesc_f16(128, C::MMA_N);
            const uint32_t bb = smem_u32(bop);
            for (int k = 0; k < Ti; ++k) {
                mbar_wait(b_full, (uint32_t)(k & 1));
                // all eight warps have consumed the blocks of step k-1: their barrier is in its next phase, which step k+1 will complete
                if (k >= 1 && k + 1 < Ti) mbar_arrive_expect_tx(&r_full[(k + 1) & 1], (RC_CL - 1) * C::BLOCK);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t d_tmem = tmem_base + 256 + j * C::MMA_N;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t blk = bb + 2 * ks * C::CH_STRIDE;
                        const uint64_t b_hi = make_smem_desc(blk, C::CH_STRIDE, 128, LAYOUT_NONE);
                        const uint64_t b_lo = make_smem_desc(blk + C::MMA_N * 16, C::CH_STRIDE, 128, LAYOUT_NONE);
                        const uint32_t a_hi = tmem_base + (2 * j) * 64 + ks * 8, a_lo = a_hi + 64;
                        mma_f16_ts(d_tmem, a_hi, b_hi, idesc, ks != 0);
                        mma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1);
                        mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                }
                mma_commit(d_full);
            }
        }
    } else {
        // ================= epilogue warp w8: cells (unit = lane, columns w8 + 8 i); drains TMEM quadrant q of tile j for peer 4j + q ====
        const int w8 = warp - 1;
        const int q = warp & 3, j = w8 >> 2;
        const uint32_t dest = (uint32_t)(4 * j + q);
        const int u = lane;
        const bool unit_ok = u < RC_U;
        const int U = (int)rank * RC_U + (unit_ok ? u : 0);
        bool ok[NC];
        size_t cellbase[NC];                 // (b0 + c) * T : row of time 0 of this cell's window
        size_t stateo[NC];                   // [dir][b][240] offset
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const long long b = b0 + w8 + 8 * i;
            ok[i] = unit_ok && b < B;
            const long long bc = b < B ? b : B - 1;
            cellbase[i] = (size_t)bc * T;
            stateo[i] = ((size_t)dir * B + bc) * TC_H + U;
        }
        const int e_up = [&] {
            const unsigned bits = p.range[1];
            const int ex = (int)((bits >> 23) & 255u) - 127;
            if (bits == 0 || ex == 128) return 0;
            const int e = 2 - ex;
            return e < -100 ? -100 : (e > 100 ? 100 : e);
        }();
        const float s_up = pow2f(e_up), s_dn = pow2f(-e_up);
        const float *gd = p.gates + (size_t)dir * B * T * TC_G + U;
        const size_t dgo = (size_t)dir * B * T * TC_G + U;
        const float *cd = p.cells + (size_t)dir * B * T * TC_H + U;
        const float *dd = p.d_out + dir * TC_H + U;

        // Saved forward values, loaded TWO steps ahead: a step (1.2 us) is not much longer than an HBM round trip, and the values of
        // step k + 1 are needed at the very top of that step.  sv = the step being processed, nx = the one after it.
        // [0..3] activated gates i, f, g, o; [4] c of the previous time step (c0 at the last one); [5] d_out
        float sv[NC][6], nx[NC][6];
        float c_cur[NC], dc_carry[NC], dh_rec[NC];
        float db_acc[4] = {0.f, 0.f, 0.f, 0.f};      // this thread's unit, summed over its columns and all steps
        auto t_of = [&](int k) { return dir ? k : Ti - 1 - k; };
        auto load_step = [&](int k, float (&dst)[NC][6]) {
            const int t = t_of(k);
            const int t_prev = dir ? t + 1 : t - 1;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (!ok[i]) continue;
                const size_t row = cellbase[i] + t;
                const float *g = gd + row * TC_G;
                dst[i][0] = __ldcs(g); dst[i][1] = __ldcs(g + TC_H); dst[i][2] = __ldcs(g + 2 * TC_H); dst[i][3] = __ldcs(g + 3 * TC_H);
                dst[i][5] = __ldcs(dd + row * (2 * TC_H));
                dst[i][4] = k + 1 < Ti ? __ldg(cd + (cellbase[i] + t_prev) * TC_H) : __ldg(p.c0 + stateo[i]);
            }
        };
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
            for (int v = 0; v < 6; ++v) sv[i][v] = nx[i][v] = 0.f;
            c_cur[i] = ok[i] ? __ldg(cd + (cellbase[i] + t_of(0)) * TC_H) : 0.f;
            dh_rec[i] = (ok[i] && p.d_hn) ? __ldg(p.d_hn + stateo[i]) : 0.f;
            dc_carry[i] = (ok[i] && p.d_cn) ? __ldg(p.d_cn + stateo[i]) : 0.f;
        }
        load_step(0, sv);
        if (Ti > 1) load_step(1, nx);
        const uint32_t bop_thread = smem_u32(bop) + (u >> 3) * C::CH_STRIDE + (u & 7) * 2;     // + gate * 4 chunks + plane + column * 16
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + j * C::MMA_N;

        auto reduce = [&](int k_sent) {      // the partials of step k_sent, summed over the eight source CTAs and scaled back
            mbar_wait_cluster(&r_full[k_sent & 1], (uint32_t)((k_sent >> 1) & 1));
            const float *rb = recv + (size_t)(k_sent & 1) * RC_CL * (C::BLOCK / 4) + lane;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const float *rc = rb + (w8 + 8 * i) * 32;
                float s = 0.f;
#pragma unroll
                for (int src = 0; src < RC_CL; ++src) s += rc[src * (C::BLOCK / 4)];
                dh_rec[i] = s * s_dn;
            }
        };

        for (int k = 0; k < Ti; ++k) {
            const int t = t_of(k);
            // ---- everything that does not depend on the incoming gradient ----
            float a1[NC], ko[NC], ki[NC], kf[NC], kg[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const float gi = sv[i][0], gf = sv[i][1], gg = sv[i][2], go = sv[i][3];
                const float tc = tanhf(c_cur[i]);
                a1[i] = go * (1.f - tc * tc);
                ko[i] = tc * go * (1.f - go);
                ki[i] = gg * gi * (1.f - gi);
                kf[i] = sv[i][4] * gf * (1.f - gf);
                kg[i] = gi * (1.f - gg * gg);
            }
            if (k > 0) reduce(k - 1);
            // ---- the dependent chain: dL/dh -> dG -> B operand ----
            float da[NC][4];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const float dh = sv[i][5] + dh_rec[i];
                const float dc = fmaf(dh, a1[i], dc_carry[i]);
                da[i][0] = dc * ki[i]; da[i][1] = dc * kf[i]; da[i][2] = dc * kg[i]; da[i][3] = dh * ko[i];
                dc_carry[i] = dc * sv[i][1];
                if (ok[i]) {
                    const uint32_t cell = bop_thread + (w8 + 8 * i) * 16;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        __half hh, hl;
                        split_f16(fminf(fmaxf(da[i][g] * s_up, -60000.f), 60000.f), hh, hl);
                        sts_b16(cell + g * 4 * C::CH_STRIDE, hh);
                        sts_b16(cell + g * 4 * C::CH_STRIDE + C::MMA_N * 16, hl);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full);
            // ---- off the chain: rotate the saved values (step k + 1 arrived during this step), send the loads of step k + 2 on
            //      their way BEFORE the stores (which wait on no scoreboard), then dG to memory ----
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                c_cur[i] = sv[i][4];
#pragma unroll
                for (int v = 0; v < 6; ++v) sv[i][v] = nx[i][v];
            }
            if (k + 2 < Ti) load_step(k + 2, nx);
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (ok[i]) {
                    const size_t o = dgo + (cellbase[i] + t) * TC_G;
#pragma unroll
                    for (int g = 0; g < 4; ++g) db_acc[g] += da[i][g];
                    if (p.dg) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) __stcs(p.dg + o + g * TC_H, da[i][g]);
                    }
                    if (p.dg_hi) {
                        const size_t o2 = ((cellbase[i] + t) * 2 + dir) * TC_G + U;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float hi, lo;
                            tf32_split(da[i][g], hi, lo);
                            __stcs(p.dg_hi + o2 + g * TC_H, hi);
                            __stcs(p.dg_lo + o2 + g * TC_H, lo);
                        }
                    }
                }
            // ---- partial dL/dh_prev of peer `dest`'s units: TMEM -> staging -> one bulk copy into its receive buffer ----
            mbar_wait(d_full, (uint32_t)(k & 1));
            tc_fence_after();
            uint32_t v[NB];
            tmem_ld_cols(taddr, v);
            tmem_ld_wait();
            tc_fence_before();
            float *mine = recv + ((size_t)(k & 1) * RC_CL + rank) * (C::BLOCK / 4);      // slot `rank` of a receive buffer
            if (dest == rank) {
                // my own block: straight into my receive buffer (a shared::cta -> shared::cluster bulk copy must target ANOTHER CTA);
                // the block barrier below orders it against the other warps' reads of this buffer one step earlier and one step later
#pragma unroll
                for (int c = 0; c < NB; ++c) mine[c * 32 + lane] = __uint_as_float(v[c]);
            } else {
                float *sb = stage + ((size_t)(k & 1) * RC_CL + w8) * (C::BLOCK / 4);
#pragma unroll
                for (int c = 0; c < NB; ++c) sb[c * 32 + lane] = __uint_as_float(v[c]);
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) bulk_copy_to_cta(mine, sb, C::BLOCK, &r_full[k & 1], dest);
            }
            named_barrier(1, 8 * 32);        // the eight epilogue warps: all arrive within the drain of the same accumulator
        }
        reduce(Ti - 1);
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (ok[i]) {
                p.dh0[stateo[i]] = dh_rec[i];
                p.dc0[stateo[i]] = dc_carry[i];
            }
        if (p.db && unit_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) atomicAdd(p.db + dir * TC_G + g * TC_H + U, db_acc[g]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                          // nobody leaves while a peer's copy into this CTA could still be in flight
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// torch W_hh [960][240] (rows gate*240 + unit) -> [dir][rank][plane][256 slots][128 k] fp16 hi / lo, zero in the padding
__global__ void pack_whhT_kernel(const float *__restrict__ w, int dir, __half *__restrict__ dst)
{
    const int rank = blockIdx.x >> 8, m = blockIdx.x & 255, k = threadIdx.x;
    const int r_dst = m >> 5, u_dst = m & 31, g = k >> 5, ku = k & 31;
    float v = 0.f;
    if (u_dst < RC_U && ku < RC_U) v = w[(size_t)(g * TC_H + rank * RC_U + ku) * TC_H + r_dst * RC_U + u_dst];
    __half hi, lo;
    split_f16(v, hi, lo);
    const size_t o = ((((size_t)dir * RC_CL + rank) * 2) * 256 + m) * 128 + k;
    dst[o] = hi;
    dst[o + (size_t)256 * 128] = lo;
}

__global__ void input_amax2_kernel(const float *__restrict__ a, long long na, const float *__restrict__ b, long long nb, unsigned *__restrict__ range)
{
    unsigned m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(i < na ? __ldg(a + i) : __ldg(b + (i - na)))));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(range + 1, m);
}

namespace {
constexpr size_t BPTT_W_BYTES = sizeof(__half) * 2 * RC_CL * 2 * 256 * 128;      // both directions
constexpr size_t BPTT_WS_BYTES = BPTT_W_BYTES + 256;

template <int NB>
int bptt_prepare(int *max_clusters_out)
{
    using C = BpCfg<NB>;
    static PerDeviceInt cached;          // 0 = not asked yet
    int n = cached.get();
    if (!n) {
        cudaError_t e = cudaFuncSetAttribute(tc_bptt_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_bptt_kernel)");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(RC_CL, 1, 1);
        cfg.blockDim = dim3(C::THREADS, 1, 1);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, tc_bptt_kernel<NB>, &cfg);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveClusters(tc_bptt_kernel)");
        if (n < 2) return fail(HSSB_E_DEVICE, "tc_bptt_kernel: fewer than two 8-CTA clusters fit on this device");
        cached.set(n);
    }
    *max_clusters_out = n;
    return 0;
}

template <int NB>
int bptt_launch(BpttParams prm, long long rem, int max_clusters, long long *done, cudaStream_t st)
{
    using C = BpCfg<NB>;
    const long long groups = std::min<long long>((rem + NB - 1) / NB, max_clusters / 2);
    *done = std::min<long long>(rem, groups * NB);
    ProfScope prof("tc_bptt", st);
    tc_bptt_kernel<NB><<<dim3((unsigned)(groups * 2 * RC_CL)), C::THREADS, C::SMEM_BYTES, st>>>(prm);
    HSSB_LAUNCH_OK("tc_bptt_kernel");
    return 0;
}
}  // namespace

size_t bptt_tc_workspace_bytes() { return BPTT_WS_BYTES; }

int bptt_tc_prepare()
{
    int n = 0;
    if (int rc = bptt_prepare<8>(&n)) return rc;
    if (int rc = bptt_prepare<16>(&n)) return rc;
    return bptt_prepare<32>(&n);
}

int bptt_tc_backward(const float *gates, float *dg, float *dg_hi, float *dg_lo, float *db, const float *cells, const float *w_fwd, const float *w_rev, const float *c0, const float *d_out,
                     const float *d_hn, const float *d_cn, int64_t B, int64_t T, float *dh0, float *dc0, void *ws, size_t ws_bytes,
                     cudaStream_t st)
{
    if (!ws || ws_bytes < BPTT_WS_BYTES) return fail(HSSB_E_WORKSPACE, "bptt workspace %zu < %zu", ws_bytes, BPTT_WS_BYTES);
    __half *whhT = static_cast<__half *>(ws);
    unsigned *range = reinterpret_cast<unsigned *>(static_cast<char *>(ws) + BPTT_W_BYTES);
    HSSB_CUDA_OK(cudaMemsetAsync(range, 0, 8, st));
    if (db) HSSB_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * 2 * TC_G, st));
    {
        ProfScope prof("bptt_pack", st);
        pack_whhT_kernel<<<RC_CL * 256, 128, 0, st>>>(w_fwd, 0, whhT);
        pack_whhT_kernel<<<RC_CL * 256, 128, 0, st>>>(w_rev, 1, whhT);
        const long long n_out = B * T * 2 * TC_H, n_state = d_hn ? 2 * B * TC_H : 0;
        input_amax2_kernel<<<148 * 4, 256, 0, st>>>(d_out, n_out, d_hn, n_state, range);
        HSSB_LAUNCH_OK("bptt pack kernels");
    }
    int mc8 = 0, mc16 = 0, mc32 = 0;
    if (int rc = bptt_prepare<8>(&mc8)) return rc;
    if (int rc = bptt_prepare<16>(&mc16)) return rc;
    if (int rc = bptt_prepare<32>(&mc32)) return rc;
    int force_nb = 0;
    if (const char *e = getenv("HSSB_BPTT_NB")) force_nb = atoi(e);
    BpttParams prm = {};
    prm.gates = gates; prm.dg = dg; prm.dg_hi = dg_hi; prm.dg_lo = dg_lo; prm.db = db; prm.cells = cells; prm.c0 = c0; prm.d_out = d_out; prm.d_hn = d_hn; prm.d_cn = d_cn;
    prm.dh0 = dh0; prm.dc0 = dc0; prm.whhT = whhT; prm.range = range; prm.B = B; prm.T = T;
    for (long long base = 0; base < B;) {
        const long long rem = B - base;
        prm.b_base = (int)base;
        long long done = 0;
        int rc;
        // as few columns per cluster as one wave of co-resident clusters allows: the per-step latency grows with the block size
        const int nb = force_nb ? force_nb : ((rem + 7) / 8 * 2 <= mc8 ? 8 : ((rem + 15) / 16 * 2 <= mc16 ? 16 : 32));
        if (nb == 8) rc = bptt_launch<8>(prm, rem, mc8, &done, st);
        else if (nb == 16) rc = bptt_launch<16>(prm, rem, mc16, &done, st);
        else if (nb == 32) rc = bptt_launch<32>(prm, rem, mc32, &done, st);
        else return fail(HSSB_E_MODE, "HSSB_BPTT_NB=%d", nb);
        if (rc) return rc;
        base += done;
    }
    return 0;
}

}  // namespace hssb
