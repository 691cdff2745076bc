// Training recurrences with cluster-resident weights (hidden_size 240): an 8-CTA cluster owns CR batch rows of one direction,
// CTA `rank` owns hidden units rank*30 .. rank*30+29 and keeps its slice of W_hh in shared memory for all T steps
// (forward: the 120 gate columns of its units, [240][120] fp32 = 115 KB; backward: the 30 columns of W_hh that feed its
// units' dL/dh, [960][30] fp32 = 115 KB), so a step touches HBM/L2 only for that step's gate rows.  What a step exchanges --
// the new h (forward), partial dL/dh blocks (backward) -- is written straight into the peers' shared memory (DSMEM), double
// buffered by step parity, then one cluster barrier.  fp32 FMA contractions; same arithmetic and layouts as lstm_train.cu (the generic
// version for other hidden sizes).  Replaces autograd over nn.LSTM in the reference's training step (main.py:67-82).
#include "model.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace hssb {
namespace {

constexpr int CH = 240;            // hidden size this file is specialised for
constexpr int CG4 = 4 * CH;        // 960 gate columns
constexpr int CCL = 8;             // CTAs per cluster
constexpr int CU = CH / CCL;       // 30 units per CTA
constexpr int CC = 4 * CU;         // 120 gate columns per CTA
constexpr int CR = 8;              // batch rows per cluster
constexpr int CTHREADS = 480;      // forward: 4 k-quarters x 120 columns; backward: 16 n-slices x 30 units
constexpr int FWD_KQ = 4, BWD_NQ = 16;

constexpr size_t FWD_SMEM = sizeof(float) * ((size_t)CH * CC + 2 * CH * CR + FWD_KQ * CR * CC);
constexpr size_t BWD_SMEM = sizeof(float) * ((size_t)CG4 * CU + 2 * CG4 * CR + BWD_NQ * CR * CU);
static_assert(FWD_SMEM <= 227 * 1024 && BWD_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void __cluster_dims__(CCL, 1, 1) __launch_bounds__(CTHREADS, 1)
train_fwd_cluster_kernel(float *__restrict__ gates, const float *__restrict__ w0T, const float *__restrict__ w1T,
                         const float *__restrict__ h0, const float *__restrict__ c0, long long B, long long T,
                         float *__restrict__ out, float *__restrict__ cells, float *__restrict__ hn, float *__restrict__ cn)
{
    extern __shared__ __align__(16) float sm[];
    float *w_s = sm;                               // [k 240][c 120], c = gate*30 + unit
    float *h_s = w_s + CH * CC;                    // [2][k 240][CR]   (all units of the cluster's rows, double buffered)
    float *part_s = h_s + 2 * CH * CR;             // [FWD_KQ][CR][CC]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const long long b0 = (long long)(blockIdx.x / CCL) * CR;
    const float *wT = dir ? w1T : w0T;             // [240][960] = W_hh^T
    float *gd = gates + (size_t)dir * B * T * CG4;
    float *cd = cells + (size_t)dir * B * T * CH;

    for (int i = tid; i < CH * CC; i += CTHREADS) {
        const int k = i / CC, c = i % CC;
        w_s[i] = wT[(size_t)k * CG4 + (c / CU) * CH + rank * CU + (c % CU)];
    }
    for (int i = tid; i < CH * CR; i += CTHREADS) {
        const int k = i / CR, r = i % CR;
        h_s[i] = (b0 + r < B) ? h0[((size_t)dir * B + b0 + r) * CH + k] : 0.f;
    }
    // elementwise role: thread (r, ul) for tid < 240
    const int er = tid / CU, eul = tid % CU;
    const bool elem = tid < CR * CU;
    const bool live = elem && (b0 + er < B);
    const int unit = rank * CU + eul;
    float c_reg = live ? c0[((size_t)dir * B + b0 + er) * CH + unit] : 0.f;
    float h_reg = 0.f;
    cluster.sync();                                 // every CTA of the cluster is running: its shared memory may be written

    // contraction role: thread (kq, c)
    const int kq = tid / CC, cc = tid % CC;
    // Global traffic stays off the barrier's critical path: a step's results are kept in registers and stored at the top of
    // the NEXT step, its projected input is loaded there too -- both complete under the contraction, so the release of the
    // cluster barrier only ever waits for the DSMEM stores (stores issued right before it cost ~2 us per step).
    float res[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // i, f, g, o, c, h of the previous step
    auto store_step = [&](long long step) {
        const long long t = dir ? (T - 1 - step) : step;
        const size_t row = (size_t)(b0 + er) * T + t;
        float *g = gd + row * CG4 + unit;
        __stcs(g, res[0]); __stcs(g + CH, res[1]); __stcs(g + 2 * CH, res[2]); __stcs(g + 3 * CH, res[3]);
        __stcs(cd + row * CH + unit, res[4]);
        __stcs(out + row * (2 * CH) + dir * CH + unit, res[5]);
    };
    for (long long step = 0; step < T; ++step) {
        const long long t = dir ? (T - 1 - step) : step;
        float pre[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            if (step > 0) store_step(step - 1);
            const float *g = gd + ((size_t)(b0 + er) * T + t) * CG4 + unit;
#pragma unroll
            for (int q = 0; q < 4; ++q) pre[q] = __ldcs(g + q * CH);
        }
        const float *hb = h_s + (step & 1) * CH * CR;
        float acc[CR];
#pragma unroll
        for (int r = 0; r < CR; ++r) acc[r] = 0.f;
#pragma unroll 4
        for (int k = kq * (CH / FWD_KQ); k < (kq + 1) * (CH / FWD_KQ); ++k) {
            const float w = w_s[k * CC + cc];
            const float4 ha = *reinterpret_cast<const float4 *>(hb + k * CR);
            const float4 hc = *reinterpret_cast<const float4 *>(hb + k * CR + 4);
            acc[0] = fmaf(ha.x, w, acc[0]); acc[1] = fmaf(ha.y, w, acc[1]); acc[2] = fmaf(ha.z, w, acc[2]); acc[3] = fmaf(ha.w, w, acc[3]);
            acc[4] = fmaf(hc.x, w, acc[4]); acc[5] = fmaf(hc.y, w, acc[5]); acc[6] = fmaf(hc.z, w, acc[6]); acc[7] = fmaf(hc.w, w, acc[7]);
        }
#pragma unroll
        for (int r = 0; r < CR; ++r) part_s[(kq * CR + r) * CC + cc] = acc[r];
        __syncthreads();
        if (elem) {
            float a[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float s = pre[q];
#pragma unroll
                for (int p = 0; p < FWD_KQ; ++p) s += part_s[(p * CR + er) * CC + q * CU + eul];
                a[q] = s;
            }
            const float ig = sigmoid_f(a[0]), fg = sigmoid_f(a[1]), gg = tanhf(a[2]), og = sigmoid_f(a[3]);
            c_reg = fg * c_reg + ig * gg;
            h_reg = og * tanhf(c_reg);
            res[0] = ig; res[1] = fg; res[2] = gg; res[3] = og; res[4] = c_reg; res[5] = h_reg;
            float *dst = h_s + ((step + 1) & 1) * CH * CR + unit * CR + er;
#pragma unroll
            for (int p = 0; p < CCL; ++p) *cluster.map_shared_rank(dst, p) = h_reg;
        }
        cluster.sync();                             // h of this step is in every CTA's buffer; part_s may be overwritten
    }
    if (live) {
        if (T > 0) store_step(T - 1);
        hn[((size_t)dir * B + b0 + er) * CH + unit] = T > 0 ? h_reg : h0[((size_t)dir * B + b0 + er) * CH + unit];
        cn[((size_t)dir * B + b0 + er) * CH + unit] = c_reg;
    }
}

__global__ void __cluster_dims__(CCL, 1, 1) __launch_bounds__(CTHREADS, 1)
train_bwd_cluster_kernel(float *__restrict__ gates, const float *__restrict__ cells, const float *__restrict__ w0,
                         const float *__restrict__ w1, const float *__restrict__ c0, const float *__restrict__ d_out,
                         const float *__restrict__ d_hn, const float *__restrict__ d_cn, long long B, long long T,
                         float *__restrict__ dh0, float *__restrict__ dc0)
{
    extern __shared__ __align__(16) float sm[];
    float *w_s = sm;                               // [n 960][j 30]: W_hh[n][rank*30 + j]
    float *dg_s = w_s + CG4 * CU;                  // [2][n 960][CR]  (dG of all units of the cluster's rows, double buffered)
    float *part_s = dg_s + 2 * CG4 * CR;           // [BWD_NQ][CR][CU]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const long long b0 = (long long)(blockIdx.x / CCL) * CR;
    const float *w = dir ? w1 : w0;                // [960][240] torch layout
    float *gd = gates + (size_t)dir * B * T * CG4;
    const float *cd = cells + (size_t)dir * B * T * CH;

    for (int i = tid; i < CG4 * CU; i += CTHREADS) {
        const int n = i / CU, j = i % CU;
        w_s[i] = w[(size_t)n * CH + rank * CU + j];
    }
    const int er = tid / CU, eul = tid % CU;       // elementwise role: thread (r, ul) for tid < 240
    const bool elem = tid < CR * CU;
    const bool live = elem && (b0 + er < B);
    const int unit = rank * CU + eul;
    float dh_rec = (live && d_hn) ? d_hn[((size_t)dir * B + b0 + er) * CH + unit] : 0.f;
    float dc_carry = (live && d_cn) ? d_cn[((size_t)dir * B + b0 + er) * CH + unit] : 0.f;
    cluster.sync();

    const int nq = tid / CU, jl = tid % CU;        // contraction role: thread (n-slice, unit)
    // this step's saved values, loaded one step ahead: i, f, g, o, c, c_prev, d_out
    float sv[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto load_step = [&](long long step) {
        const long long t = dir ? (T - 1 - step) : step;
        const long long t_prev = dir ? t + 1 : t - 1;
        const size_t row = (size_t)(b0 + er) * T + t;
        const float *g = gd + row * CG4 + unit;
        sv[0] = __ldcs(g); sv[1] = __ldcs(g + CH); sv[2] = __ldcs(g + 2 * CH); sv[3] = __ldcs(g + 3 * CH);
        sv[4] = __ldcs(cd + row * CH + unit);
        sv[5] = step > 0 ? __ldg(cd + ((size_t)(b0 + er) * T + t_prev) * CH + unit) : __ldg(c0 + ((size_t)dir * B + b0 + er) * CH + unit);
        sv[6] = __ldcs(d_out + row * (2 * CH) + dir * CH + unit);
    };
    if (live) load_step(T - 1);
    for (long long step = T - 1; step >= 0; --step) {
        const long long t = dir ? (T - 1 - step) : step;
        float *db = dg_s + (step & 1) * CG4 * CR;
        if (elem) {
            float da[4] = {0.f, 0.f, 0.f, 0.f};
            if (live) {
                const float ig = sv[0], fg = sv[1], gg = sv[2], og = sv[3], c = sv[4], c_prev = sv[5];
                const float tc = tanhf(c);
                const float dh = sv[6] + dh_rec;
                const float dc = dc_carry + dh * og * (1.f - tc * tc);
                da[0] = dc * gg * ig * (1.f - ig);
                da[1] = dc * c_prev * fg * (1.f - fg);
                da[2] = dc * ig * (1.f - gg * gg);
                da[3] = dh * tc * og * (1.f - og);
                dc_carry = dc * fg;
                float *g = gd + ((size_t)(b0 + er) * T + t) * CG4 + unit;
                __stcs(g, da[0]); __stcs(g + CH, da[1]); __stcs(g + 2 * CH, da[2]); __stcs(g + 3 * CH, da[3]);
                if (step > 0) load_step(step - 1);  // in flight during the exchange and the contraction
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float *dst = db + (size_t)(q * CH + unit) * CR + er;
#pragma unroll
                for (int p = 0; p < CCL; ++p) *cluster.map_shared_rank(dst, p) = da[q];
            }
        }
        cluster.sync();                             // dG of this step is in every CTA's buffer
        {
            float acc[CR];
#pragma unroll
            for (int r = 0; r < CR; ++r) acc[r] = 0.f;
#pragma unroll 4
            for (int n = nq * (CG4 / BWD_NQ); n < (nq + 1) * (CG4 / BWD_NQ); ++n) {
                const float wv = w_s[n * CU + jl];
                const float4 ga = *reinterpret_cast<const float4 *>(db + n * CR);
                const float4 gc = *reinterpret_cast<const float4 *>(db + n * CR + 4);
                acc[0] = fmaf(ga.x, wv, acc[0]); acc[1] = fmaf(ga.y, wv, acc[1]); acc[2] = fmaf(ga.z, wv, acc[2]); acc[3] = fmaf(ga.w, wv, acc[3]);
                acc[4] = fmaf(gc.x, wv, acc[4]); acc[5] = fmaf(gc.y, wv, acc[5]); acc[6] = fmaf(gc.z, wv, acc[6]); acc[7] = fmaf(gc.w, wv, acc[7]);
            }
#pragma unroll
            for (int r = 0; r < CR; ++r) part_s[(nq * CR + r) * CU + jl] = acc[r];
        }
        __syncthreads();
        if (elem) {
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < BWD_NQ; ++p) s += part_s[(p * CR + er) * CU + eul];
            dh_rec = s;
        }
        __syncthreads();                            // part_s is rewritten by the next step's contraction
    }
    if (live) {
        dh0[((size_t)dir * B + b0 + er) * CH + unit] = dh_rec;
        dc0[((size_t)dir * B + b0 + er) * CH + unit] = dc_carry;
    }
    cluster.sync();                                 // nobody leaves while a peer could still address this CTA's shared memory
}


// Backward, reduce-scatter form (the default; 15.4 ms against the all-gather form's 49 ms per step of batch 50 x 2000): CTA `rank` keeps the W_hh ROWS of its own units' gates ([120][240] fp32), turns its own
// dG (local, no exchange) into a partial dL/dh_prev for all 240 units, and sends each peer the 30 x CR block that peer owns:
// 7.7 KB leave the SM per step instead of the 30 KB of the all-gather form above (DSMEM egress is ~18 B/cycle/SM).
constexpr int RS_NQ = 2;
constexpr size_t BWD_RS_SMEM = sizeof(float) * ((size_t)CC * CH + CC * CR + RS_NQ * CR * CH + 2 * CCL * CR * CU);
static_assert(BWD_RS_SMEM <= 227 * 1024, "shared memory budget");

__global__ void __cluster_dims__(CCL, 1, 1) __launch_bounds__(CTHREADS, 1)
train_bwd_cluster_rs_kernel(float *__restrict__ gates, const float *__restrict__ cells, const float *__restrict__ w0,
                            const float *__restrict__ w1, const float *__restrict__ c0, const float *__restrict__ d_out,
                            const float *__restrict__ d_hn, const float *__restrict__ d_cn, long long B, long long T,
                            float *__restrict__ dh0, float *__restrict__ dc0)
{
    extern __shared__ __align__(16) float sm[];
    float *w_s = sm;                               // [nl 120][j 240]: W_hh[gate*240 + rank*30 + unit][j], nl = gate*30 + unit
    float *dgl_s = w_s + CC * CH;                  // [nl 120][CR]     this step's dG of my units
    float *part_s = dgl_s + CC * CR;               // [RS_NQ][CR][j 240]
    float *recv_s = part_s + RS_NQ * CR * CH;      // [2][src 8][CR][jl 30]  partial dL/dh_prev of my units from every CTA
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const long long b0 = (long long)(blockIdx.x / CCL) * CR;
    const float *w = dir ? w1 : w0;                // [960][240] torch layout
    float *gd = gates + (size_t)dir * B * T * CG4;
    const float *cd = cells + (size_t)dir * B * T * CH;

    for (int i = tid; i < CC * CH; i += CTHREADS) {
        const int nl = i / CH, j = i % CH;
        w_s[i] = w[(size_t)((nl / CU) * CH + rank * CU + (nl % CU)) * CH + j];
    }
    const int er = tid / CU, eul = tid % CU;       // elementwise role: thread (r, ul) for tid < 240
    const bool elem = tid < CR * CU;
    const bool live = elem && (b0 + er < B);
    const int unit = rank * CU + eul;
    float dh_rec = (live && d_hn) ? d_hn[((size_t)dir * B + b0 + er) * CH + unit] : 0.f;
    float dc_carry = (live && d_cn) ? d_cn[((size_t)dir * B + b0 + er) * CH + unit] : 0.f;
    cluster.sync();

    const int nq = tid / CH, cj = tid % CH;        // contraction role: thread (half of my 120 gate rows, unit j of all 240)
    float sv[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // this step's i, f, g, o, c, c_prev, d_out -- loaded one step ahead
    auto load_step = [&](long long step) {
        const long long t = dir ? (T - 1 - step) : step;
        const long long t_prev = dir ? t + 1 : t - 1;
        const size_t row = (size_t)(b0 + er) * T + t;
        const float *g = gd + row * CG4 + unit;
        sv[0] = __ldcs(g); sv[1] = __ldcs(g + CH); sv[2] = __ldcs(g + 2 * CH); sv[3] = __ldcs(g + 3 * CH);
        sv[4] = __ldcs(cd + row * CH + unit);
        sv[5] = step > 0 ? __ldg(cd + ((size_t)(b0 + er) * T + t_prev) * CH + unit) : __ldg(c0 + ((size_t)dir * B + b0 + er) * CH + unit);
        sv[6] = __ldcs(d_out + row * (2 * CH) + dir * CH + unit);
    };
    auto gather_dh = [&](long long step_done) {    // sum of the eight partials the step `step_done` sent me
        const float *rb = recv_s + (step_done & 1) * CCL * CR * CU + er * CU + eul;
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < CCL; ++p) s += rb[p * CR * CU];
        return s;
    };
    if (live) load_step(T - 1);
    for (long long step = T - 1; step >= 0; --step) {
        const long long t = dir ? (T - 1 - step) : step;
        if (elem) {
            if (step < T - 1) dh_rec = gather_dh(step + 1);
            float da[4] = {0.f, 0.f, 0.f, 0.f};
            if (live) {
                const float ig = sv[0], fg = sv[1], gg = sv[2], og = sv[3], c = sv[4], c_prev = sv[5];
                const float tc = tanhf(c);
                const float dh = sv[6] + dh_rec;
                const float dc = dc_carry + dh * og * (1.f - tc * tc);
                da[0] = dc * gg * ig * (1.f - ig);
                da[1] = dc * c_prev * fg * (1.f - fg);
                da[2] = dc * ig * (1.f - gg * gg);
                da[3] = dh * tc * og * (1.f - og);
                dc_carry = dc * fg;
                float *g = gd + ((size_t)(b0 + er) * T + t) * CG4 + unit;
                __stcs(g, da[0]); __stcs(g + CH, da[1]); __stcs(g + 2 * CH, da[2]); __stcs(g + 3 * CH, da[3]);
                if (step > 0) load_step(step - 1);  // in flight during the contraction and the exchange
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) dgl_s[(q * CU + eul) * CR + er] = da[q];
        }
        __syncthreads();
        {
            float acc[CR];
#pragma unroll
            for (int r = 0; r < CR; ++r) acc[r] = 0.f;
#pragma unroll 4
            for (int nl = nq * (CC / RS_NQ); nl < (nq + 1) * (CC / RS_NQ); ++nl) {
                const float wv = w_s[nl * CH + cj];
                const float4 ga = *reinterpret_cast<const float4 *>(dgl_s + nl * CR);
                const float4 gc = *reinterpret_cast<const float4 *>(dgl_s + nl * CR + 4);
                acc[0] = fmaf(ga.x, wv, acc[0]); acc[1] = fmaf(ga.y, wv, acc[1]); acc[2] = fmaf(ga.z, wv, acc[2]); acc[3] = fmaf(ga.w, wv, acc[3]);
                acc[4] = fmaf(gc.x, wv, acc[4]); acc[5] = fmaf(gc.y, wv, acc[5]); acc[6] = fmaf(gc.z, wv, acc[6]); acc[7] = fmaf(gc.w, wv, acc[7]);
            }
#pragma unroll
            for (int r = 0; r < CR; ++r) part_s[(nq * CR + r) * CH + cj] = acc[r];
        }
        __syncthreads();
        for (int i = tid; i < CR * CH; i += CTHREADS) {         // (r, j): both halves summed, sent to the CTA that owns unit j
            const int r = i / CH, j = i % CH;
            const float v = part_s[r * CH + j] + part_s[(CR + r) * CH + j];
            float *dst = recv_s + (step & 1) * CCL * CR * CU + (rank * CR + r) * CU + (j % CU);
            *cluster.map_shared_rank(dst, j / CU) = v;
        }
        cluster.sync();                             // every partial of this step has landed
    }
    if (live) {
        dh0[((size_t)dir * B + b0 + er) * CH + unit] = gather_dh(0);
        dc0[((size_t)dir * B + b0 + er) * CH + unit] = dc_carry;
    }
}

}  // namespace

bool train_cluster_supported(int H) { return H == CH; }

int train_fwd_cluster_launch(float *gates, const float *w0T, const float *w1T, const float *h0, const float *c0, int64_t B, int64_t T,
                             float *out, float *cells, float *hn, float *cn, cudaStream_t st)
{
    HSSB_CUDA_OK(cudaFuncSetAttribute(train_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    dim3 grid((unsigned)((B + CR - 1) / CR) * CCL, 2);
    ProfScope prof("lstm_train_fwd", st);
    train_fwd_cluster_kernel<<<grid, CTHREADS, FWD_SMEM, st>>>(gates, w0T, w1T, h0, c0, B, T, out, cells, hn, cn);
    HSSB_LAUNCH_OK("train_fwd_cluster_kernel");
    return 0;
}

int train_bwd_cluster_launch(float *gates, const float *cells, const float *w0, const float *w1, const float *c0, const float *d_out,
                             const float *d_hn, const float *d_cn, int64_t B, int64_t T, float *dh0, float *dc0, bool all_gather,
                             cudaStream_t st)
{
    dim3 grid((unsigned)((B + CR - 1) / CR) * CCL, 2);
    ProfScope prof("lstm_train_bwd", st);
    if (all_gather) {
        HSSB_CUDA_OK(cudaFuncSetAttribute(train_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        train_bwd_cluster_kernel<<<grid, CTHREADS, BWD_SMEM, st>>>(gates, cells, w0, w1, c0, d_out, d_hn, d_cn, B, T, dh0, dc0);
    } else {
        HSSB_CUDA_OK(cudaFuncSetAttribute(train_bwd_cluster_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_RS_SMEM));
        train_bwd_cluster_rs_kernel<<<grid, CTHREADS, BWD_RS_SMEM, st>>>(gates, cells, w0, w1, c0, d_out, d_hn, d_cn, B, T, dh0, dc0);
    }
    HSSB_LAUNCH_OK("train_bwd_cluster_kernel");
    return 0;
}

}  // namespace hssb
