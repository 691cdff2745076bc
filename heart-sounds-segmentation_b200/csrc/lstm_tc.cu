// tcgen05 / TMEM / TMA kernels of the BiLSTM segmenter (hidden_size = 240).
//
// Precision: every gate contraction runs as a split-fp16 "3-pass" product on the 5th-gen tensor
// cores with fp32 accumulation in TMEM:  x = hi + lo (hi = fp16(x), lo = fp16(x - hi), 22 mantissa
// bits together), and  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (the dropped lo*lo term is 2^-22
// relative).  That keeps log-probabilities within ~1e-6 of the fp32 reference, which is what the
// bit-identical-labels requirement needs, at 3 fp16 MMAs per product.
//
// "Cluster gate order": the 960 gate rows of one direction are permuted to g' = r*120 + q*30 + u
// (r = CTA rank in the 8-CTA recurrence cluster, q = gate i/f/g/o, u = unit 0..29 of that rank;
// torch row = q*240 + 30*r + u), so that every recurrence CTA owns one contiguous 120-wide slice.
//
// K4  tc_inproj_kernel : xproj[dir][t][b][g'] = A[b,t,:] . W_ih[g',:] + (b_ih + b_hh)[g']
//     128(t) x 192(g') output tile per CTA, K blocks of 64 through a 2-stage TMA ring (SW128),
//     accumulators in TMEM, epilogue TMEM -> regs -> swizzled smem -> TMA store.
#include "model.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <mutex>

namespace hssb {

using namespace ptx;

constexpr int TC_H = 240;
constexpr int TC_G = 960;          // gate rows per direction
constexpr int TC_NG = 2 * TC_G;    // both directions

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
__global__ void pack_whh_kernel(const float *__restrict__ w, int dir, __half *__restrict__ dst);   // defined with K5
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
__global__ void pack_wih_kernel(const float *__restrict__ w, const float *__restrict__ b_ih, const float *__restrict__ b_hh, int Kin,
                                int Kp, int dir, __half *__restrict__ hi, __half *__restrict__ lo, float *__restrict__ bias)
{
    const int gp = blockIdx.x;                     // g' = r*120 + q*30 + u
    const int r = gp / 120, q = (gp % 120) / 30, u = gp % 30;
    const int row = q * TC_H + 30 * r + u;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        __half h = __float2half_rn(0.f), l = h;
        if (k < Kin) split_f16(w[(size_t)row * Kin + k], h, l);
        hi[((size_t)dir * TC_G + gp) * Kp + k] = h;
        lo[((size_t)dir * TC_G + gp) * Kp + k] = l;
    }
    if (threadIdx.x == 0) bias[dir * TC_G + gp] = b_ih[row] + b_hh[row];
}

// ------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------
constexpr int IP_BM = 128, IP_BN = 192, IP_BK = 64, IP_STAGES = 2;
constexpr int IP_A_BYTES = IP_BM * IP_BK * 2;            // one fp16 plane of the A stage (16 KB)
constexpr int IP_B_BYTES = IP_BN * IP_BK * 2;            // one fp16 plane of the B stage (24 KB)
constexpr int IP_STAGE_BYTES = 2 * IP_A_BYTES + 2 * IP_B_BYTES;   // 80 KB
constexpr int IP_OUT_BYTES = IP_BM * 32 * 4;             // epilogue staging tile 128 x 32 fp32 (16 KB)
constexpr int IP_SMEM_BYTES = IP_STAGES * IP_STAGE_BYTES + 2 * IP_OUT_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int IP_TMEM_COLS = 256;

struct InprojParams {
    CUtensorMap a_hi, a_lo;   // [k, t, b] fp16, box (64,128,1), SW128
    CUtensorMap w_hi, w_lo;   // [k, g'(1920)] fp16, box (64,192), SW128
    CUtensorMap out;          // [g'(960), b, t, dir] fp32, box (32,1,128,1), SW128
    const float *bias;        // [1920]
    int k_real;               // true K (44->64 padded planes use 64; 480)
    int t_tiles;              // ceil(T/128)
};

__global__ void __launch_bounds__(192, 1) tc_inproj_kernel(const __grid_constant__ InprojParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    unsigned char *out_base = smem + IP_STAGES * IP_STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_base + 2 * IP_OUT_BYTES);
    uint64_t *full = bars, *empty = bars + IP_STAGES, *tmem_full = bars + 2 * IP_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * IP_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.t_tiles;
    const int t0 = (blockIdx.x % p.t_tiles) * IP_BM;
    const int n0 = blockIdx.y * IP_BN;              // over 1920 = both directions
    const int kblocks = (p.k_real + IP_BK - 1) / IP_BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.a_hi); prefetch_tmap(&p.a_lo); prefetch_tmap(&p.w_hi); prefetch_tmap(&p.w_lo); prefetch_tmap(&p.out);
        for (int s = 0; s < IP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<IP_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % IP_STAGES;
                const uint32_t ph = (kb / IP_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char *st = stage_base + s * IP_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], IP_STAGE_BYTES);
                tma_load_3d(st, &p.a_hi, &full[s], kb * IP_BK, t0, b);
                tma_load_3d(st + IP_A_BYTES, &p.a_lo, &full[s], kb * IP_BK, t0, b);
                tma_load_2d(st + 2 * IP_A_BYTES, &p.w_hi, &full[s], kb * IP_BK, n0);
                tma_load_2d(st + 2 * IP_A_BYTES + IP_B_BYTES, &p.w_lo, &full[s], kb * IP_BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(IP_BM, IP_BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % IP_STAGES;
                const uint32_t ph = (kb / IP_STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage_base + s * IP_STAGE_BYTES);
                const uint32_t a_lo = a_hi + IP_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * IP_A_BYTES;
                const uint32_t b_lo = b_hi + IP_B_BYTES;
                const int ksteps = min(IP_BK, p.k_real - kb * IP_BK + 15) / 16;   // skip all-padding K16 steps
                for (int ks = 0; ks < ksteps && ks < IP_BK / 16; ++ks) {
                    const uint32_t off = ks * 32;   // 16 fp16 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t da_hi = make_smem_desc(a_hi + off, 16, 1024, LAYOUT_SW128);
                    const uint64_t da_lo = make_smem_desc(a_lo + off, 16, 1024, LAYOUT_SW128);
                    const uint64_t db_hi = make_smem_desc(b_hi + off, 16, 1024, LAYOUT_SW128);
                    const uint64_t db_lo = make_smem_desc(b_lo + off, 16, 1024, LAYOUT_SW128);
                    mma_f16_ss(tmem_base, da_hi, db_hi, idesc, (kb | ks) != 0);
                    mma_f16_ss(tmem_base, da_lo, db_hi, idesc, 1);
                    mma_f16_ss(tmem_base, da_hi, db_lo, idesc, 1);
                }
                mma_commit(&empty[s]);          // frees the smem stage when these MMAs retire
            }
            mma_commit(tmem_full);              // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;          // tile row = time index t0 + row
        const int et = threadIdx.x - 64;        // 0..127
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int dir = n0 / TC_G, nl0 = n0 % TC_G;
        for (int c = 0; c < IP_BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
            tmem_ld_wait();
            unsigned char *ob = out_base + (c & 1) * IP_OUT_BYTES;
            if (c >= 2 && et == 0) tma_store_wait_read<1>();     // the store that last used this buffer has read it
            named_barrier(1, 128);
            const float *bias = p.bias + n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = __uint_as_float(v[4 * j + 0]) + __ldg(bias + 4 * j + 0);
                o.y = __uint_as_float(v[4 * j + 1]) + __ldg(bias + 4 * j + 1);
                o.z = __uint_as_float(v[4 * j + 2]) + __ldg(bias + 4 * j + 2);
                o.w = __uint_as_float(v[4 * j + 3]) + __ldg(bias + 4 * j + 3);
                *reinterpret_cast<float4 *>(ob + row * 128 + ((j ^ (row & 7)) << 4)) = o;   // 128B swizzle
            }
            fence_proxy_async_smem();
            named_barrier(1, 128);
            if (et == 0) {
                tma_store_4d(&p.out, ob, nl0 + c * 32, b, t0, dir);
                tma_store_commit();
            }
        }
        if (et == 0) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<IP_TMEM_COLS>(tmem_base);
}

// xproj[dir][t][b][g'] -> canonical [dir][b*T + t][q*240 + unit]   (debug / validation only)
__global__ void unpermute_xproj_kernel(const float *__restrict__ src, long long B, long long T, float *__restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 2 * B * T * TC_G;
    if (i >= total) return;
    const int gp = (int)(i % TC_G);
    const long long rest = i / TC_G;
    const long long b = rest % B, t = (rest / B) % T, dir = rest / (B * T);
    const int r = gp / 120, q = (gp % 120) / 30, u = gp % 30;
    dst[((size_t)dir * B * T + b * T + t) * TC_G + q * TC_H + 30 * r + u] = src[i];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int kp_of_layer(int layer, int F) { return layer == 0 ? 64 : 512; }
static int kreal_of_layer(int layer, int F) { return layer == 0 ? ((F + 15) / 16) * 16 : 2 * TC_H; }

size_t tc_pack_bytes(int F, int H)
{
    if (H != TC_H || F > 64) return 0;
    size_t n = 0;
    for (int l = 0; l < 2; ++l) {
        n += align_up(sizeof(__half) * 2 * TC_NG * kp_of_layer(l, F), 256);   // wih hi+lo
        n += align_up(sizeof(float) * TC_NG, 256);                             // bias
        n += align_up(sizeof(__half) * 2 * 8 * 2 * 128 * 256, 256);            // whh planes [dir][rank][plane][128][256]
    }
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
    for (int l = 0; l < 2 && !rc; ++l) {
        const int Kp = kp_of_layer(l, m->F);
        m->tc_wih[l] = reinterpret_cast<__half *>(base + off);
        off += align_up(sizeof(__half) * 2 * TC_NG * Kp, 256);
        m->tc_bias[l] = reinterpret_cast<float *>(base + off);
        off += align_up(sizeof(float) * TC_NG, 256);
        m->tc_whh[l] = reinterpret_cast<__half *>(base + off);
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
            pack_wih_kernel<<<TC_G, 128, 0, st>>>(w, bi, bh, kin[l], Kp, d, hi, lo, m->tc_bias[l]);
            if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "pack_wih_kernel"); break; }
            if ((e = cudaMemcpyAsync(w, p->w_hh[l][d], sizeof(float) * TC_G * TC_H, cudaMemcpyDefault, st)) != cudaSuccess) {
                rc = cuda_fail(e, "cudaMemcpyAsync(tc_pack w_hh)");
                break;
            }
            pack_whh_kernel<<<8 * 128, 128, 0, st>>>(w, d, m->tc_whh[l]);
            if ((e = cudaGetLastError()) != cudaSuccess) rc = cuda_fail(e, "pack_whh_kernel");
        }
    }
    cudaFreeAsync(tmp, st);
    if (rc) return rc;
    m->tc_ready = true;
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
        const uint32_t box[3] = {IP_BK, IP_BM, 1};
        if (int rc = make_tmap(&prm.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        if (int rc = make_tmap(&prm.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)TC_NG};
        const uint64_t strides[1] = {(uint64_t)Kp * 2};
        const uint32_t box[2] = {IP_BK, IP_BN};
        const __half *hi = m->tc_wih[layer], *lo = hi + (size_t)TC_NG * Kp;
        if (int rc = make_tmap(&prm.w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        if (int rc = make_tmap(&prm.w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)TC_G, (uint64_t)B, (uint64_t)T, 2};
        const uint64_t strides[3] = {(uint64_t)TC_G * 4, (uint64_t)B * TC_G * 4, (uint64_t)T * B * TC_G * 4};
        const uint32_t box[4] = {32, 1, IP_BM, 1};
        if (int rc = make_tmap(&prm.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, xproj, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    prm.bias = m->tc_bias[layer];
    prm.k_real = kreal;
    prm.t_tiles = (int)((T + IP_BM - 1) / IP_BM);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(tc_inproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IP_SMEM_BYTES); });
    dim3 grid((unsigned)(B * prm.t_tiles), TC_NG / IP_BN);
    ProfScope prof("tc_inproj", st);
    tc_inproj_kernel<<<grid, 192, IP_SMEM_BYTES, st>>>(prm);
    HSSB_LAUNCH_OK("tc_inproj_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K5: recurrence.  One 8-CTA cluster per (direction, group of S*NB batch columns).
//
// Orientation: gates are the MMA M dimension and stay put, the batch is N:
//     G^T[g' (128 lanes), b (NB cols)] = W_hh,slice[g', k] . h_{t-1}^T[k, b]  (+ xproj^T added in the epilogue)
// CTA rank r owns units 30r..30r+29 -> gate rows q*32+u (q = i,f,g,o; u < 30; rows 30,31 of every
// quadrant are zero padding) so every gate of a unit sits in one warp's TMEM lane quadrant.
//   * W_hh slice (hi and lo fp16 planes, K padded 240 -> 8*32) is loaded ONCE into TMEM columns
//     [0,256) and is the A operand of every MMA (tcgen05.mma with A in TMEM) -- weights never move.
//   * h_{t-1}^T lives in shared memory as the B operand (K-major, no swizzle, [rank][plane][k-chunk][b][8]).
//     After its epilogue each CTA owns 30 fresh h values per batch column; it writes them as an fp16
//     hi/lo "image" and its producer thread pushes that image into the B buffer of all 8 CTAs with
//     cp.async.bulk shared::cta -> shared::cluster, completing on the receiver's mbarrier
//     (the all-gather of the recurrence, no global memory round trip, no cluster barrier).
//   * Epilogue per step: tcgen05.ld the 128 x NB accumulator, add xproj (TMA-prefetched tile),
//     sigmoid/tanh per quadrant, exchange the activated gates through smem, then c/h update with the
//     cell state in registers.
// Sub-tiles: S independent groups of NB batch columns are interleaved per cluster so that the tensor
// pipe (sub-tile A's MMAs) overlaps the MUFU/LSU work of sub-tile B's epilogue and the DSMEM hops.
// ------------------------------------------------------------------------------------------------
constexpr int RC_CL = 8;            // CTAs per cluster
constexpr int RC_U = 30;            // real units per CTA
constexpr int RC_KP = 256;          // padded K (8 ranks x 32 slots)
constexpr int RC_XW = 4 * RC_U;     // xproj floats per (t, b) owned by one CTA (120)

template <int NB, int S>
struct RcCfg {
    static constexpr int HBUF_BYTES = NB * RC_KP * 2 * 2;       // one B-operand buffer: hi+lo planes (NB KB)
    static constexpr int SLICE_BYTES = NB * 32 * 2 * 2;         // one rank's image: [plane][4 chunks][NB][8] fp16
    static constexpr int XP_BYTES = NB * RC_XW * 4;
    static constexpr int GBUF_BYTES = 4 * NB * 32 * 4;
    static constexpr int PER_SUB = 2 * HBUF_BYTES + 2 * SLICE_BYTES + XP_BYTES + GBUF_BYTES;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = S * PER_SUB + BAR_BYTES + 1024;
    static constexpr int THREADS = 32 * (1 + S) + 128 * S;
    static_assert(NB % 16 == 0 && NB <= 64, "NB must be 16, 32, 48 or 64");
    static_assert(S * NB <= 256, "accumulators must fit in the TMEM columns left of the weights");
};

struct RecurParams {
    CUtensorMap xproj;          // [g'(960), b, t, dir] fp32, box (120, NB, 1, 1), no swizzle
    const __half *whh;          // [dir][rank][plane][128][256] fp16 (cluster gate order, zero padded)
    const float *h0, *c0;       // [2][B][240]
    float *hn, *cn;             // [2][B][240]  raw final state
    __half *out_hi, *out_lo;    // layer 1: relu(h) planes [B*T][480]  (nullptr for layer 2)
    float *out_f32;             // layer 2: relu(h) [B*T][480]         (nullptr for layer 1)
    long long B, T;
    int b_base;                 // first batch column handled by this launch
};

__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }

template <int NB, int S>
__global__ void __launch_bounds__(RcCfg<NB, S>::THREADS, 1) tc_recurrent_kernel(const __grid_constant__ RecurParams p)
{
    using C = RcCfg<NB, S>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    // per sub-tile regions
    auto hbuf = [&](int s, int par) { return smem + s * C::PER_SUB + par * C::HBUF_BYTES; };
    auto image = [&](int s, int par) { return smem + s * C::PER_SUB + 2 * C::HBUF_BYTES + par * C::SLICE_BYTES; };
    auto xpst = [&](int s) { return reinterpret_cast<float *>(smem + s * C::PER_SUB + 2 * C::HBUF_BYTES + 2 * C::SLICE_BYTES); };
    auto gbuf = [&](int s) { return reinterpret_cast<float *>(smem + s * C::PER_SUB + 2 * C::HBUF_BYTES + 2 * C::SLICE_BYTES + C::XP_BYTES); };
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * C::PER_SUB);
    uint64_t *h_full = bars;                 // [S][2]
    uint64_t *d_full = bars + 2 * S;         // [S]
    uint64_t *xp_full = bars + 3 * S;        // [S]
    uint64_t *xp_empty = bars + 4 * S;       // [S]
    uint64_t *img_ready = bars + 5 * S;      // [S]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6 * S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x / RC_CL;
    const int dir = cid & 1;
    const int group = cid >> 1;
    const long long T = p.T, B = p.B;

    if (threadIdx.x == 0) {
        prefetch_tmap(&p.xproj);
        for (int s = 0; s < S; ++s) {
            mbar_init(&h_full[2 * s], 1); mbar_init(&h_full[2 * s + 1], 1);
            mbar_init(&d_full[s], 1); mbar_init(&xp_full[s], 1); mbar_init(&xp_empty[s], 4); mbar_init(&img_ready[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    // zero the images (padding slots u = 30, 31 must be finite zeros forever)
    for (int i = threadIdx.x; i < S * C::PER_SUB / 16; i += C::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync();     // every CTA's barriers are initialised before any remote copy can target them

    constexpr int EPI0 = 1 + S;   // first epilogue warp
    if (warp == 0) {
        // ================= MMA issuer =================
        // wait until the weights are in TMEM (loaded by epilogue group 0, signalled through a named barrier)
        named_barrier(8, 32 + 128);
        tc_fence_after();
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, NB);
            for (long long t = 0; t < T; ++t) {
                for (int s = 0; s < S; ++s) {
                    const int par = (int)(t & 1);
                    mbar_wait_cluster(&h_full[2 * s + par], (uint32_t)((t >> 1) & 1));
                    tc_fence_after();
                    const uint32_t hb = smem_u32(hbuf(s, par));
                    const uint32_t d_tmem = tmem_base + 256 + s * NB;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t blk = hb + (j >> 1) * (NB * 128) + (j & 1) * (NB * 32);
                        const uint64_t b_hi = make_smem_desc(blk, NB * 16, 128, LAYOUT_NONE);
                        const uint64_t b_lo = make_smem_desc(blk + NB * 64, NB * 16, 128, LAYOUT_NONE);
                        const uint32_t a_hi = tmem_base + j * 8, a_lo = tmem_base + 128 + j * 8;
                        mma_f16_ts(d_tmem, a_hi, b_hi, idesc, j != 0);
                        mma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1);
                        mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                    mma_commit(&d_full[s]);
                }
            }
        }
    } else if (warp < EPI0) {
        // ================= producer for sub-tile s: xproj TMA prefetch + all-gather pushes =================
        const int s = warp - 1;
        if (lane == 0) {
            const int b0 = p.b_base + (group * S + s) * NB;
            for (long long t = -1; t + 1 < T; ++t) {
                const long long n = t + 1;                         // the step being prepared
                const int tt = (int)(dir ? (T - 1 - n) : n);
                mbar_wait(&xp_empty[s], (uint32_t)((n & 1) ^ 1));
                mbar_arrive_expect_tx(&xp_full[s], C::XP_BYTES);
                tma_load_4d(xpst(s), &p.xproj, &xp_full[s], rank * RC_XW, b0, tt, dir);
                const int par = (int)(n & 1);                      // buffer that receives h_t for step n
                mbar_arrive_expect_tx(&h_full[2 * s + par], C::HBUF_BYTES);
                mbar_wait(&img_ready[s], (uint32_t)(n & 1));
                const unsigned char *img = image(s, (int)(t & 1));
                unsigned char *dst = hbuf(s, par) + rank * C::SLICE_BYTES;
#pragma unroll
                for (uint32_t j = 0; j < RC_CL; ++j) bulk_copy_to_cta(dst, img, C::SLICE_BYTES, &h_full[2 * s + par], j);
            }
        }
    } else {
        // ================= epilogue group s =================
        const int s = (warp - EPI0) >> 2;
        const int q = warp & 3;                  // TMEM lane quadrant == gate (i, f, g, o)
        const int u = lane;
        const bool unit_ok = u < RC_U;
        const int b0 = p.b_base + (group * S + s) * NB;
        const int hcol = dir * TC_H + (int)rank * RC_U + u;      // column in the [.., 480] outputs

        if (s == 0) {
            // one-time: W_hh slice -> TMEM.  Thread (q,u) owns lane q*32+u; column c holds k' = 2c, 2c+1.
            const __half *wrow = p.whh + ((((size_t)dir * RC_CL + rank) * 2) * 128 + (q * 32 + u)) * RC_KP;
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
            named_barrier(8, 32 + 128);
        }

        float c_state[NB / 4];
        float *gb = gbuf(s);
        const float *xp = xpst(s);
        for (long long t = -1; t < T; ++t) {
            const int tt = (int)(dir ? (T - 1 - t) : t);
            if (t >= 0) {
                mbar_wait(&d_full[s], (uint32_t)(t & 1));
                tc_fence_after();
                mbar_wait(&xp_full[s], (uint32_t)(t & 1));
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + s * NB;
#pragma unroll
                for (int c16 = 0; c16 < NB / 16; ++c16) {
                    uint32_t v[16];
                    tmem_ld_x16(taddr + c16 * 16, v);
                    tmem_ld_wait();
                    if (unit_ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int b = c16 * 16 + i;
                            const float pre = __uint_as_float(v[i]) + xp[b * RC_XW + q * RC_U + u];
                            gb[(q * NB + b) * 32 + u] = (q == 2) ? fast_tanh(pre) : fast_sigmoid(pre);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xp_empty[s]);
                tc_fence_before();
                named_barrier(1 + s, 128);
            }
            // ---- phase 2: thread (q,u) updates unit u for batch columns b = q, q+4, ... ----
            __half *img_hi = reinterpret_cast<__half *>(image(s, (int)(t & 1)));
            __half *img_lo = img_hi + NB * 32;
#pragma unroll
            for (int i = 0; i < NB / 4; ++i) {
                const int b = q + 4 * i;
                const long long bg = b0 + b;
                float h = 0.f;
                if (t < 0) {
                    const bool ok = unit_ok && bg < B;
                    h = ok ? __ldg(p.h0 + ((size_t)dir * B + bg) * TC_H + rank * RC_U + u) : 0.f;
                    c_state[i] = ok ? __ldg(p.c0 + ((size_t)dir * B + bg) * TC_H + rank * RC_U + u) : 0.f;
                } else if (unit_ok) {
                    const float ig = gb[(0 * NB + b) * 32 + u], fg = gb[(1 * NB + b) * 32 + u];
                    const float gg = gb[(2 * NB + b) * 32 + u], og = gb[(3 * NB + b) * 32 + u];
                    const float c = fmaf(fg, c_state[i], ig * gg);
                    c_state[i] = c;
                    h = og * fast_tanh(c);
                    if (bg < B) {
                        const size_t o = ((size_t)bg * T + tt) * (2 * TC_H) + hcol;
                        const float hr = fmaxf(h, 0.f);
                        if (p.out_f32) p.out_f32[o] = hr;
                        else { __half hh, hl; split_f16(hr, hh, hl); p.out_hi[o] = hh; p.out_lo[o] = hl; }
                        if (t == T - 1) {
                            p.hn[((size_t)dir * B + bg) * TC_H + rank * RC_U + u] = h;
                            p.cn[((size_t)dir * B + bg) * TC_H + rank * RC_U + u] = c;
                        }
                    }
                }
                if (unit_ok) {
                    __half hh, hl;
                    split_f16(h, hh, hl);
                    const int off = (u >> 3) * (NB * 8) + b * 8 + (u & 7);
                    img_hi[off] = hh;
                    img_lo[off] = hl;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&img_ready[s]);
            // gbuf is rewritten in the next step's phase 1 only after d_full, i.e. after every thread's
            // phase-2 reads of this step (the image of this step gates the next MMA).
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// torch W_hh[960][240] -> planes [dir][rank][plane][128 rows q*32+u][256 k' = 32 r' + u']
__global__ void pack_whh_kernel(const float *__restrict__ w, int dir, __half *__restrict__ dst)
{
    const int rank = blockIdx.x / 128, row = blockIdx.x % 128;
    const int q = row / 32, u = row % 32;
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

template <int NB, int S>
static int launch_recurrent(const RecurParams &prm_in, int groups, float *xproj, cudaStream_t st)
{
    using C = RcCfg<NB, S>;
    RecurParams prm = prm_in;
    {
        const uint64_t dims[4] = {(uint64_t)TC_G, (uint64_t)prm.B, (uint64_t)prm.T, 2};
        const uint64_t strides[3] = {(uint64_t)TC_G * 4, (uint64_t)prm.B * TC_G * 4, (uint64_t)prm.T * prm.B * TC_G * 4};
        const uint32_t box[4] = {RC_XW, NB, 1, 1};
        if (int rc = make_tmap(&prm.xproj, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, xproj, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(tc_recurrent_kernel<NB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_recurrent_kernel)");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * groups * RC_CL));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RC_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof("tc_recurrent", st);
    e = cudaLaunchKernelEx(&cfg, tc_recurrent_kernel<NB, S>, prm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_recurrent_kernel)");
    return 0;
}

// One layer's recurrence for batch columns [0, B): picks the sub-tile geometry from B.
static int tc_recurrent(const hssb_model *m, int layer, float *xproj, const float *h0, const float *c0, float *hn, float *cn,
                        __half *out_hi, __half *out_lo, float *out_f32, int64_t B, int64_t T, cudaStream_t st)
{
    RecurParams prm = {};
    prm.whh = m->tc_whh[layer];
    prm.h0 = h0; prm.c0 = c0; prm.hn = hn; prm.cn = cn;
    prm.out_hi = out_hi; prm.out_lo = out_lo; prm.out_f32 = out_f32;
    prm.B = B; prm.T = T;
    // at most 8 groups per direction are co-resident (16 clusters of 8 CTAs on 148 SMs); larger batches
    // run as successive launches over blocks of batch columns
    for (int64_t base = 0; base < B;) {
        const int64_t rem = B - base;
        prm.b_base = (int)base;
        int rc;
        if (rem <= 8 * 16) { const int g = (int)((rem + 15) / 16); rc = launch_recurrent<16, 1>(prm, g, xproj, st); base += (int64_t)g * 16; }
        else if (rem <= 8 * 32) { const int g = (int)((rem + 31) / 32); rc = launch_recurrent<32, 1>(prm, g, xproj, st); base += (int64_t)g * 32; }
        else { const int g = (int)std::min<int64_t>(8, (rem + 63) / 64); rc = launch_recurrent<32, 2>(prm, g, xproj, st); base += (int64_t)g * 64; }
        if (rc) return rc;
    }
    return 0;
}

namespace {
struct TcWs { size_t xhi, xlo, xproj, o1hi, o1lo, out2, hn, cn, total; };
TcWs tc_ws_layout(int64_t B, int64_t T)
{
    const size_t M = (size_t)B * T;
    TcWs w{};
    size_t off = 0;
    w.xhi = off;   off += align_up(sizeof(__half) * M * 64, 1024);
    w.xlo = off;   off += align_up(sizeof(__half) * M * 64, 1024);
    w.xproj = off; off += align_up(sizeof(float) * 2 * M * TC_G, 1024);
    w.o1hi = off;  off += align_up(sizeof(__half) * M * 2 * TC_H, 1024);
    w.o1lo = off;  off += align_up(sizeof(__half) * M * 2 * TC_H, 1024);
    w.out2 = off;  off += align_up(sizeof(float) * M * 2 * TC_H, 1024);
    w.hn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
    w.cn = off;    off += align_up(sizeof(float) * 2 * B * TC_H, 1024);
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
    const int64_t M = B * T;
    {
        ProfScope prof("split_planes", st);
        split_planes_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(x, M, m->F, 64, xhi, xlo);
        HSSB_LAUNCH_OK("split_planes_kernel");
    }
    if (int rc = tc_inproj(m, 0, xhi, xlo, 64, B, T, xproj, st)) return rc;
    if (int rc = tc_recurrent(m, 0, xproj, h0, c0, hn, cn, o1hi, o1lo, nullptr, B, T, st)) return rc;
    if (int rc = tc_inproj(m, 1, o1hi, o1lo, 2 * TC_H, B, T, xproj, st)) return rc;
    if (int rc = tc_recurrent(m, 1, xproj, hn, cn, hn, cn, nullptr, nullptr, out2, B, T, st)) return rc;
    return head_forward(out2, M, 2 * TC_H, m->lin_w, m->lin_b, logp, labels, st);
}

}  // namespace hssb

// ------------------------------------------------------------------------------------------------
// Diagnostic entry point: layer-1 input projection only, canonical layout, for kernel-level parity
// tests (impl 0 = tcgen05 kernel, 1 = SIMT kernel).  xproj: [2][B*T][960] fp32 (torch gate order).
// workspace: 2*B*T*960*4 + 2*B*T*64*2*2 bytes.
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
    const size_t need = sizeof(float) * 2 * M * TC_G + sizeof(__half) * 2 * M * 64;
    if (!workspace || workspace_bytes < need) return fail(HSSB_E_WORKSPACE, "hssb_debug_inproj: workspace %zu < %zu", workspace_bytes, need);
    float *raw = static_cast<float *>(workspace);
    __half *hi = reinterpret_cast<__half *>(raw + 2 * M * TC_G), *lo = hi + M * 64;
    split_planes_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(x, M, m->F, 64, hi, lo);
    HSSB_LAUNCH_OK("split_planes_kernel");
    if (int rc = tc_inproj(m, 0, hi, lo, 64, B, T, raw, st)) return rc;
    unpermute_xproj_kernel<<<(unsigned)((2 * M * TC_G + 255) / 256), 256, 0, st>>>(raw, B, T, xproj);
    HSSB_LAUNCH_OK("unpermute_xproj_kernel");
    return 0;
}
