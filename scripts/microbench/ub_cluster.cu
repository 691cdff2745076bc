// Micro-benchmarks behind the recurrence kernel's design (B200, sm_100a).  One mode per process:
//   ub_cluster a2a  : all-gather step by cp.async.bulk shared::cta -> shared::cluster (DSMEM), the recurrence's protocol
//                     (S independent chains, double-buffered by step parity, round r+1 sent after round r arrived)
//   ub_cluster st   : the same all-gather by st.shared::cluster.v4 + remote mbarrier arrives
//   ub_cluster l2mc : the same through L2: bulk store smem -> global, wait, multicast bulk load global -> every CTA
//   ub_cluster ping : one-way latency of a DSMEM bulk copy
//   ub_cluster mma  : rate of tcgen05.mma kind::f16, A in TMEM, B in smem (no swizzle), cta_group 1 / 2
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I heart-sounds-segmentation_b200/csrc -o build/ub_cluster scripts/microbench/ub_cluster.cu
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

using namespace hssb::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CL = 8;
constexpr int MAXS = 4;
constexpr int SLOT = 4096;                     // bytes reserved per (chain, parity, source)

struct P { int size, fan, S, iters, mode, warps; long long *out; unsigned char *gbuf; };

__device__ __forceinline__ uint32_t dest_of(int d, int fan, uint32_t rank) { return fan == 4 ? (uint32_t)(2 * d + (rank & 1)) : (uint32_t)d; }

// mode 0: bulk copies, one issuing thread per chain;  mode 1: st.shared::cluster from `warps` warps per chain;  mode 2: L2 multicast
__global__ void __cluster_dims__(CL, 1, 1) k_gather(P p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);                    // [S][2]
    unsigned char *src = smem + 1024;                                       // [S][SLOT]
    unsigned char *buf = src + MAXS * SLOT;                                 // [S][2][CL][SLOT]
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsrc = p.mode == 2 ? CL : p.fan;
    const int wpc = p.mode == 1 ? p.warps : 1;                              // warps per chain
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * p.S; ++i) mbar_init(&bars[i], p.mode == 1 ? (uint32_t)(nsrc * wpc) : 1u);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < MAXS * SLOT / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(src)[i] = i;
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync();
    const int c = warp / wpc, wi = warp % wpc;
    if (c < p.S && p.mode != 1 && lane == 0) { mbar_arrive_expect_tx(&bars[2 * c], (uint32_t)(nsrc * p.size)); mbar_arrive_expect_tx(&bars[2 * c + 1], (uint32_t)(nsrc * p.size)); }
    __syncthreads();
    cluster_sync();
    long long t0 = clock64();
    if (c < p.S) {
        unsigned char *g = p.gbuf + ((size_t)blockIdx.x * MAXS + c) * 2 * SLOT;
        for (int r = 0; r < p.iters; ++r) {
            const int par = r & 1;
            uint64_t *bar = &bars[2 * c + par];
            unsigned char *dst = buf + (((size_t)c * 2 + par) * CL + rank) * SLOT;
            if (p.mode == 0) {
                if (lane == 0)
                    for (int d = 0; d < p.fan; ++d) bulk_copy_to_cta(dst, src + c * SLOT, (uint32_t)p.size, bar, dest_of(d, p.fan, rank));
            } else if (p.mode == 1) {
                for (int d = 0; d < p.fan; ++d) {
                    const uint32_t dest = dest_of(d, p.fan, rank);
                    const uint32_t base = mapa(smem_u32(dst), dest);
                    for (int o = (wi * 32 + lane) * 16; o < p.size; o += wpc * 32 * 16)
                        asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + o), "r"(r), "r"(lane), "r"(3), "r"(4) : "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar, dest);
                }
            } else if (lane == 0) {
                unsigned char *gs = g + par * SLOT;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gs), "r"(smem_u32(src + c * SLOT)), "r"(p.size) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                    "l"(gs), "r"(p.size), "r"(smem_u32(bar)), "h"((uint16_t)0xFF)
                    : "memory");
            }
            if (p.mode == 1) {
                mbar_wait_cluster(bar, (uint32_t)((r >> 1) & 1));
            } else if (lane == 0) {
                mbar_wait_cluster(bar, (uint32_t)((r >> 1) & 1));
                if (r + 2 < p.iters) mbar_arrive_expect_tx(bar, (uint32_t)(nsrc * p.size));
            }
            __syncwarp();
        }
    }
    long long t1 = clock64();
    if (lane == 0 && c < p.S && wi == 0) p.out[blockIdx.x * MAXS + c] = t1 - t0;
    __syncthreads();
    cluster_sync();
}

// ---- ping: rank 0 <-> rank `peer`, one bulk copy each way per iteration ----
__global__ void __cluster_dims__(CL, 1, 1) k_ping(int size, int peer, int iters, long long *out)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    unsigned char *src = smem + 1024, *dst = src + 32768;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    cluster_sync();
    long long t0 = clock64();
    if (threadIdx.x == 0 && (rank == 0 || rank == (uint32_t)peer)) {
        const uint32_t other = rank == 0 ? (uint32_t)peer : 0u;
        for (int r = 0; r < iters; ++r) {
            mbar_arrive_expect_tx(bar, (uint32_t)size);
            if (rank == 0) bulk_copy_to_cta(dst, src, (uint32_t)size, bar, other);
            mbar_wait_cluster(bar, (uint32_t)(r & 1));
            if (rank != 0) bulk_copy_to_cta(dst, src, (uint32_t)size, bar, other);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && rank == 0) out[0] = t1 - t0;
    __syncthreads();
    cluster_sync();
}

// ---- mma: issue rate, A in TMEM, B in smem ----
template <int CG>
__global__ void k_mma(int N, int count, int reps, long long *out)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    unsigned char *b = smem + 1024;          // 32 KB of zeros as the B operand
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(b)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (CG == 2) cluster_sync();
    if (threadIdx.x < 32) { if (CG == 2) tmem_alloc2<512>(slot); else tmem_alloc<512>(slot); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *slot;
    if (CG == 2) cluster_sync();
    long long best = 1ll << 60;
    const int NH = CG == 2 ? N / 2 : N;
    for (int rep = 0; rep < reps; ++rep) {
        long long t0 = clock64();
        if (threadIdx.x < 32 && rank == 0 && elect_one()) {
            const uint32_t idesc = make_idesc_f16(CG == 2 ? 256 : 128, N);
            for (int i = 0; i < count; ++i) {
                const uint64_t bd = make_smem_desc(smem_u32(b) + (i & 3) * 2 * (NH * 16), NH * 16, 128, LAYOUT_NONE);
                if (CG == 2) mma_f16_ts2(tb + 256, tb + (i & 15) * 8, bd, idesc, i != 0);
                else mma_f16_ts(tb + 256, tb + (i & 15) * 8, bd, idesc, i != 0);
            }
            if (CG == 2) mma_commit2_mc(bar, (uint16_t)3); else mma_commit(bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(bar, (uint32_t)(rep & 1));
        long long t2 = clock64();
        if (threadIdx.x == 0 && rank == 0) { best = (t2 - t0 < best) ? (t2 - t0) : best; out[1] = t1 - t0; }
        __syncthreads();
    }
    if (threadIdx.x == 0 && rank == 0) out[0] = best;
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (threadIdx.x < 32) { if (CG == 2) tmem_dealloc2<512>(tb); else tmem_dealloc<512>(tb); }
}


// ---- tmem: which (lane, column) does each thread's register get for the 16x256b / 16x128b / 16x64b load shapes? ----
__global__ void k_tmem_probe(int *out)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<64>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    // lane L (= 32*warp + lane), column c holds the value 1000*L + c
    for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = 1000u * (32 * warp + lane) + 8 * c8 + i;
        tmem_st_x8(tb + ((uint32_t)(32 * warp) << 16) + 8 * c8, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {          // quadrant 1: lanes 32..63
        uint32_t v[8];
        // 16x256b.x2: 16 lanes x 16 columns, 8 registers
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tb + (32u << 16)));
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[lane * 8 + i] = (int)v[i];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tb + (48u << 16) + 16));
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[256 + lane * 8 + i] = (int)v[i];
        uint32_t w[4];
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(tb + (32u << 16)));
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) out[512 + lane * 4 + i] = (int)w[i];
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(tb + (32u << 16)));
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) out[640 + lane * 4 + i] = (int)w[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tb);
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "a2a";
    const int clusters = argc > 2 ? atoi(argv[2]) : 12;
    long long *out;
    CK(cudaMallocManaged(&out, sizeof(long long) * 4096));
    unsigned char *gbuf;
    CK(cudaMalloc(&gbuf, (size_t)16 * CL * MAXS * 2 * SLOT));
    const int SMEM = 1024 + 1024 + MAXS * SLOT + MAXS * 2 * CL * SLOT + 1024;     // 1 KB + 16 KB + 256 KB?  -> see below
    const int iters = 300;
    if (!strcmp(mode, "a2a") || !strcmp(mode, "st") || !strcmp(mode, "l2mc")) {
        const int m = !strcmp(mode, "a2a") ? 0 : (!strcmp(mode, "st") ? 1 : 2);
        for (int S : {1, 2, 3})
            for (int size : {512, 1024, 2048, 4096})
                for (int fan : {4, 8}) {
                    if (m == 2 && fan == 4) continue;
                    for (int warps : {1, 4}) {
                        if (m != 1 && warps != 1) continue;
                        const int smem = 1024 + 1024 + MAXS * SLOT + S * 2 * CL * SLOT + 1024;
                        CK(cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                        for (int ncl : {1, clusters}) {
                            P p{size, fan, S, iters, m, warps, out, gbuf};
                            k_gather<<<ncl * CL, 32 * S * (m == 1 ? warps : 1), smem>>>(p);
                            CK(cudaDeviceSynchronize());
                            long long mx = 0;
                            for (int i = 0; i < ncl * CL; ++i)
                                for (int c = 0; c < S; ++c) mx = std::max(mx, out[i * MAXS + c]);
                            const double sent = (double)iters * S * (m == 2 ? 1 : fan) * size, recv = (double)iters * S * (m == 2 ? CL : fan) * size;
                            printf("%-4s S %d size %4d fan %d warps %d clusters %2d : %7.0f cyc/step  sent %5.2f B/cyc/CTA  received %5.2f B/cyc/CTA\n", mode, S,
                                   size, fan, warps, ncl, (double)mx / iters, sent / mx, recv / mx);
                            fflush(stdout);
                        }
                    }
                }
    } else if (!strcmp(mode, "ping")) {
        CK(cudaFuncSetAttribute(k_ping, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 1024 + 65536 + 1024));
        for (int peer : {1, 2, 7})
            for (int size : {16, 512, 1024, 4096, 16384}) {
                k_ping<<<CL, 32, 1024 + 1024 + 65536 + 1024>>>(size, peer, 200, out);
                CK(cudaDeviceSynchronize());
                printf("ping peer %d size %5d : one-way %6.0f cyc\n", peer, size, (double)out[0] / 400.0);
            }
    } else if (!strcmp(mode, "tmem")) {
        int *o;
        CK(cudaMallocManaged(&o, sizeof(int) * 1024));
        k_tmem_probe<<<1, 128>>>(o);
        CK(cudaDeviceSynchronize());
        const char *names[4] = {"16x256b.x2 @lane32,col0", "16x256b.x2 @lane48,col16", "16x128b.x2 @lane32", "16x64b.x4 @lane32"};
        const int base[4] = {0, 256, 512, 640}, nreg[4] = {8, 8, 4, 4};
        for (int k = 0; k < 4; ++k) {
            printf("%s  (thread: reg -> lane.col)\n", names[k]);
            for (int t = 0; t < 32; ++t) {
                printf("  t%02d:", t);
                for (int i = 0; i < nreg[k]; ++i) printf(" %d.%02d", o[base[k] + t * nreg[k] + i] / 1000, o[base[k] + t * nreg[k] + i] % 1000);
                printf("\n");
            }
        }
    } else if (!strcmp(mode, "mma")) {
        CK(cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
        CK(cudaFuncSetAttribute(k_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
        for (int N : {16, 32, 64, 128, 256})
            for (int count : {48, 192}) {
                k_mma<1><<<1, 64, 40000>>>(N, count, 5, out);
                CK(cudaDeviceSynchronize());
                printf("mma cta_group 1 M128 N %3d x%3d : %6lld cyc total, %5.1f cyc/MMA (issue loop %lld)\n", N, count, out[0], (double)out[0] / count, out[1]);
                if (N >= 32) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(2); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 40000;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    CK(cudaLaunchKernelEx(&cfg, k_mma<2>, N, count, 5, out));
                    CK(cudaDeviceSynchronize());
                    printf("mma cta_group 2 M256 N %3d x%3d : %6lld cyc total, %5.1f cyc/MMA (issue loop %lld)\n", N, count, out[0], (double)out[0] / count, out[1]);
                }
                fflush(stdout);
            }
    }
    (void)SMEM;
    return 0;
}
