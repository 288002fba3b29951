// tcgen05.mma issue-rate microbenchmark for CTA pairs (cta_group::2, M = 256) against single CTAs (cta_group::1, M = 128):
// cycles per K = 16 bf16 MMA for the accumulator widths of the conv2 kernels, and the cost of a tcgen05.commit every
// `per_commit` MMAs (the per-stage release of the operand ring).  SWIZZLE_128B K-major operands (zeros), 148 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate2 mma_rate2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int pair, N, iters, per_commit, mc_commit; };

__device__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <bool PAIR>
__global__ void k_rate(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
        const uint64_t ad0 = desc_sw128(base), bd0 = desc_sw128(base + 48 * 1024);
        const long long t0 = clock64();
        int since = 0;
        for (int i = 0; i < c.iters; ++i) {
            const uint64_t ad = ad0 + (uint64_t)((i & 3) * 2), bd = bd0 + (uint64_t)((i & 3) * 2);
            if (PAIR)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i));
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i));
            if (c.per_commit && ++since == c.per_commit) {
                since = 0;
                if (PAIR) {
                    if (c.mc_commit) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                                  ::"r"(smem_u32(&bar[1])), "h"((uint16_t)3));
                    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])));
                } else
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])));
            }
        }
        if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])));
        else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])));
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[0])));
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- the issue pattern of k_conv_tma's MMA thread: groups of (4 x SWIZZLE_128B K steps + `tail` x SWIZZLE_32B step) over a
// ring of 5 stages of distinct shared-memory tiles, one multicast commit per group; optionally `pollers` other warps spin
// on an mbarrier that never completes (as waiting producer / epilogue warps do)
struct GCfg { int N, groups, tail, pollers, n64, random_data; };

__device__ uint64_t desc_k(uint32_t addr, uint32_t sw_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8u * sw_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(sw_bytes == 128u ? 2 : (sw_bytes == 64u ? 4 : 6)) << 61;
    return d;
}

__global__ void __launch_bounds__(512, 1) k_groups(GCfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c.random_data) {        // bf16 values of magnitude ~1 with random mantissas and signs (zeros keep the datapath idle)
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
            uint32_t w[4];
            for (int j = 0; j < 4; ++j) { h = h * 1664525u + 1013904223u; w[j] = (h & 0x807F807Fu) | 0x3F803F80u; }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4*>(smem)[i] = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const long long t0 = clock64();
        int s = 0;
        for (int g = 0; g < c.groups; ++g) {
            const uint32_t a_s = base + (uint32_t)s * 40960u, b_s = a_s + 20480u;
            const uint64_t ad = desc_k(a_s, 128), bd = desc_k(b_s, 128);
            for (int k = 0; k < c.n64; ++k)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad + 2 * (k & 3) + (k >> 2) * 1024), "l"(bd + 2 * (k & 3)), "r"(idesc), "r"(g | k));
            for (int k = 0; k < c.tail; ++k) {
                const uint64_t ad2 = desc_k(a_s + 16384u, 32), bd2 = desc_k(b_s + 15360u, 32);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad2), "l"(bd2), "r"(idesc), "r"(1));
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar[1])), "h"((uint16_t)3));
            if (++s == 5) s = 0;
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])));
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[0])));
        out[blockIdx.x] = clock64() - t0;
        *reinterpret_cast<volatile uint32_t*>(&tslot) = 0xFFFFFFFFu;      // stop the pollers of this CTA
    } else if (warp >= 1 && warp <= c.pollers && (threadIdx.x & 31) == 0 && rank == 0) {
        // a waiting role: polls a barrier that stays incomplete until the issuer is done
        while (*reinterpret_cast<volatile uint32_t*>(&tslot) != 0xFFFFFFFFu) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[2])) : "memory");
            (void)clock64();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static void run_groups(long long* out, GCfg c) {
    cudaFuncSetAttribute(k_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    cudaMemset(out, 0, 148 * sizeof(long long));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 206 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_groups, c, out) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        printf("k_groups failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return;
    }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0; int n = 0;
    for (int i = 0; i < 148; ++i) if (h[i]) { sum += h[i]; ++n; }
    printf("groups: N=%3d  %d x sw128 + %d x sw32 per group, %2d polling warps, %s operands: %7.1f cycles / group  (%.1f per MMA)\n", c.N, c.n64, c.tail, c.pollers,
           c.random_data ? "random" : "zero  ",
           sum / n / c.groups, sum / n / c.groups / (c.n64 + c.tail));
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_rate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_rate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    {
        GCfg cs[] = {{240, 400, 1, 0, 4, 0}, {240, 400, 1, 0, 4, 1}, {240, 400, 0, 0, 5, 1}, {240, 400, 1, 12, 4, 1}, {80, 400, 0, 0, 8, 0}, {80, 400, 0, 0, 8, 1},
                     {256, 400, 0, 0, 4, 0}, {256, 400, 0, 0, 4, 1}};
        for (auto& c : cs) run_groups(out, c);
    }
    printf("%-6s %4s %10s %10s  cycles/MMA\n", "group", "N", "per_commit", "multicast");
    const int Ns[3] = {80, 128, 240};
    const int pcs[4] = {0, 8, 5, 1};
    if (getenv("SKIP_TABLE")) return 0;
    for (int pair = 0; pair <= 1; ++pair)
        for (int ni = 0; ni < 3; ++ni)
            for (int pi = 0; pi < 4; ++pi)
                for (int mc = 0; mc <= pair; ++mc) {
                    if (mc && pcs[pi] == 0) continue;
                    Cfg c{pair, Ns[ni], 2000, pcs[pi], mc};
                    cudaMemset(out, 0, 148 * sizeof(long long));
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 99 * 1024;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, k_rate<true>, c, out) : cudaLaunchKernelEx(&cfg, k_rate<false>, c, out);
                    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    long long h[148];
                    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    double s = 0; int n = 0;
                    for (int i = 0; i < 148; ++i) if (h[i]) { s += h[i]; ++n; }
                    printf("%-6s %4d %10d %10d  %9.1f   (M*N/256 per SM = %d)\n", pair ? "pair" : "single", Ns[ni], pcs[pi], mc, s / n / c.iters, 128 * Ns[ni] / 256);
                }
    return 0;
}
