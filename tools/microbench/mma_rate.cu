// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per MMA (K = 16, bf16 -> fp32) as a function of the
// shared-memory layout of A and B (SWIZZLE_NONE "interleave" vs SWIZZLE_128B, K-major vs MN-major), M and N.
// One CTA per SM issues ITER dependent-free MMAs into one accumulator and waits for the commit; operands are
// whatever shared memory holds (zeros).   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int a_mode, b_mode, M, N, a_shift, iters, vary; };   // mode: 0 K-major none, 1 K-major sw128, 2 MN-major none, 3 MN-major sw128

__device__ uint64_t make_desc(uint32_t addr, int mode, int rows) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 46;
    if (mode == 0) { d |= (uint64_t)(rows * 16 >> 4) << 16; d |= (uint64_t)8 << 32; }                 // LBO = plane, SBO = 128 B
    else if (mode == 1) { d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)2 << 61; }
    else if (mode == 2) { d |= (uint64_t)8 << 16; d |= (uint64_t)(256 >> 4) << 32; }                  // LBO = 128 B (8 k-rows), SBO = atom stride 256 B
    else { d |= (uint64_t)(8192 >> 4) << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)2 << 61; }
    return d;
}

__global__ void k_rate(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const bool a_mn = c.a_mode >= 2, b_mn = c.b_mode >= 2;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? 1u << 15 : 0) | (b_mn ? 1u << 16 : 0) |
                               ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        const uint64_t ad = make_desc(base + c.a_shift * 16, c.a_mode, c.M);
        const uint64_t bd = make_desc(base + 48 * 1024, c.b_mode, c.N);
        const long long t0 = clock64();
        // K-advance per iteration as a real main loop does it (distinct operand addresses: no operand reuse)
        const uint32_t astep = c.a_mode == 0 ? 1u : c.a_mode == 1 ? 2u : c.a_mode == 2 ? 16u : 128u;
        const uint32_t bstep = c.b_mode == 0 ? 1u : c.b_mode == 1 ? 2u : c.b_mode == 2 ? 16u : 128u;
        const uint32_t mask = c.vary ? (c.a_mode == 1 || c.a_mode == 3 ? 3u : 7u) : 0u;
        const uint32_t maskb = c.vary ? (c.b_mode == 1 || c.b_mode == 3 ? 3u : 7u) : 0u;
        const uint64_t ad_base = ad, bd_base = bd;
        for (int i = 0; i < c.iters; ++i) {
            const uint64_t ad = ad_base + (uint64_t)((i & mask) * astep);
            const uint64_t bd = bd_base + (uint64_t)((i & maskb) * bstep);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(i));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)));
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[4] = {"K-major none ", "K-major sw128", "MN-major none", "MN-major sw128"};
    printf("%-14s %-14s %4s %4s %6s  cycles/MMA  floor(M*N/256)\n", "A layout", "B layout", "M", "N", "vary");
    const int Ns[4] = {48, 80, 144, 256};
    for (int am = 0; am < 4; ++am)
        for (int bm = 0; bm < 4; ++bm) {
            if ((am >= 2) != (bm >= 2)) continue;          // keep the table small: both K-major or both MN-major
            for (int M = 64; M <= 128; M += 64)
                for (int ni = 0; ni < 4; ++ni)
                    for (int vary = 0; vary <= 1; ++vary) {
                        const int shift = 0;
                        Cfg c{am, bm, M, Ns[ni], shift, 2000, vary};
                        if (M == 64 && (Ns[ni] % 8)) continue;
                        k_rate<<<148, 128, 99 * 1024>>>(c, out);
                        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                        long long h[148];
                        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                        double s = 0;
                        for (int i = 0; i < 148; ++i) s += h[i];
                        printf("%-14s %-14s %4d %4d %6d  %9.1f  %6d\n", names[am], names[bm], M, Ns[ni], vary, s / 148 / c.iters,
                               (M < 128 ? 128 : M) * Ns[ni] / 256);
                    }
        }
    return 0;
}
