// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, bulk / tensor-map TMA copies, tcgen05 alloc / mma /
// commit / ld, shared-memory matrix descriptors, thread-block-cluster helpers.  Every mbarrier wait is bounded by a
// clock64 watchdog; a kernel whose watchdog fires records a code and traps (nn_pipeline_abort).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr long long UM_TIMEOUT = 4000000000LL;       // ~2 s of SM clocks

// A pipeline watchdog fired (a barrier did not complete within UM_TIMEOUT clocks): record the code and TRAP.  The
// accumulators are incomplete, so letting the epilogue and the optimizer run on them would silently corrupt the
// model; the trap turns the failure into a launch error that the next CUDA call of the process reports.
__device__ __noinline__ void nn_pipeline_abort(int* err_flag, int code) {
    if (err_flag) atomicExch(err_flag, code);
    __threadfence_system();
    printf("noisynet_b200: tcgen05 pipeline watchdog fired (code %d, block %d,%d)\n", code, (int)blockIdx.x, (int)blockIdx.y);
    __trap();
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    // (polling with mbarrier.test_wait instead of try_wait was measured: no difference)
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > UM_TIMEOUT) return false;
    }
    return true;
}
// waiting roles that are not on the critical path (producers waiting for a free stage, epilogue warps waiting for an
// accumulator): sleep between polls instead of competing with the MMA thread for issue slots and the barrier unit
__device__ __forceinline__ bool mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if (clock64() - t0 > UM_TIMEOUT) return false;
    }
    return true;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {     // acquire at cluster scope (remote arrivals)
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > UM_TIMEOUT) return false;
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// multicast variants (thread-block cluster of 2): the bulk copy lands at the same CTA-relative offset in every CTA
// of the mask and completes bytes on the mbarrier at the same offset there; the commit arrives on every CTA's barrier
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// CTA pair (cta_group::2): both SMs of a TPC execute ONE MMA of M = 256 -- each CTA contributes its 128 rows of A and
// HALF of the B rows from its own shared memory, each CTA's TMEM receives its 128 rows of D
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta_rank) {
    // arrive on the barrier at the same shared-memory offset in CTA `cta_rank` of the cluster
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(local_bar), "r"(cta_rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1): 8-row groups of
// 1024 bytes (SBO), swizzle atom = 8 rows x 128 B.  Advancing 16 bf16 along K = +32 bytes = +2 units.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void st_global_f32(float* ptr, float v) {
    asm volatile("st.global.f32 [%0], %1;" ::"l"(ptr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4x2(uint32_t ta, uint32_t tb, float a[4], float b[4]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%8];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4,%5,%6,%7}, [%9];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(ta), "r"(tb) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[4 + i]); }
}
__device__ __forceinline__ void tmem_ld4x4(uint32_t ta, uint32_t tb, uint32_t tc, uint32_t td, float a[4], float b[4], float c[4], float d[4]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4,%5,%6,%7}, [%17];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8,%9,%10,%11}, [%18];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12,%13,%14,%15}, [%19];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta), "r"(tb), "r"(tc), "r"(td) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[4 + i]); c[i] = __uint_as_float(r[8 + i]); d[i] = __uint_as_float(r[12 + i]);
    }
}
// K-major SWIZZLE_NONE descriptor: core matrix = 8 rows x 16 B contiguous; lbo = distance between the two
// K chunks of one MMA, sbo = distance between 8-row groups (both in 16-byte units).
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo_units & 0x3FFF) << 16;
    d |= (uint64_t)(sbo_units & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}


// Swizzled K-major descriptor for a tile whose rows are `sw_bytes` (32 / 64 / 128) wide: 8-row atoms of 8 * sw_bytes
// (SBO), layout type 6 / 4 / 2.  Advancing 16 bf16 along K = +32 bytes = +2 units inside a row.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t sw_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8u * sw_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(sw_bytes == 128u ? 2 : (sw_bytes == 64u ? 4 : 6)) << 61;
    return d;
}

// im2col-mode tensor-map copy (cuTensorMapEncodeIm2col): `pixelsPerColumn` output pixels x `channelsPerPixel` channels
// of one filter tap land as rows of the swizzled K-major tile.  Coordinates: first channel, then the INPUT position of
// the tile's first output pixel for tap (0, 0) -- (ow * stride - pad, oh * stride - pad, image) -- and the tap as offsets.
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h, int n,
                                              uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
// CTA-pair variants (cta_group::2): issued by BOTH CTAs of a pair for their own shared memory, but the transaction bytes
// are counted on the barrier of the pair's LEADER (the even CTA: bit 24 of the shared::cluster address cleared), so the
// leader's MMA thread waits on ONE barrier for both CTAs' operands.
constexpr uint32_t NN_PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_im2col_4d_2cta(uint32_t dst, const void* map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 ::"r"(dst), "l"(map), "r"(bar & NN_PEER_BIT_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_tile_2d_2cta(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar & NN_PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
// One lane of a CONVERGED warp.  The single-thread instructions of this file (tcgen05.mma / commit, TMA copies) take
// their operands from uniform registers; issued under `if (lane == 0)` inside divergent code the compiler cannot prove the
// operands warp-uniform and moves every one of them through R2UR in an ELECT loop on each use (5 per MMA, ~250 cycles per
// MMA measured).  Role loops therefore run on the whole warp with warp-uniform control flow and elect the issuing lane.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace

