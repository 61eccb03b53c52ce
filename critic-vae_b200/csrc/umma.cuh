// Blackwell (sm_100a) primitives shared by the conv kernels: mbarrier, bulk async copy,
// TMEM allocation, tcgen05.mma / commit / ld, and the no-swizzle shared-memory matrix
// descriptor that the halo-plane layout is built around.
//
// Everything here is inline PTX; nothing is borrowed from a library at run time.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cvae {

// ---------------------------------------------------------------------------------------------
// Bounded waits.  A pipeline bug must never hang the GPU box: every mbarrier wait gives up after
// kSpinLimit probes, raises a device-side flag and lets the kernel run to completion with garbage
// results; cvae_check_device_fault() turns the flag (a device int owned by api.cu) into CVAE_EDEVICE.
// ---------------------------------------------------------------------------------------------
static constexpr uint32_t kSpinLimit = 1u << 22;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Returns false (and raises *fault) when the barrier never flips.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* fault) {
#pragma unroll 1
    for (uint32_t i = 0; i < kSpinLimit; ++i) {
        if (mbar_try_wait(bar, parity)) return true;
    }
    atomicExch(fault, 1);
    return false;
}

// The same for roles that are NOT on the critical path (epilogue warps waiting for an accumulator, producers waiting
// for a free slot): back off between probes so the spinning warp does not take issue slots from the MMA thread that
// shares its scheduler.
__device__ __forceinline__ bool mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int* fault) {
#pragma unroll 1
    for (uint32_t i = 0; i < kSpinLimit; ++i) {
        if (mbar_try_wait(bar, parity)) return true;
        __nanosleep(i < 64 ? 32 : 256);
    }
    atomicExch(fault, 1);
    return false;
}

// One lane of a fully active warp.  tcgen05.mma / commit / bulk copies run on the uniform datapath: issued
// from `if (lane == 0)` code the compiler wraps each of them in an ELECT + BRA.U.ANY serialisation loop
// (hundreds of cycles per MMA); under elect.sync it emits them directly.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// Generic-proxy writes (st.shared) must be fenced before the async proxy (UMMA, bulk copy) reads.
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared, completion counted on an mbarrier (bytes % 16 == 0).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One thread: D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One thread: arrive on `bar` once every previously issued tcgen05.mma has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}

// tcgen05.ld 32 lanes x 32 bit, 16 consecutive columns: thread `lane` of the warp receives
// row (lane_base + lane), columns [col, col+16).  taddr = (lane_base << 16) | col.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// Descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") layout, Blackwell version field 1.
// The operand is a grid of 8x16-byte core matrices (8 rows 16 B apart = 128 contiguous bytes).
//   K-major : rows are M/N, the 16 bytes are 8 K-elements.  sbo = bytes between core matrices
//             adjacent in M/N, lbo = bytes between the two core matrices of one K=16 step.
//   MN-major: the 16 bytes are 8 M/N-elements, rows are K.   sbo = bytes between core matrices
//             adjacent in M/N, lbo = bytes between the two K-groups of 8 of one K=16 step.
// Only 16-byte alignment of `saddr` is needed: this is what lets one halo tile in shared memory
// serve all 25 taps of a 5x5 filter, each tap being the same descriptor with a shifted start.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (((lbo >> 4) & 0x3FFFu) << 16);
    uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D, M = 128.
static constexpr uint32_t kMajorK = 0, kMajorMN = 1;
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t n, uint32_t a_major,
                                                       uint32_t b_major, uint32_t m = 128) {
    return (1u << 4)              // D format fp32
           | (1u << 7)            // A format bf16
           | (1u << 10)           // B format bf16
           | (a_major << 15) | (b_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace cvae
