// Blackwell (sm_100a) building blocks used by the tcgen05 kernels of this library: mbarrier, bulk-copy TMA,
// TMEM allocation / loads, UMMA shared-memory + instruction descriptors and the tcgen05.mma / commit wrappers.
// Everything is inline PTX; there is no CUTLASS dependency.  The descriptor bit layouts follow the PTX ISA's
// "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma / bulk copies read through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bulk copy (TMA engine, 1-D): global -> shared, completion counted on an mbarrier ------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- tensor memory ---------------------------------------------------------------------------------------------------
// whole warp; writes the TMEM base address (lane 0, first column) of `cols` (power of two >= 32) columns to *slot
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp receives row (lane-quarter base + lane), columns
// [col, col+32).  The warp may only touch the TMEM lane quarter 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// same, 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors ----------------------------------------------------------------------------------------------
// Operand tiles are K-major ("row r holds consecutive k") with 32 fp32/tf32 elements = 128 bytes of k per row, stored as
// the canonical SWIZZLE_128B layout: row r at byte r*128, its 16-byte chunk c at position (c ^ (r & 7)); 8-row groups are
// 1024 bytes apart (= the descriptor's stride byte offset).  Tile bases must be 1024-byte aligned.  One tcgen05.mma of
// kind::tf32 consumes k = 8 (32 bytes): k-slice s of the tile is addressed by advancing the start address by 32*s bytes.
constexpr uint32_t ROW_BYTES = 128;

// byte offset of element (r, k) of a K-major SWIZZLE_128B tile, k in [0, 32)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t k) {
    return r * ROW_BYTES + ((((k >> 2) ^ (r & 7u)) & 7u) << 4) + ((k & 3u) << 2);
}

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);          // [0,14)  start address >> 4
    d |= (uint64_t)0 << 16;                           // [16,30) leading byte offset >> 4 (unused: one swizzle atom along k)
    d |= (uint64_t)(1024u >> 4) << 32;                // [32,46) stride byte offset >> 4: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                           // [46,48) descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // [61,64) layout: SWIZZLE_128B
    return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4)            // [4,6)   D format: F32
           | (2u << 7)          // [7,10)  A format: TF32
           | (2u << 10)         // [10,13) B format: TF32
           | (0u << 15)         // A major: K
           | (0u << 16)         // B major: K
           | ((N >> 3) << 17)   // [17,23) N >> 3
           | ((M >> 4) << 24);  // [24,29) M >> 4
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 3xTF32 split: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (fp32 accumulate)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    const float r = x - __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    hi = __uint_as_float(h);
    lo = __uint_as_float(l);
}

}  // namespace tc5
