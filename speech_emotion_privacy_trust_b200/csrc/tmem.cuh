// Tensor memory (TMEM, sm_100a) as a lane-private constant store.
//
// The extraction kernel is bound by the shared-memory pipe (one 128-byte wavefront per clock per SM).  A quarter of
// its shared-memory reads fetch values that depend only on the LANE, never on the item: the lane's window samples,
// its split twiddles and its mel gather program.  TMEM is 128 lanes x 512 32-bit columns per SM with its own
// register data path (tcgen05.ld / tcgen05.st, SASS LDTM / STTM); a warp reaches the 32 TMEM lanes of its quarter
// (warp id % 4), thread i <-> TMEM lane 32 (warp % 4) + i.  That is exactly a per-lane table: the kernel fills it once
// and every item reads it from there instead of from shared memory.  No tensor-core instruction is involved.
#pragma once
#include <cstdint>

namespace sept {
namespace tmem {

// one warp allocates `cols` columns (power of two >= 32) for the CTA; the base address lands in *slot (shared memory)
__device__ __forceinline__ void alloc(uint32_t* slot, uint32_t cols) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of column `col` in the calling warp's lane quarter
__device__ __forceinline__ uint32_t quarter_addr(uint32_t base, int warp, int col) {
    return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
}

// 32x32b shape: every thread moves N consecutive columns of its own TMEM lane
__device__ __forceinline__ void ld2(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ld4(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ld8(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
}
__device__ __forceinline__ void ld16(uint32_t a, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
}
__device__ __forceinline__ void st2(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// N columns (a multiple of 2) starting at `a`, largest shapes first
template <int N>
__device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    static_assert(N % 2 == 0 && N > 0, "even column counts");
    if constexpr (N >= 16) { ld16(a, r); if constexpr (N > 16) ld<N - 16>(a + 16, r + 16); }
    else if constexpr (N >= 8) { ld8(a, r); if constexpr (N > 8) ld<N - 8>(a + 8, r + 8); }
    else if constexpr (N >= 4) { ld4(a, r); if constexpr (N > 4) ld<N - 4>(a + 4, r + 4); }
    else ld2(a, r);
}
template <int N>
__device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    static_assert(N % 2 == 0 && N > 0, "even column counts");
    if constexpr (N >= 8) { st8(a, r); if constexpr (N > 8) st<N - 8>(a + 8, r + 8); }
    else { st2(a, r); if constexpr (N > 2) st<N - 2>(a + 2, r + 2); }
}

}  // namespace tmem
}  // namespace sept
