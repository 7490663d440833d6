// pk2: two fp32 values that always take the same arithmetic -- here, the same spectral sample of two
// adjacent frames.  On sm_100a it is one 64-bit register pair driven by the packed FP32 instructions
// (add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2, with splatted constants folded to immediates), which
// halves the issue slots of the FFT butterflies.  On the host (tests/hostsim only) it is a plain struct, so
// the very same phase functions can be executed lane by lane on a CPU-only box.
#pragma once

#if defined(__CUDACC__)
#define SEPT_HD __host__ __device__ __forceinline__
#else
#define SEPT_HD inline
#endif

namespace sept {

#if defined(__CUDA_ARCH__)

struct pk2 { unsigned long long v; };

SEPT_HD pk2 pk(float a, float b) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
SEPT_HD float lo(pk2 p) { return __uint_as_float((unsigned)(p.v & 0xffffffffull)); }
SEPT_HD float hi(pk2 p) { return __uint_as_float((unsigned)(p.v >> 32)); }
SEPT_HD pk2 operator+(pk2 a, pk2 b) { pk2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SEPT_HD pk2 operator-(pk2 a, pk2 b) { pk2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SEPT_HD pk2 operator*(pk2 a, pk2 b) { pk2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SEPT_HD pk2 fma2(pk2 a, pk2 b, pk2 c) {
    pk2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r;
}

#else

struct pk2 { float a, b; };

SEPT_HD pk2 pk(float a, float b) { return pk2{a, b}; }
SEPT_HD float lo(pk2 p) { return p.a; }
SEPT_HD float hi(pk2 p) { return p.b; }
SEPT_HD pk2 operator+(pk2 x, pk2 y) { return pk2{x.a + y.a, x.b + y.b}; }
SEPT_HD pk2 operator-(pk2 x, pk2 y) { return pk2{x.a - y.a, x.b - y.b}; }
SEPT_HD pk2 operator*(pk2 x, pk2 y) { return pk2{x.a * y.a, x.b * y.b}; }
SEPT_HD pk2 fma2(pk2 x, pk2 y, pk2 z) { return pk2{x.a * y.a + z.a, x.b * y.b + z.b}; }

#endif

// 8-byte shared-memory store that the assembler may not fuse with a neighbour into a 16-byte store: fusing forces
// the two register pairs into one aligned quad, which costs four MOVs per store in the DFT epilogues
#if defined(__CUDA_ARCH__)
SEPT_HD void st_shared(pk2* p, pk2 v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "l"(v.v) : "memory");
}
#else
SEPT_HD void st_shared(pk2* p, pk2 v) { *p = v; }
#endif

// warp shuffles of a packed pair (device only; the host emulation walks lanes explicitly and never calls them)
#if defined(__CUDA_ARCH__)
SEPT_HD pk2 shfl_up(pk2 x, int d) { pk2 r; r.v = __shfl_up_sync(0xffffffffu, x.v, d); return r; }
SEPT_HD pk2 shfl_lane(pk2 x, int l) { pk2 r; r.v = __shfl_sync(0xffffffffu, x.v, l); return r; }
#else
SEPT_HD pk2 shfl_up(pk2 x, int) { return x; }
SEPT_HD pk2 shfl_lane(pk2 x, int) { return x; }
#endif

SEPT_HD pk2 splat(float c) { return pk(c, c); }
SEPT_HD pk2 neg(pk2 x) { return splat(0.f) - x; }
// a - b*c
SEPT_HD pk2 fnma2(pk2 b, pk2 c, pk2 a) { return fma2(neg(b), c, a); }

}  // namespace sept
