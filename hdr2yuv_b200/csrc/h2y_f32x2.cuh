// h2y_f32x2.cuh -- Blackwell packed-fp32 (f32x2) helpers and the magic-number floor used by the fast kernels.
//
// FFMA2 / FADD2 / FMUL2 (PTX fma/add/mul.f32x2, sm_100+) process two fp32 lanes per issue slot.  Operands are
// 64-bit register pairs; pk()/unpk() are register renames (mov.b64), not arithmetic.
// floor(x) for |x| < 2^22 is taken with a round-down add of 1.5*2^23: the integer lands in the low mantissa
// bits (bits(x + MAGIC) - MAGIC_BITS == floor(x)), with no F2I conversion (16 lanes/clk/SM on B200).
#pragma once

#include <cuda_runtime.h>

namespace h2y {

constexpr float MAGIC = 12582912.0f;                      // 1.5 * 2^23
constexpr int MAGIC_BITS = 0x4B400000;

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(u64 v, int &lo, int &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ float plo(u64 v) { int a, b; unpk(v, a, b); return __int_as_float(a); }
__device__ __forceinline__ float phi(u64 v) { int a, b; unpk(v, a, b); return __int_as_float(b); }
__device__ __forceinline__ int ilo(u64 v) { int a, b; unpk(v, a, b); return a; }
__device__ __forceinline__ int ihi(u64 v) { int a, b; unpk(v, a, b); return b; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fsub2(u64 a, u64 b)         // FADD2 with a negated operand
{
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// {a.lo*m, a.hi*m}, each product rounded on its own.  NOT mul.rn.f32x2: ptxas 12.9 contracts mul.rn.f32x2 +
// add.rn.f32x2 into one FFMA2 even with explicit .rn and -fmad=false (checked in the SASS), which would drop
// the reference's separate rounding of x*maxVR before +minVR (convert.cpp:1141).  Scalar FMULs are left alone.
__device__ __forceinline__ u64 fmul2s(float lo, float hi, float m) { return pk(__fmul_rn(lo, m), __fmul_rn(hi, m)); }

__device__ __forceinline__ u64 fadd2_rm(u64 a, u64 b)      // round toward -inf
{
    u64 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ int clamp3(int v, int lo, int hi) { return min(max(v, lo), hi); }
// two 16-bit codes per register, clamped with one VIMNMX pair (lo / hi hold the bound in both halves)
__device__ __forceinline__ unsigned clamp_u16x2(unsigned v, unsigned lo, unsigned hi)
{
    asm("max.u16x2 %0, %0, %1;" : "+r"(v) : "r"(lo));
    asm("min.u16x2 %0, %0, %1;" : "+r"(v) : "r"(hi));
    return v;
}
__device__ __forceinline__ unsigned clamp_s16x2(unsigned v, unsigned lo, unsigned hi)
{
    asm("max.s16x2 %0, %0, %1;" : "+r"(v) : "r"(lo));
    asm("min.s16x2 %0, %0, %1;" : "+r"(v) : "r"(hi));
    return v;
}

}   // namespace h2y
