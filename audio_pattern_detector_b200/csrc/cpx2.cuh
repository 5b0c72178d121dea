// Packed complex arithmetic for sm_100a: a complex64 lives in one 64-bit register pair and is
// added / multiplied with Blackwell's two-wide fp32 instructions (PTX add/sub/mul/fma .f32x2 ->
// SASS FADD2 / FMUL2 / FFMA2).  ptxas folds the half swaps, sign patterns and scalar broadcasts
// written below as pack/unpack moves into operand modifiers (R.F32x2.LO_HI.NP, R.F32), so
//   complex add / sub / a +- i*b : 1 instruction      complex multiply : 2 instructions
// which halves the issue slots of an FFT butterfly (the fp32 pipe time is unchanged).
//
// The memory image of a c2 is that of a float2 (re in the low word), so spectra are loaded and
// stored directly as 64-bit words.
#pragma once
#include <cuda_runtime.h>

namespace apd {

struct c2 {
    unsigned long long v;
};

__device__ __forceinline__ c2 mk(float re, float im)
{
    c2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ void split(c2 a, float& re, float& im)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(a.v));
}
__device__ __forceinline__ c2 from_f2(float2 a) { return mk(a.x, a.y); }
__device__ __forceinline__ float2 to_f2(c2 a)
{
    float2 r;
    split(a, r.x, r.y);
    return r;
}

__device__ __forceinline__ c2 operator+(c2 a, c2 b)
{
    c2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ c2 operator-(c2 a, c2 b)
{
    c2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ c2 mul2(c2 a, c2 b)          // lane-wise product (not a complex product)
{
    c2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c)    // lane-wise a*b + c
{
    c2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// i*a and -i*a (pure register renaming + sign pattern once folded into the consumer)
__device__ __forceinline__ c2 rot_p(c2 a)
{
    float x, y;
    split(a, x, y);
    return mk(-y, x);
}
__device__ __forceinline__ c2 rot_m(c2 a)
{
    float x, y;
    split(a, x, y);
    return mk(y, -x);
}
template <int SIGN> __device__ __forceinline__ c2 rot(c2 a) { return SIGN > 0 ? rot_p(a) : rot_m(a); }   // SIGN*i*a

// a + s*b, a - s*b with a real scalar s (one FFMA2 each; s may be an immediate)
__device__ __forceinline__ c2 axpy(c2 a, float s, c2 b) { return fma2(b, mk(s, s), a); }

// complex products with a twiddle held as two scalars (2 registers per twiddle)
__device__ __forceinline__ c2 cmul(c2 x, float wr, float wi)          // x * (wr + i wi)
{
    return fma2(rot_p(x), mk(wi, wi), mul2(x, mk(wr, wr)));
}
__device__ __forceinline__ c2 cmul(c2 x, float2 w) { return cmul(x, w.x, w.y); }
__device__ __forceinline__ c2 cmul(c2 x, c2 w)
{
    float wr, wi;
    split(w, wr, wi);
    return cmul(x, wr, wi);
}

}  // namespace apd
