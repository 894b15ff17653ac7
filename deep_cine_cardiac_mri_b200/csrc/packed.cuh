// Value types of the generated DFT codelets (codelets.cuh).
//
//   float  one transform per thread (scalar FADD / FMUL / FFMA)
//   f2     TWO independent transforms per thread, one in each half of an aligned 64-bit register pair: on sm_100a
//          every butterfly operation is then ONE packed instruction (add/sub/mul/fma.f32x2 -> SASS FADD2 / FMUL2 /
//          FFMA2, literal twiddles stay immediates), i.e. half the issue slots per transformed point.  The fp32 pipe
//          does the same work either way; what the packing buys is issue bandwidth for the loads, stores and
//          shared-memory traffic that have to run beside the arithmetic.
//
// On the host (tests/host_emul) f2 is simply a pair of floats.
#pragma once
#ifndef B2S_HD
#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD inline
#endif
#endif
#include <math.h>

namespace b2s {

B2S_HD float vadd(float a, float b) { return a + b; }
B2S_HD float vsub(float a, float b) { return a - b; }
B2S_HD float vmul(float c, float a) { return c * a; }
B2S_HD float vfma(float c, float a, float b) { return fmaf(c, a, b); }
B2S_HD float vneg(float a) { return -a; }

struct alignas(8) f2 { float x, y; };
B2S_HD f2 make_f2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }

#if defined(__CUDA_ARCH__)
#define B2S_U64(v) (*reinterpret_cast<const unsigned long long*>(&(v)))
__device__ __forceinline__ f2 vadd(f2 a, f2 b) { f2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(B2S_U64(a)), "l"(B2S_U64(b))); return r; }
__device__ __forceinline__ f2 vsub(f2 a, f2 b) { f2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(B2S_U64(a)), "l"(B2S_U64(b))); return r; }
__device__ __forceinline__ f2 vmul2(f2 a, f2 b) { f2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(B2S_U64(a)), "l"(B2S_U64(b))); return r; }
__device__ __forceinline__ f2 vfma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(B2S_U64(a)), "l"(B2S_U64(b)), "l"(B2S_U64(c))); return r; }
#undef B2S_U64
#else
inline f2 vadd(f2 a, f2 b) { return make_f2(a.x + b.x, a.y + b.y); }
inline f2 vsub(f2 a, f2 b) { return make_f2(a.x - b.x, a.y - b.y); }
inline f2 vmul2(f2 a, f2 b) { return make_f2(a.x * b.x, a.y * b.y); }
inline f2 vfma2(f2 a, f2 b, f2 c) { return make_f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
B2S_HD f2 vmul(float c, f2 a) { return vmul2(a, make_f2(c, c)); }
B2S_HD f2 vfma(float c, f2 a, f2 b) { return vfma2(a, make_f2(c, c), b); }
B2S_HD f2 vneg(f2 a) { return vmul2(a, make_f2(-1.f, -1.f)); }

}  // namespace b2s
