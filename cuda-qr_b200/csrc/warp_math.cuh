// warp_math.cuh -- packed-fp32 (FFMA2) helpers and the reflector scalar approximations shared by the warp-resident
// Householder kernels (tsqr_flat.cu, tsqr_mma.cu).
#pragma once
#include "common.cuh"

namespace cqr {

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fpack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void funpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fsum2(f32x2 v) { float lo, hi; funpack2(v, lo, hi); return lo + hi; }
#ifdef CQR_DOT_SCALAR
// dot-product accumulate on two scalar FFMAs (experiment: an FFMA2 with three fresh 64-bit sources issues every 3 cycles)
__device__ __forceinline__ f32x2 dfma2(f32x2 a, f32x2 b, f32x2 c) {
  float al, ah, bl, bh, cl, ch;
  funpack2(a, al, ah); funpack2(b, bl, bh); funpack2(c, cl, ch);
  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(cl) : "f"(al), "f"(bl));
  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(ch) : "f"(ah), "f"(bh));
  return fpack2(cl, ch);
}
#else
__device__ __forceinline__ f32x2 dfma2(f32x2 a, f32x2 b, f32x2 c) { return ffma2(a, b, c); }
#endif

__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_newton(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}


}  // namespace cqr
