// panel_wb_common.cuh -- device helpers shared by the warp-block panel kernels (panel_wb.cu, panel_wb2.cu): packed
// f32x2 arithmetic, MUFU rsqrt / rcp with one Newton step, cluster rank / barrier, st.async pushes that credit the
// receiver's mbarrier, bounded mbarrier wait, {value, tag} flag slots in global memory for the two-cluster exchange.
#pragma once
#include "common.cuh"

namespace cqr {
namespace {

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 wpk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void wupk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#ifdef CQR_WB_SCALAR_FMA   // experiment: two FFMA instead of one FFMA2 (an FFMA2 with three distinct 64-bit sources issues every 3 cycles)
__device__ __forceinline__ f32x2 wfma2(f32x2 a, f32x2 b, f32x2 c) {
  float al, ah, bl, bh, cl, ch;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(al), "=f"(ah) : "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(bl), "=f"(bh) : "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(cl), "=f"(ch) : "l"(c));
  float dl, dh;
  asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(dl) : "f"(al), "f"(bl), "f"(cl));
  asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(dh) : "f"(ah), "f"(bh), "f"(ch));
  f32x2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(dl), "f"(dh));
  return d;
}
#else
__device__ __forceinline__ f32x2 wfma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif
__device__ __forceinline__ f32x2 wmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float wsum2(f32x2 v) { float lo, hi; wupk(v, lo, hi); return lo + hi; }
__device__ __forceinline__ float wrsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float wrcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return fmaf(r, fmaf(-x, r, 1.f), r); }

__device__ __forceinline__ unsigned wb_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned wb_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void wb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void wb_st_async_v4(float* local_dst, unsigned long long* local_bar, unsigned rank, float4 v) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local_dst), lb = (unsigned)__cvta_generic_to_shared(local_bar), ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rb)
               : "memory");
}
__device__ __forceinline__ void wb_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void wb_mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wb_st_flag(uint2* p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 wb_ld_flag(const uint2* p) {
  uint2 r;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
// four consecutive {value, tag} slots, polled until all carry `tag` (bounded like the mbarrier wait below)
__device__ __forceinline__ float4 wb_poll4(const uint2* p, unsigned tag, int* err) {
  long long t0 = 0;
  for (;;) {
    const uint2 a = wb_ld_flag(p), b = wb_ld_flag(p + 1), c = wb_ld_flag(p + 2), d = wb_ld_flag(p + 3);
    if (a.y == tag && b.y == tag && c.y == tag && d.y == tag)
      return make_float4(__uint_as_float(a.x), __uint_as_float(b.x), __uint_as_float(c.x), __uint_as_float(d.x));
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 2000000000LL) { atomicExch(err, 1); return make_float4(0.f, 0.f, 0.f, 0.f); }
  }
}
// bounded wait: a protocol error must not hang the device; *err is set and the caller's results are void
__device__ __forceinline__ void wb_mbar_wait(unsigned long long* bar, unsigned parity, int* err) {
  unsigned ok, a = (unsigned)__cvta_generic_to_shared(bar);
  long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 2000000000LL) { atomicExch(err, 1); break; }
  }
}

// Compact-WY T of a panel from G = striu(V^T V) (gs[i][k], i < k) and tau, by blocks, on all `nt` threads of the CTA:
//   T(c,c) = tau_c,  T(i,c) = -tau_i sum_{k=i+1..c} G(i,k) T(k,c)      (T^-1 = diag(1/tau) + striu(G); tau_i = 0 -> zero row)
// restricted to the 8-column diagonal blocks (one thread per column, <= 28 dependent steps), then three merge levels
//   T12 = -T11 (G12 T22)                                                (larft's recurrence on blocks of 8, 16, 32 columns)
// with one thread per output element.  The column-at-a-time back substitution this replaces took 40 K cycles of the
// panel kernel's 250 K (2016 dependent shared-memory FMAs in the last column); this form takes about 5 K.
// On return T's upper triangle is in gs (G is overwritten), ts is scratch.  Must be called by all threads of the CTA.
__device__ __forceinline__ void wb_build_t(float (*gs)[65], float (*ts)[65], const float* staus, int nb, int tid, int nt) {
  for (int c = tid; c < nb; c += nt) {
    const int c0 = c & ~7;
    ts[c][c] = staus[c];
    for (int i = c - 1; i >= c0; --i) {
      float acc = 0.f;
      for (int k = i + 1; k <= c; ++k) acc = fmaf(gs[i][k], ts[k][c], acc);
      ts[i][c] = -staus[i] * acc;
    }
  }
  __syncthreads();
  for (int o = tid; o < 64 * 8; o += nt) {       // diagonal blocks of T into gs (their G entries are dead)
    const int c = o & 63, i = (c & ~7) + (o >> 6);
    if (i <= c && c < nb) gs[i][c] = ts[i][c];
  }
  __syncthreads();
  for (int s = 8; s < 64; s *= 2) {
    const int sh = (s == 8) ? 3 : (s == 16 ? 4 : 5);
    const int nout = 32 * s;                      // (64 / 2s) merges x s x s outputs, consecutive threads on consecutive columns
    for (int o = tid; o < nout; o += nt) {        // Y(r,c) = sum_{k = mid..c} G(r,k) T(k,c)
      const int cc = o & (s - 1), rr = (o >> sh) & (s - 1), base = (o >> (2 * sh)) * 2 * s;
      const int r = base + rr, mid = base + s, c = mid + cc;
      if (c < nb) {
        float acc = 0.f;
        for (int k = mid; k <= c; ++k) acc = fmaf(gs[r][k], gs[k][c], acc);
        ts[r][c] = acc;
      }
    }
    __syncthreads();
    for (int o = tid; o < nout; o += nt) {        // T12(r,c) = -sum_{k = r..mid-1} T(r,k) Y(k,c)   (over the dead G12)
      const int cc = o & (s - 1), rr = (o >> sh) & (s - 1), base = (o >> (2 * sh)) * 2 * s;
      const int r = base + rr, mid = base + s, c = mid + cc;
      if (c < nb) {
        float acc = 0.f;
        for (int k = r; k < mid; ++k) acc = fmaf(gs[r][k], ts[k][c], acc);
        gs[r][c] = -acc;
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace cqr
