// panel_hh.cu -- multi-CTA Householder factorisation of one m_p x b panel (b <= 64) in ONE launch.
//
// Replaces, for the blocked square/rectangular path, the reference's serial window sweep over a
// column block (qr.cu:505-546: one 1-CTA panelHouseholderKernel launch per 64 x 4 window) and this
// library's own first-generation panel (TSQR tree -> explicit thin Q -> Householder reconstruction,
// ~10 launches and ~0.5 ms per panel at m_p = 16384).  Here the panel is distributed by row slabs over
// P co-resident CTAs that hold their slab in REGISTERS for the whole factorisation; a column step is
//   1. owner warp broadcasts the pivot column x (rows >= j) through shared memory,
//   2. every warp forms the slab-local dots x^T a_c for its 4 columns with warp-shuffle reductions
//      (north_star item 1; reflector maths as qr.c:144-167: beta = -sign*norm, u = x0 - beta, tau = -u/beta),
//   3. the P slab partials are exchanged through a flag-tagged slot array in global memory (8-byte
//      {value, tag} stores, readers poll the tag -- no fence, no atomics, no separate counter), lane i
//      of every warp fetching CTA i's partial so the cross-CTA sum is one more shuffle reduction
//      (fixed order => bitwise reproducible),
//   4. every CTA derives beta/u/tau redundantly and updates its rows: a_c -= tau v (v^T a_c).
// The same dots taken against the already finished columns c < j give G(c, j) = v_c^T v_j, i.e. the
// strict upper triangle of V^T V, from which build_t_kernel forms the panel's compact-WY T (qr.c:170-213's
// W/Y accumulation in LAPACK larft form) -- no separate Gram GEMM.  Output is directly LAPACK geqrf
// storage (R on/above the diagonal, v below, tau) plus the explicit unit-lower V the tensor-core update
// reads, so no tree, no thin Q and no reconstruction are needed.
//
// Layout: 512 threads = 16 warps.  Warp w owns columns {w, w+16, w+32, w+48}; lane l owns slab rows
// {l, l+32, ...} (RI of them).  CTAs spin on each other: the grid must be co-resident (P <= #SMs, one
// CTA per SM); spins are bounded by a clock64 timeout that raises an error flag instead of hanging.
#include "common.cuh"
#include <stdlib.h>

namespace cqr {

namespace {

__device__ __forceinline__ void st_flag(uint2* p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_flag(const uint2* p) {
  uint2 r;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}

constexpr long long kSpinTimeout = 6000000000LL;   // ~3 s of SM clocks: co-residency never arrived

}  // namespace

// Optional phase trace (debug builds only: -DCQR_HH_TRACE): clock64 of lane 0 of every warp of CTA 0 and the
// last CTA at six points of each column step -> g_hh_trace[cta01][warp][step][6].
#ifdef CQR_HH_TRACE
__device__ long long g_hh_trace[2][16][64][6];
#define HH_TRACE(k)                                                                         \
  do {                                                                                      \
    if (l == 0 && (cta == 0 || cta == P - 1)) g_hh_trace[cta == 0 ? 0 : 1][w][j][k] = clock64(); \
  } while (0)
#else
#define HH_TRACE(k) do { } while (0)
#endif

template <int RI>
__global__ void __launch_bounds__(512, 1) panel_hh_kernel(PanelHHParams p) {
  constexpr int TH = 32 * RI;
  __shared__ float xs[2][TH];
  __shared__ float sc[2][4];            // {beta, 1/u, tau, u} of the current step, written by the owner warp
  __shared__ float gs[64][65];          // CTA 0: G(c, j) = v_c^T v_j (c < j), then reused for T
  __shared__ float ts[64][65];
  __shared__ float staus[64];
  const int P = gridDim.x, cta = blockIdx.x;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const long long row0 = (long long)cta * TH;
  const long long rem = p.mp - row0;
  const int rows = rem >= TH ? TH : (int)rem;
  const int b = p.b;
  float* __restrict__ A = p.a + row0;
  uint2* slots = p.slots;                                   // [2][pmax][64]
  uint2* prow = p.slots + 2 * (size_t)p.pmax * 64;          // [2][64]  row j of the panel, published by CTA 0

  float a[4][RI];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = w + 16 * k;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      a[k][ri] = (c < b && r < rows) ? A[r + (long long)c * p.lda] : 0.f;
    }
  }

#pragma unroll
  for (int sj = 0; sj < 4; ++sj) {
    for (int wj = 0; wj < 16; ++wj) {
      const int j = 16 * sj + wj;          // pivot column: warp wj, register slot sj
      if (j >= b) break;
      const int buf = j & 1;
      const int rj = sj >> 1;              // panel row j lives in CTA 0, lane j % 32, register slot j / 32
      const unsigned tag = p.epoch * 64u + (unsigned)j + 1u;
      if (w == wj) {
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const int r = l + 32 * ri;
          xs[buf][r] = (row0 + r >= j) ? a[sj][ri] : 0.f;
        }
      }
      __syncthreads();
      HH_TRACE(0);
      float x[RI];
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) x[ri] = xs[buf][l + 32 * ri];
      // slab-local dots of x with all four owned columns (finished ones feed G, live ones the update; the
      // owner's own column gives x^T x)
      float tot[4], ajc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tot[k] = 0.f;
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) tot[k] = fmaf(x[ri], a[k][ri], tot[k]);
      }
      warp_sum_n(tot);
      HH_TRACE(1);
      if (P > 1) {
        uint2* myslot = slots + ((size_t)buf * p.pmax + cta) * 64;
        if (l < 4) {
          const float v = l == 0 ? tot[0] : (l == 1 ? tot[1] : (l == 2 ? tot[2] : tot[3]));
          st_flag(myslot + w + 16 * l, v, tag);
        }
        if (cta == 0 && l == (j & 31)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) st_flag(prow + buf * 64 + w + 16 * k, a[k][rj], tag);
        }
        // gather: lane i sums the partials of CTAs i, i+32, ... for this warp's four columns, and every lane
        // fetches the warp's four pivot-row entries.  All eight slots are polled TOGETHER (one L2 round trip
        // per attempt).  Each slot is read by one warp per CTA only: no hot addresses.
#pragma unroll
        for (int k = 0; k < 4; ++k) tot[k] = 0.f;
        const uint2* pr = prow + buf * 64;
        uint2 q[4];
        for (int i0 = 0; i0 < P; i0 += 32) {
          const int i = i0 + l;
          const bool have = i < P;
          const uint2* src = slots + ((size_t)buf * p.pmax + (have ? i : 0)) * 64;
          uint2 r[4];
          long long t0 = 0;
          for (;;) {
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = ld_flag(src + w + 16 * k);
            bool ok = true;
            if (i0 == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k) q[k] = ld_flag(pr + w + 16 * k);
#pragma unroll
              for (int k = 0; k < 4; ++k) ok = ok && (q[k].y == tag);
            }
            if (have) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ok = ok && (r[k].y == tag);
            }
            if (__all_sync(kFull, ok)) break;
            if (t0 == 0) t0 = clock64();
            if (*(volatile int*)p.err != 0) break;
            if (clock64() - t0 > kSpinTimeout) { atomicExch(p.err, 1); break; }
          }
          if (have) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tot[k] += __uint_as_float(r[k].x);
          }
        }
        warp_sum_n(tot);
#pragma unroll
        for (int k = 0; k < 4; ++k) ajc[k] = __uint_as_float(q[k].x);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) ajc[k] = __shfl_sync(kFull, a[k][rj], j & 31);
      }
      HH_TRACE(2);
      // the owner warp turns (x^T x, alpha) into the reflector scalars once and hands them to the other
      // fifteen warps through shared memory (no redundant sqrt / divisions on the issue slots)
      if (w == wj) {
        const float sjt = tot[sj], alpha = ajc[sj];
        float beta = 0.f, tau = 0.f, inv_u = 0.f, u = 1.f;
        if (sjt != 0.f) {
          const float nrm = sqrtf(sjt);
          beta = (alpha < 0.f) ? nrm : -nrm;
          u = alpha - beta;
          inv_u = 1.f / u;
          tau = -u / beta;
        }
        if (l == 0) { sc[buf][0] = beta; sc[buf][1] = inv_u; sc[buf][2] = tau; sc[buf][3] = u; }
      }
      HH_TRACE(3);
      __syncthreads();
      HH_TRACE(4);
      const float beta = sc[buf][0], inv_u = sc[buf][1], tau = sc[buf][2], u = sc[buf][3];
      const bool nz = inv_u != 0.f;
      if (RI >= 16) {   // big slabs: do not carry x across the exchange (register pressure), re-read it
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) x[ri] = xs[buf][l + 32 * ri];
      }
      // v = x / u with v_j = 1: fold 1/u into the column scalars and patch x_j := u
      if (cta == 0 && l == (j & 31)) x[rj] = u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = w + 16 * k;
        const float d = nz ? (tot[k] - beta * ajc[k]) * inv_u : ajc[k];   // v^T a_c (zero column: v = e_j)
        const bool live = (k > sj) || (k == sj && w > wj);
        if (live) {
          const float wc = tau * d * inv_u;
#pragma unroll
          for (int ri = 0; ri < RI; ++ri) a[k][ri] = fmaf(-wc, x[ri], a[k][ri]);
        } else if (c < j && cta == 0 && l == 0) {
          gs[c][j] = d;   // G(c, j) = v_c^T v_j
        }
      }
      if (w == wj) {
        if (nz) {
#pragma unroll
          for (int ri = 0; ri < RI; ++ri) {
            const long long gr = row0 + l + 32 * ri;
            if (gr > j) a[sj][ri] = x[ri] * inv_u;
            else if (gr == j) a[sj][ri] = beta;
          }
        }
        if (cta == 0 && l == 0) { p.tau[j] = tau; staus[j] = tau; }
      }
      HH_TRACE(5);
    }
  }

  float* __restrict__ V = p.vbuf + row0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = w + 16 * k;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      if (c < b && r < rows) {
        const long long gr = row0 + r;
        A[r + (long long)c * p.lda] = a[k][ri];
        V[r + (long long)c * p.ldv] = gr > c ? a[k][ri] : (gr == c ? 1.f : 0.f);
      }
    }
  }

  // Compact-WY T of the panel (replaces qr.c:170-213's W accumulation), CTA 0 only.  With G = striu(V^T V):
  //   T(c, c) = tau_c,   T(i, c) = -tau_i * sum_{k = i+1..c} G(i, k) T(k, c)     (i = c-1 .. 0)
  // i.e. T = (striu(G) + diag(1/tau))^-1 by back substitution, tau_i = 0 giving a zero row/column (H_i = I).
  // Column c is handled by 8 lanes of one warp (dot over k split 8 ways) and only ever reads its own column
  // of T, so the 64 columns run without any block-level synchronisation.
  if (cta == 0 && p.t != nullptr) {
    __syncthreads();
    const int c = threadIdx.x >> 3, q8 = threadIdx.x & 7;     // 64 columns x 8 lanes; a warp holds columns 4w .. 4w+3
    if (q8 == 0 && c < b) ts[c][c] = staus[c];
    __syncwarp();
    for (int i = 4 * w + 2; i >= 0; --i) {                    // warp-uniform trip count: shuffles stay convergent
      const bool act = i < c && c < b;
      float acc = 0.f;
      if (act)
        for (int k = i + 1 + q8; k <= c; k += 8) acc = fmaf(gs[i][k], ts[k][c], acc);
      acc += __shfl_xor_sync(kFull, acc, 1);
      acc += __shfl_xor_sync(kFull, acc, 2);
      acc += __shfl_xor_sync(kFull, acc, 4);
      if (act && q8 == 0) ts[i][c] = -staus[i] * acc;
      __syncwarp();
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < b * b; idx += 512) {
      const int i = idx % b, cc = idx / b;
      p.t[i + (long long)cc * p.ldt] = (i <= cc) ? ts[i][cc] : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster variant: the slab CTAs form one (m_p <= 8192) or two (m_p <= 16384) thread-block clusters of up to 16
// CTAs.  Inside a cluster the slab-local dots are all-gathered through distributed shared memory with one-sided
// 16-byte st.async pushes that credit the receiver's mbarrier (16 x 16 transactions per CTA and step; the first
// version pushed single floats and was bound by DSMEM transaction rate: ~1400 cycles per step, tools/hh_trace.py).
// Two clusters exchange their 64 cluster sums (and the pivot row) through flag-tagged 8-byte slots in global
// memory: one L2 round trip between exactly two parties.  All CTAs add the partials in the same order, so every CTA
// derives bitwise identical reflector scalars.
// Thread layout: a thread owns ONE column (c = tid / 8) and the rows {4 (g + 8 ch) + e} of its slab (g = tid % 8,
// RR rows in chunks of 4) held as packed f32x2 register pairs, so the dot and the rank-1 update run on FFMA2
// (two fp32 FMAs per issue slot: the 3-operand scalar FFMA issues at half rate on sm_100) and a dot product needs
// only a 3-stage shuffle reduction over the 8 row groups.
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One-sided 16-byte push into a peer CTA's shared memory that also credits 16 bytes on the peer's mbarrier (the
// receiver just waits on its own barrier: no cluster-wide barrier, no fence).
__device__ __forceinline__ void st_async_v4(float* local_dst, unsigned long long* local_bar, unsigned rank, float4 v) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local_dst), lb = (unsigned)__cvta_generic_to_shared(local_bar), ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rb)
               : "memory");
}
__device__ __forceinline__ void mbar_init_local(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_local(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_local(unsigned long long* bar, unsigned parity) {
  unsigned ok, a = (unsigned)__cvta_generic_to_shared(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// packed fp32 pairs (FFMA2)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int RR>
__global__ void __launch_bounds__(512, 1) panel_hh_cluster_kernel(PanelHHParams p) {
  constexpr int TH = 8 * RR, NCH = RR / 4;
  __shared__ __align__(16) float xs[2][TH];
  __shared__ __align__(16) float rs_in[2][16][4];    // phase 1 inbox of the column owner: [source CTA x slot][4 columns]
  __shared__ __align__(16) float prs_in[2][16][4];   // phase 1: pivot-row entries of my columns (from slab 0)
  __shared__ __align__(16) float tot_in[2][64];      // phase 2: cluster totals of all 64 columns
  __shared__ __align__(16) float prow[2][64];        // phase 2: row j of the panel (cluster 0 only)
  __shared__ unsigned long long mbar1[2], mbar2[2];
  __shared__ float sc[2][4];            // {beta, 1/u, tau, u}
  __shared__ float gs[64][65];          // CTA 0: G(c, j) = v_c^T v_j (c < j)
  __shared__ float ts[64][65];
  __shared__ float staus[64];
  const unsigned rank = cluster_ctarank(), CS = cluster_nctarank();
  const unsigned cta = blockIdx.x;                 // slab index over the whole grid
  const unsigned cl = cta / CS;                    // cluster index (0 or 1)
  const unsigned ncl = gridDim.x / CS;
  const int P = (int)gridDim.x; (void)P;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int c = threadIdx.x >> 3, g = l & 7;       // my column, my row group
  const long long row0 = (long long)cta * TH;
  const long long rem = p.mp - row0;
  const int rows = rem >= TH ? TH : (rem > 0 ? (int)rem : 0);
  const int b = p.b;
  float* __restrict__ Ac = p.a + row0 + (long long)c * p.lda;
  float* __restrict__ Vc = p.vbuf + row0 + (long long)c * p.ldv;
  const bool vec_ok = (p.lda % 4 == 0) && (p.ldv % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.vbuf) & 15) == 0);
  uint2* xslot = p.slots;                          // [2 buffers][3: cluster 0 sums, cluster 1 sums, pivot row][64]

  f32x2 a[RR / 2];                                 // a[2 ch] = rows 4(g+8ch)+{0,1}, a[2 ch + 1] = rows +{2,3}
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int r0 = 4 * (g + 8 * ch);
    float4 v;
    if (c < b && vec_ok && r0 + 3 < rows) {
      v = *reinterpret_cast<const float4*>(Ac + r0);
    } else {
      v.x = (c < b && r0 + 0 < rows) ? Ac[r0 + 0] : 0.f;
      v.y = (c < b && r0 + 1 < rows) ? Ac[r0 + 1] : 0.f;
      v.z = (c < b && r0 + 2 < rows) ? Ac[r0 + 2] : 0.f;
      v.w = (c < b && r0 + 3 < rows) ? Ac[r0 + 3] : 0.f;
    }
    a[2 * ch] = pack2(v.x, v.y);
    a[2 * ch + 1] = pack2(v.z, v.w);
  }
  if (CS > 1) {
    if (threadIdx.x == 0) {
      mbar_init_local(&mbar1[0], 1);
      mbar_init_local(&mbar1[1], 1);
      mbar_init_local(&mbar2[0], 1);
      mbar_init_local(&mbar2[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();   // every peer's barriers exist before anyone pushes
  }

#pragma unroll
  for (int sj = 0; sj < 4; ++sj) {
    for (int wj = 0; wj < 16; ++wj) {
      const int j = 16 * sj + wj;
      if (j >= b) break;
      const int buf = j & 1;
      const int chj = sj >> 1;                 // row j: chunk j / 32 (static after unrolling), group (j / 4) % 8, element j % 4
      const int gj = (j >> 2) & 7, ej = j & 3;
      const bool owner = (c == j);
      const int jrel = j - (int)row0;          // pivot row in slab-local numbering (negative below slab 0); row0 < 2^31 here
      const unsigned tag = p.epoch * 64u + (unsigned)j + 1u;
      if (owner) {
        // x = column j below (and including) the diagonal.  Only slab 0 holds rows above it, all inside chunks 0 and 1.
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float4 v;
          unpack2(a[2 * ch], v.x, v.y);
          unpack2(a[2 * ch + 1], v.z, v.w);
          if (ch < 2 && cta == 0) {
            const int r0 = 4 * (g + 8 * ch);
            v.x = (r0 + 0 >= jrel) ? v.x : 0.f;
            v.y = (r0 + 1 >= jrel) ? v.y : 0.f;
            v.z = (r0 + 2 >= jrel) ? v.z : 0.f;
            v.w = (r0 + 3 >= jrel) ? v.w : 0.f;
          }
          *reinterpret_cast<float4*>(&xs[buf][4 * (g + 8 * ch)]) = v;
        }
      }
      __syncthreads();
      HH_TRACE(0);
      // dot in batches of up to four chunks: the LDS of a batch are issued together so their latency overlaps
      f32x2 acc[4] = {0ull, 0ull, 0ull, 0ull};
      constexpr int BAT = NCH < 4 ? NCH : 4;
#pragma unroll
      for (int ch0 = 0; ch0 < NCH; ch0 += BAT) {
        ulonglong2 xv[BAT];
#pragma unroll
        for (int k = 0; k < BAT; ++k) xv[k] = *reinterpret_cast<const ulonglong2*>(&xs[buf][4 * (g + 8 * (ch0 + k))]);
#pragma unroll
        for (int k = 0; k < BAT; ++k) {
          acc[(2 * k) & 3] = fma2(xv[k].x, a[2 * (ch0 + k)], acc[(2 * k) & 3]);
          acc[(2 * k + 1) & 3] = fma2(xv[k].y, a[2 * (ch0 + k) + 1], acc[(2 * k + 1) & 3]);
        }
      }
      float s;
      {
        float s0, s1, s2, s3, s4, s5, s6, s7;
        unpack2(acc[0], s0, s1);
        unpack2(acc[1], s2, s3);
        unpack2(acc[2], s4, s5);
        unpack2(acc[3], s6, s7);
        s = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
      }
      s += __shfl_xor_sync(kFull, s, 1);
      s += __shfl_xor_sync(kFull, s, 2);
      s += __shfl_xor_sync(kFull, s, 4);
      float mine;   // a(j, c) if this thread holds row j (slab 0, g == gj)
      {
        float m0, m1, m2, m3;
        unpack2(a[2 * chj], m0, m1);
        unpack2(a[2 * chj + 1], m2, m3);
        mine = ej == 0 ? m0 : (ej == 1 ? m1 : (ej == 2 ? m2 : m3));
      }
      float ajc = __shfl_sync(kFull, mine, (l & 24) | gj);   // valid on slab 0
      HH_TRACE(1);
      if (CS > 1) {
        // All-reduce inside the cluster in two one-sided phases (st.async, 16 B, crediting the receiver's mbarrier):
        //   1. reduce-scatter: warp w sends its four column sums to the CTA that owns those columns
        //      (owner = w CS / 16; slab 0 also sends the four pivot-row entries),
        //   2. the owner's warp 0 adds the CS contributions per column (shuffle tree, fixed order) and sends the
        //      totals (and the pivot-row entries) to every peer.
        // 16 + 16 incoming messages per CTA and step instead of the 256 of a direct all-gather: DSMEM delivers
        // ~1 message per 4 cycles per CTA, so the all-gather cost ~1550 cycles per step against ~950 for this
        // (tools/probes/dsmem_probe.cu).
        const unsigned wpo = 16u / CS;                  // warps (groups of four columns) per owner CTA
        const unsigned par = (j >> 1) & 1;
        if (threadIdx.x == 0) {
          mbar_expect_tx_local(&mbar1[buf], (16 + (cl == 0 ? wpo : 0)) * 16);
          mbar_expect_tx_local(&mbar2[buf], (16 + (cl == 0 ? 16 : 0)) * 16);
        }
        const unsigned owner = ((unsigned)w * CS) >> 4, wl = (unsigned)w - owner * wpo;
        float4 sv, pv;
        sv.x = __shfl_sync(kFull, s, 0); sv.y = __shfl_sync(kFull, s, 8); sv.z = __shfl_sync(kFull, s, 16); sv.w = __shfl_sync(kFull, s, 24);
        if (l == 0) st_async_v4(&rs_in[buf][rank * wpo + wl][0], &mbar1[buf], owner, sv);
        if (cta == 0) {
          pv.x = __shfl_sync(kFull, ajc, 0); pv.y = __shfl_sync(kFull, ajc, 8); pv.z = __shfl_sync(kFull, ajc, 16); pv.w = __shfl_sync(kFull, ajc, 24);
          if (l == 1) st_async_v4(&prs_in[buf][wl][0], &mbar1[buf], owner, pv);
        }
        if (w == 0) {
          mbar_wait_local(&mbar1[buf], par);
          // lane i < 16 holds contribution i = src * wpo + slot: add over src (bits >= log2 wpo of the lane index)
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < 16) t = *reinterpret_cast<const float4*>(&rs_in[buf][l][0]);
          for (unsigned o = 8; o >= wpo; o >>= 1) {
            t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
            t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
          }
          // lane L < 16: totals of slot L % wpo to peer L / wpo; lane 16 + L: the pivot entries of that slot
          const unsigned L = (unsigned)l & 15u, slot = L % wpo, peer = L / wpo;
          float4 tv;
          tv.x = __shfl_sync(kFull, t.x, slot); tv.y = __shfl_sync(kFull, t.y, slot);
          tv.z = __shfl_sync(kFull, t.z, slot); tv.w = __shfl_sync(kFull, t.w, slot);
          const unsigned col4 = 4u * (rank * wpo + slot);
          if (l < 16) st_async_v4(&tot_in[buf][col4], &mbar2[buf], peer, tv);
          else if (cl == 0) st_async_v4(&prow[buf][col4], &mbar2[buf], peer, *reinterpret_cast<const float4*>(&prs_in[buf][slot][0]));
        }
        mbar_wait_local(&mbar2[buf], par);
        s = tot_in[buf][c];
        if (cl == 0) ajc = prow[buf][c];
      }
      if (ncl > 1) {
        // two clusters: the cluster leaders publish their 64 cluster sums (cluster 0 also the pivot row) as
        // {value, tag} pairs; every CTA polls the other cluster's slots for its own columns
        uint2* mys = xslot + ((size_t)buf * 3 + cl) * 64;
        const uint2* oth = xslot + ((size_t)buf * 3 + (cl ^ 1u)) * 64;
        uint2* piv = xslot + ((size_t)buf * 3 + 2) * 64;
        if (rank == 0 && g == 0) {
          st_flag(mys + c, s, tag);
          if (cl == 0) st_flag(piv + c, ajc, tag);
        }
        float so = 0.f, po = 0.f;
        if (g == 0) {
          long long t0 = 0;
          for (;;) {
            const uint2 r = ld_flag(oth + c);
            uint2 q; q.x = 0u; q.y = tag;
            if (cl != 0) q = ld_flag(piv + c);
            so = __uint_as_float(r.x); po = __uint_as_float(q.x);
            if (r.y == tag && q.y == tag) break;
            if (t0 == 0) t0 = clock64();
            if (*(volatile int*)p.err != 0) break;
            if (clock64() - t0 > kSpinTimeout) { atomicExch(p.err, 1); break; }
          }
        }
        so = __shfl_sync(kFull, so, l & 24);
        po = __shfl_sync(kFull, po, l & 24);
        s = (cl == 0) ? (s + so) : (so + s);   // same order on both sides: cluster 0 + cluster 1
        if (cl != 0) ajc = po;
      }
      HH_TRACE(2);
      if (owner && g == 0) {   // reflector scalars once per CTA (qr.c:144-152), handed over through shared memory
        const float sjt = s, alpha = ajc;
        float beta = 0.f, tau = 0.f, inv_u = 0.f, u = 1.f;
        if (sjt != 0.f) {
          const float nrm = sqrtf(sjt);
          beta = (alpha < 0.f) ? nrm : -nrm;
          u = alpha - beta;
          inv_u = 1.f / u;
          tau = -u / beta;
        }
        sc[buf][0] = beta; sc[buf][1] = inv_u; sc[buf][2] = tau; sc[buf][3] = u;
        if (cta == 0) { p.tau[j] = tau; staus[j] = tau; }
      }
      HH_TRACE(3);
      __syncthreads();
      HH_TRACE(4);
      const float beta = sc[buf][0], inv_u = sc[buf][1], tau = sc[buf][2], u = sc[buf][3];
      const bool nz = inv_u != 0.f;
      const float d = nz ? (s - beta * ajc) * inv_u : ajc;   // v^T a_c (zero column: v = e_j)
      if (c > j) {
        // a_c -= tau (v^T a_c) v with v = x / u and v_j = 1: 1/u is folded into the column scalar and x_j patched to u
        const float nwc = -(tau * d * inv_u);
        const f32x2 nw2 = pack2(nwc, nwc);
        constexpr int BATU = NCH < 4 ? NCH : 4;
#pragma unroll
        for (int ch0 = 0; ch0 < NCH; ch0 += BATU) {
          ulonglong2 xv[BATU];
#pragma unroll
          for (int k = 0; k < BATU; ++k) xv[k] = *reinterpret_cast<const ulonglong2*>(&xs[buf][4 * (g + 8 * (ch0 + k))]);
#pragma unroll
          for (int k = 0; k < BATU; ++k) {
            const int ch = ch0 + k;
            if (ch == chj && cta == 0 && g == gj) {
              float x0, x1, x2, x3;
              unpack2(xv[k].x, x0, x1); unpack2(xv[k].y, x2, x3);
              if (ej == 0) x0 = u; else if (ej == 1) x1 = u; else if (ej == 2) x2 = u; else x3 = u;
              xv[k].x = pack2(x0, x1); xv[k].y = pack2(x2, x3);
            }
            a[2 * ch] = fma2(xv[k].x, nw2, a[2 * ch]);
            a[2 * ch + 1] = fma2(xv[k].y, nw2, a[2 * ch + 1]);
          }
        }
      } else if (c < j) {
        if (cta == 0 && g == 0) gs[c][j] = d;   // G(c, j) = v_c^T v_j
      } else if (nz) {
        const f32x2 iu2 = pack2(inv_u, inv_u);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&xs[buf][4 * (g + 8 * ch)]);
          if (ch < 2 && cta == 0) {
            const int r0 = 4 * (g + 8 * ch);
            float xe[4], ae[4];
            unpack2(xv.x, xe[0], xe[1]); unpack2(xv.y, xe[2], xe[3]);
            unpack2(a[2 * ch], ae[0], ae[1]); unpack2(a[2 * ch + 1], ae[2], ae[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (r0 + e > jrel) ae[e] = xe[e] * inv_u;
              else if (r0 + e == jrel) ae[e] = beta;
            }
            a[2 * ch] = pack2(ae[0], ae[1]); a[2 * ch + 1] = pack2(ae[2], ae[3]);
          } else {
            a[2 * ch] = mul2(xv.x, iu2);
            a[2 * ch + 1] = mul2(xv.y, iu2);
          }
        }
      }
      HH_TRACE(5);
    }
  }

#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int r0 = 4 * (g + 8 * ch);
    const long long gr0 = row0 + r0;
    float av[4], v[4];
    unpack2(a[2 * ch], av[0], av[1]);
    unpack2(a[2 * ch + 1], av[2], av[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = gr0 + e > c ? av[e] : (gr0 + e == c ? 1.f : 0.f);
    if (c < b && vec_ok && r0 + 3 < rows) {
      *reinterpret_cast<float4*>(Ac + r0) = make_float4(av[0], av[1], av[2], av[3]);
      *reinterpret_cast<float4*>(Vc + r0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c < b && r0 + e < rows) { Ac[r0 + e] = av[e]; Vc[r0 + e] = v[e]; }
    }
  }

  // compact-WY T by back substitution (see panel_hh_kernel), CTA 0 only
  if (cta == 0 && p.t != nullptr) {
    __syncthreads();
    if (g == 0 && c < b) ts[c][c] = staus[c];
    __syncwarp();
    for (int i = 4 * w + 2; i >= 0; --i) {
      const bool act = i < c && c < b;
      float acc = 0.f;
      if (act)
        for (int k = i + 1 + g; k <= c; k += 8) acc = fmaf(gs[i][k], ts[k][c], acc);
      acc += __shfl_xor_sync(kFull, acc, 1);
      acc += __shfl_xor_sync(kFull, acc, 2);
      acc += __shfl_xor_sync(kFull, acc, 4);
      if (act && g == 0) ts[i][c] = -staus[i] * acc;
      __syncwarp();
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < b * b; idx += 512) {
      const int i = idx % b, cc = idx / b;
      p.t[i + (long long)cc * p.ldt] = (i <= cc) ? ts[i][cc] : 0.f;
    }
  }
  if (CS > 1) cluster_sync_all();   // no CTA leaves while pushes addressed to it (or by it) are in flight
}

// Rows-per-thread, cluster size and cluster count for an m_p-row panel on the cluster kernel (m_p <= 2 * 16 * 512).
bool panel_hh_cluster_plan(long long mp, int* rr, int* cs, int* ncl) {
  if (mp > 2 * 16 * 512) return false;
  if (mp > 16 * 512) { *rr = 64; *cs = 16; *ncl = 2; return true; }
  int r = 8;
  {
    // short panels: one CTA holding the whole panel (no cluster exchange at all) up to `single` rows (tuning knob)
    static const long long single = getenv("CQR_PANEL_SINGLE_ROWS") ? atoll(getenv("CQR_PANEL_SINGLE_ROWS")) : 256;
    if (mp <= single && mp <= 512) {
      while (8 * r < mp) r *= 2;
      *rr = r; *cs = 1; *ncl = 1;
      return true;
    }
  }
  while ((mp + 8 * r - 1) / (8 * r) > 16) r *= 2;
  const int P = (int)((mp + 8 * r - 1) / (8 * r));
  int c = 1;
  while (c < P) c *= 2;
  *rr = r; *cs = c; *ncl = 1;
  return true;
}

template <int RR>
static cudaError_t launch_cluster_t(const PanelHHParams& p, int cs, int ncl, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(panel_hh_cluster_kernel<RR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * ncl, 1, 1);
  cfg.blockDim = dim3(512, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, panel_hh_cluster_kernel<RR>, p);
}

bool launch_panel_hh_cluster(const PanelHHParams& p, int rr, int cs, int ncl, cudaStream_t s) {
  ++g_launches;
  cudaError_t e;
  if (rr == 8) e = launch_cluster_t<8>(p, cs, ncl, s);
  else if (rr == 16) e = launch_cluster_t<16>(p, cs, ncl, s);
  else if (rr == 32) e = launch_cluster_t<32>(p, cs, ncl, s);
  else e = launch_cluster_t<64>(p, cs, ncl, s);
  if (e != cudaSuccess) { cudaGetLastError(); --g_launches; return false; }
  return true;
}

// Slab height and CTA count for an m_p-row panel; false if it does not fit the co-residency budget.
bool panel_hh_plan(long long mp, int max_ctas, int* ri, int* ctas) {
  int r = 2;
  if (mp > 4096) r = 16;
  else if (mp > 2048) r = 8;
  else if (mp > 1024) r = 4;
  const long long P = (mp + 32 * r - 1) / (32 * r);
  if (P > max_ctas || P > kPanelHHMaxCtas) return false;
  *ri = r; *ctas = (int)P;
  return true;
}

#ifdef CQR_HH_TRACE
void panel_hh_read_trace(long long* out) { cudaMemcpyFromSymbol(out, g_hh_trace, sizeof(g_hh_trace)); }
#endif

void launch_panel_hh(const PanelHHParams& p, int ri, int ctas, cudaStream_t s) {
  ++g_launches;
  if (ri == 2) panel_hh_kernel<2><<<ctas, 512, 0, s>>>(p);
  else if (ri == 4) panel_hh_kernel<4><<<ctas, 512, 0, s>>>(p);
  else if (ri == 8) panel_hh_kernel<8><<<ctas, 512, 0, s>>>(p);
  else panel_hh_kernel<16><<<ctas, 512, 0, s>>>(p);
}

}  // namespace cqr
