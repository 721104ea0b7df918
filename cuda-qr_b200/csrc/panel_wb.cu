// panel_wb.cu -- Householder panel (m_p x 64, LAPACK storage + explicit V + compact-WY T) on the warp-block layout.
//
// Same contract as panel_hh_cluster_kernel (panel_hh.cu; replaces the reference's one-CTA panelHouseholderKernel,
// qr.cu:60-333), different data layout: every WARP holds a 64-row x 64-column block of the panel in registers in the
// layout of the batched kernel (tsqr_flat.cu): lane = 8 h + q owns columns {q, q+8, ..} (8 slots) and the rows
// 8 k + 2 h + {0,1} (8 pairs) of its block.  Consequences, measured on the old kernel's trace:
//   * x = column j is published per warp (each warp only needs its own 64 rows of it) with a __syncwarp, read ONCE per
//     lane and reused for the lane's 8 columns -- the one-column-per-thread layout re-reads x from shared memory for
//     every column, twice per step (256 .. 1024 wavefront cycles per step and CTA);
//   * a column's dot needs two shuffle stages and one CTA barrier (cross-warp sum in front of the cluster exchange);
//   * the pivot row (global row j) lives in CTA 0 / warp 0, pair j / 8, so its register index is static per step group.
// The cluster all-reduce is the two-phase st.async scheme of panel_hh.cu (reduce-scatter to the column's owner CTA, totals
// and the pivot row back to every peer), driven here by warp 0 of each CTA.  Reflector scalars follow qr.c:144-152
// (MUFU rsqrt / rcp + one Newton step); a zero column gives tau = 0 (H = I).
//
// Status (round 1): validated against the parity tests with every panel routed here (CQR_PANEL_WB_MIN_ROWS=1,
// CQR_PANEL_WB_MAX_ROWS=16384); used by default between 3072 and 8192 rows: 166 us against 197 us at 8192 rows and 152
// against 162 at 4096, but 164 against 130 at 1024, where a CTA is a single warp, and only 221 against 229 at 16384, where
// the two clusters exchange through global-memory flags -- the step is dominated by the exchange (ncu: issue slots 15 %
// busy, the warps sit in the mbarrier wait and the CTA barrier), no longer by the local work.
#include "panel_wb_common.cuh"

namespace cqr {
namespace {

template <int W>
struct WbShared {
  float xs[W][2][64];          // per-warp x = column j of the warp's 64 rows
  float part[2][W][64];        // per-warp column sums x^T b_c
  float prl[2][64];            // CTA 0: row j of the panel
  float rs_in[2][16][4];       // phase 1 inbox of the column owner
  float prs_in[2][16][4];      // phase 1: pivot-row entries of my columns (from CTA 0)
  float tot_in[2][64];         // phase 2: cluster totals
  float prow[2][64];           // phase 2: row j
  float tot_x[2][64];          // two clusters: totals over both clusters, row j
  float prow_x[2][64];
  unsigned long long mbar1[2], mbar2[2];
  float gs[64][65];            // CTA 0: G(c, j) = v_c^T v_j (c < j)
  float ts[64][65];
  float staus[64];
};

struct WbCtx {
  int q, h, w, lane;
  bool top;                    // this warp holds the panel's first 64 rows (the pivot rows)
  unsigned rank, CS, cl, ncl;   // CTA rank in its cluster, cluster size, cluster index, number of clusters (1 or 2)
  int nb;
  int* err;
  float* tau_out;
  uint2* slots;                // two clusters: {value, tag} exchange slots in global memory [2 buffers][3][64]
  unsigned epoch;
};

// steps j = 8 I0 .. 8 I0 + 7
template <int I0, int W>
__device__ __forceinline__ void wb_steps(f32x2 (&b)[8][8], WbShared<W>& sm, const WbCtx& cx) {
  const int q = cx.q, h = cx.h, w = cx.w, lane = cx.lane;
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * I0 + jj;
    if (j >= cx.nb) break;                   // uniform over the whole cluster
    const int buf = j & 1;
    const unsigned par = (j >> 1) & 1;
    const int hj = jj >> 1;
    const bool hi_half = (jj & 1) != 0;
    float* xb = sm.xs[w][buf];
    // ---- publish x (this warp's rows of column j; the top warp masks the rows on and above the diagonal) and row j
    if (q == jj) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f32x2 v = b[I0][k];
        if (cx.top) {
          if (k < I0) v = 0ull;
          else if (k == I0) {
            float lo, hi;
            wupk(v, lo, hi);
            v = wpk(2 * h > jj ? lo : 0.f, 2 * h + 1 > jj ? hi : 0.f);
          }
        }
        *reinterpret_cast<f32x2*>(xb + 8 * k + 2 * h) = v;
      }
    }
    if (cx.top && h == hj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float lo, hi;
        wupk(b[i][I0], lo, hi);
        sm.prl[buf][q + 8 * i] = hi_half ? hi : lo;
      }
    }
    __syncwarp();
    f32x2 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = *reinterpret_cast<const f32x2*>(xb + 8 * k + 2 * h);
    // ---- warp-level column sums for all 64 columns (finished columns give V^T V for T)
    {
      f32x2 d2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d2[i] = 0ull;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) d2[i] = wfma2(x[k], b[i][k], d2[i]);
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = wsum2(d2[i]);
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __shfl_xor_sync(kFull, d[i], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] += t[i];
      }
      if (h == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.part[buf][w][q + 8 * i] = d[i];
      }
    }
    __syncthreads();
    // ---- CTA sums, then the cluster all-reduce (warp 0)
    if (w == 0) {
      float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), pv = make_float4(0.f, 0.f, 0.f, 0.f);
      const int l16 = lane & 15;
      if (lane < 16) {
#pragma unroll
        for (int ww = 0; ww < W; ++ww) {
          const float4 t = *reinterpret_cast<const float4*>(&sm.part[buf][ww][4 * l16]);
          sv.x += t.x; sv.y += t.y; sv.z += t.z; sv.w += t.w;
        }
      } else {
        pv = *reinterpret_cast<const float4*>(&sm.prl[buf][4 * l16]);
      }
      if (cx.CS == 1) {
        if (lane < 16) *reinterpret_cast<float4*>(&sm.tot_in[buf][4 * l16]) = sv;
        else *reinterpret_cast<float4*>(&sm.prow[buf][4 * l16]) = pv;
      } else {
        const unsigned CS = cx.CS, rank = cx.rank;
        const unsigned wpo = 16u / CS;                  // groups of four columns per owner CTA
        if (lane == 0) {
          wb_mbar_expect(&sm.mbar1[buf], (16 + (cx.cl == 0 ? wpo : 0)) * 16);
          wb_mbar_expect(&sm.mbar2[buf], (16 + (cx.cl == 0 ? 16 : 0)) * 16);
        }
        const unsigned owner = ((unsigned)l16 * CS) >> 4, wl = (unsigned)l16 - owner * wpo;
        if (lane < 16) wb_st_async_v4(&sm.rs_in[buf][rank * wpo + wl][0], &sm.mbar1[buf], owner, sv);
        else if (rank == 0 && cx.cl == 0) wb_st_async_v4(&sm.prs_in[buf][wl][0], &sm.mbar1[buf], owner, pv);
        wb_mbar_wait(&sm.mbar1[buf], par, cx.err);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 16) t = *reinterpret_cast<const float4*>(&sm.rs_in[buf][lane][0]);
        for (unsigned o = 8; o >= wpo; o >>= 1) {
          t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
          t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
        }
        const unsigned L = (unsigned)lane & 15u, slot = L % wpo, peer = L / wpo;
        float4 tv;
        tv.x = __shfl_sync(kFull, t.x, slot); tv.y = __shfl_sync(kFull, t.y, slot);
        tv.z = __shfl_sync(kFull, t.z, slot); tv.w = __shfl_sync(kFull, t.w, slot);
        const unsigned col4 = 4u * (rank * wpo + slot);
        if (lane < 16) wb_st_async_v4(&sm.tot_in[buf][col4], &sm.mbar2[buf], peer, tv);
        else if (cx.cl == 0) wb_st_async_v4(&sm.prow[buf][col4], &sm.mbar2[buf], peer, *reinterpret_cast<const float4*>(&sm.prs_in[buf][slot][0]));
      }
    }
    if (cx.CS == 1) __syncthreads();
    else wb_mbar_wait(&sm.mbar2[buf], par, cx.err);
    const float* tot = sm.tot_in[buf];
    const float* prw = sm.prow[buf];
    if (cx.ncl > 1) {
      // two clusters: the leaders publish their 64 cluster sums (cluster 0 also row j) as {value, tag} pairs in global
      // memory, warp 0 of every CTA polls the other cluster's slots (two columns per lane) and adds in the same order on
      // both sides (cluster 0 + cluster 1), so all 32 CTAs hold bit-identical totals
      if (w == 0) {
        const unsigned tag = cx.epoch * 64u + (unsigned)j + 1u;
        uint2* mys = cx.slots + ((size_t)buf * 3 + cx.cl) * 64;
        const uint2* oth = cx.slots + ((size_t)buf * 3 + (cx.cl ^ 1u)) * 64;
        uint2* piv = cx.slots + ((size_t)buf * 3 + 2) * 64;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = lane + 32 * e;
          const float mine = sm.tot_in[buf][c];
          if (cx.rank == 0) {
            wb_st_flag(mys + c, mine, tag);
            if (cx.cl == 0) wb_st_flag(piv + c, sm.prow[buf][c], tag);
          }
          float so = 0.f, po = 0.f;
          long long t0 = 0;
          for (;;) {
            const uint2 r = wb_ld_flag(oth + c);
            uint2 qv; qv.x = 0u; qv.y = tag;
            if (cx.cl != 0) qv = wb_ld_flag(piv + c);
            so = __uint_as_float(r.x); po = __uint_as_float(qv.x);
            if (r.y == tag && qv.y == tag) break;
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 2000000000LL) { atomicExch(cx.err, 1); break; }
          }
          sm.tot_x[buf][c] = (cx.cl == 0) ? (mine + so) : (so + mine);
          sm.prow_x[buf][c] = (cx.cl == 0) ? sm.prow[buf][c] : po;
        }
      }
      __syncthreads();
      tot = sm.tot_x[buf];
      prw = sm.prow_x[buf];
    }
    // ---- reflector scalars (redundant in every thread: bit-identical inputs)
    const float sig = tot[j], alpha = prw[j];
    const float sj = fmaf(alpha, alpha, sig);
    const bool ok = sj >= 1.2e-38f;          // like panel_hh.cu and the reference: a length-1 reflector (x = 0, alpha != 0) flips the sign, tau = 2
    const float sjs = ok ? sj : 1.f;
    const float rs = wrsqrt(sjs);
    float nrm = sjs * rs;
    nrm = fmaf(fmaf(-nrm, nrm, sjs), 0.5f * rs, nrm);
    const float bc = (alpha < 0.f) ? nrm : -nrm;
    const float u = alpha - bc;
    const float inv_u = ok ? wrcp(u) : 0.f;
    const float tau = ok ? -u * wrcp(bc) : 0.f;
    const float c2s = -tau * inv_u;
    if (cx.top && lane == 0) { cx.tau_out[j] = tau; sm.staus[j] = tau; }
    // v = x / u below the diagonal and v_j = 1: patch x's pivot entry to u so the update also writes a(j, c) -= tau s
    if (cx.top && h == hj) {
      float lo, hi;
      wupk(x[I0], lo, hi);
      x[I0] = hi_half ? wpk(lo, u) : wpk(u, hi);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = q + 8 * i;
      const float s = fmaf(tot[c], inv_u, prw[c]);      // v_j^T a_c  (for c < j: v_c^T v_j)
      if (i <= I0 && cx.top && h == 0 && c < j) sm.gs[c][j] = s;
      if (i >= I0) {
        const bool act = (i > I0) || (q > jj);
        const float nwv = act ? c2s * s : 0.f;
        const f32x2 nw2 = wpk(nwv, nwv);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[i][k] = wfma2(nw2, x[k], b[i][k]);
      }
    }
    if (q == jj && ok) {                     // store the reflector: x / u below the diagonal, beta on it, R above untouched
      const f32x2 iu2 = wpk(inv_u, inv_u);
      if (cx.top) {
        float lo, hi, xlo, xhi;
        wupk(b[I0][I0], lo, hi);
        wupk(x[I0], xlo, xhi);
        const int r0 = 2 * h, r1 = 2 * h + 1;
        lo = (r0 > jj) ? xlo * inv_u : (r0 == jj ? bc : lo);
        hi = (r1 > jj) ? xhi * inv_u : (r1 == jj ? bc : hi);
        b[I0][I0] = wpk(lo, hi);
#pragma unroll
        for (int k = I0 + 1; k < 8; ++k) b[I0][k] = wmul2(b[I0][k], iu2);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) b[I0][k] = wmul2(b[I0][k], iu2);
      }
    }
  }
}

template <int I0, int W>
struct WbGroups {
  static __device__ __forceinline__ void run(f32x2 (&b)[8][8], WbShared<W>& sm, const WbCtx& cx) {
    wb_steps<I0, W>(b, sm, cx);
    WbGroups<I0 + 1, W>::run(b, sm, cx);
  }
};
template <int W>
struct WbGroups<8, W> {
  static __device__ __forceinline__ void run(f32x2 (&)[8][8], WbShared<W>&, const WbCtx&) {}
};

template <int W>
__global__ void __launch_bounds__(32 * W, 1) panel_wb_kernel(PanelHHParams p) {
  __shared__ __align__(16) WbShared<W> sm;
  WbCtx cx;
  cx.lane = threadIdx.x & 31; cx.w = threadIdx.x >> 5; cx.q = cx.lane & 7; cx.h = cx.lane >> 3;
  cx.rank = wb_ctarank(); cx.CS = wb_nctarank();
  cx.cl = blockIdx.x / cx.CS; cx.ncl = gridDim.x / cx.CS;
  cx.nb = p.b; cx.err = p.err; cx.tau_out = p.tau; cx.slots = p.slots; cx.epoch = p.epoch;
  const int q = cx.q, h = cx.h;
  const long long gw = (long long)blockIdx.x * W + cx.w;        // warp block index: rows 64 gw .. 64 gw + 63
  cx.top = (gw == 0);
  const long long row0 = 64 * gw;
  const int b_cols = p.b;
  const bool vec = (p.lda % 2 == 0) && (p.ldv % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.vbuf) & 7) == 0);
  f32x2 b[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = q + 8 * i;
    const float* col = p.a + row0 + (long long)c * p.lda + 2 * h;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long r0 = row0 + 8 * k + 2 * h;
      if (c < b_cols && vec && r0 + 1 < p.mp) {
        b[i][k] = *reinterpret_cast<const f32x2*>(col + 8 * k);
      } else {
        const float lo = (c < b_cols && r0 < p.mp) ? col[8 * k] : 0.f;
        const float hi = (c < b_cols && r0 + 1 < p.mp) ? col[8 * k + 1] : 0.f;
        b[i][k] = wpk(lo, hi);
      }
    }
  }
  if (cx.CS > 1) {
    if (threadIdx.x == 0) {
      wb_mbar_init(&sm.mbar1[0], 1); wb_mbar_init(&sm.mbar1[1], 1);
      wb_mbar_init(&sm.mbar2[0], 1); wb_mbar_init(&sm.mbar2[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    wb_cluster_sync();   // every peer's barriers exist before anyone pushes
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) sm.staus[threadIdx.x] = 0.f;

  WbGroups<0, W>::run(b, sm, cx);

  // ---- results: LAPACK storage into the panel, explicit V (unit diagonal, zeros above) into vbuf
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = q + 8 * i;
    float* acol = p.a + row0 + (long long)c * p.lda + 2 * h;
    float* vcol = p.vbuf + row0 + (long long)c * p.ldv + 2 * h;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long r0 = row0 + 8 * k + 2 * h;
      float lo, hi;
      wupk(b[i][k], lo, hi);
      const float vlo = r0 > c ? lo : (r0 == c ? 1.f : 0.f);
      const float vhi = r0 + 1 > c ? hi : (r0 + 1 == c ? 1.f : 0.f);
      if (c < b_cols && vec && r0 + 1 < p.mp) {
        *reinterpret_cast<f32x2*>(acol + 8 * k) = b[i][k];
        *reinterpret_cast<f32x2*>(vcol + 8 * k) = wpk(vlo, vhi);
      } else {
        if (c < b_cols && r0 < p.mp) { acol[8 * k] = lo; vcol[8 * k] = vlo; }
        if (c < b_cols && r0 + 1 < p.mp) { acol[8 * k + 1] = hi; vcol[8 * k + 1] = vhi; }
      }
    }
  }
  // ---- compact-WY T (CTA 0): T^-1 = diag(1 / tau) + striu(V^T V), inverted by blocks on all threads (wb_build_t)
  if (blockIdx.x == 0 && p.t != nullptr) {
    __syncthreads();
    const int nb = p.b, nt = 32 * W;
    wb_build_t(sm.gs, sm.ts, sm.staus, nb, threadIdx.x, nt);
    for (int idx = threadIdx.x; idx < nb * nb; idx += nt) {
      const int i = idx % nb, cc = idx / nb;
      p.t[i + (long long)cc * p.ldt] = (i <= cc) ? sm.gs[i][cc] : 0.f;
    }
  }
  if (cx.CS > 1) wb_cluster_sync();   // no CTA leaves while pushes addressed to it (or by it) are in flight
}

template <int W>
cudaError_t launch_wb_t(const PanelHHParams& p, int cs, int ncl, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(panel_wb_kernel<W>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * ncl, 1, 1);
  cfg.blockDim = dim3(32 * W, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, panel_wb_kernel<W>, p);
}

}  // namespace

// Warps per CTA, cluster size and cluster count for an m_p-row panel: one cluster of up to 16 CTAs x 8 warps x 64 rows up
// to 8192 rows, two such clusters (exchange through global-memory flags) up to 16384.
bool panel_wb_plan(long long mp, int* wpc, int* cs, int* ncl) {
  if (mp < 1 || mp > 16384) return false;
  *ncl = 1;
  if (mp > 8192) { *wpc = 8; *cs = 16; *ncl = 2; return true; }
  const int nw = (int)((mp + 63) / 64);
  // CQR_PANEL_WB_MIN_WPC: fewest warps per CTA above 512 rows (1, 2, 4, 8): wider CTAs mean fewer CTAs in the exchange
  static const int min_wpc = getenv("CQR_PANEL_WB_MIN_WPC") ? atoi(getenv("CQR_PANEL_WB_MIN_WPC")) : 1;
  int w = 1;
  if (nw > 8 && (min_wpc == 2 || min_wpc == 4 || min_wpc == 8)) w = min_wpc;
  if (nw <= 8) {                 // up to 512 rows: one CTA, no cluster exchange at all
    while (w < nw) w *= 2;
    *wpc = w; *cs = 1;
    return true;
  }
  while ((nw + w - 1) / w > 16) w *= 2;
  const int P = (nw + w - 1) / w;
  int c = 1;
  while (c < P) c *= 2;
  *wpc = w; *cs = c;
  return true;
}

bool launch_panel_wb(const PanelHHParams& p, int wpc, int cs, int ncl, cudaStream_t s) {
  ++g_launches;
  cudaError_t e;
  if (wpc == 1) e = launch_wb_t<1>(p, cs, ncl, s);
  else if (wpc == 2) e = launch_wb_t<2>(p, cs, ncl, s);
  else if (wpc == 4) e = launch_wb_t<4>(p, cs, ncl, s);
  else e = launch_wb_t<8>(p, cs, ncl, s);
  if (e != cudaSuccess) { cudaGetLastError(); --g_launches; return false; }
  return true;
}

}  // namespace cqr
