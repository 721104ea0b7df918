// panel_wb2.cu -- Householder panel (m_p x 64) on the warp-block layout with TWO pivot columns per cluster exchange.
//
// Default for panels of 2048 .. 8192 rows (CQR_PANEL_PAIR=0 goes back to one column per exchange, =2 forces the
// cancellation fallback in every pair, which exercises the single-step code of this file).  Same contract,
// register layout and st.async all-reduce as panel_wb.cu (which replaces the reference's one-CTA panelHouseholderKernel,
// qr.cu:60-333); what changes is the number of exchanges: panel_wb.cu's ncu capture shows the 64 exchanges, not the local
// work, are the run time (profiles/r01_panel_wb8192_summary.txt), so here one exchange serves the reflectors j and j+1.
//
// Algebra (fp32 numpy spec with the failure cases: tools/two_column_step.py).  x = column j, y = column j+1, both BEFORE
// reflector j touches anything and both restricted to the rows below j+1.  One exchange delivers, for every column c,
// p_c = x^T a_c and q_c = y^T a_c plus the rows j and j+1 of the panel (r1, r2).  Then
//   reflector j:    sigma_1 = p_j + r2_j^2, alpha_1 = r1_j               -> beta_1, u_1, tau_1   (qr.c:144-152 convention)
//   any column c:   d1_c = r1_c + (p_c + r2_j r2_c) / u_1,  t_c = tau_1 d1_c,  e_c = t_c / u_1
//   column j+1:     a1 = e_{j+1};  y' = y - a1 x;  alpha_2 = r2_{j+1} - a1 r2_j;
//                   sigma_2 = q_{j+1} - 2 a1 p_{j+1} + a1^2 p_j                             -> beta_2, u_2, tau_2
//   any column c:   a'(j+1,c) = r2_c - e_c r2_j;  y'^T a'_c = q_c - e_c p_{j+1} - a1 p_c + a1 e_c p_j;
//                   d2_c = a'(j+1,c) + y'^T a'_c / u_2,  f_c = tau_2 d2_c / u_2
//   update:         a_c -= (e_c - f_c a1) x + f_c y   below row j+1 (two FFMA2 per register pair, the work of two single
//                   steps);  a(j,c) = r1_c - t_c,  a(j+1,c) = a'(j+1,c) - tau_2 d2_c
//   V^T V for T:    the same d1_c, d2_c of the finished columns (with e_c = 0), and G(j,j+1) = r2_j/u_1 + (p_{j+1} - a1 p_j)/(u_1 u_2).
// Cancellation guard (mandatory): when sigma_2 < 0.1 q_{j+1} the expansion loses too much (neighbouring
// columns nearly dependent); the pair then finishes step j from the data it has and runs step j+1 as a single step with
// one more exchange of the true y'.  The decision is taken on bit-identical totals in every thread of the cluster, so it is
// uniform.  b must be 64.  One cluster up to 8192 rows (default); two clusters up to 16384 rows exchange through
// global-memory flags exactly like panel_wb.cu (CQR_PANEL_PAIR_MAX_ROWS=16384, not yet measured: off by default).  Measured on B200 (tools/panel_bench.py, zero fill + panel):
// 8192 rows 166 -> 132 us, 4096 rows 152 -> 121 us, 2048 rows 120 us; the lane-level numpy replay of this file is
// tools/emulate_pair_panel.py.
#include "panel_wb_common.cuh"

namespace cqr {
namespace {

template <int W>
struct Wb2Shared {
  float xs[W][2][64];          // per-warp x (column j, or the lone column j+1 of a fallen-back pair) of the warp's 64 rows
  float ys[W][2][64];          // per-warp y (column j+1)
  float part[2][W][128];       // per-warp column sums: [0,64) x^T a_c, [64,128) y^T a_c
  float prl[2][128];           // CTA 0: rows j and j+1 of the panel
  float rs_in[2][32][4];       // phase 1 inbox of the group owner (one float4 per sender and owned group)
  float prs_in[2][32][4];      // phase 1: pivot-row entries of my groups (from CTA 0)
  float tot_in[2][128];        // phase 2: cluster totals p (0..63), q (64..127)
  float prow[2][128];          // phase 2: rows j (0..63), j+1 (64..127)
  unsigned long long mbar1[2], mbar2[2];
  float gs[64][65];            // CTA 0: G(c, j) = v_c^T v_j (c < j)
  float ts[64][65];
  float staus[64];
};

struct Wb2Ctx {
  int q, h, w, lane;
  bool top;                    // this warp holds the panel's first 64 rows (the pivot rows)
  unsigned rank, CS, cl, ncl;  // CTA rank in its cluster, cluster size, cluster index, number of clusters (1 or 2)
  uint2* slots;                // two clusters: {value, tag} exchange slots in global memory [2 buffers][32 CTAs + pivot rows][128]
  unsigned epoch;
  int nb, mode;                // mode 2: every pair takes the fallback
  int* err;
  float* tau_out;
};

// Cancellation guard: sigma_2 and every y'^T a'_c come from an expansion whose relative error grows like
// q_{j+1} / sigma_2.  Measured on the fp32 spec with EVERY pair at the same angle (tools/two_column_step.py `guard_study`):
// orthogonality 2.4 n eps at sigma_2 / q = 0.1, 8 at 0.03, 24 at 0.01, 177 at 0.0015 (acceptance: 10) -- so pairs below 0.1
// take the fallback; uniform[0,1) data sits at 0.25 in its first pair (the common mean) and near 1 afterwards.
constexpr float kPairGuard = 0.1f;

// Optional phase trace (debug builds only: make TRACE=1 -> -DCQR_HH_TRACE; tools/wb2_trace.py): clock64 of lane 0 of every
// warp of CTA 0 and of the last CTA at eight points of each exchange step -> g_wb2_trace[cta01][warp][exchange][8], and at
// five points of the kernel -> g_wb2_marks[cta01][warp][5] (entry, panel loaded, steps done, results stored, T done).
#ifdef CQR_HH_TRACE
__device__ long long g_wb2_trace[2][8][64][8];
__device__ long long g_wb2_marks[2][8][5];
#define WB2_TRACE(k)                                                                                          \
  do {                                                                                                        \
    if (cx.lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                    \
      g_wb2_trace[blockIdx.x == 0 ? 0 : 1][cx.w][(ex - 1) & 63][k] = clock64();                              \
  } while (0)
#define WB2_MARK(k)                                                                                           \
  do {                                                                                                        \
    if (cx.lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                    \
      g_wb2_marks[blockIdx.x == 0 ? 0 : 1][cx.w][k] = clock64();                                             \
  } while (0)
#else
#define WB2_TRACE(k) do { } while (0)
#define WB2_MARK(k) do { } while (0)
#endif

struct Refl { float bc, inv_u, tau; bool ok; };
// beta = -sign(alpha) norm, u = alpha - beta, tau = -u / beta (qr.c:144-152); MUFU rsqrt / rcp + one Newton step as in
// panel_wb.cu; a zero column gives tau = 0 (H = I)
__device__ __forceinline__ Refl wb2_scalars(float alpha, float sig) {
  Refl r;
  const float sj = fmaf(alpha, alpha, sig);
  r.ok = sj >= 1.2e-38f;
  const float sjs = r.ok ? sj : 1.f;
  const float rs = wrsqrt(sjs);
  float nrm = sjs * rs;
  nrm = fmaf(fmaf(-nrm, nrm, sjs), 0.5f * rs, nrm);
  const bool len1 = (sig == 0.f);             // nothing below the pivot: exact sign flip, tau = 2 like the reference's division
  r.bc = len1 ? -alpha : ((alpha < 0.f) ? nrm : -nrm);
  const float u = alpha - r.bc;
  r.inv_u = r.ok ? wrcp(u) : 0.f;
  r.tau = r.ok ? (len1 ? 2.f : -u * wrcp(r.bc)) : 0.f;
  return r;
}

// Columns j = 8 I0 .. 8 I0 + 7, two per exchange; ex counts the exchanges of this launch (buffer and mbarrier parity).
// ONE instantiation serves all eight column groups: the caller rotates the register block between groups so that the
// pivot group is always b[0][.] (register group i holds the columns q + 8 ((i + I0) & 7)) and, in the warp that owns
// the pivot rows, the pivot row block is always b[.][0] (register row block k holds the rows 8 ((k + I0) & 7) ..).  With
// I0 a template parameter the kernel was 17 K instructions (276 KB), every group ran its own copy four times, and the
// ncu capture had 31-34 % of the stall samples in `no_instruction` (instruction-cache misses).  Rows above the pivot
// are zero in the published x and y, so the dots see the same terms in the same order as before.
template <int W, bool X2>
__device__ __forceinline__ void wb2_steps(f32x2 (&b)[8][8], Wb2Shared<W>& sm, const Wb2Ctx& cx, unsigned& ex, const int I0) {
  const int q = cx.q, h = cx.h, w = cx.w, lane = cx.lane;
  const int nact = 8 - I0;                   // register groups / (pivot warp) row blocks still in play: indices 0 .. nact - 1
  int jj = 0;
  bool second = false;                       // this pass is the lone step j+1 of a pair that failed the guard
#pragma unroll 1
  while (jj < 8) {
    const int j = 8 * I0 + jj;               // even
    if (j >= cx.nb) break;                   // uniform over the whole cluster
    const int buf = ex & 1;
    const unsigned par = (ex >> 1) & 1;
    ++ex;
    WB2_TRACE(0);
    const int hj = jj >> 1;                  // rows j, j+1 are the (lo, hi) of row block 0 in the lanes h == hj of the pivot warp
    const int qx = second ? jj + 1 : jj, qy = jj + 1;
    float* xb = sm.xs[w][buf];
    float* yb = sm.ys[w][buf];
    // ---- publish x and y (this warp's rows; the pivot warp masks rows 0 .. j+1) and rows j, j+1
    if (q == qx || q == qy) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f32x2 v = b[0][k];
        if (cx.top && (k >= nact || (k == 0 && h <= hj))) v = 0ull;
        if (q == qx) *reinterpret_cast<f32x2*>(xb + 8 * k + 2 * h) = v;
        if (q == qy) *reinterpret_cast<f32x2*>(yb + 8 * k + 2 * h) = v;
      }
    }
    if (cx.top && h == hj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float lo, hi;
        wupk(b[i][0], lo, hi);
        sm.prl[buf][q + 8 * i] = lo;           // exchange vectors are indexed by the ROTATED column position q + 8 i
        sm.prl[buf][64 + q + 8 * i] = hi;
      }
    }
    __syncwarp();
    f32x2 x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = *reinterpret_cast<const f32x2*>(xb + 8 * k + 2 * h);
      y[k] = *reinterpret_cast<const f32x2*>(yb + 8 * k + 2 * h);
    }
    WB2_TRACE(1);
    // ---- warp-level column sums for all 64 columns against x and y.  First shuffle stage swaps halves (odd h keeps the
    // y sums, even h the x sums), so the 16 sums cost 16 shuffles; lanes h == 0 end with x^T a_c, lanes h == 1 with y^T a_c
    {
      f32x2 d2[8], e2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { d2[i] = 0ull; e2[i] = 0ull; }
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          d2[i] = wfma2(x[k], b[i][k], d2[i]);
          e2[i] = wfma2(y[k], b[i][k], e2[i]);
        }
      const bool odd = (h & 1) != 0;
      float keep[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dx = wsum2(d2[i]), dy = wsum2(e2[i]);
        const float give = odd ? dx : dy;
        keep[i] = (odd ? dy : dx) + __shfl_xor_sync(kFull, give, 8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) keep[i] += __shfl_xor_sync(kFull, keep[i], 16);
      if (h < 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.part[buf][w][64 * h + q + 8 * i] = keep[i];
      }
    }
    WB2_TRACE(2);
    __syncthreads();
    WB2_TRACE(3);
    // ---- CTA sums, then the cluster all-reduce (warp 0): lane g handles the float4 group g of the 128 sums
    if (w == 0) {
      float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ww = 0; ww < W; ++ww) {
        const float4 t = *reinterpret_cast<const float4*>(&sm.part[buf][ww][4 * lane]);
        sv.x += t.x; sv.y += t.y; sv.z += t.z; sv.w += t.w;
      }
      const float4 pv = *reinterpret_cast<const float4*>(&sm.prl[buf][4 * lane]);   // meaningful in CTA 0 only
      if (cx.CS == 1) {
        *reinterpret_cast<float4*>(&sm.tot_in[buf][4 * lane]) = sv;
        *reinterpret_cast<float4*>(&sm.prow[buf][4 * lane]) = pv;
      } else {
        const unsigned CS = cx.CS, rank = cx.rank;
        const unsigned wpo = 32u / CS;                  // float4 groups per owner CTA
        if (lane == 0) {
          wb_mbar_expect(&sm.mbar1[buf], (32 + ((!X2 || cx.cl == 0) ? wpo : 0)) * 16);
          wb_mbar_expect(&sm.mbar2[buf], (32 + 32) * 16);
        }
        const unsigned owner = ((unsigned)lane * CS) >> 5, wl = (unsigned)lane - owner * wpo;
        wb_st_async_v4(&sm.rs_in[buf][rank * wpo + wl][0], &sm.mbar1[buf], owner, sv);
        if (rank == 0 && (!X2 || cx.cl == 0)) wb_st_async_v4(&sm.prs_in[buf][wl][0], &sm.mbar1[buf], owner, pv);
        // Two clusters: the same CTA sums also go to global memory as {value, tag} slots (one row of 128 per CTA and
        // buffer, row 32 = the pivot rows), where the OWNER of each column group in the other cluster collects them while
        // its own cluster's reduce-scatter is in flight: the cross-cluster hop (an L2 round trip, ~900 cycles) overlaps
        // phase 1 instead of following phase 2 (it used to add ~1800 cycles to each of the 32 steps of a 16384-row panel).
        unsigned tag = 0;
        if constexpr (X2) {
          tag = cx.epoch * 64u + ex;                     // ex = 1 .. 64 inside a launch, epoch unique per launch
          uint2* mine = cx.slots + ((size_t)buf * 33 + cx.cl * CS + rank) * 128 + 4 * lane;
          wb_st_flag(mine + 0, sv.x, tag); wb_st_flag(mine + 1, sv.y, tag);
          wb_st_flag(mine + 2, sv.z, tag); wb_st_flag(mine + 3, sv.w, tag);
          if (cx.cl == 0 && rank == 0) {
            uint2* pvs = cx.slots + ((size_t)buf * 33 + 32) * 128 + 4 * lane;
            wb_st_flag(pvs + 0, pv.x, tag); wb_st_flag(pvs + 1, pv.y, tag);
            wb_st_flag(pvs + 2, pv.z, tag); wb_st_flag(pvs + 3, pv.w, tag);
          }
        }
        const unsigned slot = (unsigned)lane % wpo, peer = (unsigned)lane / wpo;   // after the reduction a lane holds the total of group `slot`
        const unsigned col4 = 4u * (rank * wpo + slot);
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f), pg = u;
        if constexpr (X2) {
          // the other cluster's share of my groups (lane = sender * wpo + group, as in rs_in) and, in cluster 1, rows j, j+1:
          // polled BEFORE the wait for the own cluster's contributions, so the two latencies overlap
          u = wb_poll4(cx.slots + ((size_t)buf * 33 + (cx.cl ^ 1u) * CS + peer) * 128 + col4, tag, cx.err);
          if (cx.cl != 0) pg = wb_poll4(cx.slots + ((size_t)buf * 33 + 32) * 128 + col4, tag, cx.err);
        }
        wb_mbar_wait(&sm.mbar1[buf], par, cx.err);
        WB2_TRACE(4);
        float4 t = *reinterpret_cast<const float4*>(&sm.rs_in[buf][lane][0]);    // entry = sender * wpo + my group
        for (unsigned o = 16; o >= wpo; o >>= 1) {
          t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
          t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
        }
        float4 pr4 = *reinterpret_cast<const float4*>(&sm.prs_in[buf][slot][0]);   // rows j, j+1 of my groups (cluster 0)
        if constexpr (X2) {
          // same shuffle tree as for the own cluster's sums, so both clusters form bit-identical per-cluster sums and
          // add them in the same order (cluster 0 + cluster 1)
          for (unsigned o = 16; o >= wpo; o >>= 1) {
            u.x += __shfl_xor_sync(kFull, u.x, o); u.y += __shfl_xor_sync(kFull, u.y, o);
            u.z += __shfl_xor_sync(kFull, u.z, o); u.w += __shfl_xor_sync(kFull, u.w, o);
          }
          if (cx.cl == 0) { t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
          else { t.x = u.x + t.x; t.y = u.y + t.y; t.z = u.z + t.z; t.w = u.w + t.w; }
          if (cx.cl != 0) pr4 = pg;
        }
        wb_st_async_v4(&sm.tot_in[buf][col4], &sm.mbar2[buf], peer, t);
        wb_st_async_v4(&sm.prow[buf][col4], &sm.mbar2[buf], peer, pr4);
      }
    }
    if (cx.CS == 1) __syncthreads();
    else wb_mbar_wait(&sm.mbar2[buf], par, cx.err);
    const float* P = sm.tot_in[buf];
    const float* R1 = sm.prow[buf];
    const float* Q = P + 64;
    const float* R2 = R1 + 64;
    WB2_TRACE(5);

    if (!second) {
      // ---- reflector j and what it does to column j+1 (redundant in every thread: bit-identical inputs)
      // (column j sits at rotated position jj)
      const float xj1 = R2[jj], yj1 = R2[jj + 1];
      const float pj = P[jj], pj1 = P[jj + 1], qj1 = Q[jj + 1];
      const Refl s1 = wb2_scalars(R1[jj], fmaf(xj1, xj1, pj));
      const float iu1 = s1.inv_u, t1 = s1.tau;
      const float d1n = fmaf(fmaf(xj1, yj1, pj1), iu1, R1[jj + 1]);
      const float tn = t1 * d1n;
      const float a1 = tn * iu1;
      const float alpha2 = fmaf(-a1, xj1, yj1);
      const float sig2 = fmaf(a1 * a1, pj, fmaf(-2.f * a1, pj1, qj1));
      const bool fb = (cx.mode == 2) || (sig2 < kPairGuard * qj1);
      if (cx.top && lane == 0) { cx.tau_out[j] = t1; sm.staus[j] = t1; }
      WB2_TRACE(6);
      if (!fb) {
        const Refl s2 = wb2_scalars(alpha2, fmaxf(sig2, 0.f));
        const float iu2 = s2.inv_u, t2 = s2.tau;
        if (cx.top && lane == 0) { cx.tau_out[j + 1] = t2; sm.staus[j + 1] = t2; }
        // pass 1, all eight groups without a branch (loads and scalar chains of the groups overlap): the two update
        // coefficients per column, the patched rows j, j+1 (x and y are zero there, so the FMAs below leave them alone)
        // and the new entries of V^T V; pass 2, the FMAs of the groups that are still in play
        float cxs[8], cys[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int cp = q + 8 * i;                                           // rotated position of column c
          const float pc = P[cp], qc = Q[cp], r1c = R1[cp], r2v = R2[cp];
          const float d1 = fmaf(fmaf(xj1, r2v, pc), iu1, r1c);                // v_j^T a_c  (c < j: v_c^T v_j)
          if ((i == 0 || i >= nact) && cx.top && h == 0) {
            const int c = q + 8 * ((i + I0) & 7);
            if (c < j) { sm.gs[c][j] = d1; sm.gs[c][j + 1] = fmaf(fmaf(-a1, pc, qc), iu2, r2v); }
            else if (c == j) sm.gs[j][j + 1] = fmaf(iu1 * iu2, fmaf(-a1, pj, pj1), xj1 * iu1);
          }
          const bool act = (i < nact) && ((i > 0) || (q > jj + 1));
          const float tc = t1 * d1, ec = tc * iu1;
          const float r2c = fmaf(-ec, xj1, r2v);                              // a'(j+1, c)
          const float inner = fmaf(a1 * ec, pj, fmaf(-a1, pc, fmaf(-ec, pj1, qc)));
          const float sc2 = t2 * fmaf(inner, iu2, r2c), fc = sc2 * iu2;
          cxs[i] = act ? fmaf(fc, a1, -ec) : 0.f;
          cys[i] = act ? -fc : 0.f;
          if (act && cx.top && h == hj) b[i][0] = wpk(r1c - tc, r2c - sc2);   // rows j, j+1
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nact) {
            const f32x2 cx2 = wpk(cxs[i], cxs[i]), cy2 = wpk(cys[i], cys[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) b[i][k] = wfma2(cx2, x[k], wfma2(cy2, y[k], b[i][k]));
          }
        }
        if (q == jj + 1) {                   // column j+1: R(j, j+1), beta_2, v = (y - a1 x) / u_2 below
          const float m2 = s2.ok ? iu2 : 1.f;
          const f32x2 m22 = wpk(m2, m2), na2 = wpk(-a1, -a1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const f32x2 nv = wmul2(wfma2(na2, x[k], b[0][k]), m22);
            if (!cx.top || (k > 0 && k < nact) || (k == 0 && h > hj)) b[0][k] = nv;
          }
          if (cx.top && h == hj) b[0][0] = wpk(R1[jj + 1] - tn, s2.ok ? s2.bc : alpha2);
        }
      } else {
        // ---- cancellation fallback, first half: step j alone from the data of this exchange (x^T a_c over the rows
        // below j is p_c + r2_j r2_c); step j+1 follows as a single step with its own exchange
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int cp = q + 8 * i;
          const float r1c = R1[cp], r2v = R2[cp];
          const float d1 = fmaf(fmaf(xj1, r2v, P[cp]), iu1, r1c);
          if ((i == 0 || i >= nact) && cx.top && h == 0) {
            const int c = q + 8 * ((i + I0) & 7);
            if (c < j) sm.gs[c][j] = d1;
          }
          if (i < nact) {
            const bool act = (i > 0) || (q > jj);
            const float tc = t1 * d1, ec = tc * iu1;
            const float cxv = act ? -ec : 0.f;
            const f32x2 cx2 = wpk(cxv, cxv);
#pragma unroll
            for (int k = 0; k < 8; ++k) b[i][k] = wfma2(cx2, x[k], b[i][k]);
            if (act && cx.top && h == hj) b[i][0] = wpk(r1c - tc, fmaf(-ec, xj1, r2v));
          }
        }
      }
      if (q == jj && s1.ok) {                // column j: beta_1 on the diagonal, v = x / u_1 below (row j+1 included)
        const f32x2 iu12 = wpk(iu1, iu1);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!cx.top || (k > 0 && k < nact) || (k == 0 && h > hj)) b[0][k] = wmul2(b[0][k], iu12);
        if (cx.top && h == hj) b[0][0] = wpk(s1.bc, xj1 * iu1);
      }
      if (fb) second = true;
      else jj += 2;
      WB2_TRACE(7);
    } else {
      // ---- lone step j+1: x is the true column j+1 below row j+1, R2 the current row j+1
      const Refl s2 = wb2_scalars(R2[jj + 1], P[jj + 1]);
      const float iu2 = s2.inv_u, t2 = s2.tau;
      if (cx.top && lane == 0) { cx.tau_out[j + 1] = t2; sm.staus[j + 1] = t2; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cp = q + 8 * i;
        const float r2v = R2[cp];
        const float d2 = fmaf(P[cp], iu2, r2v);
        if ((i == 0 || i >= nact) && cx.top && h == 0) {
          const int c = q + 8 * ((i + I0) & 7);
          if (c < j + 1) sm.gs[c][j + 1] = d2;
        }
        if (i < nact) {
          const bool act = (i > 0) || (q > jj + 1);
          const float sc2 = t2 * d2, fc = sc2 * iu2;
          const float cxv = act ? -fc : 0.f;
          const f32x2 cx2 = wpk(cxv, cxv);
#pragma unroll
          for (int k = 0; k < 8; ++k) b[i][k] = wfma2(cx2, x[k], b[i][k]);
          if (act && cx.top && h == hj) {
            float lo, hi;
            wupk(b[i][0], lo, hi);
            b[i][0] = wpk(lo, r2v - sc2);
          }
        }
      }
      if (q == jj + 1 && s2.ok) {
        const f32x2 iu22 = wpk(iu2, iu2);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!cx.top || (k > 0 && k < nact) || (k == 0 && h > hj)) b[0][k] = wmul2(b[0][k], iu22);
        if (cx.top && h == hj) {
          float lo, hi;
          wupk(b[0][0], lo, hi);
          b[0][0] = wpk(lo, s2.bc);
        }
      }
      second = false;
      jj += 2;
      WB2_TRACE(7);
    }
  }
}

// Next column group: register group i takes over from i + 1 (the finished group goes to the end), and in the pivot warp
// row block k from k + 1 as well.  After eight rotations the block is back in its original order.
__device__ __forceinline__ void wb2_rotate(f32x2 (&b)[8][8], bool top) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const f32x2 t = b[0][k];
#pragma unroll
    for (int i = 0; i < 7; ++i) b[i][k] = b[i + 1][k];
    b[7][k] = t;
  }
  if (top) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const f32x2 t = b[i][0];
#pragma unroll
      for (int k = 0; k < 7; ++k) b[i][k] = b[i][k + 1];
      b[i][7] = t;
    }
  }
}

extern __shared__ __align__(16) unsigned char wb2_smem[];

template <int W, bool X2>
__global__ void __launch_bounds__(32 * W, 1) panel_wb2_kernel(PanelHHParams p) {
  Wb2Shared<W>& sm = *reinterpret_cast<Wb2Shared<W>*>(wb2_smem);
  Wb2Ctx cx;
  cx.lane = threadIdx.x & 31; cx.w = threadIdx.x >> 5; cx.q = cx.lane & 7; cx.h = cx.lane >> 3;
  cx.rank = wb_ctarank(); cx.CS = wb_nctarank();
  cx.cl = blockIdx.x / cx.CS; cx.ncl = gridDim.x / cx.CS;
  cx.nb = p.b; cx.err = p.err; cx.tau_out = p.tau; cx.mode = p.pmax; cx.slots = p.slots; cx.epoch = p.epoch;
  const int q = cx.q, h = cx.h;
  const long long gw = (long long)blockIdx.x * W + cx.w;        // warp block index: rows 64 gw .. 64 gw + 63
  cx.top = (gw == 0);
  const long long row0 = 64 * gw;
  const int b_cols = p.b;
  const bool vec = (p.lda % 2 == 0) && (p.ldv % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 7) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.vbuf) & 7) == 0);
  WB2_MARK(0);
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

  unsigned ex = 0;
  WB2_MARK(1);
#pragma unroll 1
  for (int I0 = 0; I0 < 8; ++I0) {
    wb2_steps<W, X2>(b, sm, cx, ex, I0);
    wb2_rotate(b, cx.top);
  }
  WB2_MARK(2);

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
  WB2_MARK(3);
  // ---- compact-WY T (CTA 0), as in panel_wb.cu: T^-1 = diag(1 / tau) + striu(V^T V), inverted by blocks (wb_build_t)
  if (blockIdx.x == 0 && p.t != nullptr) {
    __syncthreads();
    const int nb = p.b, nt = 32 * W;
    wb_build_t(sm.gs, sm.ts, sm.staus, nb, threadIdx.x, nt);
    for (int idx = threadIdx.x; idx < nb * nb; idx += nt) {
      const int i = idx % nb, cc = idx / nb;
      p.t[i + (long long)cc * p.ldt] = (i <= cc) ? sm.gs[i][cc] : 0.f;
    }
  }
  WB2_MARK(4);
  if (cx.CS > 1) wb_cluster_sync();   // no CTA leaves while pushes addressed to it (or by it) are in flight
}

template <int W, bool X2>
cudaError_t launch_wb2_t(const PanelHHParams& p, int cs, int ncl, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(panel_wb2_kernel<W, X2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(panel_wb2_kernel<W, X2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Wb2Shared<W>));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * ncl, 1, 1);
  cfg.blockDim = dim3(32 * W, 1, 1);
  cfg.dynamicSmemBytes = sizeof(Wb2Shared<W>);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, panel_wb2_kernel<W, X2>, p);
}

}  // namespace

#ifdef CQR_HH_TRACE
void panel_wb2_read_trace(long long* steps, long long* marks) {
  cudaMemcpyFromSymbol(steps, g_wb2_trace, sizeof(g_wb2_trace));
  cudaMemcpyFromSymbol(marks, g_wb2_marks, sizeof(g_wb2_marks));
}
#endif

// Same plan as panel_wb.cu (wpc warps per CTA, one cluster of cs CTAs); mode 1 = pairs with the cancellation guard,
// 2 = every pair falls back to two single steps.  Returns false when the shape is not covered (b != 64, more than one
// cluster) or the launch failed; the caller then takes the one-column-per-exchange kernels.
bool launch_panel_wb2(const PanelHHParams& p, int wpc, int cs, int ncl, int mode, cudaStream_t s) {
  if (p.b != 64 || ncl < 1 || ncl > 2 || cs > 16) return false;
  if (ncl == 2 && (wpc != 8 || cs != 16)) return false;
  if (ncl == 2 && panel_hh_slot_bytes() < (size_t)2 * 33 * 128 * sizeof(uint2)) return false;
  PanelHHParams pp = p;
  pp.pmax = mode;
  ++g_launches;
  cudaError_t e;
  if (ncl == 2) e = launch_wb2_t<8, true>(pp, cs, ncl, s);       // panel_wb_plan: two clusters are always 16 CTAs x 8 warps
  else if (wpc == 1) e = launch_wb2_t<1, false>(pp, cs, ncl, s);
  else if (wpc == 2) e = launch_wb2_t<2, false>(pp, cs, ncl, s);
  else if (wpc == 4) e = launch_wb2_t<4, false>(pp, cs, ncl, s);
  else e = launch_wb2_t<8, false>(pp, cs, ncl, s);
  if (e != cudaSuccess) { cudaGetLastError(); --g_launches; return false; }
  return true;
}

}  // namespace cqr
