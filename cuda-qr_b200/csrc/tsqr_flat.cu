// tsqr_flat.cu -- R-only leaf of the tall-skinny QR (BASELINE config 3): warp-resident flat-tree Householder.
//
// The reference factors a tall panel by sweeping a PR x PC window bottom-to-top, carrying the running R in the
// rows of overlap (qr.c:68-73, 109-141: "triangle on top of a square" -- the reflector of column `col` spans the
// fresh rows plus one diagonal entry of the carried R).  This kernel is that same flat tree, re-cut for a B200:
// every WARP owns one contiguous chain of 64-row blocks and one running 64 x 64 R, thousands of chains run at once,
// and the chains' R factors feed the existing binary/4-ary tile tree (tile_qr.cu).
//
// Per block B (64 x 64) and pivot column j the structured reflector is v = [e_j ; x / u] with x = B(:, j):
//   sigma = x^T x, alpha = R(j,j), beta = -sign(alpha) sqrt(alpha^2 + sigma), u = alpha - beta, tau = -u / beta  (qr.c:144-152)
//   for c > j:  s = R(j,c) + (x^T B(:,c)) / u ;  R(j,c) -= tau s ;  B(:,c) -= (tau s / u) x
// Only row j of R and the block change, so the flop count is the Householder minimum 2 m n^2 -- no stacked-R overhead.
//
// Lane layout (lane = 8 h + q): column group q in 0..7 owns columns {q, q+8, .., q+56} (register slot i = c / 8),
// row part h in 0..3 owns rows 16 h .. 16 h + 15 of the block.  A lane holds 8 slots x 16 rows as 64 packed f32x2
// pairs, so every dot product and rank-1 update is FFMA2; a column's dot needs two shuffle stages (over h) instead of
// the five of a lane-per-row layout, and x reaches the lanes as four broadcast LDS.128.  Slots die as the pivot moves
// right (slot i is skipped once 8 i + 7 <= j): the step body exists in eight statically specialised versions
// (template I0 = j / 8) and loops dynamically over j % 8, which keeps every register index static while the whole
// sweep stays ~2 K instructions (a fully unrolled 64-step body would not fit the instruction cache).
// R lives in shared memory (row-major, 16 KB per warp): a step touches only row j, one conflict-free word per slot.
// There is no block-level barrier anywhere: warps are independent, __syncwarp orders the x broadcast.
#include "common.cuh"
#include "warp_math.cuh"
#include <stdlib.h>

namespace cqr {

constexpr int kFlatWarpFloats = 64 * 64 + 2 * 64;   // R (row-major) + double-buffered x
constexpr int kFlatDefaultCfg = 2;

// Steps j = 8 I0 .. 8 I0 + 7 of one block: slots < I0 are finished columns and are not touched.
template <int I0>
__device__ __forceinline__ void flat_steps(f32x2 (&b)[8][8], float* __restrict__ Rs, float* __restrict__ xs, const int q,
                                           const int h, const int n) {
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * I0 + jj;
    if (j >= n) break;                       // warp-uniform
    float* xb = xs + (jj & 1) * 64;
    float* Rj = Rs + j * 64;
    // row j of R (last written one block ago) is read before the warp barrier and written after it
    float r[8];
#pragma unroll
    for (int i = I0; i < 8; ++i) r[i] = Rj[q + 8 * i];
    const float alpha = Rj[j];
    if (q == jj) {                           // the four lanes holding column j publish x = B(:, j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ulonglong2 v; v.x = b[I0][2 * k]; v.y = b[I0][2 * k + 1];
        *reinterpret_cast<ulonglong2*>(xb + 16 * h + 4 * k) = v;
      }
    }
    __syncwarp();
    f32x2 x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xb + 16 * h + 4 * k);
      x[2 * k] = v.x; x[2 * k + 1] = v.y;
    }
    // partial sums over this lane's 16 rows: sigma and the live columns' dots; then two butterfly stages over h.
    // Every lane ends with bit-identical totals (commutative pairings), so the scalars below are warp-uniform.
    float d[8];
    f32x2 s2 = 0ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) s2 = ffma2(x[k], x[k], s2);
    float sig = fsum2(s2);
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      f32x2 d2 = 0ull;
#pragma unroll
      for (int k = 0; k < 8; ++k) d2 = ffma2(x[k], b[i][k], d2);
      d[i] = fsum2(d2);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float ts = __shfl_xor_sync(kFull, sig, o);
      float t[8];
#pragma unroll
      for (int i = I0; i < 8; ++i) t[i] = __shfl_xor_sync(kFull, d[i], o);
      sig += ts;
#pragma unroll
      for (int i = I0; i < 8; ++i) d[i] += t[i];
    }
    // Reflector scalars on the MUFU approximations plus one Newton step each (~1 ulp, no slow-path calls): the IEEE
    // sqrt / divide sequences cost 55 instructions per step.  Columns whose norm underflows fp32 count as zero.
    float tau = 0.f, inv_u = 0.f, beta = alpha;
    const float sj = fmaf(alpha, alpha, sig);
    if (sig != 0.f && sj >= 1.2e-38f) {      // x == 0: H = I (LAPACK convention; the reference would flip a sign or NaN)
      const float rs = rsqrt_approx(sj);
      float nrm = sj * rs;
      nrm = fmaf(fmaf(-nrm, nrm, sj), 0.5f * rs, nrm);
      beta = (alpha < 0.f) ? nrm : -nrm;
      const float u = alpha - beta;
      inv_u = rcp_newton(u);
      tau = -u * rcp_newton(beta);
    }
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      const bool act = (i > I0) || (q > jj);           // column q + 8 i is to the right of the pivot
      const float s = fmaf(d[i], inv_u, r[i]);
      const float wv = act ? tau * s : 0.f;
      if (act && h == 0) Rj[q + 8 * i] = r[i] - wv;
      const float nwu = -(wv * inv_u);
      const f32x2 nw2 = fpack2(nwu, nwu);
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = ffma2(nw2, x[k], b[i][k]);
    }
    if (q == jj && h == 0) Rj[j] = beta;
  }
}

template <int I0>
struct FlatGroups {
  static __device__ __forceinline__ void run(f32x2 (&b)[8][8], float* Rs, float* xs, int q, int h, int n) {
    flat_steps<I0>(b, Rs, xs, q, h, n);
    FlatGroups<I0 + 1>::run(b, Rs, xs, q, h, n);
  }
};
template <>
struct FlatGroups<8> {
  static __device__ __forceinline__ void run(f32x2 (&)[8][8], float*, float*, int, int, int) {}
};

// ---- two pivot columns per reduction (R only) ------------------------------------------------------------------------
// The leaf is bound by its 64 dependent column steps per block (publish x, dots, two shuffle stages, rsqrt / rcp chain,
// update: ~600 clk each), not by arithmetic.  Here one pass serves the reflectors j and j + 1, with
// the algebra of panel_wb2.cu specialised to the structured reflector v_j = [e_j ; x / u] (the unit entry sits on R's row
// j, so reflector j never touches R's row j + 1).  x = B(:, j), y = B(:, j + 1), both BEFORE reflector j; one reduction
// delivers p_c = x^T a_c and q_c = y^T a_c for every live column.  Then (fp32 spec: tools/two_column_step.py)
//   reflector j:   sigma_1 = p_j, alpha_1 = R(j,j) -> beta_1, u_1, tau_1;   d1_c = R(j,c) + p_c / u_1, t_c = tau_1 d1_c, e_c = t_c / u_1
//   column j + 1:  a1 = e_{j+1};  y' = y - a1 x;  sigma_2 = q_{j+1} - 2 a1 p_{j+1} + a1^2 p_j,  alpha_2 = R(j+1,j+1) -> beta_2, u_2, tau_2
//   any column c:  y'^T a'_c = q_c - e_c p_{j+1} - a1 p_c + a1 e_c p_j;  d2_c = R(j+1,c) + y'^T a'_c / u_2,  f_c = tau_2 d2_c / u_2
//   update:        a_c -= (e_c - f_c a1) x + f_c y;   R(j,c) -= t_c;   R(j+1,c) -= tau_2 d2_c
// The expansion of sigma_2 (and of every y'^T a'_c) cancels when y is nearly parallel to x, so a pair with
// sigma_2 < 0.1 q_{j+1} finishes reflector j alone and column j + 1 takes an ordinary single step (guard and threshold as
// in panel_wb2.cu; decided on bit-identical totals, uniform over the warp).  One shared x / y buffer: a __syncwarp in
// front of the publish orders it against the previous step's readers.
constexpr float kFlatPairGuard = 0.1f;

struct FlatRefl { float bc, inv_u, tau; bool ok; };
__device__ __forceinline__ FlatRefl flat_scalars(float alpha, float sig) {   // as in flat_steps: a zero / underflowing column gives tau = 0
  FlatRefl r;
  const float sj = fmaf(alpha, alpha, sig);
  r.ok = (sig != 0.f) && (sj >= 1.2e-38f);
  const float sjs = r.ok ? sj : 1.f;
  const float rs = rsqrt_approx(sjs);
  float nrm = sjs * rs;
  nrm = fmaf(fmaf(-nrm, nrm, sjs), 0.5f * rs, nrm);
  r.bc = (alpha < 0.f) ? nrm : -nrm;
  const float u = alpha - r.bc;
  r.inv_u = r.ok ? rcp_newton(u) : 0.f;
  r.tau = r.ok ? -u * rcp_newton(r.bc) : 0.f;
  return r;
}

// One ordinary column step, pivot j = 8 I0 + jj (the odd column of a pair that failed the guard, or the last column of an odd n).
template <int I0>
__device__ __forceinline__ void flat_one(f32x2 (&b)[8][8], float* __restrict__ Rs, float* __restrict__ xb, const int q, const int h, const int jj) {
  const int j = 8 * I0 + jj;
  float* Rj = Rs + j * 64;
  float r[8];
#pragma unroll
  for (int i = I0; i < 8; ++i) r[i] = Rj[q + 8 * i];
  const float alpha = Rj[j];
  __syncwarp();
  if (q == jj) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ulonglong2 v; v.x = b[I0][2 * k]; v.y = b[I0][2 * k + 1];
      *reinterpret_cast<ulonglong2*>(xb + 16 * h + 4 * k) = v;
    }
  }
  __syncwarp();
  f32x2 x[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xb + 16 * h + 4 * k);
    x[2 * k] = v.x; x[2 * k + 1] = v.y;
  }
  float d[8];
  f32x2 s2 = 0ull;
#pragma unroll
  for (int k = 0; k < 8; ++k) s2 = ffma2(x[k], x[k], s2);
  float sig = fsum2(s2);
#pragma unroll
  for (int i = I0; i < 8; ++i) {
    f32x2 d2 = 0ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) d2 = ffma2(x[k], b[i][k], d2);
    d[i] = fsum2(d2);
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float ts = __shfl_xor_sync(kFull, sig, o);
    float t[8];
#pragma unroll
    for (int i = I0; i < 8; ++i) t[i] = __shfl_xor_sync(kFull, d[i], o);
    sig += ts;
#pragma unroll
    for (int i = I0; i < 8; ++i) d[i] += t[i];
  }
  const FlatRefl f = flat_scalars(alpha, sig);
#pragma unroll
  for (int i = I0; i < 8; ++i) {
    const bool act = (i > I0) || (q > jj);
    const float s = fmaf(d[i], f.inv_u, r[i]);
    const float wv = act ? f.tau * s : 0.f;
    if (act && h == 0) Rj[q + 8 * i] = r[i] - wv;
    const float nwu = -(wv * f.inv_u);
    const f32x2 nw2 = fpack2(nwu, nwu);
#pragma unroll
    for (int k = 0; k < 8; ++k) b[i][k] = ffma2(nw2, x[k], b[i][k]);
  }
  if (q == jj && h == 0) Rj[j] = f.ok ? f.bc : alpha;
}

// Pivots j = 8 I0 + jj and j + 1 (jj even) in one pass; false if the guard tripped (reflector j is then done, j + 1 is not).
template <int I0>
__device__ __forceinline__ bool flat_pair(f32x2 (&b)[8][8], float* __restrict__ Rs, float* __restrict__ xb, const int q, const int h, const int jj) {
  const int j = 8 * I0 + jj;
  float* Rj = Rs + j * 64;
  float* Rk = Rj + 64;
  float r1[8], r2[8];
#pragma unroll
  for (int i = I0; i < 8; ++i) { r1[i] = Rj[q + 8 * i]; r2[i] = Rk[q + 8 * i]; }
  const float alpha1 = Rj[j], alpha2 = Rk[j + 1], r1n = Rj[j + 1];
  __syncwarp();
  if (q == jj || q == jj + 1) {                 // x to xb[0..63], y to xb[64..127]
    float* dst = xb + (q - jj) * 64 + 16 * h;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ulonglong2 v; v.x = b[I0][2 * k]; v.y = b[I0][2 * k + 1];
      *reinterpret_cast<ulonglong2*>(dst + 4 * k) = v;
    }
  }
  __syncwarp();
  f32x2 x[8], y[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xb + 16 * h + 4 * k);
    const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(xb + 64 + 16 * h + 4 * k);
    x[2 * k] = v.x; x[2 * k + 1] = v.y;
    y[2 * k] = u.x; y[2 * k + 1] = u.y;
  }
  float p[8], qd[8];
#pragma unroll
  for (int i = I0; i < 8; ++i) {
    f32x2 p2 = 0ull, q2 = 0ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) { p2 = ffma2(x[k], b[i][k], p2); q2 = ffma2(y[k], b[i][k], q2); }
    p[i] = fsum2(p2); qd[i] = fsum2(q2);
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    float tp[8], tq[8];
#pragma unroll
    for (int i = I0; i < 8; ++i) { tp[i] = __shfl_xor_sync(kFull, p[i], o); tq[i] = __shfl_xor_sync(kFull, qd[i], o); }
#pragma unroll
    for (int i = I0; i < 8; ++i) { p[i] += tp[i]; qd[i] += tq[i]; }
  }
  // pivot sums from the lanes that own columns j and j + 1 (slot I0): x^T x, x^T y, y^T y -- identical in every lane
  const float pj = __shfl_sync(kFull, p[I0], jj), pj1 = __shfl_sync(kFull, p[I0], jj + 1), qj1 = __shfl_sync(kFull, qd[I0], jj + 1);
  const FlatRefl f1 = flat_scalars(alpha1, pj);
  const float a1 = f1.tau * fmaf(pj1, f1.inv_u, r1n) * f1.inv_u;
  const float sig2 = fmaf(a1, fmaf(a1, pj, -2.f * pj1), qj1);
  if (sig2 < kFlatPairGuard * qj1) {            // warp-uniform: finish reflector j from what is here, leave column j + 1 to a single step
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      const bool act = (i > I0) || (q > jj);
      const float t = act ? f1.tau * fmaf(p[i], f1.inv_u, r1[i]) : 0.f;
      if (act && h == 0) Rj[q + 8 * i] = r1[i] - t;
      const float ne = -(t * f1.inv_u);
      const f32x2 ne2 = fpack2(ne, ne);
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = ffma2(ne2, x[k], b[i][k]);
    }
    if (q == jj && h == 0) Rj[j] = f1.ok ? f1.bc : alpha1;
    return false;
  }
  const FlatRefl f2 = flat_scalars(alpha2, sig2 > 0.f ? sig2 : 0.f);
  const float a1pj = a1 * pj;
#pragma unroll
  for (int i = I0; i < 8; ++i) {
    const bool act1 = (i > I0) || (q > jj);      // right of j (column j + 1 included)
    const bool act2 = (i > I0) || (q > jj + 1);  // right of j + 1
    const float t = act1 ? f1.tau * fmaf(p[i], f1.inv_u, r1[i]) : 0.f;
    const float e = t * f1.inv_u;
    if (act1 && h == 0) Rj[q + 8 * i] = r1[i] - t;
    const float yta = fmaf(e, a1pj - pj1, fmaf(-a1, p[i], qd[i]));   // q_c - e p_{j+1} - a1 p_c + a1 e p_j
    const float t2 = act2 ? f2.tau * fmaf(yta, f2.inv_u, r2[i]) : 0.f;
    const float fc = t2 * f2.inv_u;
    if (act2 && h == 0) Rk[q + 8 * i] = r2[i] - t2;
    const float cx = fmaf(fc, a1, -e), cy = -fc;
    const f32x2 cx2 = fpack2(cx, cx), cy2 = fpack2(cy, cy);
#pragma unroll
    for (int k = 0; k < 8; ++k) b[i][k] = ffma2(cy2, y[k], ffma2(cx2, x[k], b[i][k]));
  }
  if (q == jj && h == 0) Rj[j] = f1.ok ? f1.bc : alpha1;
  if (q == jj + 1 && h == 0) Rk[j + 1] = f2.ok ? f2.bc : alpha2;
  return true;
}

template <int I0>
struct FlatGroupsPair {
  static __device__ __forceinline__ void run(f32x2 (&b)[8][8], float* Rs, float* xs, int q, int h, int n) {
#pragma unroll 1
    for (int jj = 0; jj < 8; jj += 2) {
      const int j = 8 * I0 + jj;
      if (j >= n) break;                         // warp-uniform
      if (j + 1 < n) {
        if (!flat_pair<I0>(b, Rs, xs, q, h, jj)) flat_one<I0>(b, Rs, xs, q, h, jj + 1);
      } else {
        flat_one<I0>(b, Rs, xs, q, h, jj);
      }
    }
    FlatGroupsPair<I0 + 1>::run(b, Rs, xs, q, h, n);
  }
};
template <>
struct FlatGroupsPair<8> {
  static __device__ __forceinline__ void run(f32x2 (&)[8][8], float*, float*, int, int, int) {}
};

// ---- software-pipelined step body ---------------------------------------------------------------------------------
// The step's serial chain is sigma -> two shuffles -> sqrt / reciprocals -> per-column scalars; only the rank-1 update
// depends on it.  Here the update of reflector j-1 is deferred into step j: the pivot column's slot is updated first
// (so x_j can be published), then sigma_j's chain is started and the remaining updates of reflector j-1 (independent
// FFMA2 work) and the dots of step j are issued under its latency.  State carried between steps: x_{j-1} (registers)
// and nw[i] = -(tau s_i / u) per slot (0 = nothing pending).  Iteration (I0, jj) touches slots >= I0 only: the pending
// reflector of the group's first iteration has its pivot in slot I0-1's last column, whose remaining active columns all
// lie in slots >= I0.
// KEEP: the reflector is stored -- column j of the block becomes x / u (the part of v below the carried R; v's unit entry
// sits on R's row j and is implicit) and tau goes to tau_blk[j] -- so tsqr_flat_apply_kernel can expand the implicit Q.
template <int I0, bool KEEP>
__device__ __forceinline__ void flat_steps_pipe(f32x2 (&b)[8][8], f32x2 (&xp)[8], float (&nw)[8], float* __restrict__ Rs,
                                                float* __restrict__ xs, const int q, const int h, const int n,
                                                float* __restrict__ tau_blk) {
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * I0 + jj;
    if (j >= n) break;                       // warp-uniform; what is still pending only touches zero columns
    float* xb = xs + (jj & 1) * 64;
    float* Rj = Rs + j * 64;
    {                                        // pending reflector on the pivot's slot
      const f32x2 c2 = fpack2(nw[I0], nw[I0]);
#pragma unroll
      for (int k = 0; k < 8; ++k) b[I0][k] = ffma2(c2, xp[k], b[I0][k]);
    }
    float r[8];
#pragma unroll
    for (int i = I0; i < 8; ++i) r[i] = Rj[q + 8 * i];
    const float alpha = Rj[j];
    if (q == jj) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ulonglong2 v; v.x = b[I0][2 * k]; v.y = b[I0][2 * k + 1];
        *reinterpret_cast<ulonglong2*>(xb + 16 * h + 4 * k) = v;
      }
    }
    __syncwarp();
    f32x2 x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xb + 16 * h + 4 * k);
      x[2 * k] = v.x; x[2 * k + 1] = v.y;
    }
    f32x2 s2a = 0ull, s2b = 0ull;            // two chains: sigma heads the critical path
#pragma unroll
    for (int k = 0; k < 4; ++k) { s2a = ffma2(x[2 * k], x[2 * k], s2a); s2b = ffma2(x[2 * k + 1], x[2 * k + 1], s2b); }
    float sig = fsum2(s2a) + fsum2(s2b);
    sig += __shfl_xor_sync(kFull, sig, 8);
    sig += __shfl_xor_sync(kFull, sig, 16);
    // branch-free scalars (see flat_steps): a zero or underflowing column gives tau = 0, H = I.  Written here, ahead of
    // the FFMA2 streams below, so the MUFU chain runs under them (a warp issues in order).
    const float sj = fmaf(alpha, alpha, sig);
    const bool ok = (sig != 0.f) && (sj >= 1.2e-38f);
    const float sjs = ok ? sj : 1.f;
    const float rs = rsqrt_approx(sjs);
    float nrm = sjs * rs;
    nrm = fmaf(fmaf(-nrm, nrm, sjs), 0.5f * rs, nrm);
    const float bc = (alpha < 0.f) ? nrm : -nrm;
    const float u = alpha - bc;
    const float inv_u = ok ? rcp_newton(u) : 0.f;
    const float tau = ok ? -u * rcp_newton(bc) : 0.f;
    const float c2s = -tau * inv_u;
#pragma unroll
    for (int i = I0 + 1; i < 8; ++i) {       // pending reflector on the other live slots
      const f32x2 c2 = fpack2(nw[i], nw[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = ffma2(c2, xp[k], b[i][k]);
    }
    // dots in two halves of the live slots: the second half's FFMA2s cover the first half's shuffle latency
    // (k outer / slot inner so consecutive FFMA2s share x[k]: an FFMA2 with three fresh 64-bit sources issues every
    // 3 cycles on sm_100, one with two every 2 -- tools/probes/ffma2_probe.cu)
    float d[8];
    constexpr int IM = (I0 + 8 + 1) / 2;     // first half: slots I0 .. IM-1
    {
      f32x2 d2[8];
#pragma unroll
      for (int i = I0; i < IM; ++i) d2[i] = 0ull;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = I0; i < IM; ++i) d2[i] = dfma2(x[k], b[i][k], d2[i]);
#pragma unroll
      for (int i = I0; i < IM; ++i) d[i] = fsum2(d2[i]);
#pragma unroll
      for (int i = I0; i < IM; ++i) d[i] += __shfl_xor_sync(kFull, d[i], 8);
#pragma unroll
      for (int i = IM; i < 8; ++i) d2[i] = 0ull;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = IM; i < 8; ++i) d2[i] = dfma2(x[k], b[i][k], d2[i]);
#pragma unroll
      for (int i = I0; i < IM; ++i) d[i] += __shfl_xor_sync(kFull, d[i], 16);
#pragma unroll
      for (int i = IM; i < 8; ++i) d[i] = fsum2(d2[i]);
#pragma unroll
      for (int i = IM; i < 8; ++i) d[i] += __shfl_xor_sync(kFull, d[i], 8);
#pragma unroll
      for (int i = IM; i < 8; ++i) d[i] += __shfl_xor_sync(kFull, d[i], 16);
    }
    if (q == jj && h == 0) Rj[j] = ok ? bc : alpha;
    if (KEEP) {
      if (q == jj) {                         // nothing pending touches this column any more (its nw stays 0)
        const f32x2 iu2 = fpack2(inv_u, inv_u);
#pragma unroll
        for (int k = 0; k < 8; ++k) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b[I0][k]) : "l"(b[I0][k]), "l"(iu2));
        if (h == 0) tau_blk[j] = tau;
      }
    }
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      const bool act = (i > I0) || (q > jj);
      const float s = fmaf(d[i], inv_u, r[i]);
      nw[i] = act ? c2s * s : 0.f;
      if (act && h == 0) Rj[q + 8 * i] = fmaf(-tau, s, r[i]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) xp[k] = x[k];
  }
}

template <int I0, bool KEEP>
struct FlatGroupsPipe {
  static __device__ __forceinline__ void run(f32x2 (&b)[8][8], f32x2 (&xp)[8], float (&nw)[8], float* Rs, float* xs, int q, int h, int n,
                                             float* tau_blk) {
    flat_steps_pipe<I0, KEEP>(b, xp, nw, Rs, xs, q, h, n, tau_blk);
    FlatGroupsPipe<I0 + 1, KEEP>::run(b, xp, nw, Rs, xs, q, h, n, tau_blk);
  }
};
template <bool KEEP>
struct FlatGroupsPipe<8, KEEP> {
  static __device__ __forceinline__ void run(f32x2 (&)[8][8], f32x2 (&)[8], float (&)[8], float*, float*, int, int, int, float*) {}
};

template <int WPC, int MINB, bool PIPE, bool KEEP = false, bool PAIR = false>
__global__ void __launch_bounds__(32 * WPC, MINB) tsqr_flat_r_kernel(FlatTsqrParams p) {
  static_assert(PIPE || !KEEP, "the reflector store lives in the pipelined step body");
  static_assert(!PAIR || (!PIPE && !KEEP), "two pivot columns per reduction: R-only, unpipelined");
  extern __shared__ __align__(16) float flat_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 7, h = lane >> 3;
  const long long chain = (long long)blockIdx.x * WPC + w;
  if (chain >= p.chains) return;             // warps never meet at a block barrier
  if (p.gate != nullptr) {             // gated launch = programmatic dependent launch (FlatCfg::launch)
    pdl_trigger();
    pdl_wait();
    if (*(const volatile int*)p.gate == 0) return;   // the Gram leaf produced R: nothing to do
  }
  float* Rs = flat_smem + w * kFlatWarpFloats;
  float* xs = Rs + 64 * 64;
  for (int i = lane; i < 64 * 64 / 4; i += 32) reinterpret_cast<float4*>(Rs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  if (KEEP) {                                // the chain's first block is already factored (launch_tsqr_flat_first_blocks): R = its upper triangle
    const long long r0 = chain * p.rows_per_chain;
    for (int c = 0; c < p.n; ++c) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = lane + 32 * rr;
        if (r <= c && r0 + r < p.m) Rs[r * 64 + c] = p.a[r0 + r + (long long)c * p.lda];
      }
    }
    __syncwarp();
  }

  f32x2 xp[8];
  float nw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) xp[k] = 0ull;
  const int n = p.n;
  const long long row0 = chain * p.rows_per_chain;
  const long long row1 = (row0 + p.rows_per_chain < p.m) ? row0 + p.rows_per_chain : p.m;
  const bool aligned = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0);
  for (long long rb = row0 + (KEEP ? 64 : 0); rb < row1; rb += 64) {
    f32x2 b[8][8];
    const float* src = p.a + rb + 16 * h;
    if (aligned && rb + 64 <= row1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
        if (c < n) {
          const ulonglong2* s4 = reinterpret_cast<const ulonglong2*>(src + (long long)c * p.lda);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const ulonglong2 v = __ldcs(s4 + k);
            b[i][2 * k] = v.x; b[i][2 * k + 1] = v.y;
          }
          if (rb + 128 <= row1) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (long long)c * p.lda + 64));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) b[i][k] = 0ull;
        }
      }
    } else {                                 // ragged last block or unaligned source: guarded scalar loads, zero fill
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long r = rb + 16 * h + 2 * k;
          const float lo = (c < n && r < row1) ? src[(long long)c * p.lda + 2 * k] : 0.f;
          const float hi = (c < n && r + 1 < row1) ? src[(long long)c * p.lda + 2 * k + 1] : 0.f;
          b[i][k] = fpack2(lo, hi);
        }
      }
    }
    // the x buffers are reused from block to block: for odd n the previous block's last step read buffer 0, which this
    // block's first step rewrites (racecheck, 20011 x 17) -- order them
    __syncwarp();
    if (PIPE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) nw[i] = 0.f;       // nothing pending at the top of a block
      FlatGroupsPipe<0, KEEP>::run(b, xp, nw, Rs, xs, q, h, n, KEEP ? p.tau_out + (rb >> 6) * 64 : nullptr);
    } else if (PAIR) {
      FlatGroupsPair<0>::run(b, Rs, xs, q, h, n);
    } else {
      FlatGroups<0>::run(b, Rs, xs, q, h, n);
    }
    if (KEEP) {                              // V block over the source rows (same addressing as the load)
      float* dstv = p.a_out + rb + 16 * h;
      if (aligned && rb + 64 <= row1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = q + 8 * i;
          if (c < n) {
            ulonglong2* d4 = reinterpret_cast<ulonglong2*>(dstv + (long long)c * p.lda);
#pragma unroll
            for (int k = 0; k < 4; ++k) { ulonglong2 v; v.x = b[i][2 * k]; v.y = b[i][2 * k + 1]; d4[k] = v; }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = q + 8 * i;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const long long r = rb + 16 * h + 2 * k;
            float lo, hi;
            funpack2(b[i][k], lo, hi);
            if (c < n && r < row1) dstv[(long long)c * p.lda + 2 * k] = lo;
            if (c < n && r + 1 < row1) dstv[(long long)c * p.lda + 2 * k + 1] = hi;
          }
        }
      }
    }
  }
  __syncwarp();
  // chain k's R goes to slot (k % fan) of parent tile (k / fan); a full 64 x 64 slot is written (zeros below the diagonal
  // and in columns >= n) so the tree above never sees stale workspace
  float* dst = p.r_out + (chain / p.fan) * p.r_tile_stride + (chain % p.fan) * CQR_SLOT;
  for (int c = 0; c < 64; ++c) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = lane + 32 * rr;
      dst[r + (long long)c * p.r_ld] = (r <= c && c < n) ? Rs[r * 64 + c] : 0.f;
    }
  }
}

// out = Q_leaf * [X; 0] for the implicit Q of the flat leaf (KEEP): every chain starts from its 64 x nc seed X (rows of the
// carried R: what the tree level above hands down) and walks its blocks last to first; in a block the reflectors are
// applied in reverse, s = Y_R(j, c) + v_j^T Y_B(:, c), Y_R(j, c) -= tau s, Y_B(:, c) -= tau s v_j, and the finished
// Y_B is the block's 64 rows of the output.  Same lane layout as the factorisation (no slot ever dies here: every
// reflector touches all nc columns); v_j comes straight from global memory one step ahead of its use (all eight column
// groups read the same 256 bytes), the next block is prefetched into L2.  The chain's first block is a dense QR of real rows (no virtual rows): it is expanded last.
template <int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) tsqr_flat_apply_kernel(FlatApplyParams p) {
  extern __shared__ __align__(16) float flat_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 7, h = lane >> 3;
  const long long chain = (long long)blockIdx.x * WPC + w;
  if (chain >= p.chains) return;
  float* Ys = flat_smem + w * (64 * 64);     // Y_R, row-major: Ys[j * 64 + c]
  const int n = p.n, nc = p.nc;
  {
    const float* xs = p.x ? p.x + (chain / p.fan) * p.x_tile_stride + (chain % p.fan) * CQR_SLOT : nullptr;
    for (int c = 0; c < 64; ++c) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = lane + 32 * rr;
        float v = 0.f;
        if (c < nc && r < p.x_rows) v = xs ? xs[r + (long long)c * p.x_ld] : (r == c ? 1.f : 0.f);
        Ys[r * 64 + c] = v;
      }
    }
  }
  __syncwarp();
  const long long row0 = chain * p.rows_per_chain;
  const long long row1 = (row0 + p.rows_per_chain < p.m) ? row0 + p.rows_per_chain : p.m;
  const bool aligned = (p.ldv % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.v) & 15) == 0);
  const bool oaligned = (p.ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.q) & 15) == 0);
  const long long nblk = (row1 - row0 + 63) / 64;
  for (long long t = nblk - 1; t >= 1; --t) {
    const long long rb = row0 + 64 * t;
    const bool full = rb + 64 <= row1;
    const float* vsrc = p.v + rb + 16 * h;
    const float* tau_blk = p.tau + (rb >> 6) * 64;
    if (aligned) {                           // next block (one up) into L2: 64 columns x 256 B, four 128-byte lines per lane
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;       // column idx / 2, half idx % 2
        if ((idx >> 1) < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.v + rb - 64 + (long long)(idx >> 1) * p.ldv + 32 * (idx & 1)));
      }
    }
    f32x2 b[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = 0ull;
    auto load_v = [&](int j, f32x2 (&v)[8]) {
      const float* col = vsrc + (long long)j * p.ldv;
      if (aligned && full) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const ulonglong2 t4 = *reinterpret_cast<const ulonglong2*>(col + 4 * k);
          v[2 * k] = t4.x; v[2 * k + 1] = t4.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long r = rb + 16 * h + 2 * k;
          v[k] = fpack2(r < row1 ? col[2 * k] : 0.f, r + 1 < row1 ? col[2 * k + 1] : 0.f);
        }
      }
    };
    f32x2 vn[8];
    load_v(n - 1, vn);
    float taun = tau_blk[n - 1];
#pragma unroll 1
    for (int j = n - 1; j >= 0; --j) {
      f32x2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = vn[k];
      const float tau = taun;
      if (j > 0) { load_v(j - 1, vn); taun = tau_blk[j - 1]; }
      float* Yj = Ys + j * 64;
      float yr[8], d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) yr[i] = Yj[q + 8 * i];
      {
        f32x2 d2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d2[i] = 0ull;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int i = 0; i < 8; ++i) d2[i] = ffma2(v[k], b[i][k], d2[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fsum2(d2[i]);
      }
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = __shfl_xor_sync(kFull, d[i], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] += t8[i];
      }
      __syncwarp();                          // every lane has read row j of Y_R before part 0 rewrites it
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float wv = tau * (yr[i] + d[i]);
        if (h == 0) Yj[q + 8 * i] = yr[i] - wv;
        const f32x2 nw2 = fpack2(-wv, -wv);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[i][k] = ffma2(nw2, v[k], b[i][k]);
      }
    }
    float* dst = p.q + rb + 16 * h;
    if (oaligned && full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
        if (c < nc) {
          ulonglong2* d4 = reinterpret_cast<ulonglong2*>(dst + (long long)c * p.ldq);
#pragma unroll
          for (int k = 0; k < 4; ++k) { ulonglong2 t4; t4.x = b[i][2 * k]; t4.y = b[i][2 * k + 1]; d4[k] = t4; }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long r = rb + 16 * h + 2 * k;
          float lo, hi;
          funpack2(b[i][k], lo, hi);
          if (c < nc && r < row1) dst[(long long)c * p.ldq + 2 * k] = lo;
          if (c < nc && r + 1 < row1) dst[(long long)c * p.ldq + 2 * k + 1] = hi;
        }
      }
    }
  }
  // The chain's first block holds a dense Householder QR (launch_tsqr_flat_first_blocks): the carried rows ARE this
  // block's rows, so Y_R moves into registers and the block's reflectors v_j = [0; 1; v(j+1:)] are applied in reverse.
  {
    const long long rb = row0;
    const bool full = rb + 64 <= row1;
    const float* vsrc = p.v + rb + 16 * h;
    const float* tau_blk = p.tau + (rb >> 6) * 64;
    __syncwarp();
    f32x2 b[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) b[i][k] = fpack2(Ys[(16 * h + 2 * k) * 64 + q + 8 * i], Ys[(16 * h + 2 * k + 1) * 64 + q + 8 * i]);
#pragma unroll 1
    for (int j = n - 1; j >= 0; --j) {
      const float tau = tau_blk[j];
      if (tau == 0.f) continue;              // warp-uniform
      const float* col = vsrc + (long long)j * p.ldv;
      f32x2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rl = 16 * h + 2 * k;       // block-local rows rl, rl + 1
        float lo = 0.f, hi = 0.f;
        if (rl > j && rb + rl < row1) lo = col[2 * k];
        if (rl + 1 > j && rb + rl + 1 < row1) hi = col[2 * k + 1];
        if (rl == j) lo = 1.f;
        if (rl + 1 == j) hi = 1.f;
        v[k] = fpack2(lo, hi);
      }
      float d[8];
      {
        f32x2 d2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d2[i] = 0ull;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int i = 0; i < 8; ++i) d2[i] = ffma2(v[k], b[i][k], d2[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fsum2(d2[i]);
      }
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = __shfl_xor_sync(kFull, d[i], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] += t8[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float wv = tau * d[i];
        const f32x2 nw2 = fpack2(-wv, -wv);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[i][k] = ffma2(nw2, v[k], b[i][k]);
      }
    }
    float* dst = p.q + rb + 16 * h;
    if (oaligned && full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
        if (c < nc) {
          ulonglong2* d4 = reinterpret_cast<ulonglong2*>(dst + (long long)c * p.ldq);
#pragma unroll
          for (int k = 0; k < 4; ++k) { ulonglong2 t4; t4.x = b[i][2 * k]; t4.y = b[i][2 * k + 1]; d4[k] = t4; }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = q + 8 * i;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long r = rb + 16 * h + 2 * k;
          float lo, hi;
          funpack2(b[i][k], lo, hi);
          if (c < nc && r < row1) dst[(long long)c * p.ldq + 2 * k] = lo;
          if (c < nc && r + 1 < row1) dst[(long long)c * p.ldq + 2 * k + 1] = hi;
        }
      }
    }
  }
}

// Variants: CQR_FLAT_CFG = 0: plain step body, 3 CTAs x 4 warps per SM (168 registers, a few spills);  1: plain, 2 x 4
// (236 registers);  2 (default): software-pipelined, 2 x 4 (218 registers).  Registers are allotted per 4 warps, so 8
// or 12 resident warps per SM are the only useful occupancies (a 2 x 5 variant compiles to the 168-register budget).
template <int WPC, int MINB, bool PIPE>
struct FlatCfg {
  static constexpr size_t smem = (size_t)WPC * kFlatWarpFloats * sizeof(float);
  static int per_sm() {
    cudaFuncSetAttribute(tsqr_flat_r_kernel<WPC, MINB, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tsqr_flat_r_kernel<WPC, MINB, PIPE>, 32 * WPC, smem) != cudaSuccess || nb < 1) {
      cudaGetLastError();
      nb = 1;
    }
    return nb * WPC;
  }
  static void launch(const FlatTsqrParams& p, cudaStream_t s) {
    static size_t pad = (size_t)-1;          // CQR_FLAT_SMEM_PAD=bytes: occupancy experiments (extra dynamic shared memory)
    if (pad == (size_t)-1) {
      const char* e = getenv("CQR_FLAT_SMEM_PAD");
      pad = e ? (size_t)atol(e) : 0;
      if (pad) cudaFuncSetAttribute(tsqr_flat_r_kernel<WPC, MINB, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + pad));
    }
    static PerDeviceOnce once;             // per_sm() set the attribute on the device that ran flat_init() only
    if (once.first()) cudaFuncSetAttribute(tsqr_flat_r_kernel<WPC, MINB, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + pad));
    if (p.gate != nullptr) {             // behind the Gram leaf: programmatic dependent launch, ordinary launch if the attribute is refused
      if (launch_pdl(tsqr_flat_r_kernel<WPC, MINB, PIPE>, dim3((unsigned)((p.chains + WPC - 1) / WPC)), dim3(32 * WPC), smem + pad, s, p) == cudaSuccess) return;
      cudaGetLastError();
    }
    tsqr_flat_r_kernel<WPC, MINB, PIPE><<<(p.chains + WPC - 1) / WPC, 32 * WPC, smem + pad, s>>>(p);
  }
};

static int g_flat_cfg = -1, g_flat_per_sm = 0;
static void flat_init() {
  if (g_flat_cfg >= 0) return;
  const char* e = getenv("CQR_FLAT_CFG");
  g_flat_cfg = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : kFlatDefaultCfg;
  switch (g_flat_cfg) {
    case 0: g_flat_per_sm = FlatCfg<4, 3, false>::per_sm(); break;
    case 1: g_flat_per_sm = FlatCfg<4, 2, false>::per_sm(); break;
    default: g_flat_per_sm = FlatCfg<4, 2, true>::per_sm(); break;
  }
}

// Chains the device keeps resident at once (one wave): SMs x resident CTAs x warps per CTA.
int flat_tsqr_max_chains(int sm_count) {
  flat_init();
  return sm_count * g_flat_per_sm;
}

// Implicit-Q variant (reflectors over A, taus to p.tau_out) and its expansion: 2 CTAs x 4 warps per SM, so the chain
// geometry is the one flat_tsqr_max_chains reports for the pipelined R-only kernel.
void launch_tsqr_flat_keep(const FlatTsqrParams& p, cudaStream_t s) {
  if (p.chains <= 0) return;
  ++g_launches;
  constexpr size_t smem = (size_t)4 * kFlatWarpFloats * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(tsqr_flat_r_kernel<4, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tsqr_flat_r_kernel<4, 2, true, true><<<(p.chains + 3) / 4, 128, smem, s>>>(p);
}

void launch_tsqr_flat_apply(const FlatApplyParams& p, cudaStream_t s) {
  if (p.chains <= 0) return;
  ++g_launches;
  constexpr size_t smem = (size_t)4 * 64 * 64 * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(tsqr_flat_apply_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tsqr_flat_apply_kernel<4, 2><<<(p.chains + 3) / 4, 128, smem, s>>>(p);
}

void launch_tsqr_flat_r(const FlatTsqrParams& p, cudaStream_t s, bool pair) {
  if (p.chains <= 0) return;
  flat_init();
  ++g_launches;
  if (pair) {                                  // two pivot columns per reduction: same 2 x 4 warps per SM geometry
    constexpr size_t smem = (size_t)4 * kFlatWarpFloats * sizeof(float);
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(tsqr_flat_r_kernel<4, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tsqr_flat_r_kernel<4, 2, false, false, true><<<(p.chains + 3) / 4, 128, smem, s>>>(p);
    return;
  }
  switch (g_flat_cfg) {
    case 0: FlatCfg<4, 3, false>::launch(p, s); break;
    case 1: FlatCfg<4, 2, false>::launch(p, s); break;
    default: FlatCfg<4, 2, true>::launch(p, s); break;
  }
}

// ================================================================================================================
// Batched small QR (BASELINE config 4: 65 536 independent 64 x 64 matrices), one WARP per matrix.
//
// Same machinery as the flat leaf above -- a dense Householder step is the flat step with "row j of R" being the
// matrix's own row j -- with three differences:
//   * rows are dealt to the four row parts in pairs (lane part h holds rows 8 k + 2 h + {0, 1}, k = 0..7), so the rows
//     above pivot group I0 are the register pairs k < I0 of EVERY lane and are skipped statically (rows shrink);
//   * the pivot-row entries a(j, c) come from a shuffle (they live in part h = (j % 8) / 2, pair I0) instead of shared
//     memory, and are written back by the rank-1 update itself: x's pivot entry is patched to u, so v_j = 1;
//   * the reflector is stored: column j below the diagonal becomes x / u, the diagonal beta, tau goes to tau[b n + j]
//     (LAPACK geqrf storage, qr.c:144-167 scalars; an all-zero tail gives tau = 0 where the reference gives 2 or NaN).
// 65 536 matrices: 3.5 ms with the two-threads-per-column CTA kernel (tile_qr.cu), see DESIGN.md for this one.
template <int I0>
__device__ __forceinline__ void dense_steps(f32x2 (&b)[8][8], float* __restrict__ xs, float* __restrict__ staus, const int q,
                                            const int h, const int n) {
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * I0 + jj;
    if (j >= n) break;                       // warp-uniform
    float* xb = xs + (jj & 1) * 72;
    const int hj = jj >> 1;                  // row part holding row j (pair I0)
    const bool hi_half = (jj & 1) != 0;      // ... in the high half of the pair
    if (q == jj) {                           // publish x = column j strictly below the diagonal (zeros on and above it)
      float lo, hi;
      funpack2(b[I0][I0], lo, hi);
      if (h == hj) xb[64] = hi_half ? hi : lo;                       // alpha = a(j, j)
      const float zlo = (2 * h > jj) ? lo : 0.f, zhi = (2 * h + 1 > jj) ? hi : 0.f;
      *reinterpret_cast<float2*>(xb + 8 * I0 + 2 * h) = make_float2(zlo, zhi);
#pragma unroll
      for (int k = I0 + 1; k < 8; ++k) *reinterpret_cast<f32x2*>(xb + 8 * k + 2 * h) = b[I0][k];
    }
    // pivot-row entries of my columns: from the lane of my column group in part hj
    float r[8];
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      float lo, hi;
      funpack2(b[i][I0], lo, hi);
      r[i] = __shfl_sync(kFull, hi_half ? hi : lo, q + 8 * hj);
    }
    __syncwarp();
    f32x2 x[8];
#pragma unroll
    for (int k = I0; k < 8; ++k) x[k] = *reinterpret_cast<const f32x2*>(xb + 8 * k + 2 * h);
    const float alpha = xb[64];
    f32x2 s2 = 0ull;
#pragma unroll
    for (int k = I0; k < 8; ++k) s2 = ffma2(x[k], x[k], s2);
    float sig = fsum2(s2);
    float d[8];
    {
      f32x2 d2[8];
#pragma unroll
      for (int i = I0; i < 8; ++i) d2[i] = 0ull;
#pragma unroll
      for (int k = I0; k < 8; ++k)
#pragma unroll
        for (int i = I0; i < 8; ++i) d2[i] = dfma2(x[k], b[i][k], d2[i]);
#pragma unroll
      for (int i = I0; i < 8; ++i) d[i] = fsum2(d2[i]);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float ts = __shfl_xor_sync(kFull, sig, o);
      float t[8];
#pragma unroll
      for (int i = I0; i < 8; ++i) t[i] = __shfl_xor_sync(kFull, d[i], o);
      sig += ts;
#pragma unroll
      for (int i = I0; i < 8; ++i) d[i] += t[i];
    }
    const float sj = fmaf(alpha, alpha, sig);
    const bool ok = (sig != 0.f) && (sj >= 1.2e-38f);
    const float sjs = ok ? sj : 1.f;
    const float rs = rsqrt_approx(sjs);
    float nrm = sjs * rs;
    nrm = fmaf(fmaf(-nrm, nrm, sjs), 0.5f * rs, nrm);
    const float bc = (alpha < 0.f) ? nrm : -nrm;
    const float u = alpha - bc;
    const float inv_u = ok ? rcp_newton(u) : 0.f;
    const float tau = ok ? -u * rcp_newton(bc) : 0.f;
    const float c2s = -tau * inv_u;
    if (q == 0 && h == 0) staus[j] = tau;
    // v = x / u below the diagonal and v_j = 1: patch x's pivot entry to u so the update also writes a(j, c) -= tau s
    if (h == hj) {
      float lo, hi;
      funpack2(x[I0], lo, hi);
      x[I0] = hi_half ? fpack2(lo, u) : fpack2(u, hi);
    }
#pragma unroll
    for (int i = I0; i < 8; ++i) {
      const bool act = (i > I0) || (q > jj);
      const float s = fmaf(d[i], inv_u, r[i]);
      const float nwv = act ? c2s * s : 0.f;
      const f32x2 nw2 = fpack2(nwv, nwv);
#pragma unroll
      for (int k = I0; k < 8; ++k) b[i][k] = ffma2(nw2, x[k], b[i][k]);
    }
    if (q == jj && ok) {                     // store the reflector: beta on the diagonal, x / u below, R above untouched
      float lo, hi, xlo, xhi;
      funpack2(b[I0][I0], lo, hi);
      funpack2(x[I0], xlo, xhi);             // the patched entry is never selected below (row == j takes beta)
      const int r0 = 2 * h, r1 = 2 * h + 1;
      lo = (r0 > jj) ? xlo * inv_u : (r0 == jj ? bc : lo);
      hi = (r1 > jj) ? xhi * inv_u : (r1 == jj ? bc : hi);
      b[I0][I0] = fpack2(lo, hi);
      const f32x2 iu2 = fpack2(inv_u, inv_u);
#pragma unroll
      for (int k = I0 + 1; k < 8; ++k) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b[I0][k]) : "l"(b[I0][k]), "l"(iu2));
    }
  }
}

template <int I0>
struct DenseGroups {
  static __device__ __forceinline__ void run(f32x2 (&b)[8][8], float* xs, float* staus, int q, int h, int n) {
    dense_steps<I0>(b, xs, staus, q, h, n);
    DenseGroups<I0 + 1>::run(b, xs, staus, q, h, n);
  }
};
template <>
struct DenseGroups<8> {
  static __device__ __forceinline__ void run(f32x2 (&)[8][8], float*, float*, int, int, int) {}
};

constexpr int kDenseWarpFloats = 2 * 72 + 64;   // double-buffered x (+ alpha) and the taus

template <int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) batched_qr_warp_kernel(float* __restrict__ base, long long stride, long long lda,
                                                                         int m, int n, int batch, float* __restrict__ tau_out,
                                                                         long long tau_stride, long long rows_total) {
  __shared__ __align__(16) float dsm[WPC * kDenseWarpFloats];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 7, h = lane >> 3;
  const long long mat = (long long)blockIdx.x * WPC + w;
  if (mat >= batch) return;
  float* xs = dsm + w * kDenseWarpFloats;
  float* staus = xs + 2 * 72;
  float* A = base + mat * stride;
  if (rows_total > 0) {                      // matrices are row blocks `stride` rows apart of one tall matrix: the last may be short
    const long long rem = rows_total - mat * stride;
    if (rem < m) m = (int)rem;
  }
  const bool vec = (lda % 2 == 0) && (stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(base) & 7) == 0);
  f32x2 b[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = q + 8 * i;
    const float* col = A + (long long)c * lda + 2 * h;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r0 = 8 * k + 2 * h;
      if (c < n && vec && r0 + 1 < m) {
        b[i][k] = *reinterpret_cast<const f32x2*>(col + 8 * k);
      } else {
        const float lo = (c < n && r0 < m) ? col[8 * k] : 0.f;
        const float hi = (c < n && r0 + 1 < m) ? col[8 * k + 1] : 0.f;
        b[i][k] = fpack2(lo, hi);
      }
    }
  }
  staus[lane] = 0.f; staus[lane + 32] = 0.f;
  __syncwarp();
  DenseGroups<0>::run(b, xs, staus, q, h, n);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = q + 8 * i;
    float* col = A + (long long)c * lda + 2 * h;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r0 = 8 * k + 2 * h;
      if (c < n && vec && r0 + 1 < m) {
        *reinterpret_cast<f32x2*>(col + 8 * k) = b[i][k];
      } else {
        float lo, hi;
        funpack2(b[i][k], lo, hi);
        if (c < n && r0 < m) col[8 * k] = lo;
        if (c < n && r0 + 1 < m) col[8 * k + 1] = hi;
      }
    }
  }
  for (int j = lane; j < n; j += 32) tau_out[mat * tau_stride + j] = staus[j];
}

void launch_batched_qr_warp(float* base, long long stride, long long lda, int m, int n, int batch, float* tau, cudaStream_t s) {
  if (batch <= 0) return;
  ++g_launches;
  constexpr int WPC = 4;
  batched_qr_warp_kernel<WPC, 2><<<(batch + WPC - 1) / WPC, 32 * WPC, 0, s>>>(base, stride, lda, m, n, batch, tau, (long long)n, 0);
}

// First 64-row block of every chain of the implicit-Q flat leaf: a dense Householder QR in place (R in the block's upper
// triangle, v below it, tau into the chain's first row of the [block][64] table).  Starting the chains from a true QR of
// real rows -- instead of from R = 0 on 64 virtual rows -- keeps the thin Q orthonormal for rank-deficient and badly
// conditioned inputs: with virtual rows, [0; A] = Q R puts part of Q's columns into the virtual rows whenever R is
// (nearly) singular, and dropping those rows loses orthogonality (1.5e3 n eps on a graded matrix with a duplicated column).
void launch_tsqr_flat_first_blocks(float* a, long long lda, long long m, int n, long long rows_per_chain, int chains, float* tau,
                                   cudaStream_t s) {
  if (chains <= 0) return;
  ++g_launches;
  constexpr int WPC = 4;
  batched_qr_warp_kernel<WPC, 2><<<(chains + WPC - 1) / WPC, 32 * WPC, 0, s>>>(a, rows_per_chain, lda, 64, n, chains, tau, rows_per_chain, m);
}

}  // namespace cqr
