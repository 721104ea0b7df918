// tile_qr_core.cuh -- the register-resident Householder QR of one (32*RI) x 64 tile by a 256-thread CTA, as a device
// function: shared by tile_qr_kernel (tile_qr.cu) and the peer-memory R-tree kernel (rtree_peer.cu), which runs it
// between a flag wait and a remote store.  Layout and maths: see tile_qr.cu.
#pragma once
#include "common.cuh"

namespace cqr {

// a[ci][ri] = tile(row l + 32 ri, column w + 8 ci); vs: two TH-float broadcast buffers in shared memory; tau_out (may be
// null): the tile's 64 taus.  On exit a holds R on and above the diagonal and the reflectors below it.
template <int RI>
__device__ __forceinline__ void tile_qr_core(float (&a)[8][RI], const int nc, float (*vs)[32 * RI], float* __restrict__ tau_out, const int w,
                                             const int l) {
  // One __syncthreads and ONE round of warp reductions per column: the pivot column x (rows >= j,
  // zero above) is broadcast through shared memory, every warp forms s_c = x^T a_c for its live
  // columns together with s_j = x^T x, and derives the reflector scalars redundantly:
  //   beta = -sign(alpha) sqrt(s_j), u = alpha - beta, tau = -u / beta      (qr.c:149-152)
  //   v = (x - beta e_j) / u  =>  v^T a_c = (s_c - beta a_jc) / u
  // so no second reduction (norm first, then dots) and no second barrier is needed.
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    for (int wj = 0; wj < 8; ++wj) {
      const int j = 8 * ci + wj;   // pivot column; owned by warp wj, register slot ci
      if (j >= nc) break;
      const int buf = j & 1;
      const int rj = ci >> 2;      // row j lives in lane j%32, register slot j/32 = ci/4 (static after unroll)
      if (w == wj) {
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const int r = l + 32 * ri;
          vs[buf][r] = (r >= j) ? a[ci][ri] : 0.f;
        }
      }
      __syncthreads();
      float x[RI];
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) x[ri] = vs[buf][l + 32 * ri];
      const float alpha = vs[buf][j];
      // Branch-free over the (compile-time) live register slots c2 >= ci so the dot products and
      // the shuffle chains of all columns interleave; slot ci is live only in warps w > wj and is
      // masked out of the update below.  red[0] = x^T x, red[1 + c2 - ci] = x^T a_c2.
      constexpr int NLIVE = 8;   // upper bound; entries below ci are never touched after unrolling
      float red[NLIVE + 1], ajc[NLIVE];
      red[0] = 0.f;
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) red[0] = fmaf(x[ri], x[ri], red[0]);
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2) {
        red[1 + c2] = 0.f;
        ajc[c2] = 0.f;
        if (c2 >= ci) {
#pragma unroll
          for (int ri = 0; ri < RI; ++ri) red[1 + c2] = fmaf(x[ri], a[c2][ri], red[1 + c2]);
          ajc[c2] = __shfl_sync(kFull, a[c2][rj], j & 31);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {   // stage-wise so all live chains overlap; guards fold after unrolling ci
        float t[NLIVE + 1];
        t[0] = __shfl_xor_sync(kFull, red[0], o);
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2)
          if (c2 >= ci) t[1 + c2] = __shfl_xor_sync(kFull, red[1 + c2], o);
        red[0] += t[0];
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2)
          if (c2 >= ci) red[1 + c2] += t[1 + c2];
      }
      const float sj = red[0];
      float beta = 0.f, tau = 0.f, inv_u = 0.f, u = 1.f;
      if (sj != 0.f) {
        const float nrm = sqrtf(sj);
        beta = (alpha < 0.f) ? nrm : -nrm;
        u = alpha - beta;
        inv_u = 1.f / u;
        tau = -u / beta;
      }
      // a_c -= w_c v with v = x / u except v_j = 1: fold 1/u into w_c and patch x_j := u instead
      if ((j & 31) == l) x[rj] = u;
      const float scale = tau * inv_u * inv_u;
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2) {
        if (c2 >= ci) {
          float wc = scale * (red[1 + c2] - beta * ajc[c2]);
          if (c2 == ci && w <= wj) wc = 0.f;
#pragma unroll
          for (int ri = 0; ri < RI; ++ri) a[c2][ri] = fmaf(-wc, x[ri], a[c2][ri]);
        }
      }
      if (w == wj) {
        if (sj != 0.f) {
#pragma unroll
          for (int ri = 0; ri < RI; ++ri) {
            const int r = l + 32 * ri;
            if (r > j) a[ci][ri] = x[ri] * inv_u;
            else if (r == j) a[ci][ri] = beta;
          }
        }
        if (tau_out != nullptr && l == 0) tau_out[j] = tau;
      }
    }
  }

}

}  // namespace cqr
