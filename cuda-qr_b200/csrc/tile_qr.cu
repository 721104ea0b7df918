// tile_qr.cu -- register-resident Householder QR of one (32*RI) x 64 tile per CTA.
//
// This one kernel is the leaf of the TSQR tree, every interior tree node (a tile of four
// stacked 64 x 64 R factors), the per-matrix worker of the batched 64 x 64 path, and the
// per-rank / cross-rank combine step of the multi-GPU R tree.  It replaces the reference's
// one-CTA-per-window panelHouseholderKernel (qr.cu:60-333) -- the reference walks a flat tree
// of 60-row windows serially with one launch each; here every tile of a level runs at once.
//
// Layout: 256 threads = 8 warps.  Warp w owns columns {w, w+8, ..., w+56}; lane l owns rows
// {l, l+32, ...}.  A column's dot products are warp-shuffle reductions (north_star: "reflectors
// generated with warp-shuffle norm/dot reductions"); the only block-level traffic is the
// 1 KB reflector broadcast through shared memory, one __syncthreads per column.
// Reflector maths follow qr.c:144-167 (beta = -sign*norm, u = x0 + sign*norm, tau = sign*u/norm;
// a length-1 reflector gets tau = 2 like the reference) except that an all-zero column gives
// tau = 0 (H = I) instead of the reference's NaN.
#include "common.cuh"
#include "tile_qr_core.cuh"

namespace cqr {

// rows of tile t: a negative rows_total means "every tile has -rows_total rows" (batched mode)
__device__ __forceinline__ int tile_rows_of(const TileSrc& s, int t, int th) {
  if (s.rows_total < 0) return (int)(-s.rows_total);
  const long long rem = s.rows_total - (long long)t * th;
  return rem >= th ? th : (rem > 0 ? (int)rem : 0);
}

template <int RI>
__global__ void __launch_bounds__(256, (RI <= 4 ? 3 : 2)) tile_qr_kernel(TileQRParams p) {
  constexpr int TH = 32 * RI;
  const int t = blockIdx.x;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __shared__ float vs[2][TH];
  if (p.gate != nullptr) {             // gated launches are programmatic dependent launches (launch_tile_qr)
    pdl_trigger();
    pdl_wait();
    if (*(const volatile int*)p.gate == 0) return;   // the Gram leaf produced R: nothing to do
  }

  const int rows = tile_rows_of(p.a, t, TH);
  float* src = p.a.base + (long long)t * p.a.tile_stride;
  const long long ld = p.a.ld;
  const int nc = p.ncols;

  float a[8][RI];
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    const int c = w + 8 * ci;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      a[ci][ri] = (c < nc && r < rows) ? src[r + c * ld] : 0.f;
    }
  }

  tile_qr_core<RI>(a, nc, vs, p.tau != nullptr ? p.tau + (long long)t * p.tau_stride : nullptr, w, l);

  if (p.write_back) {
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const int c = w + 8 * ci;
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) {
        const int r = l + 32 * ri;
        if (c < nc && r < rows) src[r + c * ld] = a[ci][ri];
      }
    }
  }
  if (p.r_out != nullptr) {
    float* dst = p.r_out + (long long)(t / p.fan) * p.r_tile_stride + (long long)(t % p.fan) * CQR_SLOT;
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const int c = w + 8 * ci;
#pragma unroll
      for (int ri = 0; ri < 2; ++ri) {
        const int r = l + 32 * ri;
        if (c < nc && r < p.r_rows) dst[r + c * p.r_ld] = (r <= c) ? a[ci][ri] : 0.f;
      }
    }
  }
}

// out_tile = Q_tile * [X; 0]: the tile's reflectors (from tile_qr_kernel with write_back)
// applied in reverse order to a block whose only non-zero rows are the 64-row seed X.
// Walking the TSQR tree root -> leaves with this kernel expands the implicit Q into the
// explicit thin Q (north_star item 4, "explicit Q formation").  Columns are independent:
// no block-level synchronisation inside the reflector loop.
template <int RI>
__global__ void __launch_bounds__(256, (RI <= 4 ? 3 : 2)) tile_apply_q_kernel(TileApplyParams p) {
  constexpr int TH = 32 * RI;
  extern __shared__ float sv[];   // TH x nref reflector tile (ld = TH), then 64 taus
  float* stau = sv + TH * CQR_SLOT;
  const int t = blockIdx.x;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int rows = tile_rows_of(p.v, t, TH);
  const float* vsrc = p.v.base + (long long)t * p.v.tile_stride;
  const int nref = p.nref, nc = p.nc;

  for (int idx = threadIdx.x; idx < TH * nref; idx += 256) {
    const int r = idx % TH, c = idx / TH;
    sv[idx] = (r < rows) ? vsrc[r + c * p.v.ld] : 0.f;
  }
  if (threadIdx.x < CQR_SLOT) stau[threadIdx.x] = (threadIdx.x < nref) ? p.tau[(long long)t * CQR_SLOT + threadIdx.x] : 0.f;

  const float* xs = p.x ? p.x + (long long)(t / p.fan) * p.x_tile_stride + (long long)(t % p.fan) * CQR_SLOT : nullptr;
  float y[8][RI];
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    const int c = w + 8 * ci;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      float val = 0.f;
      if (c < nc && r < CQR_SLOT) {
        if (xs) val = (r < p.x_rows) ? xs[r + c * p.x_ld] : 0.f;
        else val = (r == c) ? 1.f : 0.f;
      }
      y[ci][ri] = val;
    }
  }
  __syncthreads();

  for (int j = nref - 1; j >= 0; --j) {
    const float tau = stau[j];
    if (tau == 0.f) continue;
    float v[RI];
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      v[ri] = (r > j) ? sv[r + j * TH] : (r == j ? 1.f : 0.f);
    }
    // all 8 register columns unconditionally (columns >= nc hold zeros): no branches, so the eight
    // dot-product / shuffle chains overlap
    float d[8];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      d[ci] = 0.f;
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) d[ci] = fmaf(v[ri], y[ci][ri], d[ci]);
    }
    warp_sum_n(d);
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const float dc = d[ci] * tau;
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) y[ci][ri] = fmaf(-dc, v[ri], y[ci][ri]);
    }
  }

  const int orows = tile_rows_of(p.out, t, TH);
  float* dst = p.out.base + (long long)t * p.out.tile_stride;
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    const int c = w + 8 * ci;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri;
      if (c < nc && r < orows) dst[r + c * p.out.ld] = y[ci][ri];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Batched small QR (BASELINE config 4: independent 64 x 64 matrices, one CTA per matrix): a column lives in the
// registers of ONE thread pair, so a Householder step needs a single shuffle instead of a warp-wide reduction:
// the owner thread of column j forms norm / beta / tau / v from its own registers (qr.c:144-167), publishes v through
// shared memory (double-buffered: one barrier per step), and every thread to its right applies the reflector to its own
// column with a private dot product.  The step loop is fully unrolled so every register index is static and each step
// only touches rows >= j.  The tile kernel above spends its issue slots on 5-stage shuffle reductions for 16 elements
// per thread (6.9 ms for 65 536 matrices, 4.8 % of the HBM roof); here a matrix costs ~13 K warp-instructions.
// Two threads per column (h = tid & 1; thread h owns the 4-row chunks 8 i + 4 h .. +3, i = 0..7): 32 registers of
// matrix per thread instead of 64 doubles the resident warps (the per-step critical path is the owner's serial
// norm -> sqrt -> divide chain, hidden only by other matrices), and a column's dot needs one shuffle.
template <int J>
__device__ __forceinline__ void qr_col_step(float (&a)[32], const int c, const int h, const int n, float (*vs)[64], float* stau,
                                            float* __restrict__ tau_row) {
  constexpr int buf = J & 1, iJ = J / 8, hJ = (J / 4) & 1, eJ = J & 3;
  const int lane = threadIdx.x & 31;
  // row(i, e) = 8 i + 4 h + e > J  <=>  i > iJ, or i == iJ and 4 h + e > 4 hJ + eJ
  const int bnd = 4 * hJ + eJ - 4 * h;   // in chunk iJ: element e is below the diagonal iff e > bnd
  const unsigned pair_mask = 3u << (lane & ~1);                       // the two threads of my column
  const unsigned live_mask = __ballot_sync(kFull, c > J && c < n);    // threads that update in this step (whole pairs)
  if (c == J) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) s[e] = (e > bnd) ? a[4 * iJ + e] * a[4 * iJ + e] : 0.f;
#pragma unroll
    for (int i = iJ + 1; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = fmaf(a[4 * i + e], a[4 * i + e], s[e]);
    float sig = (s[0] + s[1]) + (s[2] + s[3]);
    sig += __shfl_xor_sync(pair_mask, sig, 1);
    const float alpha = __shfl_sync(pair_mask, a[4 * iJ + eJ], (lane & ~1) | hJ);
    const float sj = fmaf(alpha, alpha, sig);
    float tau = 0.f;
    if (sj != 0.f) {
      const float nrm = sqrtf(sj);
      const float beta = (alpha < 0.f) ? nrm : -nrm;
      const float u = alpha - beta;
      const float inv_u = 1.f / u;
      tau = -u / beta;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e > bnd) a[4 * iJ + e] *= inv_u;
      if (h == hJ) a[4 * iJ + eJ] = beta;
#pragma unroll
      for (int i = iJ + 1; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) a[4 * i + e] *= inv_u;
    }
    // publish v (rows > J are meaningful; v_J = 1 implicit) and tau
#pragma unroll
    for (int i = iJ; i < 8; ++i)
      *reinterpret_cast<float4*>(&vs[buf][8 * i + 4 * h]) = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    if (h == 0) { stau[buf] = tau; tau_row[J] = tau; }
  }
  __syncthreads();
  if (c > J && c < n) {   // both threads of a column take the same branch (shuffle below is pair-convergent)
    const float tau = stau[buf];
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const float4 t = *reinterpret_cast<const float4*>(&vs[buf][8 * iJ + 4 * h]);
      const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] = (e > bnd) ? tv[e] * a[4 * iJ + e] : 0.f;
      if (h == hJ) d[eJ] = a[4 * iJ + eJ];   // v_J = 1
    }
#pragma unroll
    for (int i = iJ + 1; i < 8; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(&vs[buf][8 * i + 4 * h]);
      d[0] = fmaf(t.x, a[4 * i], d[0]); d[1] = fmaf(t.y, a[4 * i + 1], d[1]);
      d[2] = fmaf(t.z, a[4 * i + 2], d[2]); d[3] = fmaf(t.w, a[4 * i + 3], d[3]);
    }
    float dot = (d[0] + d[1]) + (d[2] + d[3]);
    dot += __shfl_xor_sync(live_mask, dot, 1);
    const float w = tau * dot;
    if (w != 0.f) {
      const float4 t = *reinterpret_cast<const float4*>(&vs[buf][8 * iJ + 4 * h]);
      const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e > bnd) a[4 * iJ + e] = fmaf(-w, tv[e], a[4 * iJ + e]);
      if (h == hJ) a[4 * iJ + eJ] -= w;
#pragma unroll
      for (int i = iJ + 1; i < 8; ++i) {
        const float4 t2 = *reinterpret_cast<const float4*>(&vs[buf][8 * i + 4 * h]);
        a[4 * i] = fmaf(-w, t2.x, a[4 * i]); a[4 * i + 1] = fmaf(-w, t2.y, a[4 * i + 1]);
        a[4 * i + 2] = fmaf(-w, t2.z, a[4 * i + 2]); a[4 * i + 3] = fmaf(-w, t2.w, a[4 * i + 3]);
      }
    }
  }
}

template <int J>
struct QrColSteps {
  static __device__ __forceinline__ void run(float (&a)[32], int c, int h, int n, float (*vs)[64], float* stau, float* tau_row) {
    if (J < n) {   // uniform
      qr_col_step<J>(a, c, h, n, vs, stau, tau_row);
      QrColSteps<J + 1>::run(a, c, h, n, vs, stau, tau_row);
    }
  }
};
template <>
struct QrColSteps<64> {
  static __device__ __forceinline__ void run(float (&)[32], int, int, int, float (*)[64], float*, float*) {}
};

__global__ void __launch_bounds__(128) batched_qr_col_kernel(float* __restrict__ base, long long stride, long long lda, int m, int n,
                                                             float* __restrict__ tau_out) {
  __shared__ __align__(16) float vs[2][64];
  __shared__ float stau[2];
  const int c = threadIdx.x >> 1, h = threadIdx.x & 1;
  float* A = base + (long long)blockIdx.x * stride + (long long)c * lda + 4 * h;
  float a[32];
  const bool vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (stride % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r0 = 8 * i + 4 * h;
    if (c < n && vec && r0 + 3 < m) {
      const float4 v = *reinterpret_cast<const float4*>(A + 8 * i);
      a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) a[4 * i + e] = (c < n && r0 + e < m) ? A[8 * i + e] : 0.f;
    }
  }
  QrColSteps<0>::run(a, c, h, n, vs, stau, tau_out + (long long)blockIdx.x * n);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r0 = 8 * i + 4 * h;
    if (c < n && vec && r0 + 3 < m) {
      *reinterpret_cast<float4*>(A + 8 * i) = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c < n && r0 + e < m) A[8 * i + e] = a[4 * i + e];
    }
  }
}

void launch_batched_qr_col(float* base, long long stride, long long lda, int m, int n, int batch, float* tau, cudaStream_t s) {
  if (batch <= 0) return;
  ++g_launches;
  batched_qr_col_kernel<<<batch, 128, 0, s>>>(base, stride, lda, m, n, tau);
}

void launch_tile_qr(const TileQRParams& p, int tiles, int tile_rows, cudaStream_t s) {
  if (tiles <= 0) return;
  ++g_launches;
  if (p.gate != nullptr) {               // behind the Gram leaf: overlap the launch with the predecessor (see pdl_wait)
    cudaError_t e;
    if (tile_rows == 64) e = launch_pdl(tile_qr_kernel<2>, dim3(tiles), dim3(256), 0, s, p);
    else if (tile_rows == 128) e = launch_pdl(tile_qr_kernel<4>, dim3(tiles), dim3(256), 0, s, p);
    else e = launch_pdl(tile_qr_kernel<8>, dim3(tiles), dim3(256), 0, s, p);
    if (e == cudaSuccess) return;
    cudaGetLastError();                   // attribute refused: launch the ordinary way (pdl_wait is then a no-op)
  }
  if (tile_rows == 64) tile_qr_kernel<2><<<tiles, 256, 0, s>>>(p);
  else if (tile_rows == 128) tile_qr_kernel<4><<<tiles, 256, 0, s>>>(p);
  else tile_qr_kernel<8><<<tiles, 256, 0, s>>>(p);
}

void launch_tile_apply_q(const TileApplyParams& p, int tiles, int tile_rows, cudaStream_t s) {
  if (tiles <= 0) return;
  ++g_launches;
  const size_t smem = (size_t)tile_rows * CQR_SLOT * sizeof(float) + CQR_SLOT * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(tile_apply_q_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * CQR_SLOT * 4 + 256);
    cudaFuncSetAttribute(tile_apply_q_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * CQR_SLOT * 4 + 256);
  }
  if (tile_rows == 64) tile_apply_q_kernel<2><<<tiles, 256, smem, s>>>(p);
  else if (tile_rows == 128) tile_apply_q_kernel<4><<<tiles, 256, smem, s>>>(p);
  else tile_apply_q_kernel<8><<<tiles, 256, smem, s>>>(p);
}

}  // namespace cqr
