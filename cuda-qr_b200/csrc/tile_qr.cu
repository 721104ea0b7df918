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
// Reflector maths follow qr.c:144-167 (beta = -sign*norm, u = x0 + sign*norm, tau = sign*u/norm)
// except that a zero tail gives tau = 0 (H = I) instead of the reference's NaN.
#include "common.cuh"

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
  __shared__ float stau[2];

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

#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    for (int wj = 0; wj < 8; ++wj) {
      const int j = 8 * ci + wj;   // pivot column; owned by warp wj, register slot ci
      if (j >= nc) break;
      const int buf = j & 1;
      if (w == wj) {
        const int rj = ci >> 2;   // row j lives in lane j%32, register slot j/32 = ci/4 (static after unroll)
        float ss = 0.f;
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const float x = a[ci][ri];
          ss += (l + 32 * ri > j) ? x * x : 0.f;
        }
        ss = warp_sum(ss);
        const float alpha = __shfl_sync(kFull, a[ci][rj], j & 31);
        float beta = alpha, tau = 0.f, u = 1.f;
        if (ss != 0.f) {
          const float nrm = sqrtf(alpha * alpha + ss);
          beta = (alpha < 0.f) ? nrm : -nrm;
          u = alpha - beta;
          tau = (beta - alpha) / beta;
        }
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const int r = l + 32 * ri;
          const float x = a[ci][ri];
          const float vv = (r > j) ? x / u : (r == j ? 1.f : 0.f);
          vs[buf][r] = vv;
          if (r > j) a[ci][ri] = vv;
          else if (r == j) a[ci][ri] = beta;
        }
        if (l == 0) {
          stau[buf] = tau;
          p.tau[(long long)t * p.tau_stride + j] = tau;
        }
      }
      __syncthreads();
      const float tau = stau[buf];
      if (tau != 0.f) {
        float v[RI];
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) v[ri] = vs[buf][l + 32 * ri];
#pragma unroll
        for (int c2 = ci; c2 < 8; ++c2) {
          if (c2 > ci || w > wj) {
            float d = 0.f;
#pragma unroll
            for (int ri = 0; ri < RI; ++ri) d = fmaf(v[ri], a[c2][ri], d);
            d = warp_sum(d) * tau;
#pragma unroll
            for (int ri = 0; ri < RI; ++ri) a[c2][ri] = fmaf(-d, v[ri], a[c2][ri]);
          }
        }
      }
    }
  }

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
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      if (w + 8 * ci < nc) {
        float d = 0.f;
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) d = fmaf(v[ri], y[ci][ri], d);
        d = warp_sum(d) * tau;
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) y[ci][ri] = fmaf(-d, v[ri], y[ci][ri]);
      }
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

void launch_tile_qr(const TileQRParams& p, int tiles, int tile_rows, cudaStream_t s) {
  if (tiles <= 0) return;
  ++g_launches;
  if (tile_rows == 64) tile_qr_kernel<2><<<tiles, 256, 0, s>>>(p);
  else if (tile_rows == 128) tile_qr_kernel<4><<<tiles, 256, 0, s>>>(p);
  else tile_qr_kernel<8><<<tiles, 256, 0, s>>>(p);
}

void launch_tile_apply_q(const TileApplyParams& p, int tiles, int tile_rows, cudaStream_t s) {
  if (tiles <= 0) return;
  ++g_launches;
  const size_t smem = (size_t)tile_rows * CQR_SLOT * sizeof(float) + CQR_SLOT * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tile_apply_q_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * CQR_SLOT * 4 + 256);
    cudaFuncSetAttribute(tile_apply_q_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * CQR_SLOT * 4 + 256);
    attr_done = true;
  }
  if (tile_rows == 64) tile_apply_q_kernel<2><<<tiles, 256, smem, s>>>(p);
  else if (tile_rows == 128) tile_apply_q_kernel<4><<<tiles, 256, smem, s>>>(p);
  else tile_apply_q_kernel<8><<<tiles, 256, smem, s>>>(p);
}

}  // namespace cqr
