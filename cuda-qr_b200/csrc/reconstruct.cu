// reconstruct.cu -- Householder reconstruction of a TSQR panel and compact-WY T assembly.
//
// The TSQR tree gives a panel's R and an implicit Q that only tile kernels can apply.  The
// trailing update wants ONE pair (Y, T) per panel so that (I - Y T Y^T)^T C is two GEMMs
// (north_star items 2-3; the reference's per-window W/Y accumulation, qr.c:170-213, is the
// thing being replaced).  Following Ballard et al., "Reconstructing Householder vectors from
// TSQR" (IPDPS 2014): with the panel's explicit thin Q (mp x b),
//      Q - [S; 0] = Y * U          (LU without pivoting; S = diag(+-1), s_j = -sign of the pivot)
// gives unit-lower-trapezoidal Y, and  T = -U S Y1^{-T},  R_house = S R_tsqr,  tau_j = T_jj.
// Pivots are 1 + |q_jj| >= 1, so no pivoting is needed and the LU is backward stable.
//   hr_top : b x b LU of Q's top block, T, U^{-1}, S*R       (one CTA; the serial part)
//   hr_rows: Y2 = Q2 * U^{-1} for the remaining mp - b rows   (row-parallel)
//   build_t: T of an aggregated block of panels from the Gram matrix V^T V
//            (Joffrain et al.: T^{-1} + T^{-T} = V^T V), block recurrence on 64-wide panels.
#include "common.cuh"

namespace cqr {

// One CTA, 256 threads = 64 rows x 4 column phases.  One barrier per elimination step: row j is
// final before step j starts and only rows > j are written during it.
__global__ void __launch_bounds__(256) hr_top_kernel(HrParams p) {
  extern __shared__ float hr_smem[];   // M, L, T, U^-1 (64 x 65 each) + pivots + signs
  float (*M)[65] = reinterpret_cast<float (*)[65]>(hr_smem);
  float (*L)[65] = reinterpret_cast<float (*)[65]>(hr_smem + 64 * 65);
  float (*Tm)[65] = reinterpret_cast<float (*)[65]>(hr_smem + 2 * 64 * 65);
  float (*Ui)[65] = reinterpret_cast<float (*)[65]>(hr_smem + 3 * 64 * 65);
  float* piv = hr_smem + 4 * 64 * 65;
  float* sgn = piv + 64;
  const int b = p.b, tid = threadIdx.x;
  const int ti = tid & 63, tk = tid >> 6;
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx & 63, j = idx >> 6;
    M[i][j] = (i < b && j < b && i < p.mp) ? p.q[i + j * p.ldq] : 0.f;
    L[i][j] = 0.f;
    Ui[i][j] = (i == j) ? 1.f : 0.f;
  }
  // LU of (Q1 - S): s_j = -sign(pivot candidate), pivot = d - s_j with |pivot| >= 1
  for (int j = 0; j < b; ++j) {
    __syncthreads();
    const float d = M[j][j];
    const float sg = (d >= 0.f) ? -1.f : 1.f;
    const float pv = d - sg;
    if (tid == 0) { sgn[j] = sg; piv[j] = pv; }
    if (ti > j && ti < b) {
      const float li = M[ti][j] / pv;
      if (tk == 0) L[ti][j] = li;
      // stage through registers: the compiler cannot prove M[ti][.] and M[j][.] do not alias and would
      // otherwise serialise load -> fma -> store per element
      float mi[16], mj[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int k = j + 1 + tk + 4 * it;
        mi[it] = (k < b) ? M[ti][k] : 0.f;
        mj[it] = (k < b) ? M[j][k] : 0.f;
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int k = j + 1 + tk + 4 * it;
        if (k < b) M[ti][k] = fmaf(-li, mj[it], mi[it]);
      }
    }
  }
  __syncthreads();
  // B = -U S (U = triu(M) with the pivots on its diagonal); T solves T Y1^T = B, U^-1 by back substitution.
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx & 63, k = idx >> 6;
    float u = 0.f;
    if (i < b && k < b && i <= k) u = (i == k) ? piv[i] : M[i][k];
    Tm[i][k] = (k < b) ? -u * sgn[k] : 0.f;
  }
  // Rank-1 sweeps, two barriers per step: threads tk < 2 advance T (column k is final at step k),
  // threads tk >= 2 advance U^-1 (row q = b-1-k is final at step k).
  for (int k = 0; k < b; ++k) {
    const int q = b - 1 - k;
    __syncthreads();
    if (tk >= 2 && tk == 2 && ti >= q && ti < b) Ui[q][ti] = Ui[q][ti] / piv[q];
    __syncthreads();
    if (tk < 2) {
      if (ti <= k) {
        const float tik = Tm[ti][k];
        float tv[32], lv[32];
#pragma unroll
        for (int it = 0; it < 32; ++it) {
          const int k2 = k + 1 + tk + 2 * it;
          tv[it] = (k2 < b) ? Tm[ti][k2] : 0.f;
          lv[it] = (k2 < b) ? L[k2][k] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < 32; ++it) {
          const int k2 = k + 1 + tk + 2 * it;
          if (k2 < b) Tm[ti][k2] = fmaf(-tik, lv[it], tv[it]);
        }
      }
    } else if (ti >= q && ti < b) {
      const float xq = Ui[q][ti];
      float uv[32], mv[32];
#pragma unroll
      for (int it = 0; it < 32; ++it) {
        const int i = tk - 2 + 2 * it;
        uv[it] = (i < q) ? Ui[i][ti] : 0.f;
        mv[it] = (i < q) ? M[i][q] : 0.f;
      }
#pragma unroll
      for (int it = 0; it < 32; ++it) {
        const int i = tk - 2 + 2 * it;
        if (i < q) Ui[i][ti] = fmaf(-mv[it], xq, uv[it]);
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx & 63, j = idx >> 6;
    p.uinv[i + j * 64] = (i <= j && j < b) ? Ui[i][j] : 0.f;
    if (i < b && j < b) {
      p.t[i + j * p.ldt] = (i <= j) ? Tm[i][j] : 0.f;
      const float y = (i > j) ? L[i][j] : (i == j ? 1.f : 0.f);
      if (i < p.mp) {
        p.a[i + j * p.lda] = (i > j) ? L[i][j] : sgn[i] * p.rt[i + j * p.ldrt];
        p.vbuf[i + j * p.ldv] = y;
      }
    }
  }
  if (tid < b) p.tau[tid] = Tm[tid][tid];
}

// Y2 = Q2 U^-1: 64 rows per CTA, 4 threads per row (16 output columns each).
__global__ void __launch_bounds__(256) hr_rows_kernel(HrParams p) {
  __shared__ float Ui[64][64];
  const int b = p.b, tid = threadIdx.x;
  for (int idx = tid; idx < 64 * 64; idx += 256) Ui[idx & 63][idx >> 6] = p.uinv[idx];
  __syncthreads();
  const int rr = tid & 63, j0 = (tid >> 6) * 16;
  const long long r = (long long)b + (long long)blockIdx.x * 64 + rr;
  if (r >= p.mp || j0 >= b) return;
  const float* q = p.q + r;
  float acc[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.f;
  const int kend = min(b, j0 + 16);
  for (int k = 0; k < kend; ++k) {
    const float qk = q[k * p.ldq];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) acc[jj] = fmaf(qk, Ui[k][j0 + jj], acc[jj]);
  }
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    const int j = j0 + jj;
    if (j < b) {
      p.a[r + j * p.lda] = acc[jj];
      p.vbuf[r + j * p.ldv] = acc[jj];
    }
  }
}

void launch_hr_top(const HrParams& p, cudaStream_t s) {
  ++g_launches;
  constexpr size_t smem = (4 * 64 * 65 + 128) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(hr_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  hr_top_kernel<<<1, 256, smem, s>>>(p);
}

void launch_hr_rows(const HrParams& p, cudaStream_t s) {
  const long long rest = p.mp - p.b;
  if (rest <= 0) return;
  ++g_launches;
  hr_rows_kernel<<<(unsigned)((rest + 63) / 64), 256, 0, s>>>(p);
}

// Compact-WY T of kb aggregated reflectors from the Gram matrix G = V^T V (upper part read) and tau.
//   diagonal 64 x 64 blocks: larft column recurrence  T[0:i,i] = -tau_i T[0:i,0:i] G[0:i,i]  (build_t_diag_kernel,
//                            one CTA per block; skipped when the panel kernels already wrote them: have_diag)
//   block column J > 0     : T[0:c0, J] = -T[0:c0, 0:c0] (G[0:c0, J] T_JJ),  c0 = 64 J     (build_t_offdiag_kernel)
// The off-diagonal kernel is ONE cluster of 8 CTAs; CTA r owns columns 8r .. 8r+7 of every block column, thread i
// owns row i.  Block columns are sequential (column J needs T[0:c0, 0:c0]) and separated by a cluster barrier.
// An earlier single-CTA version spent ~80 us per outer block here, FMA-issue bound on one SM and on the critical path
// of every block (panel chain -> T -> look-ahead slice).
__global__ void __launch_bounds__(64) build_t_diag_kernel(const float* __restrict__ g, long long ldg,
                                                          const float* __restrict__ tau, float* __restrict__ t,
                                                          long long ldt, int kb) {
  __shared__ float Tjj[64][65], Gjj[64][65];
  const int tid = threadIdx.x, c0 = 64 * blockIdx.x, bj = min(64, kb - c0);
  for (int j = 0; j < 64; ++j) {
    const bool in = tid < bj && j < bj;
    Gjj[tid][j] = in ? g[(c0 + tid) + (long long)(c0 + j) * ldg] : 0.f;
    Tjj[tid][j] = 0.f;
  }
  __syncthreads();
  for (int i = 0; i < bj; ++i) {
    const float ti = tau[c0 + i];
    if (tid < i) {
      float acc = 0.f;
      for (int k = tid; k < i; ++k) acc = fmaf(Tjj[tid][k], Gjj[k][i], acc);
      Tjj[tid][i] = -ti * acc;
    } else if (tid == i) {
      Tjj[i][i] = ti;
    }
    __syncthreads();
  }
  if (tid < bj)
    for (int j = 0; j < bj; ++j) t[(c0 + tid) + (long long)(c0 + j) * ldt] = (tid <= j) ? Tjj[tid][j] : 0.f;
}

__global__ void __launch_bounds__(512) build_t_offdiag_kernel(const float* __restrict__ g, long long ldg, float* t,
                                                              long long ldt, int kb) {
  __shared__ float Tj8[64][8];       // T_JJ[:, my 8 columns]
  __shared__ __align__(16) float tmp8[512][8];   // (G[0:c0, J] T_JJ)[:, my 8 columns]
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  const int i = threadIdx.x;
  const int nb = (kb + 63) / 64;
  for (int J = 0; J < nb; ++J) {
    const int c0 = 64 * J, bj = min(64, kb - c0);
    const int col0 = c0 + 8 * (int)r;                       // my columns: col0 .. col0+7 (those < kb)
    // zero everything below the diagonal block in my columns (T is applied as a dense kb x kb operand)
    for (int idx = i; idx < (kb - c0 - bj) * 8; idx += 512) {
      const int row = c0 + bj + idx % (kb - c0 - bj), cc = idx / (kb - c0 - bj);
      if (col0 + cc < kb) t[row + (long long)(col0 + cc) * ldt] = 0.f;
    }
    if (J > 0) {
      for (int idx = i; idx < 64 * 8; idx += 512) {
        const int k = idx & 63, cc = idx >> 6;
        Tj8[k][cc] = (k < bj && col0 + cc < kb) ? __ldcg(t + (c0 + k) + (long long)(col0 + cc) * ldt) : 0.f;
      }
      __syncthreads();
      float acc[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.f;
      if (i < c0) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
          const float gv = (k < bj) ? g[i + (long long)(c0 + k) * ldg] : 0.f;
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) acc[cc] = fmaf(gv, Tj8[k][cc], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) tmp8[i][cc] = acc[cc];
      }
      __syncthreads();
      if (i < c0) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.f;
        // T[0:c0, 0:c0] is upper triangular: row i couples with k >= i; warp-uniform start keeps loads coalesced
        const int kstart = i & ~31;
#pragma unroll 4
        for (int k = kstart; k < c0; ++k) {
          const float tv = __ldcg(t + i + (long long)k * ldt);   // zero below the diagonal
          const float4 w0 = *reinterpret_cast<const float4*>(&tmp8[k][0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&tmp8[k][4]);
          acc[0] = fmaf(tv, w0.x, acc[0]); acc[1] = fmaf(tv, w0.y, acc[1]);
          acc[2] = fmaf(tv, w0.z, acc[2]); acc[3] = fmaf(tv, w0.w, acc[3]);
          acc[4] = fmaf(tv, w1.x, acc[4]); acc[5] = fmaf(tv, w1.y, acc[5]);
          acc[6] = fmaf(tv, w1.z, acc[6]); acc[7] = fmaf(tv, w1.w, acc[7]);
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
          if (col0 + cc < kb) t[i + (long long)(col0 + cc) * ldt] = -acc[cc];
      }
    }
    // block column J complete on every CTA before anyone reads it as part of T[0:c0', 0:c0']
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

void launch_build_t(const float* g, long long ldg, const float* tau, float* t, long long ldt, int kb,
                    int have_diag, cudaStream_t s) {
  const int nb = (kb + 63) / 64;
  if (!have_diag) {
    ++g_launches;
    build_t_diag_kernel<<<nb, 64, 0, s>>>(g, ldg, tau, t, ldt, kb);
  }
  ++g_launches;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8, 1, 1);
  cfg.blockDim = dim3(512, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, build_t_offdiag_kernel, g, ldg, t, ldt, kb);
}

}  // namespace cqr
