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
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(hr_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  hr_top_kernel<<<1, 256, smem, s>>>(p);
}

void launch_hr_rows(const HrParams& p, cudaStream_t s) {
  const long long rest = p.mp - p.b;
  if (rest <= 0) return;
  ++g_launches;
  hr_rows_kernel<<<(unsigned)((rest + 63) / 64), 256, 0, s>>>(p);
}

// One CTA, 256 threads as a 16 x 16 grid of 4 x 4 register tiles over 64 x 64 blocks.
// G (kb x kb Gram matrix, upper part read), tau (kb), T (kb x kb upper, zero below).  Block column J:
//     T_JJ        : larft column recurrence  T[0:i,i] = -tau_i T[0:i,0:i] G[0:i,i]   (unless have_diag)
//     tmp[rb]     = G[rb, J] T_JJ                              rb < J
//     T[rb, J]    = -sum_{kb = rb}^{J-1} T[rb, kb] tmp[kb]
__device__ __forceinline__ void block_mm_acc(float (&acc)[4][4], const float (*As)[65], const float (*Bs)[65], int tx,
                                             int ty) {
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { av[i] = As[tx + 16 * i][k]; bv[i] = Bs[k][ty + 16 * i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], bv[jj], acc[i][jj]);
  }
}

__global__ void __launch_bounds__(256) build_t_kernel(float* g, long long ldg, const float* tau, float* t,
                                                      long long ldt, int kb, int have_diag) {
  extern __shared__ float bt_smem[];
  float (*Tjj)[65] = reinterpret_cast<float (*)[65]>(bt_smem);
  float (*As)[65] = reinterpret_cast<float (*)[65]>(bt_smem + 64 * 65);
  float (*Bs)[65] = reinterpret_cast<float (*)[65]>(bt_smem + 2 * 64 * 65);
  float* tmp = bt_smem + 3 * 64 * 65;   // (nb-1) blocks of 64 x 64 (row-major within a block, ld 64)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nb = (kb + 63) / 64;
  for (int J = 0; J < nb; ++J) {
    const int c0 = 64 * J;
    const int bj = min(64, kb - c0);
    for (int idx = tid; idx < 64 * 64; idx += 256) {
      const int i = idx & 63, j = idx >> 6;
      const bool in = (i < bj && j < bj);
      As[i][j] = in ? g[(c0 + i) + (long long)(c0 + j) * ldg] : 0.f;   // G_JJ for the recurrence
      Tjj[i][j] = (in && have_diag && i <= j) ? t[(c0 + i) + (long long)(c0 + j) * ldt] : 0.f;
    }
    __syncthreads();
    if (!have_diag) {
      for (int i = 0; i < bj; ++i) {
        const float ti = tau[c0 + i];
        if (tid < i) {
          float acc = 0.f;
          for (int k = tid; k < i; ++k) acc = fmaf(Tjj[tid][k], As[k][i], acc);
          Tjj[tid][i] = -ti * acc;
        } else if (tid == i) {
          Tjj[i][i] = ti;
        }
        __syncthreads();
      }
    }
    for (int idx = tid; idx < 64 * 64; idx += 256) {
      const int i = idx & 63, j = idx >> 6;
      if (i < bj && j < bj) t[(c0 + i) + (long long)(c0 + j) * ldt] = (i <= j) ? Tjj[i][j] : 0.f;
    }
    for (int idx = tid; idx < (kb - c0 - bj) * bj; idx += 256) {   // zero below the diagonal block
      const int i = c0 + bj + idx % (kb - c0 - bj), j = c0 + idx / (kb - c0 - bj);
      t[i + (long long)j * ldt] = 0.f;
    }
    // tmp[rb] = G[rb, J] * T_JJ
    for (int rb = 0; rb < J; ++rb) {
      __syncthreads();
      for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int i = idx & 63, j = idx >> 6;
        As[i][j] = (j < bj) ? g[(64 * rb + i) + (long long)(c0 + j) * ldg] : 0.f;
      }
      __syncthreads();
      float acc[4][4] = {};
      block_mm_acc(acc, As, Tjj, tx, ty);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) tmp[rb * 4096 + (tx + 16 * i) * 64 + ty + 16 * jj] = acc[i][jj];
    }
    // T[rb, J] = -sum_kb T[rb, kb] tmp[kb]
    for (int rb = 0; rb < J; ++rb) {
      float acc[4][4] = {};
      for (int kbk = rb; kbk < J; ++kbk) {
        __syncthreads();
        for (int idx = tid; idx < 64 * 64; idx += 256) {
          const int i = idx & 63, j = idx >> 6;
          As[i][j] = t[(64 * rb + i) + (long long)(64 * kbk + j) * ldt];
          Bs[i][j] = tmp[kbk * 4096 + i * 64 + j];
        }
        __syncthreads();
        block_mm_acc(acc, As, Bs, tx, ty);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int col = ty + 16 * jj;
          if (col < bj) t[(64 * rb + tx + 16 * i) + (long long)(c0 + col) * ldt] = -acc[i][jj];
        }
    }
    __syncthreads();
  }
}

void launch_build_t(const float* g, long long ldg, const float* tau, float* t, long long ldt, int kb,
                    int have_diag, cudaStream_t s) {
  ++g_launches;
  const size_t smem = (size_t)(3 * 64 * 65 + ((kb - 1) / 64) * 4096) * sizeof(float);   // Tjj, As, Bs + (nb-1) tmp blocks
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(build_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (3 * 64 * 65 + 7 * 4096) * 4);
    attr_done = true;
  }
  build_t_kernel<<<1, 256, smem, s>>>(const_cast<float*>(g), ldg, tau, t, ldt, kb, have_diag);
}

}  // namespace cqr
