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

__global__ void __launch_bounds__(256) hr_top_kernel(HrParams p) {
  extern __shared__ float hr_smem[];   // 3 x (64 x 65) + 64 floats: over the 48 KB static limit
  float (*M)[65] = reinterpret_cast<float (*)[65]>(hr_smem);
  float (*Tm)[65] = reinterpret_cast<float (*)[65]>(hr_smem + 64 * 65);
  float (*Ui)[65] = reinterpret_cast<float (*)[65]>(hr_smem + 2 * 64 * 65);
  float* sgn = hr_smem + 3 * 64 * 65;
  const int b = p.b, tid = threadIdx.x;
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx & 63, j = idx >> 6;
    M[i][j] = (i < b && j < b && i < p.mp) ? p.q[i + j * p.ldq] : 0.f;
    Tm[i][j] = 0.f;
    Ui[i][j] = 0.f;
  }
  __syncthreads();
  for (int j = 0; j < b; ++j) {
    if (tid == 0) {
      const float d = M[j][j];
      const float s = (d >= 0.f) ? -1.f : 1.f;
      sgn[j] = s;
      M[j][j] = d - s;
    }
    __syncthreads();
    const float piv = M[j][j];
    if (tid > j && tid < b) M[tid][j] /= piv;
    __syncthreads();
    const int nrem = b - 1 - j;
    for (int idx = tid; idx < nrem * nrem; idx += 256) {
      const int i = j + 1 + idx % nrem, k = j + 1 + idx / nrem;
      M[i][k] = fmaf(-M[i][j], M[j][k], M[i][k]);
    }
    __syncthreads();
  }
  // T Y1^T = -U S  (row i of T by forward substitution), and U^{-1} column by column.
  if (tid < b) {
    const int i = tid;
    for (int k = i; k < b; ++k) {
      float acc = -M[i][k] * sgn[k];
      for (int q = i; q < k; ++q) acc = fmaf(-Tm[i][q], M[k][q], acc);
      Tm[i][k] = acc;
    }
  } else if (tid >= 64 && tid < 64 + b) {
    const int k = tid - 64;
    Ui[k][k] = 1.f / M[k][k];
    for (int i = k - 1; i >= 0; --i) {
      float acc = 0.f;
      for (int q = i + 1; q <= k; ++q) acc = fmaf(M[i][q], Ui[q][k], acc);
      Ui[i][k] = -acc / M[i][i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx & 63, j = idx >> 6;
    p.uinv[i + j * 64] = Ui[i][j];
    if (i < b && j < b) {
      p.t[i + j * p.ldt] = Tm[i][j];
      const float y = (i > j) ? M[i][j] : (i == j ? 1.f : 0.f);
      if (i < p.mp) {
        p.a[i + j * p.lda] = (i > j) ? M[i][j] : sgn[i] * p.rt[i + j * p.ldrt];
        p.vbuf[i + j * p.ldv] = y;
        if (p.vlo) p.vlo[i + j * p.ldv] = tf32_lo(y);
      }
    }
  }
  if (tid < b) p.tau[tid] = Tm[tid][tid];
}

__global__ void __launch_bounds__(256) hr_rows_kernel(HrParams p) {
  __shared__ float Ui[64][64];
  const int b = p.b, tid = threadIdx.x;
  for (int idx = tid; idx < 64 * 64; idx += 256) Ui[idx & 63][idx >> 6] = p.uinv[idx];
  __syncthreads();
  const long long r = (long long)b + (long long)blockIdx.x * 256 + tid;
  if (r >= p.mp) return;
  const float* q = p.q + r;
#pragma unroll 1
  for (int j0 = 0; j0 < b; j0 += 16) {
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
        if (p.vlo) p.vlo[r + j * p.ldv] = tf32_lo(acc[jj]);
      }
    }
  }
}

void launch_hr_top(const HrParams& p, cudaStream_t s) {
  ++g_launches;
  constexpr size_t smem = (3 * 64 * 65 + 64) * sizeof(float);
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
  hr_rows_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(p);
}

// One CTA.  G (kb x kb, upper part read; its block columns are overwritten as scratch), tau (kb),
// T (kb x kb upper, zero below).  Block column J of T:
//     T_JJ           : larft column recurrence  T[0:i,i] = -tau_i T[0:i,0:i] G[0:i,i]
//     T[0:64J, J]    = -T[0:64J,0:64J] * (G[0:64J, J] * T_JJ)
__global__ void __launch_bounds__(256) build_t_kernel(float* g, long long ldg, const float* tau, float* t,
                                                      long long ldt, int kb, int have_diag) {
  extern __shared__ float tmp[];   // 64*(nb-1) x 64, ld = rowsJ
  __shared__ float Tjj[64][65];
  __shared__ float Gs[64][65];
  const int tid = threadIdx.x;
  const int nb = (kb + 63) / 64;
  for (int J = 0; J < nb; ++J) {
    const int c0 = 64 * J;
    const int bj = min(64, kb - c0);
    for (int idx = tid; idx < 64 * 64; idx += 256) {
      const int i = idx & 63, j = idx >> 6;
      const bool in = (i < bj && j < bj);
      Gs[i][j] = in ? g[(c0 + i) + (long long)(c0 + j) * ldg] : 0.f;
      Tjj[i][j] = (in && have_diag && i <= j) ? t[(c0 + i) + (long long)(c0 + j) * ldt] : 0.f;
    }
    __syncthreads();
    if (!have_diag) {
      for (int i = 0; i < bj; ++i) {
        const float ti = tau[c0 + i];
        if (tid < i) {
          float acc = 0.f;
          for (int k = tid; k < i; ++k) acc = fmaf(Tjj[tid][k], Gs[k][i], acc);
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
    // zero the block row below the diagonal block's columns (rows > c0+bj handled by later J) and
    // the part of this block column below the diagonal block
    for (int idx = tid; idx < (kb - c0 - bj) * bj; idx += 256) {
      const int i = c0 + bj + idx % (kb - c0 - bj), j = c0 + idx / (kb - c0 - bj);
      t[i + (long long)j * ldt] = 0.f;
    }
    if (J > 0) {
      const int rowsJ = c0;
      // tmp = G[0:rowsJ, J-block] * T_JJ   (thread per row, row held in registers)
      for (int r = tid; r < rowsJ; r += 256) {
        float grow[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) grow[k] = (k < bj) ? g[r + (long long)(c0 + k) * ldg] : 0.f;
#pragma unroll 4
        for (int c = 0; c < bj; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 64; ++k) acc = fmaf(grow[k], Tjj[k][c], acc);
          tmp[r + c * rowsJ] = acc;
        }
      }
      __syncthreads();
      // T[0:rowsJ, J-block] = -T[0:rowsJ, 0:rowsJ] * tmp  (T's explicit zeros keep the k loop uniform)
      for (int r = tid; r < rowsJ; r += 256) {
        for (int cb = 0; cb < bj; cb += 16) {
          float acc[16];
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) acc[cc] = 0.f;
          for (int k = 0; k < rowsJ; ++k) {
            const float trk = t[r + (long long)k * ldt];
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) acc[cc] = fmaf(trk, tmp[k + (cb + cc) * rowsJ], acc[cc]);
          }
#pragma unroll
          for (int cc = 0; cc < 16; ++cc)
            if (cb + cc < bj) t[r + (long long)(c0 + cb + cc) * ldt] = -acc[cc];
        }
      }
    }
    __syncthreads();
  }
}

void launch_build_t(const float* g, long long ldg, const float* tau, float* t, long long ldt, int kb,
                    int have_diag, cudaStream_t s) {
  ++g_launches;
  const size_t smem = (kb > 64) ? (size_t)((kb - 1) / 64) * 64 * 64 * sizeof(float) + 64 * 4 : 0;   // rowsJ <= 64*(nb-1)
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(build_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 448 * 64 * 4 + 256);
    attr_done = true;
  }
  build_t_kernel<<<1, 256, smem, s>>>(const_cast<float*>(g), ldg, tau, t, ldt, kb, have_diag);
}

}  // namespace cqr
