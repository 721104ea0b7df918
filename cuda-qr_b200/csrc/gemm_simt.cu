// gemm_simt.cu -- exact-fp32 SIMT GEMMs and the element-wise helpers of the QR driver.
//
// The SIMT GEMMs serve (1) the legacy dgemm entry point (qr.c:443-459 is a plain fp32 triple
// loop, so its device core stays fp32-FMA exact), (2) shapes the TMA path cannot take
// (leading dimensions not a multiple of 4 floats, e.g. the reference's 6 x 4 demo), and
// (3) the cross-check the tcgen05 3xTF32 kernels are validated against.
#include "common.cuh"

namespace cqr {

std::atomic<long long> g_launches{0};

constexpr int BT = 64;   // output tile edge
constexpr int BK = 16;   // K chunk

// D[z] = A(Kz, :)^T B(Kz, :)   A: K x M (lda), B: K x N (ldb)
__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(int M, int N, int K, const float* __restrict__ a,
                                                           long long lda, const float* __restrict__ b, long long ldb,
                                                           float* __restrict__ d, long long ldd, int kper,
                                                           long long d_split_stride) {
  __shared__ float As[BK][BT + 4];
  __shared__ float Bs[BK][BT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BT, n0 = blockIdx.y * BT;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lk = tid & 15, lc = tid >> 4;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lc + 16 * i, k = k0 + lk;
      As[lk][c] = (k < kend && m0 + c < M) ? a[k + (long long)(m0 + c) * lda] : 0.f;
      Bs[lk][c] = (k < kend && n0 + c < N) ? b[k + (long long)(n0 + c) * ldb] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][tx + 16 * i]; bv[i] = Bs[k][ty + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dz = d + (long long)blockIdx.z * d_split_stride;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tx + 16 * i, n = n0 + ty + 16 * j;
      if (m < M && n < N) dz[m + (long long)n * ldd] = acc[i][j];
    }
}

// D = alpha A B + beta D   A: M x K (lda), B: K x N (ldb)
__global__ void __launch_bounds__(256) gemm_nn_simt_kernel(int M, int N, int K, float alpha,
                                                           const float* __restrict__ a, long long lda,
                                                           const float* __restrict__ b, long long ldb, float beta,
                                                           float* __restrict__ d, long long ldd) {
  __shared__ float As[BK][BT + 4];
  __shared__ float Bs[BK][BT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BT, n0 = blockIdx.y * BT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int am = tid & 63, ak = tid >> 6;   // A: 64 rows x 4 k per pass
  const int bk = tid & 15, bc = tid >> 4;   // B: 16 k x 16 cols per pass
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ak + 4 * i;
      As[ak + 4 * i][am] = (k < K && m0 + am < M) ? a[(m0 + am) + (long long)k * lda] : 0.f;
      const int c = bc + 16 * i, kk = k0 + bk;
      Bs[bk][c] = (kk < K && n0 + c < N) ? b[kk + (long long)(n0 + c) * ldb] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][tx + 16 * i]; bv[i] = Bs[k][ty + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tx + 16 * i, n = n0 + ty + 16 * j;
      if (m < M && n < N) {
        float* dp = d + m + (long long)n * ldd;
        float v = alpha * acc[i][j];
        if (beta != 0.f) v = fmaf(beta, *dp, v);
        *dp = v;
      }
    }
}

__global__ void reduce_splits_kernel(int M, int N, const float* __restrict__ part, long long ldp, long long stride,
                                     int splits, float* __restrict__ out, long long ldo) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[m + (long long)n * ldp + (long long)z * stride];
    out[m + (long long)n * ldo] = s;
  }
}

// X(kb x nc) = op(T) * (sum_z part_z), op(T) = T^T (trans != 0) or T, with T kb x kb upper triangular (zeros below):
// the split-K reduction of W = V^T C fused with the multiplication by the compact-WY T.  Replaces three launches
// (reduce, tensor GEMM with K = kb, copy) on the latency-critical narrow updates (inner block updates, look-ahead
// slice).  One CTA per 8 columns; thread i owns row i of X.
__global__ void __launch_bounds__(256) tw_fused_kernel(int kb, int nc, const float* __restrict__ part, long long ldp,
                                                       long long stride, int splits, const float* __restrict__ t,
                                                       long long ldt, int trans, float* __restrict__ x, long long ldx) {
  __shared__ __align__(16) float Ws[256][8];
  __shared__ float Ts[32][257];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * 8;
  const int kb32 = (kb + 31) & ~31;   // rows kb .. kb32 are read by the last k chunk (against zeros of T): keep them finite
  for (int idx = tid; idx < kb32 * 8; idx += 256) {
    const int k = idx % kb32, cc = idx / kb32;
    const int n = n0 + cc;
    float s = 0.f;
    if (n < nc && k < kb)
      for (int z = 0; z < splits; ++z) s += part[k + (long long)n * ldp + (long long)z * stride];
    Ws[k][cc] = s;
  }
  float acc[8];
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.f;
  const int i = tid;
  const int wlo = tid & ~31, whi = tid | 31;   // rows of this warp
  for (int k0 = 0; k0 < kb; k0 += 32) {
    __syncthreads();
    const int kn = min(32, kb - k0);
    if (trans) {   // op(T)[i][k] = T[k][i]
      for (int idx = tid; idx < 32 * kb; idx += 256) {
        const int kk = idx & 31, ii = idx >> 5;
        Ts[kk][ii] = (kk < kn) ? t[(k0 + kk) + (long long)ii * ldt] : 0.f;
      }
    } else {       // op(T)[i][k] = T[i][k]
      for (int idx = tid; idx < 32 * kb; idx += 256) {
        const int ii = idx % kb, kk = idx / kb;
        Ts[kk][ii] = (kk < kn) ? t[ii + (long long)(k0 + kk) * ldt] : 0.f;
      }
    }
    __syncthreads();
    // triangular structure: T^T couples row i with k <= i, T with k >= i (warp-uniform skip)
    const bool need = trans ? (k0 <= whi) : (k0 + 31 >= wlo);
    if (i < kb && need) {
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const float tv = Ts[kk][i];
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k0 + kk][0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k0 + kk][4]);
        acc[0] = fmaf(tv, w0.x, acc[0]); acc[1] = fmaf(tv, w0.y, acc[1]);
        acc[2] = fmaf(tv, w0.z, acc[2]); acc[3] = fmaf(tv, w0.w, acc[3]);
        acc[4] = fmaf(tv, w1.x, acc[4]); acc[5] = fmaf(tv, w1.y, acc[5]);
        acc[6] = fmaf(tv, w1.z, acc[6]); acc[7] = fmaf(tv, w1.w, acc[7]);
      }
    }
  }
  if (i < kb) {
#pragma unroll
    for (int cc = 0; cc < 8; ++cc)
      if (n0 + cc < nc) x[i + (long long)(n0 + cc) * ldx] = acc[cc];
  }
}

// mode 1: identity; 2: zero; 3: copy; 4: extract R; 5: extract V
template <int MODE>
__global__ void elementwise_kernel(long long m, int n, const float* __restrict__ a, long long lda, float* __restrict__ b,
                                   long long ldb, int aux) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    float v;
    if (MODE == 1) v = (i == j) ? 1.f : 0.f;
    else if (MODE == 2) v = 0.f;
    else if (MODE == 3) v = a[i + j * lda];
    else if (MODE == 4) v = (i <= j && i < aux) ? a[i + j * lda] : 0.f;   // aux = rows of A
    else {                                                                 // aux = d0: diagonal offset
      const long long dj = (long long)aux + j;
      v = (i < dj) ? 0.f : (i == dj ? 1.f : a[i + j * lda]);
    }
    b[i + j * ldb] = v;
  }
}

static dim3 ew_grid(long long m, int n) {
  return dim3((unsigned)((m + 255) / 256), (unsigned)(n < 1 ? 1 : (n > 4096 ? 4096 : n)), 1);
}

void launch_gemm_tn_simt(int M, int N, int K, const float* a, long long lda, const float* b, long long ldb, float* d,
                         long long ldd, int splits, long long d_split_stride, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  if (splits < 1) splits = 1;
  int kper = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  if (kper < BK) kper = BK;
  ++g_launches;
  dim3 grid((M + BT - 1) / BT, (N + BT - 1) / BT, splits);
  gemm_tn_simt_kernel<<<grid, 256, 0, s>>>(M, N, K, a, lda, b, ldb, d, ldd, kper, d_split_stride);
}

void launch_gemm_nn_simt(int M, int N, int K, float alpha, const float* a, long long lda, const float* b,
                         long long ldb, float beta, float* d, long long ldd, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  ++g_launches;
  dim3 grid((M + BT - 1) / BT, (N + BT - 1) / BT, 1);
  gemm_nn_simt_kernel<<<grid, 256, 0, s>>>(M, N, K, alpha, a, lda, b, ldb, beta, d, ldd);
}

void launch_reduce_splits(int M, int N, const float* part, long long ldp, long long stride, int splits, float* out,
                          long long ldo, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  ++g_launches;
  reduce_splits_kernel<<<ew_grid(M, N), 256, 0, s>>>(M, N, part, ldp, stride, splits, out, ldo);
}

void launch_tw_fused(int kb, int nc, const float* part, long long ldp, long long stride, int splits, const float* t,
                     long long ldt, int trans, float* x, long long ldx, cudaStream_t s) {
  ++g_launches;
  tw_fused_kernel<<<(nc + 7) / 8, 256, 0, s>>>(kb, nc, part, ldp, stride, splits, t, ldt, trans, x, ldx);
}

void launch_set_identity(float* a, long long lda, int m, int n, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  ++g_launches;
  elementwise_kernel<1><<<ew_grid(m, n), 256, 0, s>>>(m, n, nullptr, 0, a, lda, 0);
}

void launch_fill_zero(float* a, long long lda, long long m, int n, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  ++g_launches;
  elementwise_kernel<2><<<ew_grid(m, n), 256, 0, s>>>(m, n, nullptr, 0, a, lda, 0);
}

void launch_copy_matrix(long long m, int n, const float* a, long long lda, float* b, long long ldb, cudaStream_t s) {
  if (m <= 0 || n <= 0) return;
  ++g_launches;
  elementwise_kernel<3><<<ew_grid(m, n), 256, 0, s>>>(m, n, a, lda, b, ldb, 0);
}

void launch_extract_r(const float* a, long long lda, int m, int n, float* r, long long ldr, int r_rows,
                      cudaStream_t s) {
  if (r_rows <= 0 || n <= 0) return;
  ++g_launches;
  elementwise_kernel<4><<<ew_grid(r_rows, n), 256, 0, s>>>(r_rows, n, a, lda, r, ldr, m);
}

void launch_extract_v(const float* a, long long lda, long long mp, int b, int d0, float* v, long long ldv,
                      cudaStream_t s) {
  if (mp <= 0 || b <= 0) return;
  ++g_launches;
  elementwise_kernel<5><<<ew_grid(mp, b), 256, 0, s>>>(mp, b, a, lda, v, ldv, d0);
}

// Back substitution with one kb x kb upper-triangular diagonal block (kb <= 64): X = R^-1 B in place, one thread
// per right-hand side, R staged in shared memory.  The off-diagonal part of the solve is GEMM work (driver.cu).
__global__ void __launch_bounds__(128) trsm_upper_block_kernel(const float* __restrict__ r, long long ldr, int kb, float* __restrict__ b,
                                                               long long ldb, int nrhs, int* __restrict__ singular) {
  __shared__ float rs[64][65];
  for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) {
    const int i = idx % kb, j = idx / kb;
    rs[i][j] = (i <= j) ? r[i + (long long)j * ldr] : 0.f;
  }
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nrhs) return;
  float* bc = b + (long long)col * ldb;
  float x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x[i] = (i < kb) ? bc[i] : 0.f;
#pragma unroll
  for (int i = 63; i >= 0; --i) {
    if (i < kb) {
      const float d = rs[i][i];
      if (d == 0.f && col == 0) *singular = 1;
      const float xi = x[i] / d;
      x[i] = xi;
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < i) x[k] = fmaf(-rs[k][i], xi, x[k]);
    }
  }
#pragma unroll
  for (int i = 0; i < 64; ++i)
    if (i < kb) bc[i] = x[i];
}

void launch_trsm_upper_block(const float* r, long long ldr, int kb, float* b, long long ldb, int nrhs, int* singular, cudaStream_t s) {
  if (kb <= 0 || nrhs <= 0) return;
  ++g_launches;
  trsm_upper_block_kernel<<<(nrhs + 127) / 128, 128, 0, s>>>(r, ldr, kb, b, ldb, nrhs, singular);
}

}  // namespace cqr
