// f64_qr.cu -- double-precision variant of the blocked Householder QR.
//
// The reference contemplates `Scalar double` (qr.c:9 "can be float or double", bank-size switch qr.cu:747-754) but ships
// and measures float only.  This is the same path in fp64: 32-column panels factored by ONE cooperative launch (every
// CTA owns a row slab; a column step is a pass of partial dots -- x^T a_c for all 32 columns at once, which is the norm,
// the update's inner products and the new column of V^T V in one sweep -- a grid-wide sum through double atomics and a
// grid barrier, then the rank-1 update), the panel's compact-WY T by back substitution on V^T V like the fp32 panels,
// and the trailing update (I - V T V^T)^T C as three fp64 kernels (W = V^T C, X = T^T W, C -= V X).  B200 has no
// fp64-capable tcgen05 path and its DMMA rate is below the fp64 FMA pipe's, so the contraction runs on DFMA
// (37 TFLOP/s peak); with K = 32 the update is HBM-bound, which is what these kernels are sized for.
// Scalar formulas: qr.c:144-152 in LAPACK form (beta = -sign(alpha) ||x||, u = alpha - beta, tau = -u / beta, v = x / u),
// a zero column gives tau = 0.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace cqr {

namespace {

constexpr int DB = 32;          // fp64 panel width

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

struct DPanelParams {
  double* a; long long lda;     // panel at its diagonal: mp x b, LAPACK storage on exit
  long long mp; int b;
  double* tau;                  // b
  double* v; long long ldv;     // explicit V (unit diagonal, zeros above), mp x b
  double* t; int ldt;           // b x b compact-WY T (upper)
  double* acc;                  // 3 x 32 grid-wide accumulators (zero on entry, zero on exit)
  double* g;                    // 32 x 32 scratch: strictly upper part of V^T V
};

__global__ void __launch_bounds__(256) dpanel_kernel(DPanelParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[8][DB];
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gthreads = (long long)gridDim.x * blockDim.x;
  const int b = p.b;
  for (int j = 0; j < b; ++j) {
    double* acc = p.acc + (j % 3) * DB;
    // row j of the panel (stable since the last barrier): read now, so that no barrier is needed between the totals
    // and the update that rewrites it
    double wc[DB];
#pragma unroll
    for (int c = 0; c < DB; ++c) wc[c] = (c < b) ? p.a[j + (long long)c * p.lda] : 0.0;
    const double alpha = p.a[j + (long long)j * p.lda];
    // ---- x^T a_c over rows > j for every column c of the panel (x = column j)
    double d[DB];
#pragma unroll
    for (int c = 0; c < DB; ++c) d[c] = 0.0;
    for (long long r = j + 1 + gtid; r < p.mp; r += gthreads) {   // strictly below the pivot row: row j joins in closed form
      const double x = p.a[r + (long long)j * p.lda];
#pragma unroll
      for (int c = 0; c < DB; ++c)
        if (c < b) d[c] = fma(x, p.a[r + (long long)c * p.lda], d[c]);
    }
#pragma unroll
    for (int c = 0; c < DB; ++c) {
      const double s = warp_sum_d(d[c]);
      if (l == 0) red[w][c] = s;
    }
    __syncthreads();
    if (tid < b) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][tid];
      if (s != 0.0) atomicAdd(acc + tid, s);
    }
    if (blockIdx.x == 0 && tid < DB) p.acc[((j + 1) % 3) * DB + tid] = 0.0;   // next step's accumulators
    grid.sync();
    // ---- reflector scalars (every thread, from the same totals)
    const double sub = acc[j];                       // ||x||^2 below the pivot (summed apart from alpha^2, as dlarfg does: no cancellation)
    const double sj = fma(alpha, alpha, sub);
    double beta = alpha, tau = 0.0, inv_u = 0.0;
    if (sub > 0.0 && sj > 0.0) {
      const double nrm = sqrt(sj);
      beta = alpha < 0.0 ? nrm : -nrm;
      const double u = alpha - beta;
      inv_u = 1.0 / u;
      tau = -u / beta;
    }
    // w_c = tau v^T a_c for the columns right of j; (V^T V)(c, j) for the finished ones
#pragma unroll
    for (int c = 0; c < DB; ++c) {
      const double vta = fma(acc[c], inv_u, wc[c]);               // v_j^T a_c: v_j = 1 at row j, x / u below
      wc[c] = (c > j && c < b) ? tau * vta : 0.0;
      if (c < j && blockIdx.x == 0 && tid == 0) p.g[c + j * DB] = (tau != 0.0) ? vta : 0.0;
    }
    for (long long r = j + gtid; r < p.mp; r += gthreads) {
      const double x = p.a[r + (long long)j * p.lda];
      const double vr = (r == j) ? 1.0 : x * inv_u;
#pragma unroll
      for (int c = 0; c < DB; ++c)
        if (c > j && c < b) p.a[r + (long long)c * p.lda] -= wc[c] * vr;
      p.v[r + (long long)j * p.ldv] = (tau != 0.0 || r == j) ? vr : 0.0;
      p.a[r + (long long)j * p.lda] = (r == j) ? beta : (tau != 0.0 ? vr : x);
    }
    for (long long r = gtid; r < j; r += gthreads) p.v[r + (long long)j * p.ldv] = 0.0;
    if (gtid == 0) p.tau[j] = tau;
    grid.sync();
  }
  // ---- T = (diag(1/tau) + striu(V^T V))^-1, one thread per column (CTA 0)
  if (blockIdx.x == 0 && tid < b) {
    const int k = tid;
    double tc[DB];
#pragma unroll
    for (int i = 0; i < DB; ++i) tc[i] = 0.0;
    const double tk = p.tau[k];
#pragma unroll
    for (int i = DB - 1; i >= 0; --i) {
      if (i > k || i >= b) continue;
      if (i == k) { tc[i] = tk; continue; }
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < DB; ++q)
        if (q > i && q <= k) s = fma(p.g[i + q * DB], tc[q], s);
      tc[i] = -p.tau[i] * s;
    }
    for (int i = 0; i < b; ++i) p.t[i + (long long)k * p.ldt] = (i <= k) ? tc[i] : 0.0;
  }
  if (blockIdx.x == 0 && tid < 3 * DB) p.acc[tid] = 0.0;   // every read of the accumulators is behind the last grid barrier
}

// W(kb x nc) += V(rows, kb)^T C(rows, nc) over a chunk of rows: 32 x 32 output tile per CTA, atomically accumulated.
__global__ void __launch_bounds__(256) dgemm_tn_kernel(const double* __restrict__ v, long long ldv, const double* __restrict__ c, long long ldc,
                                                       long long rows, int kb, int nc, double* __restrict__ wout, int ldw, int rows_per_cta) {
  __shared__ double sv[64][DB + 1], sc[64][DB + 1];
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * DB;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  const int ki = (tid >> 4) * 2, ci = (tid & 15) * 2;      // 2 x 2 outputs per thread
  double o00 = 0, o01 = 0, o10 = 0, o11 = 0;
  for (long long rb = r0; rb < r1; rb += 64) {
    for (int e = tid; e < 64 * DB; e += 256) {
      const int rr = e & 63, cc = e >> 6;
      const long long r = rb + rr;
      sv[rr][cc] = (r < r1 && cc < kb) ? v[r + (long long)cc * ldv] : 0.0;
      sc[rr][cc] = (r < r1 && c0 + cc < nc) ? c[r + (long long)(c0 + cc) * ldc] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) {
      const double v0 = sv[rr][ki], v1 = sv[rr][ki + 1], x0 = sc[rr][ci], x1 = sc[rr][ci + 1];
      o00 = fma(v0, x0, o00); o01 = fma(v0, x1, o01); o10 = fma(v1, x0, o10); o11 = fma(v1, x1, o11);
    }
    __syncthreads();
  }
  if (ki < kb && c0 + ci < nc) atomicAdd(wout + ki + (long long)(c0 + ci) * ldw, o00);
  if (ki < kb && c0 + ci + 1 < nc) atomicAdd(wout + ki + (long long)(c0 + ci + 1) * ldw, o01);
  if (ki + 1 < kb && c0 + ci < nc) atomicAdd(wout + ki + 1 + (long long)(c0 + ci) * ldw, o10);
  if (ki + 1 < kb && c0 + ci + 1 < nc) atomicAdd(wout + ki + 1 + (long long)(c0 + ci + 1) * ldw, o11);
}

// X = op(T) W for the kb x kb upper-triangular T (trans: T^T); W is consumed (zeroed) so the accumulator is ready again.
__global__ void dtmul_kernel(const double* __restrict__ t, int ldt, int trans, double* __restrict__ w, int ldw, int kb, int nc, double* __restrict__ x,
                             int ldx) {
  __shared__ double sw[DB];
  const int col = blockIdx.x, k = threadIdx.x;
  if (k < kb) sw[k] = w[k + (long long)col * ldw];
  __syncthreads();
  if (k < kb) {
    double s = 0.0;
    for (int i = 0; i < kb; ++i) s = fma(trans ? t[i + (long long)k * ldt] : t[k + (long long)i * ldt], sw[i], s);
    x[k + (long long)col * ldx] = s;
    w[k + (long long)col * ldw] = 0.0;
  }
}

// C(rows x nc) -= V(rows x kb) X(kb x nc): 128 x 32 tile per CTA, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256) dgemm_nn_kernel(const double* __restrict__ v, long long ldv, const double* __restrict__ x, int ldx,
                                                       double* __restrict__ c, long long ldc, long long rows, int kb, int nc) {
  __shared__ double sv[DB][128 + 1], sx[DB][DB + 1];
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * 128;
  const int c0 = blockIdx.y * DB;
  for (int e = tid; e < 128 * DB; e += 256) {
    const int rr = e & 127, kk = e >> 7;
    sv[kk][rr] = (r0 + rr < rows && kk < kb) ? v[r0 + rr + (long long)kk * ldv] : 0.0;
  }
  for (int e = tid; e < DB * DB; e += 256) {
    const int kk = e & 31, cc = e >> 5;
    sx[kk][cc] = (kk < kb && c0 + cc < nc) ? x[kk + (long long)(c0 + cc) * ldx] : 0.0;
  }
  __syncthreads();
  const int ri = (tid & 31) * 4, ci = (tid >> 5) * 4;
  double o[4][4] = {};
#pragma unroll 4
  for (int kk = 0; kk < DB; ++kk) {
    double a4[4], b4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a4[i] = sv[kk][ri + i]; b4[i] = sx[kk][ci + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) o[i][q] = fma(a4[i], b4[q], o[i][q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + ri + i < rows && c0 + ci + q < nc) c[r0 + ri + i + (long long)(c0 + ci + q) * ldc] -= o[i][q];
}

// V (unit diagonal, zeros above) of the kb-column panel stored LAPACK-style in a; T from V^T V (in g, kb x kb) and tau.
__global__ void dextract_v_kernel(const double* __restrict__ a, long long lda, long long mp, int kb, double* __restrict__ v, long long ldv) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= mp) return;
  for (int c = 0; c < kb; ++c) v[r + (long long)c * ldv] = r < c ? 0.0 : (r == c ? 1.0 : a[r + (long long)c * lda]);
}
__global__ void dbuild_t_kernel(double* __restrict__ g, int ldg, const double* __restrict__ tau, int kb, double* __restrict__ t, int ldt) {
  const int k = threadIdx.x;
  if (k >= kb) return;
  double tc[DB];
  for (int i = 0; i < DB; ++i) tc[i] = 0.0;
  for (int i = k; i >= 0; --i) {
    if (i == k) { tc[i] = tau[k]; continue; }
    double s = 0.0;
    for (int q = i + 1; q <= k; ++q) s = fma(g[i + (long long)q * ldg], tc[q], s);
    tc[i] = -tau[i] * s;
  }
  for (int i = 0; i < kb; ++i) t[i + (long long)k * ldt] = (i <= k) ? tc[i] : 0.0;
}
__global__ void dzero_kernel(double* __restrict__ p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}
__global__ void dset_identity_kernel(double* __restrict__ a, long long lda, long long m, int n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m * (long long)n) a[i % m + (i / m) * lda] = (i % m == i / m) ? 1.0 : 0.0;
}
__global__ void dextract_r_kernel(const double* __restrict__ a, long long lda, int n, double* __restrict__ r, long long ldr, int r_rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)r_rows * n) { const long long rr = i % r_rows, c = i / r_rows; r[rr + c * ldr] = (rr <= c) ? a[rr + c * lda] : 0.0; }
}

// C <- (I - V op(T) V^T) C with scratch w, x (kb x nc, ld DB; w zero on entry and exit)
void dapply_block(const double* v, long long ldv, const double* t, int trans_t, double* c, long long ldc, long long rows, int kb, int nc, double* w,
                  double* x, cudaStream_t s) {
  if (nc <= 0 || kb <= 0) return;
  const int rpc = 2048;
  dim3 g1((nc + DB - 1) / DB, (unsigned)((rows + rpc - 1) / rpc));
  g_launches += 3;
  dgemm_tn_kernel<<<g1, 256, 0, s>>>(v, ldv, c, ldc, rows, kb, nc, w, DB, rpc);
  dtmul_kernel<<<nc, DB, 0, s>>>(t, DB, trans_t, w, DB, kb, nc, x, DB);
  dim3 g3((unsigned)((rows + 127) / 128), (nc + DB - 1) / DB);
  dgemm_nn_kernel<<<g3, 256, 0, s>>>(v, ldv, x, DB, c, ldc, rows, kb, nc);
}

int dpanel_max_ctas(int sm_count) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dpanel_kernel, 256, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
  return nb * sm_count;
}

}  // namespace

size_t f64_workspace_bytes(long long m, int nc_max) {
  return (size_t)(m * DB + DB * DB * 2 + 3 * DB + 2LL * DB * (nc_max > m ? nc_max : m) + 64) * sizeof(double) + 4096;
}

// Blocked Householder QR in fp64 (LAPACK storage).  ws: f64_workspace_bytes(m, n) bytes.
int f64_geqrf(double* a, long long lda, long long m, int n, double* tau, void* ws, int sm_count, cudaStream_t s) {
  double* vb = (double*)ws;
  double* t = vb + m * DB;
  double* g = t + DB * DB;
  double* acc = g + DB * DB;
  double* w = acc + 3 * DB + 32;
  double* x = w + (long long)DB * n;
  const long long zn = (long long)(DB * DB * 2 + 3 * DB + 32) + 2LL * DB * n;
  dzero_kernel<<<(unsigned)((zn + 255) / 256), 256, 0, s>>>(t, zn);
  const int max_ctas = dpanel_max_ctas(sm_count);
  for (int j0 = 0; j0 < n; j0 += DB) {
    const int b = n - j0 < DB ? n - j0 : DB;
    const long long mp = m - j0;
    DPanelParams p{a + j0 + (long long)j0 * lda, lda, mp, b, tau + j0, vb, m, t, DB, acc, g};
    long long ctas = (mp + 255) / 256;             // a thread per row until the device is full
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    void* args[] = {&p};
    ++g_launches;
    cudaError_t e = cudaLaunchCooperativeKernel((void*)dpanel_kernel, dim3((unsigned)ctas), dim3(256), args, 0, s);
    if (e != cudaSuccess) return (int)e;
    dapply_block(vb, m, t, 1, a + j0 + (long long)(j0 + b) * lda, lda, mp, b, n - j0 - b, w, x, s);
  }
  return (int)cudaGetLastError();
}

// C <- Q^T C (trans) or Q C for the factorisation in (a, tau); (V, T) per panel are rebuilt from the LAPACK storage.
int f64_apply_q(int trans, const double* a, long long lda, long long m, int n, const double* tau, double* c, long long ldc, int nc, bool c_is_identity_start,
                void* ws, cudaStream_t s) {
  double* vb = (double*)ws;
  double* t = vb + m * DB;
  double* g = t + DB * DB;
  double* acc = g + DB * DB;
  double* w = acc + 3 * DB + 32;
  double* x = w + (long long)DB * (nc > n ? nc : n);
  const long long zn = (long long)(DB * DB * 2 + 3 * DB + 32) + 2LL * DB * (nc > n ? nc : n);
  dzero_kernel<<<(unsigned)((zn + 255) / 256), 256, 0, s>>>(t, zn);
  const int npan = (n + DB - 1) / DB;
  for (int pi = 0; pi < npan; ++pi) {
    const int j0 = (trans ? pi : npan - 1 - pi) * DB;
    const int b = n - j0 < DB ? n - j0 : DB;
    const long long mp = m - j0;
    g_launches += 3;
    dextract_v_kernel<<<(unsigned)((mp + 255) / 256), 256, 0, s>>>(a + j0 + (long long)j0 * lda, lda, mp, b, vb, m);
    dim3 g1(1, (unsigned)((mp + 2047) / 2048));
    dgemm_tn_kernel<<<g1, 256, 0, s>>>(vb, m, vb, m, mp, b, b, g, DB, 2048);
    dbuild_t_kernel<<<1, DB, 0, s>>>(g, DB, tau + j0, b, t, DB);
    dzero_kernel<<<(DB * DB + 255) / 256, 256, 0, s>>>(g, DB * DB);
    const int cskip = (c_is_identity_start && !trans) ? (j0 < nc ? j0 : nc) : 0;   // backward accumulation from [I; 0]: columns < j0 are still zero below row j0
    dapply_block(vb, m, t, trans ? 1 : 0, c + j0 + (long long)cskip * ldc, ldc, mp, b, nc - cskip, w, x, s);
  }
  return (int)cudaGetLastError();
}

void f64_set_identity(double* a, long long lda, long long m, int n, cudaStream_t s) {
  ++g_launches;
  dset_identity_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, s>>>(a, lda, m, n);
}
void f64_extract_r(const double* a, long long lda, int n, double* r, long long ldr, int r_rows, cudaStream_t s) {
  ++g_launches;
  dextract_r_kernel<<<(unsigned)(((long long)r_rows * n + 255) / 256), 256, 0, s>>>(a, lda, n, r, ldr, r_rows);
}

}  // namespace cqr
