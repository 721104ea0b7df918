// legacy_format.cu -- exporter for the reference's own storage format.
//
// The library's factorisation stores LAPACK-style reflectors and one tau per column; the reference's explicitQR
// (qr.c:330-438) instead expects what ITS sweep leaves behind: per PC-wide column block, PR x PC windows walked bottom to
// top in steps of PR - PC rows (qr.c:68-73), every window column a reflector over the row range of qr.c:109-141
// (fresh rows plus the diagonal entry of the R carried up from the window below), the non-unit parts stored in place,
// and tau[(rowPanels * pcCount + prCount) * PC + col] with windows counted from the bottom (qr.c:300-304).
// This file runs exactly that sweep on the GPU with the reference's GPU parameters PR = 64, PC = 4 (qr.cu:21-23) so that
// the unmodified reference can consume the output: one launch per column block; a thread owns one trailing column and
// carries its 64-row window in registers (the 4 rows of overlap never leave them), every CTA re-derives the block's
// windows (a 64 x 4 panel: cheap) and CTA 0 writes panel and tau.  Scalar formulas as qr.c:144-167 (sigma, sign, u,
// tau = sign u / norm, diag = -sign norm, v = x / u); a zero column gives tau = 0 where the reference divides 0 / 0.
// It is a format exporter for parity checks, not the fast path: 2 launches per window become n / 4 launches in total.
#include "common.cuh"

namespace cqr {

namespace {

constexpr int LPR = 64, LPC = 4, LTC = 128;   // window height, block width, trailing columns (threads) per CTA

__device__ __forceinline__ float block_sum64(float v, float* red, int tid) {   // sum over threads 0..63, result to all
  if (tid < 64) {
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
  }
  __syncthreads();
  const float s = red[0] + red[1];
  __syncthreads();
  return s;
}

// `pout` (m x 4, ld m) receives the factored panel: the other CTAs of the launch still read the block's original columns
// from `a` while CTA 0 is ahead of them, so the panel goes back into `a` only after the launch (launch_legacy_sweep).
__global__ void __launch_bounds__(LTC) legacy_sweep_kernel(float* __restrict__ a, long long lda, int m, int n, int pc, float* __restrict__ tau,
                                                           int rowPanels, int pcCount, float* __restrict__ pout) {
  __shared__ float panel[LPC][LPR];   // the window's own columns
  __shared__ float vv[LPC][LPR];      // reflectors, zero outside [vstart, vend), 1 at vstart
  __shared__ float taus[LPC];
  __shared__ float red[2];
  const int tid = threadIdx.x;
  const int ct = pc + LPC + blockIdx.x * LTC + tid;     // this thread's trailing column
  const bool has = ct < n;
  float t[LPR];
#pragma unroll
  for (int i = 0; i < LPR; ++i) t[i] = 0.f;
  int prCount = 0, pr_last = -1;
  for (int pr = m - LPR; pr + LPR > pc && pr >= 0; pr -= LPR - LPC, ++prCount) {
    const bool bottom = pr == m - LPR, top = pr <= pc;
    // ---- window in: the bottom window reads 64 rows, the others 60 fresh rows under the 4 carried ones
    if (!bottom) {
      if (tid < LPC * LPC) panel[tid / LPC][LPR - LPC + tid % LPC] = panel[tid / LPC][tid % LPC];
#pragma unroll
      for (int i = 0; i < LPC; ++i) t[LPR - LPC + i] = t[i];
    }
    __syncthreads();
    const int fresh = bottom ? LPR : LPR - LPC;
    for (int e = tid; e < LPC * fresh; e += LTC) panel[e / fresh][e % fresh] = a[pr + e % fresh + (long long)(pc + e / fresh) * lda];
    if (has) {
      const float* col = a + pr + (long long)ct * lda;
#pragma unroll
      for (int i = 0; i < LPR - LPC; ++i) t[i] = col[i];
      if (bottom) {
#pragma unroll
        for (int i = LPR - LPC; i < LPR; ++i) t[i] = col[i];
      }
    }
    __syncthreads();
    // ---- the window's four reflectors (qr.c:109-167) and their effect on the window's later columns (qr.c:215-235)
    for (int col = 0; col < LPC; ++col) {
      const int vstart = top ? pc - pr + col : col;
      const int vend = bottom ? LPR : LPR - LPC + col + 1;
      const bool in = tid < LPR && tid >= vstart && tid < vend;
      const float x = in ? panel[col][tid] : 0.f;
      const float sigma = block_sum64(x * x, red, tid);
      const float x0 = panel[col][vstart];
      const float norm = sqrtf(sigma);
      const float sign = x0 < 0.f ? -1.f : 1.f;
      const float u = x0 + sign * norm;
      const float tcol = norm > 0.f ? sign * u / norm : 0.f;
      __syncthreads();
      if (tid < LPR) {
        float v = 0.f;
        if (in) v = (tid == vstart) ? 1.f : (norm > 0.f ? x / u : 0.f);
        vv[col][tid] = v;
        if (in) panel[col][tid] = (tid == vstart) ? -sign * norm : v;
      }
      if (tid == 0) taus[col] = tcol;
      __syncthreads();
      for (int c2 = col + 1; c2 < LPC; ++c2) {
        const float s = block_sum64(tid < LPR ? vv[col][tid] * panel[c2][tid] : 0.f, red, tid);
        if (tid < LPR) panel[c2][tid] -= tcol * s * vv[col][tid];
        __syncthreads();
      }
    }
    // ---- trailing column of this thread: H_0 .. H_3 in order (the reference's I + Y W^T, qr.c:255-293)
    if (has) {
#pragma unroll
      for (int col = 0; col < LPC; ++col) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LPR; ++i) s = fmaf(vv[col][i], t[i], s);
        const float f = taus[col] * s;
#pragma unroll
        for (int i = 0; i < LPR; ++i) t[i] = fmaf(-f, vv[col][i], t[i]);
      }
      float* col = a + pr + (long long)ct * lda;   // rows below the overlap are final for this window
#pragma unroll
      for (int i = LPC; i < LPR; ++i) col[i] = t[i];
    }
    if (blockIdx.x == 0) {
      for (int e = tid; e < LPC * (LPR - LPC); e += LTC) {
        const int c = e / (LPR - LPC), r = LPC + e % (LPR - LPC);
        pout[pr + r + (long long)c * m] = panel[c][r];
      }
      if (tid < LPC) tau[((long long)rowPanels * pcCount + prCount) * LPC + tid] = taus[tid];
    }
    pr_last = pr;
    __syncthreads();
  }
  if (pr_last >= 0) {                              // the overlap rows of the last (topmost) window
    if (has) {
      float* col = a + pr_last + (long long)ct * lda;
#pragma unroll
      for (int i = 0; i < LPC; ++i) col[i] = t[i];
    }
    if (blockIdx.x == 0 && tid < LPC * LPC) pout[pr_last + tid % LPC + (long long)(tid / LPC) * m] = panel[tid / LPC][tid % LPC];
  }
}

}  // namespace

bool legacy_format_shape_ok(int m, int n) {
  return m >= LPR && n >= LPC && n <= m && (m - LPR) % (LPR - LPC) == 0 && n % LPC == 0;
}

// In-place sweep over the device matrix; tau is the reference-sized grid (rowPanels * colPanels * 4 floats, zeroed here).
// `scratch`: m x 4 floats.
void launch_legacy_sweep(float* a, long long lda, int m, int n, float* tau, int rowPanels, int colPanels, float* scratch, cudaStream_t s) {
  cudaMemsetAsync(tau, 0, (size_t)rowPanels * colPanels * LPC * sizeof(float), s);
  for (int pc = 0, k = 0; pc < n; pc += LPC, ++k) {
    const int ntrail = n - pc - LPC;
    const int ctas = ntrail > 0 ? (ntrail + LTC - 1) / LTC : 1;
    ++g_launches;
    legacy_sweep_kernel<<<ctas, LTC, 0, s>>>(a, lda, m, n, pc, tau, rowPanels, k, scratch);
    int pr_last = m - LPR;                        // topmost window of this block (same loop as the kernel's)
    for (int pr = m - LPR; pr + LPR > pc && pr >= 0; pr -= LPR - LPC) pr_last = pr;
    cudaMemcpy2DAsync(a + pr_last + (long long)pc * lda, (size_t)lda * sizeof(float), scratch + pr_last, (size_t)m * sizeof(float),
                      (size_t)(m - pr_last) * sizeof(float), LPC, cudaMemcpyDeviceToDevice, s);
  }
}

}  // namespace cqr
