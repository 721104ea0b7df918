// tsqr_mma.cu -- R-only tall-skinny QR (BASELINE config 3) with the block products on the tensor pipe.
//
// Same recurrence as tsqr_flat.cu -- the reference's flat tree (qr.c:68-73, 109-141): a running 64 x 64 R absorbs one
// 64-row block B at a time through structured reflectors v_j = [e_j ; x_j / u_j] -- but blocked by sub-panels of 8 columns:
//   * the sub-panel [R(J,J) ; B(:,J)] is factored column by column on the FMA pipe (scalar formulas of qr.c:144-152):
//     every 8-column group of the block lives in 8 lane groups x 4 lanes x 16 rows, so a column's dot is 8 FFMA2 and two
//     shuffles; the dots of the finished columns give striu(X^T X), from which T follows by back substitution
//     (T^-1 = diag(1/tau) + striu(X^T X));
//   * the columns C to the right get  W = R(J,C) + X^T B(:,C),  Y = T^T W,  R(J,C) -= Y,  B(:,C) -= X Y  as
//     mma.sync.m16n8k8 TF32 products with fp32 accumulation, every operand split into a TF32 head and its exact fp32
//     remainder and three products issued (lo*hi + hi*lo + hi*hi: fp32-faithful, tools/mma_leaf_spec.py is the numpy spec).
// The block is held TRANSPOSED in accumulator-fragment layout (M = the block's columns, N = its rows): then the
// accumulator registers of the block ARE the A fragments of X^T B (with a fixed permutation of the k index that the B
// fragments -- the sub-panel's own registers -- share), W's accumulators are the A fragments of T^T W and of the update,
// and the only data that moves is X~^T (64 x 8) through shared memory once per sub-panel.
//
// One chain of 64-row blocks per warp, 8 warps per CTA, R in shared memory (row-major, padded); when its chain is done
// a CTA combines its warps' R factors in a binary tree (a combine is the same block step with the peer's R as the block),
// so a launch turns m rows into 64 rows per CTA and the levels above are the same kernel on the stacked R's.
// Measured on B200 (tools/probes/mma_tf32_probe.cu): HMMA.1688.F32.TF32 issues every 8 clk per SM sub-partition with
// 20 clk latency and overlaps FFMA issue, i.e. 4/3 of the fp32 FMA rate for a 3xTF32 product -- the gain is in issue
// slots (816 HMMA per block against ~8 K FFMA2 issue cycles), not in raw arithmetic rate.
#include "common.cuh"
#include "warp_math.cuh"
#include <stdlib.h>

namespace cqr {

namespace {

constexpr int kRld = 68;                       // row stride of R in shared memory: rows 2t and 2t+1 of a lane quad fall on distinct banks
constexpr int kMmaWarpFloats = 64 * kRld + 2 * 64 + 64 * 8 + 64 + 16;   // R, double-buffered x, X~ (row-major 64 x 8), striu(X^T X), tau

__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// x = hi + lo with hi = x rounded to TF32 (11 significant bits, round to nearest on the bit pattern: ties away from zero)
// and lo = x - hi exactly.  Rounding, not truncating: a truncated head leaves a remainder of x's sign, the dropped
// lo * lo term and the pipe's own truncation of lo then bias every product the same way, and over a chain of blocks
// the bias adds up linearly (8M x 64 uniform data: Gram error 2e-4 instead of 5e-7).  cvt.rna.tf32.f32 does the same in
// five instructions (NaN handling); the integer add + mask is two, and since the head is a computed value it lands in
// its fragment slot directly -- the block's accumulator registers need the order (c0, c2, c1, c3) as an A fragment.
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ unsigned bits(float x) { return __float_as_uint(x); }

struct Lane { int g, t; };   // lane = 4 g + t: g = column within an 8-column group (fragment row), t = row-pair selector

// Block registers: blk[mt][nt][e] = B(row 8 nt + 2 t + (e & 1), column 16 mt + 8 (e >> 1) + g) -- the accumulator fragment of
// the 16 x 8 tile (columns 16 mt .. +15) x (rows 8 nt .. +7) of B^T.

// Sub-panel P (columns 8 P .. 8 P + 7): Householder column by column on the FMA pipe.  Lane group g owns column 8 P + g.
// On exit xc holds x~ = x / u of the lane's column, gs the strictly upper triangle of X~^T X, taus the 8 tau values.
__device__ __forceinline__ void subpanel(const int P, f32x2 (&xc)[8], float* __restrict__ Rs, float* __restrict__ xs, float* __restrict__ gs,
                                         float* __restrict__ taus, const Lane L) {
  const int g = L.g, t = L.t;
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * P + jj;
    float* xb = xs + (jj & 1) * 64;
    float* Rj = Rs + j * kRld;
    const float rho = Rj[8 * P + g];           // R(j, my column); last written one block (or one step) ago, read before the barrier
    const float alpha = Rj[j];
    if (g == jj) {                             // the four lanes of column j publish x in register order (16 floats per lane)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ulonglong2 v; v.x = xc[2 * k]; v.y = xc[2 * k + 1];
        *reinterpret_cast<ulonglong2*>(xb + 16 * t + 4 * k) = v;
      }
    }
    __syncwarp();
    f32x2 x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xb + 16 * t + 4 * k);
      x[2 * k] = v.x; x[2 * k + 1] = v.y;
    }
    f32x2 da = 0ull, db = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) { da = ffma2(x[2 * k], xc[2 * k], da); db = ffma2(x[2 * k + 1], xc[2 * k + 1], db); }
    float d = fsum2(da) + fsum2(db);
    d += __shfl_xor_sync(kFull, d, 1);
    d += __shfl_xor_sync(kFull, d, 2);         // x^T (my column) over the 64 rows; group jj holds sigma = x^T x
    const float sig = __shfl_sync(kFull, d, 4 * jj);
    // reflector scalars (qr.c:144-152) on the MUFU approximations plus one Newton step, as in tsqr_flat.cu; a zero or
    // underflowing column gives tau = 0, H = I
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
    if (g > jj) {                              // columns right of the pivot inside the sub-panel
      const float s = fmaf(d, inv_u, rho);
      const float f = -tau * inv_u * s;
      if (t == 0) Rj[8 * P + g] = fmaf(-tau, s, rho);
      const f32x2 f2 = fpack2(f, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) xc[k] = ffma2(f2, x[k], xc[k]);
    } else if (g == jj) {                      // the pivot column becomes x~ = x / u
      const f32x2 iu2 = fpack2(inv_u, inv_u);
#pragma unroll
      for (int k = 0; k < 8; ++k) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(xc[k]) : "l"(x[k]), "l"(iu2));
      if (t == 0) { Rj[j] = ok ? bc : alpha; taus[jj] = tau; }
    } else {                                   // finished columns: (X~^T X~)(g, jj)
      if (t == 0) gs[g * 8 + jj] = d * inv_u;
    }
  }
  __syncwarp();
}

// Column g of the sub-panel's compact-WY T by back substitution on T^-1 = diag(1/tau) + striu(X~^T X~); returned as
// the B fragment of Y = T^T W: b0 = T(2 t, g), b1 = T(2 t + 1, g).
// T's diagonal (the taus) is returned apart, as tau(2 t), tau(2 t + 1) for the lane's two W columns, and the fragment holds
// the strictly upper part only: Y = tau . W + striu(T)^T W.  W is R-sized while a chain runs (v_j ~ e_j, tau ~ 2, Y ~ 2 R),
// and the tensor pipe accumulates with round-toward-zero: a Y shaved by an ulp per block shrinks |R| by as much, and over
// a chain of 111 blocks that compounds to 1e-5.  The dominant term tau . W is therefore a rounded fp32 multiply on the
// FMA pipe; striu(T) is O(|x~|^2), small exactly when R is large.
__device__ __forceinline__ void t_fragment(const float* __restrict__ gs, const float* __restrict__ taus, const Lane L, float& b0, float& b1,
                                           float& tau0, float& tau1) {
  float G[8][8], tv[8], tc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 lo = *reinterpret_cast<const float4*>(gs + 8 * i), hi = *reinterpret_cast<const float4*>(gs + 8 * i + 4);
    G[i][0] = lo.x; G[i][1] = lo.y; G[i][2] = lo.z; G[i][3] = lo.w; G[i][4] = hi.x; G[i][5] = hi.y; G[i][6] = hi.z; G[i][7] = hi.w;
  }
  {
    const float4 lo = *reinterpret_cast<const float4*>(taus), hi = *reinterpret_cast<const float4*>(taus + 4);
    tv[0] = lo.x; tv[1] = lo.y; tv[2] = lo.z; tv[3] = lo.w; tv[4] = hi.x; tv[5] = hi.y; tv[6] = hi.z; tv[7] = hi.w;
  }
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int l = i + 1; l < 8; ++l) {
      if ((l - i) & 1) a0 = fmaf(G[i][l], tc[l], a0); else a1 = fmaf(G[i][l], tc[l], a1);
    }
    tc[i] = (i == L.g) ? tv[i] : ((i < L.g) ? -tv[i] * (a0 + a1) : 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) if (i == L.g) tc[i] = 0.f;
  b0 = (L.t == 0) ? tc[0] : (L.t == 1) ? tc[2] : (L.t == 2) ? tc[4] : tc[6];
  b1 = (L.t == 0) ? tc[1] : (L.t == 1) ? tc[3] : (L.t == 2) ? tc[5] : tc[7];
  tau0 = (L.t == 0) ? tv[0] : (L.t == 1) ? tv[2] : (L.t == 2) ? tv[4] : tv[6];
  tau1 = (L.t == 0) ? tv[1] : (L.t == 1) ? tv[3] : (L.t == 2) ? tv[5] : tv[7];
}

// Columns right of sub-panel P: W = R(J, C) + X~^T B(:, C), Y = T^T W, R(J, C) -= Y, B(:, C) -= X~ Y.
// P is a run-time value (one copy of this code instead of eight: the fully specialised version was 120 KB of
// instructions, and eight warps streaming through it at different places were instruction-fetch bound); the 16-column
// tiles keep static register indices and are skipped by warp-uniform guards.
__device__ __forceinline__ void trailing(const int P, float (&blk)[4][8][4], const f32x2 (&xc)[8], float* __restrict__ Rs, float* __restrict__ xt,
                                         const float* __restrict__ gs, const float* __restrict__ taus, const Lane L, const int n) {
  const int g = L.g, t = L.t;
  const bool even = (P & 1) == 0;              // P even: the sub-panel is the low half of tile P / 2, whose high half is trailing
  const int mtf = (P + 1) >> 1;                // first tile that takes part (for even P this is the shared tile P / 2)
  // X~^T for the update's B fragments: row-major 64 x 8 in shared memory
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float lo, hi;
    funpack2(xc[nt], lo, hi);
    xt[(8 * nt + 2 * t) * 8 + g] = lo;
    xt[(8 * nt + 2 * t + 1) * 8 + g] = hi;
  }
  // B fragments of the dots: the sub-panel's own registers, b0 = X~(8 kt + 2 t, g), b1 = X~(8 kt + 2 t + 1, g)
  unsigned xh[8][2], xl[8][2];
#pragma unroll
  for (int kt = 0; kt < 8; ++kt) {
    float lo, hi;
    funpack2(xc[kt], lo, hi);
    split_tf32(lo, xh[kt][0], xl[kt][0]);
    split_tf32(hi, xh[kt][1], xl[kt][1]);
  }
  float tb0, tb1, tau0, tau1;
  t_fragment(gs, taus, L, tb0, tb1, tau0, tau1);
  unsigned th0, th1, tl0, tl1;
  split_tf32(tb0, th0, tl0);
  split_tf32(tb1, th1, tl1);
  float* R0 = Rs + (8 * P + 2 * t) * kRld + g;  // rows 8 P + 2 t (and + 1) of R
  float* R1 = R0 + kRld;
  float yn[4][4];                              // A fragments of the update: -Y in fragment order
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    if (mt >= mtf && 16 * mt < n) {            // warp-uniform
      const bool half = even && 2 * mt == P;
      float r[4];
      r[0] = R0[16 * mt]; r[1] = R1[16 * mt]; r[2] = R0[16 * mt + 8]; r[3] = R1[16 * mt + 8];
      // three independent accumulator chains (lo*hi, hi*lo, hi*hi): an HMMA has 20 clk latency and issues every 8
      // (all three start from zero: the pipe accumulates with round-toward-zero, and with the large R(J, C) in the
      // accumulator every HMMA would shave up to an ulp of R off W -- 8 per block, always toward zero, which over a
      // chain of 111 blocks shrank R by 1e-4; R joins in one rounded fp32 add below)
      float w0[4] = {0.f, 0.f, 0.f, 0.f}, w1[4] = {0.f, 0.f, 0.f, 0.f}, w2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        unsigned h0, h1, h2, h3, l0, l1, l2, l3;
        split_tf32(blk[mt][kt][0], h0, l0);
        split_tf32(blk[mt][kt][2], h1, l1);
        split_tf32(blk[mt][kt][1], h2, l2);
        split_tf32(blk[mt][kt][3], h3, l3);
        mma_tf32(w0, l0, l1, l2, l3, xh[kt][0], xh[kt][1]);
        mma_tf32(w1, h0, h1, h2, h3, xl[kt][0], xl[kt][1]);
        mma_tf32(w2, h0, h1, h2, h3, xh[kt][0], xh[kt][1]);
      }
      float wv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) wv[e] = r[e] + ((w0[e] + w1[e]) + w2[e]);
      // Y = T^T W: A = W^T (same k permutation), B = T
      float y[4] = {0.f, 0.f, 0.f, 0.f};
      {
        unsigned h0, h1, h2, h3, l0, l1, l2, l3;
        split_tf32(wv[0], h0, l0);
        split_tf32(wv[2], h1, l1);
        split_tf32(wv[1], h2, l2);
        split_tf32(wv[3], h3, l3);
        mma_tf32(y, l0, l1, l2, l3, th0, th1);
        mma_tf32(y, h0, h1, h2, h3, tl0, tl1);
        mma_tf32(y, h0, h1, h2, h3, th0, th1);
      }
      y[0] = fmaf(tau0, wv[0], y[0]); y[1] = fmaf(tau1, wv[1], y[1]); y[2] = fmaf(tau0, wv[2], y[2]); y[3] = fmaf(tau1, wv[3], y[3]);
      if (!half) { R0[16 * mt] = r[0] - y[0]; R1[16 * mt] = r[1] - y[1]; }
      R0[16 * mt + 8] = r[2] - y[2]; R1[16 * mt + 8] = r[3] - y[3];
      // the sub-panel's own columns (low half of a shared tile) stay as they are: zero rows of -Y
      yn[mt][0] = half ? 0.f : -y[0]; yn[mt][1] = -y[2]; yn[mt][2] = half ? 0.f : -y[1]; yn[mt][3] = -y[3];
    }
  }
  __syncwarp();                                // X~^T is in shared memory
  unsigned bh[8][2], bl[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 v = *reinterpret_cast<const float2*>(xt + (8 * nt + g) * 8 + 2 * t);   // X~(8 nt + g, 2 t), X~(8 nt + g, 2 t + 1)
    split_tf32(v.x, bh[nt][0], bl[nt][0]);
    split_tf32(v.y, bh[nt][1], bl[nt][1]);
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    if (mt >= mtf && 16 * mt < n) {
      unsigned h0, h1, h2, h3, l0, l1, l2, l3;
      split_tf32(yn[mt][0], h0, l0);
      split_tf32(yn[mt][1], h1, l1);
      split_tf32(yn[mt][2], h2, l2);
      split_tf32(yn[mt][3], h3, l3);
      // -X~ Y into zero accumulators, then one rounded add per element (same reason as above)
      float u[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { u[nt][0] = 0.f; u[nt][1] = 0.f; u[nt][2] = 0.f; u[nt][3] = 0.f; }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32(u[nt], l0, l1, l2, l3, bh[nt][0], bh[nt][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32(u[nt], h0, h1, h2, h3, bl[nt][0], bl[nt][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32(u[nt], h0, h1, h2, h3, bh[nt][0], bh[nt][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) blk[mt][nt][e] += u[nt][e];
    }
  }
  __syncwarp();                                // every lane has read X~^T before the next sub-panel rewrites it
}

#ifdef CQR_MMA_TRACE
__device__ long long g_mma_trace[4];           // clock cycles of warp 0 of CTA 0: [0] loads, [1] sub-panels, [2] trailing, [3] blocks
#define CQR_TR(i, t0) do { if (tr) g_mma_trace[i] += clock64() - (t0); } while (0)
#else
#define CQR_TR(i, t0) do { } while (0)
#endif

// One 64-row block into the running R: eight sub-panels.
__device__ __forceinline__ void block_step(float (&blk)[4][8][4], float* Rs, float* xs, float* xt, float* gs, float* taus, const Lane L, const int n,
                                           const bool tr) {
#pragma unroll 1
  for (int P = 0; P < 8; ++P) {
    if (8 * P >= n) break;                     // warp-uniform
#ifdef CQR_MMA_TRACE
    long long t0 = clock64();
#endif
    f32x2 xc[8];                               // the lane's sub-panel column: tile P / 2, half P % 2 of the block registers
    switch (P) {
#define CQR_XC_CASE(PP)                                                                                              \
      case PP:                                                                                                       \
        _Pragma("unroll") for (int nt = 0; nt < 8; ++nt) xc[nt] = fpack2(blk[PP / 2][nt][2 * (PP % 2)], blk[PP / 2][nt][2 * (PP % 2) + 1]); \
        break;
      CQR_XC_CASE(0) CQR_XC_CASE(1) CQR_XC_CASE(2) CQR_XC_CASE(3) CQR_XC_CASE(4) CQR_XC_CASE(5) CQR_XC_CASE(6)
      default:
        _Pragma("unroll") for (int nt = 0; nt < 8; ++nt) xc[nt] = fpack2(blk[3][nt][2], blk[3][nt][3]);
        break;
#undef CQR_XC_CASE
    }
    subpanel(P, xc, Rs, xs, gs, taus, L);
    CQR_TR(1, t0);
#ifdef CQR_MMA_TRACE
    t0 = clock64();
#endif
    if (P < 7) trailing(P, blk, xc, Rs, xt, gs, taus, L, n);
    CQR_TR(2, t0);
  }
}

}  // namespace

template <int WPC>
__global__ void __launch_bounds__(32 * WPC, 1) tsqr_mma_r_kernel(MmaTsqrParams p) {
  extern __shared__ __align__(16) float mma_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Lane L{lane >> 2, lane & 3};
  const int g = L.g, t = L.t;
  float* Rs = mma_smem + w * kMmaWarpFloats;
  float* xs = Rs + 64 * kRld;
  float* xt = xs + 2 * 64;
  float* gs = xt + 64 * 8;
  float* taus = gs + 64;
  for (int i = lane; i < 64 * kRld / 4; i += 32) reinterpret_cast<float4*>(Rs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < 16) taus[lane] = 0.f;
  __syncwarp();
  const long long chain = (long long)blockIdx.x * WPC + w;
  const bool have = chain < p.chains;
  const int n = p.n;
  const long long row0 = chain * p.rows_per_chain;
  const long long row1 = (row0 + p.rows_per_chain < p.m) ? row0 + p.rows_per_chain : p.m;
  const bool al8 = (p.lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 7) == 0);
  const bool al16 = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0);
  const int nb = (int)(p.rows_per_chain >> 6);
  int levels = 0;
  while ((1 << levels) < WPC) ++levels;
  // every warp runs the same number of iterations (the combine levels hold CTA barriers): nb chain blocks, then the
  // binary combine tree over the CTA's warps
  for (int it = 0; it < nb + levels; ++it) {
    float blk[4][8][4];
    bool act;
    if (it < nb) {
      const long long rb = row0 + 64ll * it;
      act = have && rb < row1;
      if (act) {
        const float* src = p.a + rb + 2 * t;
        if (al8 && rb + 64 <= row1) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = 16 * mt + 8 * h + g;
              if (c < n) {
                const float* col = src + (long long)c * p.lda;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                  const float2 v = __ldcs(reinterpret_cast<const float2*>(col + 8 * nt));
                  blk[mt][nt][2 * h] = v.x; blk[mt][nt][2 * h + 1] = v.y;
                }
              } else {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) { blk[mt][nt][2 * h] = 0.f; blk[mt][nt][2 * h + 1] = 0.f; }
              }
            }
          if (al16 && rb + 128 <= row1) {        // next block into L2 through the bulk-copy (TMA) unit: 64 columns x 256 B
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int c = lane + 32 * i;
              if (c < n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(p.a + rb + 64 + (long long)c * p.lda) : "memory");
            }
          }
        } else {                               // ragged last block or unaligned source: guarded scalar loads, zero fill
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = 16 * mt + 8 * h + g;
#pragma unroll
              for (int nt = 0; nt < 8; ++nt) {
                const long long r = rb + 8 * nt + 2 * t;
                blk[mt][nt][2 * h] = (c < n && r < row1) ? src[(long long)c * p.lda + 8 * nt] : 0.f;
                blk[mt][nt][2 * h + 1] = (c < n && r + 1 < row1) ? src[(long long)c * p.lda + 8 * nt + 1] : 0.f;
              }
            }
        }
      }
    } else {
      const int s = 1 << (it - nb);
      __syncthreads();                         // the peer's R is final (and last level's readers are done)
      act = ((w & (2 * s - 1)) == 0) && (w + s < WPC) && (chain + s < p.chains);
      if (act) {
        const float* Ro = mma_smem + (w + s) * kMmaWarpFloats;   // upper triangular, zeros below the diagonal
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) blk[mt][nt][e] = Ro[(8 * nt + 2 * t + (e & 1)) * kRld + 16 * mt + 8 * (e >> 1) + g];
      }
    }
    if (act) {
      __syncwarp();
      block_step(blk, Rs, xs, xt, gs, taus, L, n, blockIdx.x == 0 && w == 0 && lane == 0);
#ifdef CQR_MMA_TRACE
      if (blockIdx.x == 0 && w == 0 && lane == 0) g_mma_trace[3] += 1;
#endif
    }
  }
  __syncwarp();
  if (w == 0) {                                // the CTA's R: a full 64-row slot (zeros below the diagonal and right of n), or the final n x n
    const int rows = p.out_rows;
    for (int c = 0; c < (rows == 64 ? 64 : n); ++c) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = lane + 32 * rr;
        if (r < rows) p.r_out[(long long)blockIdx.x * 64 + r + (long long)c * p.r_ld] = (r <= c && c < n) ? Rs[r * kRld + c] : 0.f;
      }
    }
  }
}

constexpr int kMmaWpc = 8;

#ifdef CQR_MMA_TRACE
void mma_tsqr_read_trace(long long* out) {
  cudaMemcpyFromSymbol(out, g_mma_trace, sizeof(long long) * 4);
  long long z[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(g_mma_trace, z, sizeof(z));
}
#endif

int mma_tsqr_warps_per_cta() { return kMmaWpc; }

void launch_tsqr_mma_r(const MmaTsqrParams& p, cudaStream_t s) {
  if (p.chains <= 0) return;
  ++g_launches;
  constexpr size_t smem = (size_t)kMmaWpc * kMmaWarpFloats * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(tsqr_mma_r_kernel<kMmaWpc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tsqr_mma_r_kernel<kMmaWpc><<<(p.chains + kMmaWpc - 1) / kMmaWpc, 32 * kMmaWpc, smem, s>>>(p);
}

}  // namespace cqr
