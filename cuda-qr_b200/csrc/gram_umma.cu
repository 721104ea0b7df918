// gram_umma.cu -- R-only tall-skinny leaf on the tensor pipe at HBM speed: R = chol(A^T A) with the Gram matrix formed
// to better than fp32 input precision (north_star: ">= 70 % of HBM bandwidth on 8M x 64 TSQR").
//
// Why this and not Householder: a Householder leaf has 64 dependent column steps per 64-row block and its 2 m n^2 flops
// run on the fp32 FMA pipe -- three kernels (tsqr_flat.cu single / pair step, tsqr_mma.cu) all sit at ~3.5 ms for
// 8M x 64 against a 0.33 ms read time.  The Gram matrix needs no dependent steps and is one dense contraction
// (m n^2 MACs) -- tensor-core work -- but plain fp32 / 3xTF32 accumulation loses what Cholesky then amplifies by
// cond(A)^2.  So the contraction is made ERROR-FREE (Ozaki-style slicing):
//   * rows are taken in groups of 128; per group and column c, 2^E_c > max |a(:,c)| (one pass over the group in
//     registers, two shuffles);
//   * a = s1 + s2 + s3 + rho with s_i = RN(residual to a multiple of 2^(E_c - 8 i)) -- |s1| <= 256, |s2|,|s3| <= 128
//     units, so every slice is exact in bf16 (8-bit significand) and |rho| <= 2^(E_c - 25): elements in the column's top
//     binade are represented exactly, the rest to 2^-25 of the column maximum (half an fp32 ulp of the maximum);
//   * ONE tcgen05.mma (kind::f16, bf16 x bf16 -> f32, M = 128, N = 192, K = 16) per 16 rows forms
//         [S1 | S2]^T [S1 | S2 | S3]  =  D11 D12 D13
//                                        D21 D22 D23            (Dxy = Sx^T Sy, 64 x 64 each)
//     in TMEM.  Every product is an integer < 2^16 times the pair's quantum and a group adds 128 of them: all partial sums
//     are integers <= 2^23 in that unit, i.e. EXACT in the fp32 accumulator whatever the tensor pipe's rounding mode
//     (measured: it truncates, profiles/r02_tsqr_mma_accuracy.txt);
//   * after 128 rows the epilogue warps read the accumulator (tcgen05.ld), form U = D11/2 + D12 + D13 (TMEM lanes 0-63)
//     and V = D22/2 + D23 (lanes 64-127) with two rounded fp32 adds, keep fp32 running sums over 8 groups and add those
//     into a per-CTA fp64 slab; G = sum over CTAs of U + U^T + V + V^T (the dropped D33 is 2^-32 relative);
//   * gram_finish_kernel reduces the slabs in fp64 (fixed order), and its last block factors G = R^T R in fp64, bounds
//     cond of the diagonally scaled Gram matrix by n * ||R^^-1||_F^2 (explicit triangular inverse) and either writes R
//     (fp32) or raises the gate that lets the Householder leaf behind it run (ill-conditioned, non-finite or badly
//     scaled input) -- no host synchronisation either way.
// Data flow of gram_kernel (persistent, one CTA per SM, 448 threads):
//   warp 0     TMA producer: four 32-row x 64-column fp32 boxes (128B swizzle) = one 128-row group per stage, 3 stages
//   warps 2-9  converters: stage -> registers (32 values of one column per thread), column maxima, slices -> bf16 K-major
//              128B-swizzled operand tiles S1|S2|S3 (2 slice stages), fence.proxy.async
//   warp 1     MMA issuer: 8 tcgen05.mma per group into one of two TMEM accumulators (192 columns each)
//   warps 10-13 epilogue (one per TMEM lane quarter)
// Per 128-row group and SM: 32 KB from HBM (1.4 K clk at 6.5 TB/s), 8 MMAs x 96 clk, ~76 KB of shared-memory traffic.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "umma_common.cuh"

namespace cqr {

namespace {

constexpr int kGR = 128;                          // rows per group (= per raw stage = per accumulator flush)
constexpr int kNRaw = 3;                          // raw ring depth
constexpr int kNSl = 2;                           // slice ring depth
constexpr uint32_t kRawTile = 32 * 64 * 4;        // one TMA box: 32 rows x 64 columns fp32, rows of 128 B = one column's 32 k
constexpr uint32_t kRawStage = 4 * kRawTile;      // 32 KB
constexpr uint32_t kSlTile = 64 * 128;            // 64 columns x 64 k bf16 (128 B rows), 128B swizzle
constexpr uint32_t kSlKb = 3 * kSlTile;           // S1 | S2 | S3 of one 64-deep K block
constexpr uint32_t kSlStage = 2 * kSlKb;          // 48 KB
constexpr int kConvT = 256, kEpiT = 128;
constexpr int kGramThreads = 64 + kConvT + kEpiT;
constexpr int kDump = 8;                          // groups per fp32 running sum
constexpr int kAccCols = 256;                     // TMEM column stride between the two accumulators (192 used)
constexpr size_t kGramSmem = (size_t)kNRaw * kRawStage + (size_t)kNSl * kSlStage + 1024 + 256;
constexpr int kSlabDoubles = 128 * 64;            // per-CTA fp64 slab: [j][tmem lane]

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::f16 instruction descriptor: F32 accumulate, BF16 A and B (both K-major), M = 128, N = 192.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {   // exact: both are integers <= 256 times a power of two
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct GramParams {
  long long m;
  long long groups;        // ceil(m / 128)
  double* slabs;           // gridDim.x slabs of kSlabDoubles
  int* status;             // [0] |= 1 when a column's scale is outside the range the fp32 accumulators can carry
};

__global__ void __launch_bounds__(kGramThreads, 1) gram_kernel(const __grid_constant__ CUtensorMap tm_a, GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t raw0 = smem_u32(smem);
  const uint32_t sl0 = raw0 + kNRaw * kRawStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNRaw * kRawStage + kNSl * kSlStage);
  // full[kNRaw] rawfree[kNRaw] conv[kNSl] slfree[kNSl] tfull[2] tempty[2]
  const uint32_t full0 = smem_u32(bars), rawfree0 = full0 + 8 * kNRaw, conv0 = rawfree0 + 8 * kNRaw, slfree0 = conv0 + 8 * kNSl,
                 tfull0 = slfree0 + 8 * kNSl, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNRaw + 2 * kNSl + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // groups of this CTA: contiguous, sizes differ by at most one
  const long long base = p.groups / gridDim.x, rem = p.groups % gridDim.x;
  const long long g_beg = base * blockIdx.x + (blockIdx.x < rem ? blockIdx.x : rem);
  const int ng = (int)(base + (blockIdx.x < rem ? 1 : 0));

  if (threadIdx.x == 0) {
    for (int r = 0; r < kNRaw; ++r) { mbar_init(full0 + 8 * r, 1); mbar_init(rawfree0 + 8 * r, kConvT); }
    for (int t = 0; t < kNSl; ++t) { mbar_init(conv0 + 8 * t, kConvT); mbar_init(slfree0 + 8 * t, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiT); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int g = 0; g < ng; ++g) {
        const uint32_t r = g % kNRaw, ph = (g / kNRaw) & 1;
        mbar_wait(rawfree0 + 8 * r, ph ^ 1);
        const uint32_t dst = raw0 + r * kRawStage, full = full0 + 8 * r;
        mbar_expect_tx(full, kRawStage);
        const long long row0 = (g_beg + g) * kGR;
#pragma unroll
        for (int q = 0; q < 4; ++q) tma_load_2d(dst + q * kRawTile, &tm_a, full, (int)(row0 + 32 * q), 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int g = 0; g < ng; ++g) {
        const uint32_t t = g % kNSl, a = g & 1;
        mbar_wait(conv0 + 8 * t, (g / kNSl) & 1);
        mbar_wait(tempty0 + 8 * a, ((g >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + a * kAccCols;
        const uint32_t sl = sl0 + t * kSlStage;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A = rows 0-127 ([S1 | S2]), B = rows 0-191 ([S1 | S2 | S3]) of the same K-major tile stack: one descriptor
            const uint64_t d = make_desc(sl + kb * kSlKb + k * 32, 16, 1024, 2);
            umma_bf16(tmem_d, d, d, kIdescBf16, (kb | k) != 0);
          }
        }
        umma_commit(slfree0 + 8 * t);
        umma_commit(tfull0 + 8 * a);
      }
    }
  } else if (warp < 2 + kConvT / 32) {
    // converters: thread = (column c, 32-row quarter q of the group)
    const int cw = warp - 2;
    const int c = 8 * cw + (lane & 7), q = lane >> 3;
    const uint32_t swz = (uint32_t)(c & 7);
    const uint32_t rd_off = q * kRawTile + c * 128;
    // destination: K block q / 2, bytes (q % 2) * 64 .. + 63 of row c  ->  16-byte chunks 4 (q % 2) + {0..3}
    const uint32_t wr_off = (q >> 1) * kSlKb + c * 128;
    const uint32_t chunk0 = (uint32_t)(q & 1) * 4;
    for (int g = 0; g < ng; ++g) {
      const uint32_t r = g % kNRaw, t = g % kNSl;
      mbar_wait(full0 + 8 * r, (g / kNRaw) & 1);
      float v[32];
      const uint32_t src = raw0 + r * kRawStage + rd_off;
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                     : "r"(src + ((j ^ swz) << 4)));
      }
      float mx = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) mx = fmaxf(mx, fabsf(v[k]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      // every value has been consumed (the maximum depends on all 32 loads), so the loads have completed: only now may
      // TMA refill the stage.  Arriving right after ISSUING the loads let a fast (L2-fed) refill overtake them.
      mbar_arrive(rawfree0 + 8 * r);
      // 2^E > mx with biased exponent eb + 1; slice quanta 2^(E-8), 2^(E-16), 2^(E-24); sigma_i = 1.5 * 2^(quantum + 23)
      const uint32_t eb = __float_as_uint(mx) >> 23;
      float sg1, sg2, sg3;
      if (eb == 0) {               // zero (or denormal) column in this group: slices are the values themselves (zeros)
        sg1 = sg2 = sg3 = 0.f;
      } else {
        if (eb < 87 || eb > 167) atomicOr(p.status, 1);   // |a| outside 2^-40 .. 2^40 (or Inf): leave it to the Householder leaf
        const uint32_t ec = eb < 30 ? 30 : (eb > 230 ? 230 : eb);
        sg1 = __uint_as_float(((ec + 16) << 23) | 0x400000u);
        sg2 = __uint_as_float(((ec + 8) << 23) | 0x400000u);
        sg3 = __uint_as_float((ec << 23) | 0x400000u);
      }
      mbar_wait(slfree0 + 8 * t, ((g / kNSl) & 1) ^ 1);   // the MMAs of group g - 2 are done with this slice stage
      const uint32_t dst = sl0 + t * kSlStage + wr_off;
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) {                  // one 16-byte chunk (8 k) of each slice per pass
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          float s1[2], s2[2], s3[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float x = v[8 * i + k + e];
            const float a1 = __fsub_rn(__fadd_rn(x, sg1), sg1);
            const float r1 = __fsub_rn(x, a1);
            const float a2 = __fsub_rn(__fadd_rn(r1, sg2), sg2);
            const float r2 = __fsub_rn(r1, a2);
            const float a3 = __fsub_rn(__fadd_rn(r2, sg3), sg3);
            s1[e] = a1; s2[e] = a2; s3[e] = a3;
          }
          w1[k >> 1] = pack_bf16(s1[0], s1[1]);
          w2[k >> 1] = pack_bf16(s2[0], s2[1]);
          w3[k >> 1] = pack_bf16(s3[0], s3[1]);
        }
        const uint32_t o = ((chunk0 + i) ^ swz) << 4;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + o), "r"(w1[0]), "r"(w1[1]), "r"(w1[2]), "r"(w1[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kSlTile + o), "r"(w2[0]), "r"(w2[1]), "r"(w2[2]), "r"(w2[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * kSlTile + o), "r"(w3[0]), "r"(w3[1]), "r"(w3[2]), "r"(w3[3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(conv0 + 8 * t);
    }
  } else {
    // epilogue: warp reads TMEM lanes 32 qq .. + 31.  Lanes 0-63 (row i of [S1|S2]^T = column i of S1): U(i, :);
    // lanes 64-127 (column i of S2): V(i, :).
    const int qq = warp & 3;
    const int tl = 32 * qq + lane;
    const bool upper = qq < 2;
    double* slab = p.slabs + (size_t)blockIdx.x * kSlabDoubles + tl;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    bool first = true;
    for (int g = 0; g < ng; ++g) {
      const uint32_t a = g & 1;
      mbar_wait(tfull0 + 8 * a, (g >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t td = tmem_base + a * kAccCols + ((uint32_t)(32 * qq) << 16);
      if (upper) {
#pragma unroll
        for (int jc = 0; jc < 8; ++jc) {
          float d1[8], d2[8], d3[8];
          tmem_ld8(td + 8 * jc, d1);
          tmem_ld8(td + 64 + 8 * jc, d2);
          tmem_ld8(td + 128 + 8 * jc, d3);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[8 * jc + j] += __fadd_rn(__fmaf_rn(0.5f, d1[j], d2[j]), d3[j]);
        }
      } else {
#pragma unroll
        for (int jc = 0; jc < 8; ++jc) {
          float d2[8], d3[8];
          tmem_ld8(td + 64 + 8 * jc, d2);
          tmem_ld8(td + 128 + 8 * jc, d3);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[8 * jc + j] += __fmaf_rn(0.5f, d2[j], d3[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * a);
      if ((g + 1) % kDump == 0 || g + 1 == ng) {
#pragma unroll
        for (int jb = 0; jb < 64; jb += 8) {   // eight at a time: all 64 loads in flight at once would spill
          double cur[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cur[j] = first ? 0.0 : slab[(jb + j) * 128];
#pragma unroll
          for (int j = 0; j < 8; ++j) { slab[(jb + j) * 128] = cur[j] + (double)acc[jb + j]; acc[jb + j] = 0.f; }
          asm volatile("" ::: "memory");
        }
        first = false;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- reduction of the slabs, fp64 Cholesky, condition bound, gate ---------------------------------------------------------
struct GramFinishParams {
  const double* slabs; int nslabs;
  int n;
  double* g;               // 64 x 64 fp64 scratch (row i by block i)
  float* r; long long ldr; // n x n upper triangular out (zeros below the diagonal)
  int* status;             // [0] in: scale flag from gram_kernel; out: 0 = R written, 1 = the Householder leaf must run (gate)
  unsigned* ticket;
  double* info;            // [0] n * ||R^^-1||_F^2 (bound on cond_2 of the unit-diagonal Gram matrix), [1] smallest pivot / diagonal
  double bound_max;
};

__global__ void __launch_bounds__(256) gram_finish_kernel(GramFinishParams p) {
  __shared__ double red[4][64];
  __shared__ double Gs[65][65];   // upper triangle: G -> R; X = R^^-1 goes below it, X(l, j) at Gs[j + 1][l]
  __shared__ double dg[64], rinv[64], cinv[64];
  __shared__ int s_last, s_fail;
  const int tid = threadIdx.x, j = tid & 63, part = tid >> 6, i = blockIdx.x;
  {
    double s = 0.0;
    for (int b = part; b < p.nslabs; b += 4) {
      const double* S = p.slabs + (size_t)b * kSlabDoubles;
      // U(i,j) = S[j*128 + i], V(i,j) = S[j*128 + 64 + i]
      s += (__ldcg(S + j * 128 + i) + __ldcg(S + i * 128 + j)) + (__ldcg(S + j * 128 + 64 + i) + __ldcg(S + i * 128 + 64 + j));
    }
    red[part][j] = s;
    __syncthreads();
    if (part == 0) p.g[i * 64 + j] = (red[0][j] + red[1][j]) + (red[2][j] + red[3][j]);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1);
    s_fail = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int n = p.n;
  for (int e = tid; e < 64 * 64; e += 256) Gs[e >> 6][e & 63] = __ldcg(p.g + e);
  __syncthreads();
  if (tid < 64) dg[tid] = Gs[tid][tid];
  __syncthreads();
  // right-looking Cholesky on the upper triangle, unscaled rows: after step k row k holds r(k,:) * r(k,k)
  for (int k = 0; k < n; ++k) {
    const double d = Gs[k][k];
    if (!(d > 0.0) || !(d <= 1.7e308) || !(d > 1e-30 * dg[k])) { if (tid == 0) s_fail = 1; break; }   // uniform: every thread reads the same d
    const double rd = 1.0 / d;
    if (j > k && j < n) {
      const double w = Gs[k][j] * rd;
      for (int ii = k + 1 + part; ii <= j; ii += 4) Gs[ii][j] -= Gs[k][ii] * w;
    }
    __syncthreads();
  }
  __syncthreads();
  const bool fail = s_fail != 0 || (*(volatile int*)p.status != 0);
  double bound = 0.0, minpiv = 0.0;
  if (!fail) {
    // R(k,j) = Gs[k][j] / sqrt(Gs[k][k]);   R^ = R diag(1/sqrt(g_jj))
    if (tid < n) rinv[tid] = 1.0 / sqrt(Gs[tid][tid]);
    __syncthreads();
    for (int e = tid; e < 64 * 64; e += 256) {
      const int k = e >> 6, jj = e & 63;
      if (k <= jj && jj < n) Gs[k][jj] *= rinv[k];          // R
    }
    __syncthreads();
    for (int e = tid; e < 64 * 64; e += 256) {
      const int k = e >> 6, jj = e & 63;
      float out = 0.f;
      if (k <= jj && jj < n) out = (float)Gs[k][jj];
      if (k < n && jj < n) p.r[k + (long long)jj * p.ldr] = out;
    }
    // X = R^^-1, column j by back substitution in "axpy" form: all columns advance together, row l = n-1 .. 0;
    // thread (j, part) carries the partial sums of rows part + 4 u of column j.
    double sacc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) sacc[u] = 0.0;
    if (tid < n) cinv[tid] = 1.0 / sqrt(dg[tid]);   // R^(i, l) = R(i, l) cinv[l]
    __syncthreads();
    for (int l = n - 1; l >= 0; --l) {
      if ((l & 3) == part && j < n && j >= l) {      // x(l, j) = (delta_lj - s_l) / r^(l, l)
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u) if (u == (l >> 2)) s = sacc[u];
        Gs[j + 1][l] = ((j == l ? 1.0 : 0.0) - s) / (Gs[l][l] * cinv[l]);
      }
      __syncthreads();
      if (j < n && j >= l) {
        const double x = Gs[j + 1][l] * cinv[l];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ii = part + 4 * u;
          if (ii < l) sacc[u] += Gs[ii][l] * x;
        }
      }
    }
    __syncthreads();
    // ||X||_F^2 and the smallest scaled pivot
    double f = 0.0;
    for (int e = tid; e < 64 * 64; e += 256) {
      const int k = e >> 6, jj = e & 63;
      if (k <= jj && jj < n) f += Gs[jj + 1][k] * Gs[jj + 1][k];
    }
    for (int o = 16; o; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if ((tid & 31) == 0) red[0][tid >> 5] = f;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[0][w];
      bound = (double)n * t;
      minpiv = 1e300;
      for (int k = 0; k < n; ++k) { const double r = Gs[k][k] * Gs[k][k] / dg[k]; if (r < minpiv) minpiv = r; }
    }
  }
  if (tid == 0) {
    const bool gate = fail || !(bound <= p.bound_max);
    p.info[0] = fail ? -1.0 : bound;
    p.info[1] = minpiv;
    *p.status = gate ? 1 : 0;
    *p.ticket = 0u;
    __threadfence();
  }
}

}  // namespace

bool gram_tsqr_eligible(const float* a, long long lda, long long m, int n) {
  return umma_available() && lda % 4 == 0 && aligned16(a) && n >= 1 && n <= 64 && m >= kGR && m < (1ll << 31) - 256;
}

size_t gram_tsqr_workspace_floats(int sm_count) {   // slabs + G + info (doubles), then ticket + status
  return 2 * ((size_t)sm_count * kSlabDoubles + 64 * 64 + 8) + 16;
}

// Enqueues the Gram leaf; *gate_out is the device flag that is 0 when R was written and 1 when the Householder leaf behind
// it has to produce R.  Returns false when the launch could not be made (nothing enqueued).
bool launch_tsqr_gram_r(const float* a, long long lda, long long m, int n, float* r, long long ldr, float* ws, int sm_count,
                        int max_ctas, double bound_max, int** gate_out, double** info_out, cudaStream_t s) {
  if (!gram_tsqr_eligible(a, lda, m, n)) return false;
  CUtensorMap tm;
  if (!make_map(&tm, a, m, n, lda, 32, 64)) return false;
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem) != cudaSuccess) {
      cudaGetLastError();
      once.retry();
      return false;
    }
  }
  double* slabs = reinterpret_cast<double*>(ws);
  double* g = slabs + (size_t)sm_count * kSlabDoubles;
  double* info = g + 64 * 64;
  unsigned* ticket = reinterpret_cast<unsigned*>(info + 8);
  int* status = reinterpret_cast<int*>(ticket + 4);
  const long long groups = (m + kGR - 1) / kGR;
  if (max_ctas < 1 || max_ctas > sm_count) max_ctas = sm_count;
  const int grid = (int)(groups < max_ctas ? groups : max_ctas);
  // status and ticket start at zero: the finish kernel leaves the ticket at zero and rewrites status, but the scale flag is
  // OR-ed in by gram_kernel, so clear it per call
  if (cudaMemsetAsync(ticket, 0, 32, s) != cudaSuccess) { cudaGetLastError(); return false; }
  GramParams gp{m, groups, slabs, status};
  ++g_launches;
  gram_kernel<<<grid, kGramThreads, kGramSmem, s>>>(tm, gp);
  if (cudaGetLastError() != cudaSuccess) { --g_launches; return false; }
  GramFinishParams fp{slabs, grid, n, g, r, ldr, status, ticket, info, bound_max};
  ++g_launches;
  gram_finish_kernel<<<64, 256, 0, s>>>(fp);
  if (gate_out) *gate_out = status;
  if (info_out) *info_out = info;
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace cqr
