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
//   * a = s1 + s2 + s3 + rho with s1, s2 = RN(residual to a multiple of 2^(E_c - 8), 2^(E_c - 16)) -- |s1| <= 256, |s2| <= 128
//     quanta -- and s3 = the bf16 rounding of the rest (|r2| <= 2^(E_c - 17): 8 significant bits reach 2^(E_c - 24) or finer).
//     Every slice is exact in bf16 (8-bit significand) and |rho| <= 2^(E_c - 25): elements in the column's top binade are
//     represented exactly, the rest to 2^-25 of the column maximum (half an fp32 ulp of the maximum);
//   * ONE tcgen05.mma (kind::f16, bf16 x bf16 -> f32, M = 128, N = 192, K = 16) per 16 rows forms
//         [S1 | S2]^T [S1 | S2 | S3]  =  D11 D12 D13
//                                        D21 D22 D23            (Dxy = Sx^T Sy, 64 x 64 each)
//     in TMEM.  Every product of the fixed-point slices (D11, D12, D22) is an integer < 2^16 times the pair's quantum and a
//     group adds 128 of them: all partial sums are integers <= 2^23 in that unit, i.e. EXACT in the fp32 accumulator whatever
//     the tensor pipe's rounding mode; the blocks with S3 are 2^-16 below D11, a truncated bit there is 2^-36 of G
//     (measured: it truncates, profiles/r02_tsqr_mma_accuracy.txt);
//   * after 128 rows the epilogue warps read the accumulator (tcgen05.ld) and add U = D11/2 + D12 + D13 (TMEM lanes 0-63)
//     and V = D22/2 + D23 (lanes 64-127) to fp64 running sums held in registers for the whole kernel (one store per CTA at
//     the end); G = sum over CTAs of U + U^T + V + V^T (the dropped D33 is 2^-32 relative);
//   * gram_finish_kernel reduces the slabs in fp64 (fixed order), and its last block factors G = R^T R in fp64, bounds
//     cond of the diagonally scaled Gram matrix by n * ||R^^-1||_F^2 (explicit triangular inverse) and either writes R
//     (fp32) or raises the gate that lets the Householder leaf behind it run (ill-conditioned, non-finite or badly
//     scaled input) -- no host synchronisation either way.
// Data flow of gram_kernel (persistent, one CTA per SM, 576 threads):
//   warp 0     TMA producer: four 32-row x 64-column fp32 boxes (128B swizzle) = one 128-row group per stage, 3 stages
//   warps 2-9  converters: stage -> registers (32 values of one column per thread), column maxima, slices -> bf16 K-major
//              128B-swizzled operand tiles S1|S2|S3 (2 slice stages), fence.proxy.async
//   warp 1     MMA issuer: 8 tcgen05.mma per group into one of two TMEM accumulators (192 columns each)
//   warps 10-17 epilogue (two per TMEM lane quarter, 32 columns each)
// Per 128-row group and SM: 32 KB from HBM (1.4 K clk at 6.5 TB/s), 8 MMAs x 96 clk, ~76 KB of shared-memory traffic.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
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
constexpr int kConvT = 256, kEpiT = 256;
constexpr int kGramThreads = 64 + kConvT + kEpiT;
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

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {   // FADD2: two rounded fp32 adds in one issue slot
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
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
      uint64_t v[16];                 // 32 values as fp32 pairs (k, k + 1)
      const uint32_t src = raw0 + r * kRawStage + rd_off;
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[2 * j]), "=l"(v[2 * j + 1]) : "r"(src + ((j ^ swz) << 4)));
      }
      float mx = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float lo, hi;
        unpack_f32x2(v[k], lo, hi);
        mx = fmaxf(mx, fmaxf(fabsf(lo), fabsf(hi)));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      // every value has been consumed (the maximum depends on all 32 loads), so the loads have completed: only now may
      // TMA refill the stage.  Arriving right after ISSUING the loads let a fast (L2-fed) refill overtake them.
      mbar_arrive(rawfree0 + 8 * r);
      // 2^E > mx with biased exponent eb + 1; slice quanta 2^(E-8), 2^(E-16), 2^(E-24); sigma_i = 1.5 * 2^(quantum + 23)
      const uint32_t eb = __float_as_uint(mx) >> 23;
      float sg1, sg2;
      if (eb == 0) {               // zero (or denormal) column in this group: slices are the values themselves (zeros)
        sg1 = sg2 = 0.f;
      } else {
        if (eb < 87 || eb > 167) atomicOr(p.status, 1);   // |a| outside 2^-40 .. 2^40 (or Inf): leave it to the Householder leaf
        const uint32_t ec = eb < 30 ? 30 : (eb > 230 ? 230 : eb);
        sg1 = __uint_as_float(((ec + 16) << 23) | 0x400000u);
        sg2 = __uint_as_float(((ec + 8) << 23) | 0x400000u);
      }
      const uint64_t S1 = pack_f32x2(sg1, sg1), S2 = pack_f32x2(sg2, sg2);
      mbar_wait(slfree0 + 8 * t, ((g / kNSl) & 1) ^ 1);   // the MMAs of group g - 2 are done with this slice stage
      const uint32_t dst = sl0 + t * kSlStage + wr_off;
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) {                  // one 16-byte chunk (8 k) of each slice per pass
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const uint64_t x = v[4 * i + pp];
          const uint64_t a1 = sub2(add2(x, S1), S1);     // RN to a multiple of the first quantum (both halves at once)
          const uint64_t r1 = sub2(x, a1);               // exact
          const uint64_t a2 = sub2(add2(r1, S2), S2);
          const uint64_t r2 = sub2(r1, a2);
          // third slice: the bf16 conversion itself rounds r2 (|r2| <= 2^(E-17)) to 8 significant bits -- at least as fine as
          // the fixed quantum 2^(E-24).  Its products are no longer integers of one quantum, but D13 / D23 sit 2^-16 below D11:
          // what the accumulator may truncate there is 2^-36 of G.
          float lo, hi;
          unpack_f32x2(a1, lo, hi); w1[pp] = pack_bf16(lo, hi);
          unpack_f32x2(a2, lo, hi); w2[pp] = pack_bf16(lo, hi);
          unpack_f32x2(r2, lo, hi); w3[pp] = pack_bf16(lo, hi);
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
    // epilogue: a warp may only touch TMEM lanes 32 (warp % 4) .. + 31; two warps share a lane quarter and take 32 of the
    // 64 columns each.  Lanes 0-63 (row i of [S1|S2]^T = column i of S1) give U(i, :), lanes 64-127 (column i of S2) V(i, :).
    // The running sums live in fp64 REGISTERS for the whole kernel (32 per thread).  A group's contribution D11/2 + D12 + D13
    // is rounded to fp32 once (2^-25 of the group's value, random sign: ~1e-10 of G after a few thousand groups) and then
    // added exactly; nothing is lost between groups.  One F2F per value: the conversion runs on the quarter-rate XU pipe,
    // two per value (x and y converted separately) made the epilogue the slowest stage at 36 % XU utilisation.  (First version: fp32 running sums dumped into an fp64 slab in global memory every 8 groups --
    // the read-modify-write of the slab and 15 spilled accumulators made the epilogue the slowest stage of the pipeline.)
    const int qq = warp & 3, hh = (warp - (2 + kConvT / 32)) >> 2;
    const int tl = 32 * qq + lane;
    const uint32_t offx = qq < 2 ? 0u : 64u;
    const float wy = qq < 2 ? 1.f : 0.f;
    double acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0;
    for (int g = 0; g < ng; ++g) {
      const uint32_t a = g & 1;
      mbar_wait(tfull0 + 8 * a, (g >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t td = tmem_base + a * kAccCols + ((uint32_t)(32 * qq) << 16) + 32 * hh;
      // one code path for both halves: upper lanes x = D11 / 2, y = D12 + D13; lower lanes x = D22 / 2, y = 0 * D22 + D23.
      // Four columns per pass, the next pass's twelve values already in flight (tcgen05.wait::ld waits for ALL outstanding
      // loads, so the prefetch is issued after the wait and lands while this pass converts and adds).
      float c1[4], c2[4], c3[4], d1[4], d2[4], d3[4];   // two register sets, alternating: no copies between them
      tmem_ld4(td + offx, c1);
      tmem_ld4(td + 64, c2);
      tmem_ld4(td + 128, c3);
#pragma unroll
      for (int jc = 0; jc < 8; jc += 2) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");      // set c has landed
        tmem_ld4(td + offx + 4 * (jc + 1), d1);
        tmem_ld4(td + 64 + 4 * (jc + 1), d2);
        tmem_ld4(td + 128 + 4 * (jc + 1), d3);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[4 * jc + j] += (double)__fmaf_rn(0.5f, c1[j], __fmaf_rn(c2[j], wy, c3[j]));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");      // set d has landed
        if (jc + 2 < 8) {
          tmem_ld4(td + offx + 4 * (jc + 2), c1);
          tmem_ld4(td + 64 + 4 * (jc + 2), c2);
          tmem_ld4(td + 128 + 4 * (jc + 2), c3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[4 * (jc + 1) + j] += (double)__fmaf_rn(0.5f, d1[j], __fmaf_rn(d2[j], wy, d3[j]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * a);
    }
    double* slab = p.slabs + (size_t)blockIdx.x * kSlabDoubles + (size_t)(32 * hh) * 128 + tl;
#pragma unroll
    for (int j = 0; j < 32; ++j) slab[j * 128] = acc[j];
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
  int* status;             // persistent context words: [0] scale flag (set by gram_kernel, cleared here), [1] ticket (left at 0),
                           // [2] gate out: 0 = R written, 1 = the Householder leaf must run
  unsigned* ticket;
  double* info;            // [0] n * ||R^^-1||_F^2 (bound on cond_2 of the unit-diagonal Gram matrix), [1] smallest pivot / diagonal
  double bound_max;
};

// 1 / d and 1 / sqrt(d): hardware seed on the high word (MUFU.RCP64H / RSQ64H, ~2^-20) + one Newton step in fp64 (~2^-40
// relative: a perturbation of G far below the 1e-10 it is known to).  The IEEE division and square root are ~40-instruction
// subroutines with dependent DFMA chains, and they sit on the critical path of each of the 64 elimination steps; a seed through
// fp32 (F2F, MUFU, F2F) costs two more XU round trips per step.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d * r, r, 1.0);        // 1 - d r^2
  return fma(0.5 * r, e, r);
}

constexpr int kFinT = 1024;   // threads of the finish kernel: a 32 x 32 grid of 2 x 2 register tiles over the 64 x 64 matrix

__global__ void __launch_bounds__(kFinT) gram_finish_kernel(GramFinishParams p) {
  __shared__ double red[kFinT];
  __shared__ double rowk[2][72];   // the published pivot row (double-buffered): [0..63] row k, [64] 1 / d_k, [65] 1 / sqrt(d_k), [66] fail flag
  __shared__ double dg[64], pivS[64], csqS[64];
  __shared__ int s_last, s_scale, s_bad[2];   // s_bad[k & 1]: the pivot of step k is not a positive, finite, in-range number
  const int tid = threadIdx.x, i = blockIdx.x;
  pdl_trigger();   // the first gated kernel behind may be scheduled now (it waits for this grid to complete)
  pdl_wait();      // the slabs of gram_kernel
  {
    // T(i, j) = sum over CTAs of U(j, i) + V(j, i) = S[i * 128 + j] + S[i * 128 + 64 + j]: contiguous reads only; G = T + T^T.
    // Eight slab groups in parallel (fixed assignment and order: the result does not depend on timing), loads unrolled by four.
    const int j2 = tid & 127, grp = tid >> 7;
    const double* S = p.slabs + i * 128 + j2;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = grp;
    for (; b + 24 < p.nslabs; b += 32) {
      const double v0 = __ldcg(S + (size_t)b * kSlabDoubles), v1 = __ldcg(S + (size_t)(b + 8) * kSlabDoubles),
                   v2 = __ldcg(S + (size_t)(b + 16) * kSlabDoubles), v3 = __ldcg(S + (size_t)(b + 24) * kSlabDoubles);
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; b < p.nslabs; b += 8) s0 += __ldcg(S + (size_t)b * kSlabDoubles);
    red[tid] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < 64) {
      double t = 0.0;
#pragma unroll
      for (int g = 0; g < 8; ++g) t += red[128 * g + tid] + red[128 * g + 64 + tid];
      p.g[i * 64 + tid] = t;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last block: fp64 Cholesky G = R^T R with the matrix in REGISTERS ---------------------------------------------------
  // Thread (pr, q) = (warp, lane) owns M[2 pr .. 2 pr + 1][2 q .. 2 q + 1].  On and above the diagonal M is G, reduced in place
  // (rows stay unscaled: after step k row k holds r(k,:) r(k,k)); strictly below it M carries the running sums S of the forward
  // substitution for Y = Rs^-T (Rs = R with columns scaled to unit norm, Rs^T Rs = the unit-diagonal Gram matrix):
  //     Y(k,j) = (delta_kj - S(k,j)) / rs(k,k),   S(k',j) += rs(k,k') Y(k,j)  for k' > k >= j     (S(k,k) = 0 always),
  // and ||Y||_F^2 is the trace of the inverse of the unit-diagonal Gram matrix.  Per step the warp that owns row k publishes
  // it (64 values plus 1/d_k and 1/sqrt(d_k), formed one step earlier by the thread that finished the pivot) and everybody
  // updates its tile from four published values: one barrier and ~160 shared-memory wavefronts per step.  (Two versions with
  // the matrix in shared memory took 1450-1670 cycles per step, bound by ~1400 wavefronts of LDS/STS -- in-kernel clocks.)
  const int n = p.n;
  const int pr = tid >> 5, q = tid & 31;
  const int r0 = 2 * pr, c0 = 2 * q;
  double m00, m01, m10, m11;
  {
    auto gl = [&](int r, int c) -> double { return r <= c ? __ldcg(p.g + r * 64 + c) + __ldcg(p.g + c * 64 + r) : 0.0; };
    m00 = gl(r0, c0); m01 = gl(r0, c0 + 1); m10 = gl(r0 + 1, c0); m11 = gl(r0 + 1, c0 + 1);
  }
  if (pr == q) { dg[r0] = m00; dg[r0 + 1] = m11; csqS[r0] = m00 > 0.0 ? sqrt(m00) : 0.0; csqS[r0 + 1] = m11 > 0.0 ? sqrt(m11) : 0.0; }
  __syncthreads();
  // column scales of Rs for this thread's ROW indices (0 for a zero column: the pivot test stops the elimination there)
  const double ci0 = dg[r0] > 0.0 ? 1.0 / csqS[r0] : 0.0, ci1 = dg[r0 + 1] > 0.0 ? 1.0 / csqS[r0 + 1] : 0.0;
  double rcp_next = 0.0, rsq_next = 0.0;        // of the next pivot: valid in the thread that owns it
  if (tid == 0) { rcp_next = 1.0 / m00; rsq_next = 1.0 / sqrt(m00); }
  double ysq = 0.0;
  int failed = 0;
  const long long tk0 = clock64();
  for (int k = 0; k < n; ++k) {
    double* buf = rowk[k & 1];
    if (pr == (k >> 1)) {                       // publish row k
      const bool odd = k & 1;
      const double v0 = odd ? m10 : m00, v1 = odd ? m11 : m01;
      buf[c0] = v0; buf[c0 + 1] = v1;
      if (q == pr) {
        const double d = odd ? m11 : m00;
        // 1e-37 < d < 1e37 on the high word (integer compares: an fp64 compare would sit on the step's critical path);
        // negative, zero, NaN and Inf all fall outside
        const int hi = __double2hiint(d), hg = __double2hiint(dg[k]);
        s_bad[k & 1] = !(hi > 0x38410000 && hi < 0x47A00000 && hg < 0x47A00000);
        buf[64] = rcp_next; buf[65] = rsq_next;
        pivS[k] = d;
      }
    }
    __syncthreads();
    if (s_bad[k & 1]) { failed = 1; break; }    // uniform
    const double rcp = buf[64], rs = buf[65];
    if (pr == (k >> 1)) {                       // R(k, :) = row k / sqrt(d_k), rounded to fp32; zeros left of the diagonal
      const bool odd = k & 1;
      const double v0 = odd ? m10 : m00, v1 = odd ? m11 : m01;
      if (c0 < n) p.r[k + (long long)c0 * p.ldr] = c0 >= k ? (float)(v0 * rs) : 0.f;
      if (c0 + 1 < n) p.r[k + (long long)(c0 + 1) * p.ldr] = c0 + 1 >= k ? (float)(v1 * rs) : 0.f;
    }
    if (r0 + 1 > k) {                           // warp-uniform: rows <= k are finished
      const double a0 = buf[r0], a1 = buf[r0 + 1];          // row k at this thread's row indices
      const double b0 = buf[c0], b1 = buf[c0 + 1];          // ... and at its column indices
      const double scale = rs * csqS[k];                    // 1 / rs(k,k)
      // Cholesky part (k < r <= c): m -= a_r (b_c / d)
      const double w0 = b0 * rcp, w1 = b1 * rcp;
      // forward-substitution part (c <= k < r): m += rs(k,r) Y(k,c);  rs(k,r) = a_r / sqrt(d) * ci_r
      const double y0 = c0 < k ? -b0 * scale : (c0 == k ? scale : 0.0);
      const double y1 = c0 + 1 < k ? -b1 * scale : (c0 + 1 == k ? scale : 0.0);
      const double s0 = a0 * rs * ci0, s1 = a1 * rs * ci1;
      if (r0 > k) {
        if (c0 >= r0) m00 -= a0 * w0; else if (c0 <= k) m00 += s0 * y0;
        if (c0 + 1 >= r0) m01 -= a0 * w1; else if (c0 + 1 <= k) m01 += s0 * y1;
      }
      {
        const int r1 = r0 + 1;                 // r1 > k here
        if (c0 >= r1) m10 -= a1 * w0; else if (c0 <= k) m10 += s1 * y0;
        if (c0 + 1 >= r1) m11 -= a1 * w1; else if (c0 + 1 <= k) m11 += s1 * y1;
      }
      if (pr == ((k + 1) >> 1) && q == pr && k + 1 < n) {   // the next pivot is final: its reciprocals for the next step
        const double dn = ((k + 1) & 1) ? m11 : m00;
        rcp_next = fast_rcp(dn);                 // garbage for a pivot out of range: the publish step flags it, nobody uses it
        rsq_next = fast_rsqrt(dn);
      }
    }
    if (pr == (k >> 1)) {                       // ||Y(k, :)||^2, columns j <= k, by the row's owner warp
      const double scale = rs * csqS[k];
      const double b0 = buf[c0], b1 = buf[c0 + 1];
      const double y0 = c0 < k ? -b0 * scale : (c0 == k ? scale : 0.0);
      const double y1 = c0 + 1 < k ? -b1 * scale : (c0 + 1 == k ? scale : 0.0);
      ysq += y0 * y0 + y1 * y1;
    }
  }
  __syncthreads();
  const long long tk1 = clock64();
  if (tid == 0) s_scale = *(volatile int*)p.status;
  __syncthreads();
  const bool fail = failed != 0 || s_scale != 0;
  double bound = 0.0, minpiv = 0.0;
  if (!fail) {
    for (int o = 16; o; o >>= 1) ysq += __shfl_xor_sync(0xffffffffu, ysq, o);
    if ((tid & 31) == 0) red[tid >> 5] = ysq;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kFinT / 32; ++w) t += red[w];
      bound = (double)n * t;
      minpiv = 1e300;
      for (int k = 0; k < n; ++k) { const double r = pivS[k] / dg[k]; if (r < minpiv) minpiv = r; }
    }
  }
  if (tid == 0) {
    const bool gate = fail || !(bound <= p.bound_max);
    p.info[0] = fail ? -1.0 : bound;
    p.info[1] = minpiv;
    p.info[2] = (double)(tk1 - tk0);   // clocks of the elimination loop (tools/gram_debug.py)
    p.status[2] = gate ? 1 : 0;
    p.status[0] = 0;   // scale flag and ticket are left clean for the next call (the words live in the context, not in the workspace)
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
bool launch_tsqr_gram_r(const float* a, long long lda, long long m, int n, float* r, long long ldr, float* ws, int* flags, int sm_count,
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
  double* info = reinterpret_cast<double*>(flags + 4);   // verdict (bound, smallest pivot ratio, loop clocks): persistent, see below
  // flags: zero-initialised words owned by the context: [0] scale flag, [1] ticket, [2] gate, [4..9] three doubles of verdict.
  // The finish kernel leaves [0] and [1] at zero, so no memset is needed per call, and the verdict stays readable
  // (cqr_tsqr_gram_info) whatever later calls do with the workspace, which is shared and may be reallocated.
  int* status = flags;
  unsigned* ticket = reinterpret_cast<unsigned*>(flags + 1);
  const long long groups = (m + kGR - 1) / kGR;
  if (max_ctas < 1 || max_ctas > sm_count) max_ctas = sm_count;
  const int grid = (int)(groups < max_ctas ? groups : max_ctas);
  GramParams gp{m, groups, slabs, status};
  ++g_launches;
  gram_kernel<<<grid, kGramThreads, kGramSmem, s>>>(tm, gp);
  if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
    if (getenv("CQR_DEBUG")) fprintf(stderr, "gram_kernel launch: %s\n", cudaGetErrorString(e));
    --g_launches;
    return false;
  }
  GramFinishParams fp{slabs, grid, n, g, r, ldr, status, ticket, info, bound_max};
  ++g_launches;
  if (launch_pdl(gram_finish_kernel, dim3(64), dim3(kFinT), 0, s, fp) != cudaSuccess) { cudaGetLastError(); gram_finish_kernel<<<64, kFinT, 0, s>>>(fp); }
  if (gate_out) *gate_out = flags + 2;
  if (info_out) *info_out = info;
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace cqr
