// gemm_umma.cu -- tcgen05 / TMEM / TMA GEMMs for the trailing update (north_star item 3):
//     TN:  D[z] = A(Kz,:)^T B(Kz,:)          W = V^T C, X = T^T W, Gram = V^T V   (split-K partials)
//     NN:  D    = alpha A B + beta D (+ D_lo)  C -= V X, X = T W
// fp32 fidelity on a tensor pipe that has no fp32 MMA comes from the 3xTF32 split: every
// operand is held as hi + lo (hi = top 11 significand bits, lo = exact remainder, see
// common.cuh) and each K step issues  A_hi B_hi + A_lo B_hi + A_hi B_lo  into the same TMEM
// accumulator (the dropped lo*lo term is 2^-22 relative).  The split happens INSIDE the kernel:
// TMA lands the raw fp32 tile (which the tensor core reads as `hi` -- it ignores the low 13 bits),
// four converter warps write lo = x - trunc(x) to a sibling buffer at the same swizzled offsets,
// fence.proxy.async, and hand the stage to the MMA warp.  HBM/L2 see every operand exactly once.
//   warp 0   : TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warps 2-5: converters during the main loop (ld.shared.v4 -> lo -> st.shared.v4), then the
//              epilogue (tcgen05.ld 32x32b -> registers -> coalesced column-major stores)
//   warp 1   : MMA issuer   (one elected lane, tcgen05.mma kind::tf32, M=128, N=BN, K=8)
// A is K-major for TN (columns of V are contiguous along K) and MN-major for NN (V itself);
// B is always K-major.  Ragged M/N/K edges rely on TMA zero fill plus masked stores.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "umma_common.cuh"

namespace cqr {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;               // fp32 elements per 128-byte swizzle row

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// kind::tf32 instruction descriptor: F32 accumulate, TF32 A/B, M = 128, N = BN.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct Epilogue {
  float* d;
  long long ldd;
  long long split_stride;   // TN: partial z at d + z*split_stride
  float alpha, beta;
  int prefetch_kb;          // k-block (counted from the END of the tile's K loop) at which the C tile is pulled into L2; < 0: never
};


// Persistent, warp-specialised: grid = min(tiles, #SMs); tile t -> (m-tile fastest, then n-tile, then K split).
//   warp 0       TMA producer: raw fp32 A/B k-blocks into a 2-stage ring (48 KB raw + 48 KB lo per stage at BN = 256)
//   warps 2-9    converters: lo = x - trunc(x) of the landed k-block into the stage's lo half (overlaps the
//                MMAs of the previous k-block)
//   warp 1       MMA issuer: 12 tcgen05.mma per k-block into one of two TMEM accumulators
//   warps 10-17  epilogue: drain the other accumulator (tcgen05.ld) while the next tile's MMAs run.  A warp may only
//                touch TMEM lanes 32 (warp % 4) .. +31, so two warps share a lane quarter and split the columns; each
//                walks its half in 16-column chunks with the C values of the NEXT chunk already in flight (the
//                C -= V X update reads 128 KB of C per tile: with 4 warps and no prefetch the epilogue, not the
//                tensor pipe, set the pace: 122 vs 190 TF/s at K = 256)
// Barriers: full[s] (TMA landed), conv[s] (lo written, 256 arrivals), empty[s] (MMAs done with the stage),
// tfull[a]/tempty[a] per accumulator.  All roles walk the same (tile, k-block) sequence, so phases are
// derived from two running counters: g = k-blocks so far, i = tiles so far on this CTA.
constexpr int kRing = 2;
constexpr int kConvThreads = 256;
constexpr int kEpiThreads = 256;
constexpr int kThreadsP = 64 + kConvThreads + kEpiThreads;

template <int BN, bool kAMn>
__global__ void __launch_bounds__(kThreadsP, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_d, Epilogue ep, int M, int N, int K, int kper, int tiles_m,
                 int tiles_n, int splits) {
  constexpr uint32_t kABytes = BM * BK * 4;            // 16 KB
  constexpr uint32_t kBBytes = BN * BK * 4;
  constexpr uint32_t kRawBytes = kABytes + kBBytes;    // one k-block: [A raw][B raw]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kStageBytes = 2 * kRawBytes;    // [A raw][B raw][A lo][B lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRing * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kRing + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kRing, conv0 = full0 + 16 * kRing,
                 tfull0 = full0 + 24 * kRing, tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = tiles_m * tiles_n * splits;

  if (threadIdx.x == 0) {
    for (int r = 0; r < kRing; ++r) {
      mbar_init(full0 + 8 * r, 1);
      mbar_init(empty0 + 8 * r, 1);
      mbar_init(conv0 + 8 * r, kConvThreads);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // tile decomposition shared by all roles
  auto tile_coords = [&](int t, int& m0, int& n0, int& z, int& kbeg, int& nkb) {
    const int mt = t % tiles_m;
    const int rest = t / tiles_m;
    const int nt = rest % tiles_n;
    z = rest / tiles_n;
    m0 = mt * BM; n0 = nt * BN;
    kbeg = z * kper;
    const int kend = min(K, kbeg + kper);
    nkb = (kend - kbeg + BK - 1) / BK;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, z, kbeg, nkb;
        tile_coords(t, m0, n0, z, kbeg, nkb);
        // pull the C tile into L2 shortly before the epilogue reads it: issued too early (at the start of the tile)
        // the lines are evicted again by the ~100 MB the other CTAs stream through L2 in the meantime
        const int pf_at = (ep.beta != 0.f && ep.prefetch_kb >= 0) ? max(0, nkb - 1 - ep.prefetch_kb) : -1;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          if (kb == pf_at) {
#pragma unroll
            for (int c = 0; c < BN; c += 64) tma_prefetch_2d(&tm_d, m0, n0 + c);
          }
          const uint32_t r = g % kRing, ph = (g / kRing) & 1;
          mbar_wait(empty0 + 8 * r, ph ^ 1);
          const uint32_t sa = smem_base + r * kStageBytes;
          const uint32_t full = full0 + 8 * r;
#if defined(CQR_GEMM_EXP) && CQR_GEMM_EXP == 2
          mbar_expect_tx(full, 2 * kRawBytes);
#elif defined(CQR_GEMM_EXP) && CQR_GEMM_EXP == 3
          mbar_expect_tx(full, kRawBytes + kABytes);
#else
          mbar_expect_tx(full, kRawBytes);
#endif
          const int k0 = kbeg + kb * BK;
          if (kAMn) {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), &tm_a, full, m0 + 32 * c, k0);
          } else {
            tma_load_2d(sa, &tm_a, full, k0, m0);
          }
          tma_load_2d(sa + kABytes, &tm_b, full, k0, n0);
#if defined(CQR_GEMM_EXP) && CQR_GEMM_EXP >= 2   // timing experiment: the lo halves arrive by TMA too (values are wrong)
          if (kAMn) {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + kRawBytes + c * (BK * 128), &tm_a, full, m0 + 32 * c, k0);
          } else {
            tma_load_2d(sa + kRawBytes, &tm_a, full, k0, m0);
          }
#if CQR_GEMM_EXP == 2
          tma_load_2d(sa + kRawBytes + kABytes, &tm_b, full, k0, n0);
#endif
#endif
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, kAMn);
      uint32_t g = 0, i = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        int m0, n0, z, kbeg, nkb;
        tile_coords(t, m0, n0, z, kbeg, nkb);
        const uint32_t a = i & 1;
        mbar_wait(tempty0 + 8 * a, ((i >> 1) & 1) ^ 1);          // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + a * BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t r = g % kRing, ph = (g / kRing) & 1;
          mbar_wait(conv0 + 8 * r, ph);                            // raw landed AND lo written
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + r * kStageBytes;
          const uint32_t lo_base = sa + kRawBytes;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            uint64_t a_hi, a_lo;
            if (kAMn) {   // K = 8 is two 4-deep groups (SBO = 512 B): +1024 B per step; M chunks 4096 B apart (LBO)
              a_hi = make_desc(sa + k * 1024, BK * 128, 512, 1);
              a_lo = make_desc(lo_base + k * 1024, BK * 128, 512, 1);
            } else {      // K-major: +32 B inside the swizzled 128 B row
              a_hi = make_desc(sa + k * 32, 16, 1024, 2);
              a_lo = make_desc(lo_base + k * 32, 16, 1024, 2);
            }
            const uint64_t b_hi = make_desc(sa + kABytes + k * 32, 16, 1024, 2);
            const uint64_t b_lo = make_desc(lo_base + kABytes + k * 32, 16, 1024, 2);
            umma_tf32(tmem_d, a_lo, b_hi, idesc, (kb | k) != 0);
            umma_tf32(tmem_d, a_hi, b_lo, idesc, 1);
            umma_tf32(tmem_d, a_hi, b_hi, idesc, 1);
          }
          umma_commit(empty0 + 8 * r);   // stage (raw + lo) may be refilled
        }
        umma_commit(tfull0 + 8 * a);
      }
    }
  } else if (warp < 2 + kConvThreads / 32) {
    // converters: lo = x - trunc(x), same (swizzled) offsets
    const int ct = threadIdx.x - 64;
    uint32_t g = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int m0, n0, z, kbeg, nkb;
      tile_coords(t, m0, n0, z, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const uint32_t r = g % kRing, ph = (g / kRing) & 1;
        mbar_wait(full0 + 8 * r, ph);     // implies the stage's previous MMAs are done (producer waited on empty)
        const uint32_t sa = smem_base + r * kStageBytes;
        const uint32_t lo_base = sa + kRawBytes;
#if defined(CQR_GEMM_EXP) && CQR_GEMM_EXP == 3
        constexpr uint32_t kConvFrom = kABytes;
#elif defined(CQR_GEMM_EXP)
        constexpr uint32_t kConvFrom = kRawBytes;
#else
        constexpr uint32_t kConvFrom = 0;
#endif
#pragma unroll 4
        for (uint32_t off = kConvFrom + ct * 16; off < kRawBytes; off += kConvThreads * 16) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa + off));
          v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to tcgen05.mma
        mbar_arrive(conv0 + 8 * r);
      }
    }
  } else {
    // epilogue: warp reads TMEM lanes [32q, 32q+32) = rows m0 + 32q + lane, columns [h BN/2, (h+1) BN/2).
    // Interior tiles take a lean path (pointer-increment addressing, no per-element bounds checks, chunk loop NOT
    // unrolled): the first version was fully unrolled with 64-bit index arithmetic and predicates per element and
    // was bound by its own instruction stream (ncu: 19 % stall_no_inst = I-cache misses, 21 % stall_wait).
    const int q = warp & 3, h = (warp - (2 + kConvThreads / 32)) >> 2;
    constexpr int CH = 16, NCHUNK = (BN / 2) / CH;
    static_assert(NCHUNK % 2 == 0, "chunk loop is unrolled by two (double-buffered C prefetch)");
    const bool has_c = ep.beta != 0.f;
    const long long ldd = ep.ldd;
    uint32_t i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      int m0, n0, z, kbeg, nkb;
      tile_coords(t, m0, n0, z, kbeg, nkb);
      const uint32_t a = i & 1;
      const int m = m0 + 32 * q + lane;
      const bool row_ok = m < M;
      const int nbase = n0 + h * (BN / 2);
      float* colp = ep.d + (long long)z * ep.split_stride + m + (long long)nbase * ldd;   // (m, nbase)
      const uint32_t tmem_d = tmem_base + a * BN + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * (BN / 2));
      const int nvalid = N - nbase;   // valid columns of this warp's half (may be <= 0 or >= BN / 2)
      auto load_chunk = [&](float (&buf)[CH], const float* p, int lim) {   // lim = valid columns of the chunk
        if (lim >= CH) {
#pragma unroll
          for (int j = 0; j < CH; ++j, p += ldd) buf[j] = __ldcg(p);
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j, p += ldd) buf[j] = (j < lim) ? __ldcg(p) : 0.f;
        }
      };
      auto store_chunk = [&](const float (&v)[CH], const float (&c_old)[CH], float* p, int lim) {
        if (lim >= CH) {
#pragma unroll
          for (int j = 0; j < CH; ++j, p += ldd) {
            float r = ep.alpha * v[j];
            if (has_c) r = fmaf(ep.beta, c_old[j], r);
            *p = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j, p += ldd) {
            float r = ep.alpha * v[j];
            if (has_c) r = fmaf(ep.beta, c_old[j], r);
            if (j < lim) *p = r;
          }
        }
      };
      if (m0 + BM <= M) {   // all rows of the tile exist (columns may be ragged: masked per chunk)
        float old[CH], nxt[CH];
        if (has_c) load_chunk(old, colp, nvalid);   // first chunk of C: issued before the accumulator is even complete
        mbar_wait(tfull0 + 8 * a, (i >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < NCHUNK; c += 2) {
          const int lim0 = nvalid - CH * c, lim1 = lim0 - CH, lim2 = lim1 - CH;
          if (lim0 <= 0) break;
          float v[CH];
          tmem_ld16(tmem_d + (uint32_t)(CH * c), v);
          if (has_c && lim1 > 0) load_chunk(nxt, colp + CH * ldd, lim1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          store_chunk(v, old, colp, lim0);
          if (lim1 > 0) {
            tmem_ld16(tmem_d + (uint32_t)(CH * (c + 1)), v);
            if (has_c && c + 2 < NCHUNK && lim2 > 0) load_chunk(old, colp + 2 * CH * ldd, lim2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            store_chunk(v, nxt, colp + CH * ldd, lim1);
          }
          colp += 2 * CH * ldd;
        }
        if (nvalid <= 0) { /* nothing stored: the accumulator still has to be released below */ }
      } else {
        mbar_wait(tfull0 + 8 * a, (i >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {   // bottom edge tiles: masked rows and columns, no prefetch
          float v[CH];
          tmem_ld16(tmem_d + (uint32_t)(CH * c), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int nb = nbase + CH * c;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (row_ok && nb + j < N) {
              float* p = colp + (long long)(CH * c + j) * ldd;
              float r = ep.alpha * v[j];
              if (has_c) r = fmaf(ep.beta, __ldcg(p), r);
              *p = r;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * a);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------

int sm_count() {   // per device: a process may hold contexts on several GPUs
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

template <int BN, bool kAMn>
bool launch_umma(int M, int N, int K, const float* a, long long lda, const float* b, long long ldb, const Epilogue& ep,
                 int splits, int max_ctas, cudaStream_t s) {
  CUtensorMap ta, tb, td;
  bool ok;
  if (kAMn) ok = make_map(&ta, a, M, K, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  else ok = make_map(&ta, a, K, M, lda, BK, BM);
  ok = ok && make_map(&tb, b, K, N, ldb, BK, BN);
  // D is only touched by TMA for the L2 prefetch of the C tile (beta != 0); plain (unswizzled) 128 x 64 boxes
  if (ep.beta != 0.f && ep.ldd % 4 == 0 && aligned16(ep.d)) ok = ok && make_map(&td, ep.d, M, N, ep.ldd, BM, 64, CU_TENSOR_MAP_SWIZZLE_NONE);
  else td = tb;
  if (!ok) return false;
  Epilogue e2 = ep;
  constexpr size_t smem = (size_t)kRing * 2 * (BM * BK * 4 + BN * BK * 4) + 1024 + 256;
  static PerDeviceOnce once;
  if (once.first()) {
    if (cudaFuncSetAttribute(umma_gemm_kernel<BN, kAMn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      once.retry();
      return false;
    }
  }
  int kper = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  if (kper < BK) kper = BK;
  splits = (K + kper - 1) / kper;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * splits;
  int grid;
  if (max_ctas < 0) {            // at most -max_ctas tiles per CTA: the grid may exceed the SM count (CTAs queue in hardware)
    const long long g = (total - max_ctas - 1) / -max_ctas;
    grid = (int)(g < 1 ? 1 : g);
  } else {
    if (max_ctas < 1 || max_ctas > sm_count()) max_ctas = sm_count();
    grid = (int)(total < max_ctas ? total : max_ctas);
  }
  ++g_launches;
  umma_gemm_kernel<BN, kAMn><<<grid, kThreadsP, smem, s>>>(ta, tb, td, e2, M, N, K, kper, tiles_m, tiles_n, splits);
  if (cudaGetLastError() != cudaSuccess) { --g_launches; return false; }   // launch refused: the caller falls back to the SIMT GEMM
  return true;
}

}  // namespace

bool umma_available() {   // per device
  static signed char state[64];
  static bool init = false;
  if (!init) { for (int i = 0; i < 64; ++i) state[i] = -1; init = true; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (state[dev] < 0) {
    int major = 0;
    state[dev] = (get_encode() != nullptr && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
                  major == 10) ? 1 : 0;
  }
  return state[dev] == 1;
}

// Number of K splits the kernel will really use for a request (the reduction must agree).
int umma_effective_splits(int K, int splits) {
  if (splits < 1) splits = 1;
  int kper = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  if (kper < BK) kper = BK;
  return (K + kper - 1) / kper;
}

bool launch_gemm_tn_umma(int M, int N, int K, const float* a, long long lda, const float* b, long long ldb, float* d,
                         long long ldd, int splits, long long d_split_stride, int max_ctas, cudaStream_t s) {
  if (!umma_available()) return false;
  if (lda % 4 || ldb % 4 || !aligned16(a) || !aligned16(b)) return false;
  if (M < 1 || N < 1 || K < 1) return false;
  Epilogue ep{d, ldd, d_split_stride, 1.f, 0.f, -1};
  if (N > 128) return launch_umma<256, false>(M, N, K, a, lda, b, ldb, ep, splits, max_ctas, s);
  return launch_umma<128, false>(M, N, K, a, lda, b, ldb, ep, splits, max_ctas, s);
}

bool launch_gemm_nn_umma(int M, int N, int K, float alpha, const float* a, long long lda, const float* b, long long ldb,
                         float beta, float* d, long long ldd, int max_ctas, cudaStream_t s) {
  if (!umma_available()) return false;
  if (lda % 4 || ldb % 4 || !aligned16(a) || !aligned16(b)) return false;
  if (M < 1 || N < 1 || K < 1) return false;
  static int pf = -2;
  if (pf == -2) { const char* e = getenv("CQR_CPREFETCH"); pf = e ? atoi(e) : 1000; }   // default: at the start of the tile
  Epilogue ep{d, ldd, 0, alpha, beta, pf};
  if (N > 128) return launch_umma<256, true>(M, N, K, a, lda, b, ldb, ep, 1, max_ctas, s);
  return launch_umma<128, true>(M, N, K, a, lda, b, ldb, ep, 1, max_ctas, s);
}

}  // namespace cqr
