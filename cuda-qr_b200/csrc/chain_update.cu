// chain_update.cu -- the panel chain's K = 64 inner update C <- (I - V op(T) V^T) C in ONE launch.
//
// Inside an outer block every 64-column panel is followed, on the panel stream, by the update of the block's remaining
// <= 192 columns with that panel's reflectors (the reference's trailing update, qr.c:255-293 / qr.cu:335-465, restricted
// to the block).  As three launches -- W = V^T C split-K on tcgen05, reduction + T^T W, C -= V X on tcgen05 -- it costs
// ~73 us per update, 192 times per 16384^2 factorisation and all on the critical path, although the work is tiny
// (2 x 2 m_p 64 192 flop, 29 MB at m_p = 16384): the time is launch hand-overs and the ramp-up / tail of three small
// kernels.  Here the three phases share one kernel on the panel partition's SMs with two grid-wide barriers in between
// (every CTA is resident: the grid is at most the partition's SM count and the panel stream runs nothing else):
//   1. each CTA forms its row slab's share of W = V^T C  (64 x nc) in registers and writes it out;
//   2. CTA p sums the shares for its few columns and multiplies by op(T);
//   3. each CTA updates its slab, C -= V X, with X (64 x nc) in shared memory.
// Plain fp32 on the FMA pipe (FFMA2 over row pairs): at these sizes the tensor pipe buys nothing (the update is
// latency-bound below ~8192 rows) and fp32 FMA needs no operand split.  Shared-memory layouts are chosen so that every
// inner-loop load is a conflict-free 64-bit access: V_s[k][r] with a stride of 66 floats, threads owning k = tk + 16 i.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "warp_math.cuh"

namespace cqr {

namespace {

constexpr int CU_KB = 64, CU_THREADS = 512;

// NC = widest C handled, SUB = rows per staged sub-tile.  <192, 64>: the chain's inner update (32-SM partition, long
// slabs); <256, 32>: the look-ahead slice onto the next block's 256 columns (GEMM partition, short slabs).
template <int NC, int SUB>
struct CuCfg {
  static constexpr int VLD = SUB + 2;             // row stride of V_s / C_s: 8-byte aligned, lane stride 2 banks
  static constexpr int XLD = NC + 4;              // row stride of X_s
  static constexpr int STAGE = (CU_KB + NC) * VLD;  // floats of one {V_s, C_s} stage
  static constexpr size_t SMEM = (size_t)(2 * STAGE + CU_KB * XLD) * sizeof(float);
  static constexpr int RP = SUB / 2;              // row pairs per sub-tile
  static constexpr int CL = CU_THREADS / RP;      // column lanes of the loader
  static constexpr int NJ = NC / 32;              // phase 1: columns per thread
  static constexpr int RA = SUB / 32;             // phase 3: row pairs per thread
  static constexpr int NJ3 = NC / 32;             // phase 3: columns per thread (pairs 2 tc + 64 j + {0, 1})
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 8-byte asynchronous global -> shared copy; bytes beyond `src_bytes` (0, 4 or 8) are zero-filled.
__device__ __forceinline__ void cp_async8(float* dst, const float* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// All CTAs of the (co-resident) grid meet; `target` is the counter value that means "everybody of this round arrived".
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu_u32(counter) - target) < 0) {
      if (clock64() - t0 > (3ll << 30)) { *err = 1; break; }   // ~1.6 s: the grid was not co-resident; results are void
    }
    __threadfence();
  }
  __syncthreads();
}

template <int NC, int SUB>
__global__ void __launch_bounds__(CU_THREADS, 1) chain_update_kernel(ChainUpdParams p) {
  using Cfg = CuCfg<NC, SUB>;
  constexpr int VLD = Cfg::VLD, XLD = Cfg::XLD, STAGE = Cfg::STAGE;
  extern __shared__ __align__(16) float cu_smem[];
  float* Xs = cu_smem + 2 * STAGE;              // [64][XLD] X, X_s[k][c]            (phase 3)
  float* Ts = cu_smem;                          // [64][65]  op(T)                   (phase 2, over stage 0)
  float* Ws = cu_smem + STAGE;                  // column sums                       (phase 2, over stage 1)
  const int tid = threadIdx.x;
  const int P = gridDim.x, pid = blockIdx.x;
  const int kb = p.kb, nc = p.nc;
  const long long rows_per = ((p.mp + P - 1) / P + SUB - 1) / SUB * SUB;
  const long long row_lo = (long long)pid * rows_per;
  const long long row_hi = row_lo + rows_per < p.mp ? row_lo + rows_per : p.mp;
  const int nt = row_hi > row_lo ? (int)((row_hi - row_lo + SUB - 1) / SUB) : 0;

  // V_s[k][r] = V(rb + r, k), C_s[c][r] = C(rb + r, c); rows beyond the slab, k >= kb and c >= nc are zero-filled.
  auto fetch = [&](int t, bool with_c) {
    float* Vs = cu_smem + (t & 1) * STAGE;
    float* Cs = Vs + CU_KB * VLD;
    const long long rb = row_lo + (long long)t * SUB;
    const int r2 = 2 * (tid % Cfg::RP), col0 = tid / Cfg::RP;
    const long long r = rb + r2;
    const int rbytes = r + 1 < row_hi ? 8 : (r < row_hi ? 4 : 0);
    const long long rs = r < row_hi ? r : row_lo;                // a valid address when nothing is read
#pragma unroll
    for (int i = 0; i < CU_KB / Cfg::CL; ++i) {
      const int k = col0 + Cfg::CL * i;
      const bool in = k < kb;
      cp_async8(Vs + k * VLD + r2, p.v + (in ? rs + (long long)k * p.ldv : row_lo), in ? rbytes : 0);
    }
    if (with_c) {
#pragma unroll
      for (int i = 0; i < NC / Cfg::CL; ++i) {
        const int c = col0 + Cfg::CL * i;
        const bool in = c < nc;
        cp_async8(Cs + c * VLD + r2, p.c + (in ? rs + (long long)c * p.ldc : row_lo), in ? rbytes : 0);
      }
    }
    cp_async_commit();
  };

  // ---- phase 1: this slab's share of W = V^T C.  Thread (tk, tc): k = tk + 16 i (i < 4), c = tc + 32 j (j < NC / 32).
  {
    const int tk = tid & 15, tc = tid >> 4;
    f32x2 acc[4][Cfg::NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NJ; ++j) acc[i][j] = 0ull;
    if (nt > 0) fetch(0, true);
    for (int t = 0; t < nt; ++t) {
      if (t + 1 < nt) { fetch(t + 1, true); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();
      const float* Vs = cu_smem + (t & 1) * STAGE;
      const float* Cs = Vs + CU_KB * VLD;
#pragma unroll 4
      for (int rp = 0; rp < Cfg::RP; ++rp) {
        f32x2 v2[4], c2[Cfg::NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) v2[i] = *reinterpret_cast<const f32x2*>(Vs + (tk + 16 * i) * VLD + 2 * rp);
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) c2[j] = *reinterpret_cast<const f32x2*>(Cs + (tc + 32 * j) * VLD + 2 * rp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < Cfg::NJ; ++j) acc[i][j] = ffma2(v2[i], c2[j], acc[i][j]);
      }
      __syncthreads();                            // the stage is refilled two iterations later
    }
    float* wp = p.wpart + (long long)pid * (CU_KB * NC);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NJ; ++j) wp[(tk + 16 * i) + CU_KB * (tc + 32 * j)] = fsum2(acc[i][j]);
  }
  grid_barrier(p.bar, p.bar_base + (unsigned)P, p.err);

  // ---- phase 2: X(:, my columns) = op(T) sum_q W_q(:, my columns)
  {
    const int cper = (nc + P - 1) / P;
    const int c_lo = pid * cper, c_hi = min(nc, c_lo + cper);
    const int nval = (c_hi - c_lo) * CU_KB;
    if (nval > 0) {
      for (int idx = tid; idx < CU_KB * CU_KB; idx += CU_THREADS) {   // Ts[i][k'] = op(T)(k', i): X(k') = sum_i Ts[i][k'] W(i)
        const int i = idx & 63, k2 = idx >> 6;
        float tv = 0.f;
        if (i < kb && k2 < kb) tv = p.trans ? p.t[i + (long long)k2 * p.ldt] : p.t[k2 + (long long)i * p.ldt];
        Ts[i * (CU_KB + 1) + k2] = tv;
      }
      // the P partial sums of a value are split over `np` threads (fixed by the launch shape: the order of the additions,
      // and with it the result, does not depend on timing)
      int np = CU_THREADS / nval;
      if (np > 8) np = 8;
      if (np * nval > STAGE) np = STAGE / nval;
      if (np < 1) np = 1;
      for (int idx = tid; idx < nval * np; idx += CU_THREADS) {
        const int v = idx % nval, part = idx / nval;
        const float* src = p.wpart + (v & 63) + CU_KB * (c_lo + (v >> 6));
        float s0 = 0.f;
#pragma unroll 4
        for (int q = part; q < P; q += np) s0 += __ldcg(src + (long long)q * (CU_KB * NC));
        Ws[part * nval + v] = s0;
      }
      __syncthreads();
      for (int v = tid; v < nval; v += CU_THREADS) {
        float s0 = Ws[v];
        for (int part = 1; part < np; ++part) s0 += Ws[part * nval + v];
        Ws[v] = s0;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nval; idx += CU_THREADS) {
      const int k2 = idx & 63, cl = idx >> 6;
      float s0 = 0.f;
#pragma unroll 8
      for (int i = 0; i < CU_KB; ++i) s0 = fmaf(Ts[i * (CU_KB + 1) + k2], Ws[cl * CU_KB + i], s0);
      p.x[k2 + CU_KB * (c_lo + cl)] = s0;
    }
    __syncthreads();                              // Ts / Ws live in the stages phase 3 refills
  }
  if (nt > 0) fetch(0, false);                    // V of the first sub-tile travels during the barrier
  grid_barrier(p.bar, p.bar_base + 2u * (unsigned)P, p.err);

  // ---- phase 3: C -= V X on this slab.  Thread (rt, tc): row pairs 2 rt + 32 a (a < SUB / 32), columns 2 tc + 64 j + {0, 1}.
  {
    for (int idx = tid; idx < CU_KB * NC; idx += CU_THREADS) {
      const int k = idx & 63, c = idx >> 6;
      Xs[k * XLD + c] = (c < nc) ? __ldcg(p.x + k + CU_KB * c) : 0.f;
    }
    const int rt = tid & 15, tc = tid >> 4;
    constexpr int RA = Cfg::RA, NJ3 = Cfg::NJ3;
    for (int t = 0; t < nt; ++t) {
      if (t + 1 < nt) { fetch(t + 1, false); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();
      const float* Vs = cu_smem + (t & 1) * STAGE;
      const long long rb = row_lo + (long long)t * SUB;
      // the C values this thread rewrites: issued before the products so the L2 round trip hides under them
      float2 old[RA][NJ3];
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const long long r = rb + 32 * a + 2 * rt;
#pragma unroll
        for (int j = 0; j < NJ3; ++j) {
          const int c = 2 * tc + 64 * (j >> 1) + (j & 1);
          old[a][j] = make_float2(0.f, 0.f);
          if (c < nc && r + 1 < row_hi) old[a][j] = __ldcg(reinterpret_cast<const float2*>(p.c + r + (long long)c * p.ldc));
          else if (c < nc && r < row_hi) old[a][j].x = __ldcg(p.c + r + (long long)c * p.ldc);
        }
      }
      f32x2 acc[RA][NJ3];
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int j = 0; j < NJ3; ++j) acc[a][j] = 0ull;
#pragma unroll 4
      for (int k = 0; k < CU_KB; ++k) {
        f32x2 va[RA];
#pragma unroll
        for (int a = 0; a < RA; ++a) va[a] = *reinterpret_cast<const f32x2*>(Vs + k * VLD + 32 * a + 2 * rt);
#pragma unroll
        for (int j = 0; j < NJ3 / 2; ++j) {
          const float2 xx = *reinterpret_cast<const float2*>(Xs + k * XLD + 2 * tc + 64 * j);
          const f32x2 x0 = fpack2(xx.x, xx.x), x1 = fpack2(xx.y, xx.y);
#pragma unroll
          for (int a = 0; a < RA; ++a) {
            acc[a][2 * j] = ffma2(va[a], x0, acc[a][2 * j]);
            acc[a][2 * j + 1] = ffma2(va[a], x1, acc[a][2 * j + 1]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const long long r = rb + 32 * a + 2 * rt;
#pragma unroll
        for (int j = 0; j < NJ3; ++j) {
          const int c = 2 * tc + 64 * (j >> 1) + (j & 1);
          if (c < nc && r < row_hi) {
            float lo, hi;
            funpack2(acc[a][j], lo, hi);
            float* cp = p.c + r + (long long)c * p.ldc;
            if (r + 1 < row_hi) *reinterpret_cast<float2*>(cp) = make_float2(old[a][j].x - lo, old[a][j].y - hi);
            else cp[0] = old[a][j].x - lo;
          }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace

bool chain_update_fits(int kb, int nc, const float* v, long long ldv, const float* c, long long ldc) {
  return kb >= 1 && kb <= CU_KB && nc >= 1 && nc <= 256 && ldc % 2 == 0 && ldv % 2 == 0 &&
         (reinterpret_cast<uintptr_t>(c) & 7) == 0 && (reinterpret_cast<uintptr_t>(v) & 7) == 0;
}

// Floats of partial-sum scratch one CTA needs for a C of nc columns (the kernel variant is chosen by nc).
long long chain_update_part_floats(int nc) { return (long long)CU_KB * (nc <= 192 ? 192 : 256); }

// Fewest CTAs for which phase 2's column sums fit the staging buffer.
int chain_update_min_ctas(int nc) {
  const int stage = nc <= 192 ? CuCfg<192, 64>::STAGE : CuCfg<256, 32>::STAGE;
  const int cper_max = stage / CU_KB;
  return (nc + cper_max - 1) / cper_max;
}

// wpart: ctas * chain_update_part_floats(nc) floats, x: 64 * nc floats (leading dimension 64); bar: one device counter shared
// by all launches of the context, bar_base its value when this launch starts (the caller advances its copy by 2 * ctas).
template <int NC, int SUB>
static void launch_cu(const ChainUpdParams& p, int ctas, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(chain_update_kernel<NC, SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CuCfg<NC, SUB>::SMEM);
  // Cooperative launch: the grid is gang-scheduled, so CTAs never spin at the grid barrier while the rest of the grid
  // waits for SMs that another context's kernel holds (two contexts on one device share the SMs of their partitions).
  static const bool coop = !(getenv("CQR_CHAIN_COOP") && atoi(getenv("CQR_CHAIN_COOP")) == 0);
  if (coop) {
    ChainUpdParams q = p;
    void* args[] = {&q};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)chain_update_kernel<NC, SUB>, dim3(ctas), dim3(CU_THREADS), args, CuCfg<NC, SUB>::SMEM, s);
    static const bool trace = getenv("CQR_CHAIN_COOP_TRACE") != nullptr;
    if (trace) { static int shown = 0; if (shown++ < 4) fprintf(stderr, "chain_update<%d,%d> cooperative launch, %d CTAs: %s\n", NC, SUB, ctas, cudaGetErrorString(e)); }
    if (e == cudaSuccess) return;
    cudaGetLastError();   // not supported here (or the grid does not fit): plain launch, the barrier's timeout still guards it
  }
  chain_update_kernel<NC, SUB><<<ctas, CU_THREADS, CuCfg<NC, SUB>::SMEM, s>>>(p);
}

void launch_chain_update(const ChainUpdParams& p, int ctas, cudaStream_t s) {
  ++g_launches;
  if (p.nc <= 192) launch_cu<192, 64>(p, ctas, s);
  else launch_cu<256, 32>(p, ctas, s);
}

}  // namespace cqr
