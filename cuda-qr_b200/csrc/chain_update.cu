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
#include "common.cuh"
#include "warp_math.cuh"

namespace cqr {

namespace {

constexpr int CU_KB = 64, CU_NC = 192, CU_THREADS = 512, CU_SUB = 64;
constexpr int CU_VLD = CU_SUB + 2;    // 66: rows of V_s / C_s (8-byte aligned, lane stride 2 banks)
constexpr int CU_XLD = CU_NC + 4;     // 196: rows of X_s
constexpr int CU_STAGE = (CU_KB + CU_NC) * CU_VLD;   // floats of one {V_s, C_s} stage
constexpr size_t kChainUpdSmem = (size_t)(2 * CU_STAGE + CU_KB * CU_XLD) * sizeof(float);

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

__global__ void __launch_bounds__(CU_THREADS, 1) chain_update_kernel(ChainUpdParams p) {
  extern __shared__ __align__(16) float cu_smem[];
  float* Xs = cu_smem + 2 * CU_STAGE;           // [64][196] X, X_s[k][c]            (phase 3)
  float* Ts = cu_smem;                          // [64][65]  op(T)                   (phase 2, over stage 0)
  float* Ws = cu_smem + CU_STAGE;               // [cper][64] column sums            (phase 2, over stage 1)
  const int tid = threadIdx.x;
  const int P = gridDim.x, pid = blockIdx.x;
  const int kb = p.kb, nc = p.nc;
  const long long rows_per = ((p.mp + P - 1) / P + CU_SUB - 1) / CU_SUB * CU_SUB;
  const long long row_lo = (long long)pid * rows_per;
  const long long row_hi = row_lo + rows_per < p.mp ? row_lo + rows_per : p.mp;
  const int nt = row_hi > row_lo ? (int)((row_hi - row_lo + CU_SUB - 1) / CU_SUB) : 0;

  // V_s[k][r] = V(rb + r, k), C_s[c][r] = C(rb + r, c); rows beyond the slab, k >= kb and c >= nc are zero-filled.
  auto fetch = [&](int t, bool with_c) {
    float* Vs = cu_smem + (t & 1) * CU_STAGE;
    float* Cs = Vs + CU_KB * CU_VLD;
    const long long rb = row_lo + (long long)t * CU_SUB;
    const int r2 = 2 * (tid & 31), col0 = tid >> 5;              // 32 row pairs x 16 columns per pass
    const long long r = rb + r2;
    const int rbytes = r + 1 < row_hi ? 8 : (r < row_hi ? 4 : 0);
    const long long rs = r < row_hi ? r : row_lo;                // a valid address when nothing is read
#pragma unroll
    for (int i = 0; i < CU_KB / 16; ++i) {
      const int k = col0 + 16 * i;
      const bool in = k < kb;
      cp_async8(Vs + k * CU_VLD + r2, p.v + (in ? rs + (long long)k * p.ldv : row_lo), in ? rbytes : 0);
    }
    if (with_c) {
#pragma unroll
      for (int i = 0; i < CU_NC / 16; ++i) {
        const int c = col0 + 16 * i;
        const bool in = c < nc;
        cp_async8(Cs + c * CU_VLD + r2, p.c + (in ? rs + (long long)c * p.ldc : row_lo), in ? rbytes : 0);
      }
    }
    cp_async_commit();
  };

  // ---- phase 1: this slab's share of W = V^T C.  Thread (tk, tc): k = tk + 16 i (i < 4), c = tc + 32 j (j < 6).
  {
    const int tk = tid & 15, tc = tid >> 4;
    f32x2 acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[i][j] = 0ull;
    if (nt > 0) fetch(0, true);
    for (int t = 0; t < nt; ++t) {
      if (t + 1 < nt) { fetch(t + 1, true); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();
      const float* Vs = cu_smem + (t & 1) * CU_STAGE;
      const float* Cs = Vs + CU_KB * CU_VLD;
#pragma unroll 4
      for (int rp = 0; rp < CU_SUB / 2; ++rp) {
        f32x2 v2[4], c2[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) v2[i] = *reinterpret_cast<const f32x2*>(Vs + (tk + 16 * i) * CU_VLD + 2 * rp);
#pragma unroll
        for (int j = 0; j < 6; ++j) c2[j] = *reinterpret_cast<const f32x2*>(Cs + (tc + 32 * j) * CU_VLD + 2 * rp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) acc[i][j] = ffma2(v2[i], c2[j], acc[i][j]);
      }
      __syncthreads();                            // the stage is refilled two iterations later
    }
    float* wp = p.wpart + (long long)pid * (CU_KB * CU_NC);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) wp[(tk + 16 * i) + CU_KB * (tc + 32 * j)] = fsum2(acc[i][j]);
  }
  grid_barrier(p.bar, p.bar_base + (unsigned)P, p.err);

  // ---- phase 2: X(:, my columns) = op(T) sum_q W_q(:, my columns)
  {
    const int cper = (nc + P - 1) / P;            // <= 192 columns, P >= 1: at most 192 x 64 sums fit a stage
    const int c_lo = pid * cper, c_hi = min(nc, c_lo + cper);
    if (c_hi > c_lo) {
      for (int idx = tid; idx < CU_KB * CU_KB; idx += CU_THREADS) {   // Ts[i][k'] = op(T)(k', i): X(k') = sum_i Ts[i][k'] W(i)
        const int i = idx & 63, k2 = idx >> 6;
        float tv = 0.f;
        if (i < kb && k2 < kb) tv = p.trans ? p.t[i + (long long)k2 * p.ldt] : p.t[k2 + (long long)i * p.ldt];
        Ts[i * (CU_KB + 1) + k2] = tv;
      }
      for (int idx = tid; idx < (c_hi - c_lo) * CU_KB; idx += CU_THREADS) {
        const int k = idx & 63, c = c_lo + (idx >> 6);
        const float* src = p.wpart + k + CU_KB * c;
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < P; ++q) s += __ldcg(src + (long long)q * (CU_KB * CU_NC));
        Ws[(idx >> 6) * CU_KB + k] = s;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < (c_hi - c_lo) * CU_KB; idx += CU_THREADS) {
      const int k2 = idx & 63, cl = idx >> 6;
      float s = 0.f;
#pragma unroll 8
      for (int i = 0; i < CU_KB; ++i) s = fmaf(Ts[i * (CU_KB + 1) + k2], Ws[cl * CU_KB + i], s);
      p.x[k2 + CU_KB * (c_lo + cl)] = s;
    }
    __syncthreads();                              // Ts / Ws live in the stages phase 3 refills
  }
  if (nt > 0) fetch(0, false);                    // V of the first sub-tile travels during the barrier
  grid_barrier(p.bar, p.bar_base + 2u * (unsigned)P, p.err);

  // ---- phase 3: C -= V X on this slab.  Thread (rt, tc): rows {2 rt, 2 rt + 1, 32 + 2 rt, 33 + 2 rt}, c = 2 tc + 64 j + {0, 1}.
  {
    for (int idx = tid; idx < CU_KB * CU_NC; idx += CU_THREADS) {
      const int k = idx & 63, c = idx >> 6;
      Xs[k * CU_XLD + c] = (c < nc) ? __ldcg(p.x + k + CU_KB * c) : 0.f;
    }
    const int rt = tid & 15, tc = tid >> 4;
    for (int t = 0; t < nt; ++t) {
      if (t + 1 < nt) { fetch(t + 1, false); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();
      const float* Vs = cu_smem + (t & 1) * CU_STAGE;
      const long long rb = row_lo + (long long)t * CU_SUB;
      // the C values this thread rewrites: issued before the products so the L2 round trip hides under them
      float2 old[2][6];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const long long r = rb + 32 * a + 2 * rt;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int c = 2 * tc + 64 * (j >> 1) + (j & 1);
          old[a][j] = make_float2(0.f, 0.f);
          if (c < nc && r + 1 < row_hi) old[a][j] = __ldcg(reinterpret_cast<const float2*>(p.c + r + (long long)c * p.ldc));
          else if (c < nc && r < row_hi) old[a][j].x = __ldcg(p.c + r + (long long)c * p.ldc);
        }
      }
      f32x2 acc[2][6];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[a][j] = 0ull;
#pragma unroll 4
      for (int k = 0; k < CU_KB; ++k) {
        const f32x2 va = *reinterpret_cast<const f32x2*>(Vs + k * CU_VLD + 2 * rt);
        const f32x2 vb = *reinterpret_cast<const f32x2*>(Vs + k * CU_VLD + 32 + 2 * rt);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float2 xx = *reinterpret_cast<const float2*>(Xs + k * CU_XLD + 2 * tc + 64 * j);
          const f32x2 x0 = fpack2(xx.x, xx.x), x1 = fpack2(xx.y, xx.y);
          acc[0][2 * j] = ffma2(va, x0, acc[0][2 * j]);
          acc[0][2 * j + 1] = ffma2(va, x1, acc[0][2 * j + 1]);
          acc[1][2 * j] = ffma2(vb, x0, acc[1][2 * j]);
          acc[1][2 * j + 1] = ffma2(vb, x1, acc[1][2 * j + 1]);
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const long long r = rb + 32 * a + 2 * rt;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
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
  return kb >= 1 && kb <= CU_KB && nc >= 1 && nc <= CU_NC && ldc % 2 == 0 && ldv % 2 == 0 &&
         (reinterpret_cast<uintptr_t>(c) & 7) == 0 && (reinterpret_cast<uintptr_t>(v) & 7) == 0;
}

// wpart: ctas * 64 * 192 floats, x: 64 * 192 floats; bar: one device counter shared by all launches of the context,
// bar_base its value when this launch starts (the caller advances its copy by 2 * ctas).
void launch_chain_update(const ChainUpdParams& p, int ctas, cudaStream_t s) {
  ++g_launches;
  static PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(chain_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainUpdSmem);
  chain_update_kernel<<<ctas, CU_THREADS, kChainUpdSmem, s>>>(p);
}

}  // namespace cqr
