// Probe: cycles per all-reduce round of 64 floats among the 16 CTAs (512 threads each) of one cluster, for several
// DSMEM exchange schemes.  Each round's input depends on the previous round's total (no overlap between rounds).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 dsmem_probe.cu -o dsmem_probe
#include <cuda_runtime.h>
#include <cstdio>
#define RK(x) do { cudaError_t r = (x); if (r != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r)); return 1; } } while (0)
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ unsigned ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa(const void* p, unsigned rank) {
  unsigned la = (unsigned)__cvta_generic_to_shared(p), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_async_f32(unsigned ra, unsigned rb, float v) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned ra, unsigned rb, float4 v) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rb) : "memory");
}
__device__ __forceinline__ void bulk_s2c(unsigned dst_cluster, unsigned src_cta, unsigned bytes, unsigned rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok, a = (unsigned)__cvta_generic_to_shared(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}

// mode 0: 4-byte st.async all-gather (thread (c,g) pushes to peers g, g+8)
// mode 1: 16-byte st.async all-gather (lane i of warp w pushes the warp's 4 sums to peer i)
// mode 2: stage 64 sums in local smem, __syncthreads, 16 bulk copies of 256 B (one per peer)
// mode 3: reduce-scatter + all-gather, 16-byte st.async: CTA r reduces columns 4r..4r+3
// mode 4: 16-byte plain remote stores + cluster barrier
// mode 5: no exchange at all (loop overhead reference)
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* cyc, int rounds) {
  __shared__ __align__(16) float inbox[2][16][64];
  __shared__ __align__(16) float stage[2][64];
  __shared__ __align__(16) float rs_in[2][16][4];
  __shared__ __align__(16) float tot_in[2][64];
  __shared__ unsigned long long mbar[2], mbar2[2];
  const unsigned rank = ctarank(), CS = nctarank();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, c = threadIdx.x >> 3, g = l & 7;
  if (threadIdx.x == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init(&mbar2[0], 1); mbar_init(&mbar2[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  csync();
  float s = 1.0f + 0.001f * threadIdx.x + rank;
  long long t0 = clock64();
  for (int j = 0; j < rounds; ++j) {
    const int buf = j & 1;
    const unsigned par = (j >> 1) & 1;
    // stand-in for the slab dot: 3-stage shuffle reduction over the 8 row groups
    s += __shfl_xor_sync(kFull, s, 1); s += __shfl_xor_sync(kFull, s, 2); s += __shfl_xor_sync(kFull, s, 4);
    float total;
    if (MODE == 0) {
      if (threadIdx.x == 0) mbar_expect(&mbar[buf], CS * 64 * 4);
      for (unsigned i = g; i < CS; i += 8) st_async_f32(mapa(&inbox[buf][rank][c], i), mapa(&mbar[buf], i), s);
      mbar_wait(&mbar[buf], par);
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) t += inbox[buf][i][c];
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else if (MODE == 1) {
      if (threadIdx.x == 0) mbar_expect(&mbar[buf], CS * 64 * 4);
      float4 sv;
      sv.x = __shfl_sync(kFull, s, 0); sv.y = __shfl_sync(kFull, s, 8); sv.z = __shfl_sync(kFull, s, 16); sv.w = __shfl_sync(kFull, s, 24);
      if ((unsigned)l < CS) st_async_v4(mapa(&inbox[buf][rank][4 * w], l), mapa(&mbar[buf], l), sv);
      mbar_wait(&mbar[buf], par);
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) t += inbox[buf][i][c];
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else if (MODE == 2) {
      if (threadIdx.x == 0) mbar_expect(&mbar[buf], CS * 64 * 4);
      if (g == 0) stage[buf][c] = s;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (w == 0 && (unsigned)l < CS)
        bulk_s2c(mapa(&inbox[buf][rank][0], l), (unsigned)__cvta_generic_to_shared(&stage[buf][0]), 256, mapa(&mbar[buf], l));
      mbar_wait(&mbar[buf], par);
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) t += inbox[buf][i][c];
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else if (MODE == 3) {
      // phase 1: warp w sends its 4 sums (columns 4w..4w+3) to CTA w  (needs CS == 16 == #warps)
      if (threadIdx.x == 0) { mbar_expect(&mbar[buf], CS * 16); mbar_expect(&mbar2[buf], CS * 16); }
      float4 sv;
      sv.x = __shfl_sync(kFull, s, 0); sv.y = __shfl_sync(kFull, s, 8); sv.z = __shfl_sync(kFull, s, 16); sv.w = __shfl_sync(kFull, s, 24);
      if (l == 0) st_async_v4(mapa(&rs_in[buf][rank][0], w), mapa(&mbar[buf], w), sv);
      // phase 2 (warp 0 only): reduce the 16 contributions for my 4 columns, broadcast the totals to all peers
      if (w == 0) {
        mbar_wait(&mbar[buf], par);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)l < CS) t = *reinterpret_cast<const float4*>(&rs_in[buf][l][0]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
          t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
        }
        if ((unsigned)l < CS) st_async_v4(mapa(&tot_in[buf][4 * rank], l), mapa(&mbar2[buf], l), t);
      }
      mbar_wait(&mbar2[buf], par);
      total = tot_in[buf][c];
    } else if (MODE == 4) {
      float4 sv;
      sv.x = __shfl_sync(kFull, s, 0); sv.y = __shfl_sync(kFull, s, 8); sv.z = __shfl_sync(kFull, s, 16); sv.w = __shfl_sync(kFull, s, 24);
      if ((unsigned)l < CS) {
        unsigned ra = mapa(&inbox[buf][rank][4 * w], l);
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(sv.x), "f"(sv.y), "f"(sv.z), "f"(sv.w) : "memory");
      }
      csync();
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) t += inbox[buf][i][c];
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else if (MODE == 6) {
      // stage the 64 sums locally, then warp i sends all 256 B to peer i as one warp-wide instruction (16 lanes x 16 B)
      if (threadIdx.x == 0) mbar_expect(&mbar[buf], CS * 64 * 4);
      if (g == 0) stage[buf][c] = s;
      __syncthreads();
      if ((unsigned)w < CS && l < 16) {
        const float4 v = *reinterpret_cast<const float4*>(&stage[buf][4 * l]);
        st_async_v4(mapa(&inbox[buf][rank][4 * l], w), mapa(&mbar[buf], w), v);
      }
      mbar_wait(&mbar[buf], par);
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) t += inbox[buf][i][c];
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else if (MODE == 7) {
      // like 6 but the all-gather carries per-CTA totals of a reduce step done by remote LOADS: stage locally, cluster
      // barrier-free flag: each CTA pulls peers' staged sums with ld.shared::cluster after an mbarrier handshake
      if (threadIdx.x == 0) mbar_expect(&mbar[buf], CS * 4);
      if (g == 0) stage[buf][c] = s;
      __syncthreads();
      if (threadIdx.x < CS) st_async_f32(mapa(&rs_in[buf][rank][0], threadIdx.x), mapa(&mbar[buf], threadIdx.x), 1.f);   // "my stage is ready"
      mbar_wait(&mbar[buf], par);
      float t = 0.f;
      for (unsigned i = g; i < CS; i += 8) {
        float v; unsigned ra = mapa(&stage[buf][c], i);
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
        t += v;
      }
      t += __shfl_xor_sync(kFull, t, 1); t += __shfl_xor_sync(kFull, t, 2); t += __shfl_xor_sync(kFull, t, 4);
      total = t;
    } else {
      total = s * 1.0001f;
    }
    s = total * 0.01f + 0.5f;
  }
  long long t1 = clock64();
  csync();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 512 + threadIdx.x] = s;
}

template <int MODE>
int run(const char* name, float* d_out, long long* d_cyc, int cs) {
  RK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int rounds = 256;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(512); cfg.stream = 0;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RK(cudaLaunchKernelEx(&cfg, probe<MODE>, d_out, d_cyc, rounds));
    RK(cudaDeviceSynchronize());
  }
  long long cyc[16]; float out[4];
  RK(cudaMemcpy(cyc, d_cyc, sizeof(long long) * cs, cudaMemcpyDeviceToHost));
  RK(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
  long long mx = 0; for (int i = 0; i < cs; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
  printf("CS=%2d %-46s %7.1f cycles/round   (check %.6f)\n", cs, name, (double)mx / rounds, out[1]);
  return 0;
}

int main() {
  float* d_out; long long* d_cyc;
  RK(cudaMalloc(&d_out, 16 * 512 * 4)); RK(cudaMalloc(&d_cyc, 16 * 8));
  for (int cs : {16, 8, 2}) {
    run<5>("no exchange (loop overhead)", d_out, d_cyc, cs);
    run<0>("st.async 4 B all-gather", d_out, d_cyc, cs);
    run<1>("st.async 16 B all-gather", d_out, d_cyc, cs);
    run<2>("smem stage + 256 B bulk copy per peer", d_out, d_cyc, cs);
    if (cs == 16) run<3>("reduce-scatter + all-gather (16 B st.async)", d_out, d_cyc, cs);
    run<4>("16 B remote stores + cluster barrier", d_out, d_cyc, cs);
    run<6>("smem stage + one warp per peer (16 x 16 B contiguous)", d_out, d_cyc, cs);
    run<7>("smem stage + ready flags + remote loads", d_out, d_cyc, cs);
  }
  return 0;
}
