// Probe: can a 16-SM green context host a 16-CTA cluster while the remaining SMs run a persistent kernel?
// build: nvcc -gencode arch=compute_100a,code=sm_100a green_ctx_probe.cu -o green_ctx_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { const char* s; cuGetErrorString(r, &s); printf("%s failed: %s\n", #x, s); return 1; } } while (0)
#define RK(x) do { cudaError_t r = (x); if (r != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r)); return 1; } } while (0)

__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void spin_kernel(unsigned* sm, unsigned long long* t0, unsigned long long* t1, long long ns) {
  extern __shared__ char sm_[];
  if (threadIdx.x == 0) {
    sm[blockIdx.x] = smid();
    unsigned long long a = gtime();
    t0[blockIdx.x] = a;
    while ((long long)(gtime() - a) < ns) { }
    t1[blockIdx.x] = gtime();
  }
}

int main() {
  RK(cudaSetDevice(0));
  RK(cudaFree(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  CUdevResource all; CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("total SMs %u\n", all.sm.smCount);
  for (unsigned flags : {0u, (unsigned)CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE}) {
    for (unsigned want : {16u, 32u}) {
      for (unsigned ng : {1u, 2u}) {
        CUdevResource res[2], rem; unsigned n = ng;
        CUresult r = cuDevSmResourceSplitByCount(res, &n, &all, &rem, flags, want);
        if (r != CUDA_SUCCESS) { printf("split flags %u want %u groups %u: error %d\n", flags, want, ng, (int)r); continue; }
        printf("split flags %u want %u groups %u -> n=%u sizes %u %u remaining %u\n", flags, want, ng, n, res[0].sm.smCount, n > 1 ? res[1].sm.smCount : 0, rem.sm.smCount);
      }
    }
  }
  unsigned *d_sm; unsigned long long *d_t0, *d_t1;
  RK(cudaMalloc(&d_sm, 4096 * 4)); RK(cudaMalloc(&d_t0, 4096 * 8)); RK(cudaMalloc(&d_t1, 4096 * 8));
  RK(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  RK(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (unsigned flags : {0u, (unsigned)CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE}) {
    for (unsigned ng : {1u, 2u}) {
      CUdevResource res[2], rem; unsigned n = ng;
      if (cuDevSmResourceSplitByCount(res, &n, &all, &rem, flags, 16) != CUDA_SUCCESS || n < ng) { printf("skip flags %u ng %u\n", flags, ng); continue; }
      CUdevResourceDesc dA, dB; CUgreenCtx gA, gB; CUstream sA, sB;
      CK(cuDevResourceGenerateDesc(&dA, res, ng));
      CK(cuDevResourceGenerateDesc(&dB, &rem, 1));
      CK(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
      CK(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
      CK(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
      CK(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
      const unsigned nB = rem.sm.smCount;
      RK(cudaMemset(d_sm, 0xff, 4096 * 4));
      // B: persistent-like kernel on the remaining SMs (1 CTA/SM through 190 KB smem), 3 ms
      spin_kernel<<<nB, 448, 190 * 1024, (cudaStream_t)sB>>>(d_sm, d_t0, d_t1, 3000000);
      RK(cudaGetLastError());
      // A: ng clusters of 16 CTAs x 512 threads, 1 ms, launched while B runs; then a second one right after
      for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16 * ng); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 40 * 1024; cfg.stream = (cudaStream_t)sA;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, spin_kernel, d_sm + 1024 * (rep + 1), d_t0 + 1024 * (rep + 1), d_t1 + 1024 * (rep + 1), (long long)1000000);
        if (e != cudaSuccess) { printf("cluster launch (flags %u ng %u) failed: %s\n", flags, ng, cudaGetErrorString(e)); cudaGetLastError(); }
      }
      RK(cudaDeviceSynchronize());
      std::vector<unsigned> sm(4096); std::vector<unsigned long long> t0(4096), t1(4096);
      RK(cudaMemcpy(sm.data(), d_sm, 4096 * 4, cudaMemcpyDeviceToHost));
      RK(cudaMemcpy(t0.data(), d_t0, 4096 * 8, cudaMemcpyDeviceToHost));
      RK(cudaMemcpy(t1.data(), d_t1, 4096 * 8, cudaMemcpyDeviceToHost));
      unsigned long long b0 = ~0ull, b1 = 0; std::vector<int> used(256, 0);
      for (unsigned i = 0; i < nB; ++i) { b0 = b0 < t0[i] ? b0 : t0[i]; b1 = b1 > t1[i] ? b1 : t1[i]; used[sm[i] & 255]++; }
      int multi = 0; for (int u : used) multi += u > 1;
      printf("flags %u ng %u: B %u CTAs on SMs (dups %d), ran %.3f ms\n", flags, ng, nB, multi, (b1 - b0) * 1e-6);
      for (int rep = 0; rep < 2; ++rep) {
        unsigned long long a0 = ~0ull, a1 = 0; int overlap = 0;
        printf("   A rep %d smids:", rep);
        for (unsigned i = 0; i < 16 * ng; ++i) {
          unsigned k = 1024 * (rep + 1) + i;
          a0 = a0 < t0[k] ? a0 : t0[k]; a1 = a1 > t1[k] ? a1 : t1[k];
          printf(" %u", sm[k]); if (sm[k] < 256 && used[sm[k]]) overlap++;
        }
        printf("\n   A rep %d: start %+.3f ms after B start, end %+.3f ms (B end %+.3f), SMs shared with B: %d\n", rep, ((double)a0 - (double)b0) * 1e-6, ((double)a1 - (double)b0) * 1e-6, (b1 - b0) * 1e-6, overlap);
      }
      cuStreamDestroy(sA); cuStreamDestroy(sB); cuGreenCtxDestroy(gA); cuGreenCtxDestroy(gB);
    }
  }
  return 0;
}
