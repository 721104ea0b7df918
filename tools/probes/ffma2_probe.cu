// ffma2_probe.cu -- issue rate of FFMA2 / FFMA / FADD / SHFL on one SM sub-partition as a function of resident warps
// and independent chains per warp.   nvcc -arch=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE, int CH>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
  f32x2 acc[CH]; float fa[CH];
  for (int i = 0; i < CH; ++i) { acc[i] = (unsigned long long)(threadIdx.x + i); fa[i] = seed * (threadIdx.x + i); }
  f32x2 xr[CH], br[CH]; float fb[CH], fc[CH];
  for (int i = 0; i < CH; ++i) { xr[i] = (unsigned long long)(threadIdx.x * 3 + i); br[i] = (unsigned long long)(threadIdx.x * 7 + i); fb[i] = seed + i; fc[i] = seed - i; }
  f32x2 x = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed * 0.5f);
  float fx = seed;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 0) acc[i] = ffma2(x, acc[i], acc[i]);
        else if (MODE == 1) asm volatile("fma.rn.f32 %0, %1, %0, %0;" : "+f"(fa[i]) : "f"(fx));
        else if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa[i]) : "f"(fx));
        else if (MODE == 3) fa[i] = __shfl_xor_sync(0xffffffffu, fa[i], 8);
        else if (MODE == 4) acc[i] = ffma2(xr[i], br[(i + r) % CH], acc[i]);            // three distinct 64-bit sources (dot pattern)
        else if (MODE == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(fa[i]) : "f"(fb[i]), "f"(fc[(i + r) % CH]));
        else if (MODE == 6) br[i] = ffma2(x, xr[(i + r) % CH], br[i]);                  // update pattern: one shared scalar pair
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += fa[i] + __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)br[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int CH>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096, threads = 128 * warps_per_smsp;   // one CTA on one SM: warps spread over the 4 sub-partitions
  probe<MODE, CH><<<1, threads>>>(out, cyc, iters, 1.0001f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double inst_per_warp = (double)iters * 4 * CH;
  printf("%-6s chains %d warps/SMSP %d: %.2f clk per instr per warp, %.2f clk per instr per SMSP\n", name, CH, warps_per_smsp,
         c / inst_per_warp, c / (inst_per_warp * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 3, 4, 8}) {
    if (w == 1) { run<0, 1>("FFMA2", 1); run<0, 2>("FFMA2", 1); run<0, 4>("FFMA2", 1); run<0, 8>("FFMA2", 1); run<1, 1>("FFMA", 1); run<1, 8>("FFMA", 1); run<2, 8>("FADD", 1); run<3, 1>("SHFL", 1); run<3, 8>("SHFL", 1); }
    else if (w == 2) { run<0, 8>("FFMA2", 2); run<1, 8>("FFMA", 2); run<3, 8>("SHFL", 2); }
    else if (w == 3) { run<0, 8>("FFMA2", 3); run<1, 8>("FFMA", 3); }
    else if (w == 4) { run<0, 8>("FFMA2", 4); run<1, 8>("FFMA", 4); run<2, 8>("FADD", 4); run<3, 8>("SHFL", 4); }
    else { run<0, 8>("FFMA2", 8); run<1, 8>("FFMA", 8); run<4, 8>("FFMA2-3src", 8); run<5, 8>("FFMA-3src", 8); run<6, 8>("FFMA2-upd", 8);
           run<4, 8>("FFMA2-3src", 2); run<6, 8>("FFMA2-upd", 2); run<5, 8>("FFMA-3src", 2); }
  }
  return 0;
}
