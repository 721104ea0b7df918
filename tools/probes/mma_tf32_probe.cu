// mma_tf32_probe.cu -- issue rate and latency of the legacy warp-level tensor path on sm_100a:
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (SASS HMMA.1688.F32.TF32) as a function of resident warps per SM
// sub-partition and independent accumulator chains per warp; plus the same loop interleaved 1:1 with FFMA.
//   nvcc -arch=sm_100a -O3 -o mma_tf32_probe mma_tf32_probe.cu && ./mma_tf32_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: CH independent accumulator chains, fixed A/B.  MODE 1: same + one FFMA per mma.  MODE 2: the accumulator of
// mma i feeds the A operand of mma i+1 (the leaf's dependency pattern: dots -> update).
template <int MODE, int CH>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
  float d[CH][4];
  unsigned a[4], b[2];
  float f[CH];
  for (int i = 0; i < CH; ++i) { for (int k = 0; k < 4; ++k) d[i][k] = seed * (threadIdx.x + i + k); f[i] = seed + i; }
  for (int k = 0; k < 4; ++k) a[k] = __float_as_uint(seed * (k + 1) * 1e-3f);
  b[0] = __float_as_uint(seed * 1e-3f); b[1] = __float_as_uint(seed * 2e-3f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (MODE == 2) {
          unsigned aa[4];
          for (int k = 0; k < 4; ++k) aa[k] = __float_as_uint(d[(i + CH - 1) % CH][k]);
          mma_tf32(d[i], aa, b);
        } else {
          mma_tf32(d[i], a, b);
        }
        if (MODE == 1) asm volatile("fma.rn.f32 %0, %1, %0, %0;" : "+f"(f[i]) : "f"(seed));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int CH>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2048, threads = 128 * warps_per_smsp;
  probe<MODE, CH><<<1, threads>>>(out, cyc, iters, 1.0001f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_warp = (double)iters * 4 * CH;
  printf("%-10s chains %d warps/SMSP %d: %6.2f clk per mma per warp, %6.2f clk per mma per SMSP  (%.0f MAC/clk/SM)\n", name, CH,
         warps_per_smsp, c / per_warp, c / (per_warp * warps_per_smsp), 1024.0 * 4 * per_warp * warps_per_smsp / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 1>("mma", 1); run<0, 2>("mma", 1); run<0, 4>("mma", 1); run<0, 8>("mma", 1);
  run<0, 1>("mma", 2); run<0, 4>("mma", 2); run<0, 8>("mma", 2);
  run<0, 4>("mma", 3); run<0, 8>("mma", 3); run<0, 8>("mma", 4);
  run<1, 4>("mma+ffma", 1); run<1, 8>("mma+ffma", 2); run<1, 8>("mma+ffma", 3);
  run<2, 1>("mma-dep", 1); run<2, 2>("mma-dep", 1); run<2, 4>("mma-dep", 1); run<2, 4>("mma-dep", 2);
  return 0;
}
