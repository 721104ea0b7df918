// gemm_umma.cu -- tcgen05 / TMEM / TMA 3xTF32 GEMMs (placeholder until the kernels land).
#include "common.cuh"
namespace cqr {
bool umma_available() { return false; }
bool launch_gemm_tn_umma(int, int, int, const float*, const float*, long long, const float*, const float*, long long,
                         float*, long long, int, long long, cudaStream_t) { return false; }
bool launch_gemm_nn_umma(int, int, int, float, const float*, const float*, long long, const float*, const float*,
                         long long, float, float*, long long, float*, long long, cudaStream_t) { return false; }
}  // namespace cqr
