// compare_cusolver.cu -- the comparator slot of the reference's command line.  The reference can time MAGMA's
// magma_sgeqrf2_gpu next to its own mmqr (qr.cu:555-565 magmaQR, 790-806: timing only, the result is never compared,
// compiled out by default).  MAGMA is not on this box; cuSOLVER's geqrf is, so the slot is filled with it -- loaded with
// dlopen at run time, never linked, never on the hot path, and absent libraries just report CQR_EUNSUPPORTED.
#include <cuda_runtime.h>
#include <dlfcn.h>
#pragma GCC visibility push(default)
#include "../../include/cudaqr_b200.h"
#pragma GCC visibility pop

namespace {
typedef void* Handle;
typedef int (*CreateFn)(Handle*);
typedef int (*DestroyFn)(Handle);
typedef int (*BufFn)(Handle, int, int, float*, int, int*);
typedef int (*GeqrfFn)(Handle, int, int, float*, int, float*, float*, int, int*);
struct Api { CreateFn create = nullptr; DestroyFn destroy = nullptr; BufFn buf = nullptr; GeqrfFn geqrf = nullptr; bool ok = false; };
Api& api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = nullptr;
    for (const char* name : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"})
      if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
    if (h) {
      a.create = (CreateFn)dlsym(h, "cusolverDnCreate");
      a.destroy = (DestroyFn)dlsym(h, "cusolverDnDestroy");
      a.buf = (BufFn)dlsym(h, "cusolverDnSgeqrf_bufferSize");
      a.geqrf = (GeqrfFn)dlsym(h, "cusolverDnSgeqrf");
      a.ok = a.create && a.destroy && a.buf && a.geqrf;
    }
  }
  return a;
}
}  // namespace

// Same shape as the reference's magmaQR (qr.cu:555-565): host matrix up, library geqrf on the device copy, matrix (and
// here tau[0..n)) back down.  Blocking.  Returns 0, CQR_EUNSUPPORTED without the library, or a CUDA / cuSOLVER status.
extern "C" int cqr_compare_cusolver_sgeqrf(float* mat, float* tau, int m, int n) {
  if (!mat || !tau || m < 1 || n < 1 || m < n) return CQR_EINVAL;
  Api& a = api();
  if (!a.ok) return CQR_EUNSUPPORTED;
  static Handle h = nullptr;
  if (!h && a.create(&h) != 0) return CQR_EUNSUPPORTED;
  float *dA = nullptr, *dtau = nullptr, *work = nullptr;
  int* info = nullptr;
  int lwork = 0, rc = 0;
  cudaError_t e;
  if ((e = cudaMalloc((void**)&dA, (size_t)m * n * sizeof(float))) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc((void**)&dtau, (size_t)n * sizeof(float))) != cudaSuccess) { cudaFree(dA); return (int)e; }
  cudaMalloc((void**)&info, sizeof(int));
  e = cudaMemcpy(dA, mat, (size_t)m * n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) rc = a.buf(h, m, n, dA, m, &lwork);
  if (e == cudaSuccess && rc == 0) e = cudaMalloc((void**)&work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(float));
  if (e == cudaSuccess && rc == 0) rc = a.geqrf(h, m, n, dA, m, dtau, work, lwork, info);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(mat, dA, (size_t)m * n * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(tau, dtau, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dtau); cudaFree(work); cudaFree(info);
  if (e != cudaSuccess) return (int)e;
  return rc ? 10000 + rc : 0;
}
