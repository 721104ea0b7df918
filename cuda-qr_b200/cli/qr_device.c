/* qr_device.c -- command-line driver with the reference's interface (`qr_device m n`, qr.cu:709-857) on top of
 * libcudaqr_b200.so.  Plain C99, no CUDA headers: it calls only the legacy entry points (getPanelDims, mmqr,
 * explicitQR, dgemm), exactly the calls the reference's own main() makes, so it doubles as the link-time drop-in
 * example of INTEGRATION.md.
 *
 * Behaviour kept from the reference: usage line and exit(1) without two sizes (qr.cu:715-719), the size rounding
 * m -> PR + k (PR - PC), n -> multiple of PC not above m with PR = 64, PC = 4 (qr.cu:722-734) and the
 * "Exact problem size" line, srand(12) / rand() uniform [0,1) input in column-major order (qr.cu:765-771), three timed
 * trials of the whole mmqr call with gettimeofday, input restored between trials outside the timing (qr.cu:774-788),
 * and the result line " MMQR ran QR on MxN matrix in T s (avg over 3)" (qr.cu:789).
 * The MAGMA comparator slot (qr.cu:790-806: the library's geqrf timed the same way, result not compared) is filled with
 * cuSOLVER's geqrf when a third or fourth argument "compare" is given and libcusolver is on the box.
 * Not kept: the shared-memory bank configuration (qr.cu:741-759, a Kepler setting).  Extras: QR_DEVICE_EXACT=1 skips the rounding (the library takes any m >= n);
 * a third argument "check" re-enables the reference's commented-out validation (qr.cu:822-850): explicit Q and R,
 * QR = Q*R with dgemm, and the Frobenius norm of QR - A. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "cudaqr_b200.h"

#define PR 64
#define PC 4
#define TRIALS 3

int main(int argc, const char** argv) {
  if (argc < 3) {
    puts("Usage: ./qr_device m n");
    exit(1);
  }
  int m = atoi(argv[1]);
  int n = atoi(argv[2]);
  int check = 0, compare = 0;
  for (int i = 3; i < argc; i++) {
    if (strcmp(argv[i], "check") == 0) check = 1;
    if (strcmp(argv[i], "compare") == 0) compare = 1;
  }
  const char* exact = getenv("QR_DEVICE_EXACT");
  if (!(exact && exact[0] == '1')) { /* make m, n fit the reference's window grid */
    int numPanels = (int)((double)(m - PR) / (PR - PC) + 0.5);
    m = PR + numPanels * (PR - PC);
    numPanels = (int)((double)n / PC + 0.5);
    if (numPanels == 0) numPanels = 1;
    n = numPanels * PC;
    while (n > m) n -= PC;
  }
  printf("Exact problem size: %dx%d\n", m, n);
  if (!(m > 0 && n > 0 && m >= n)) {
    puts("need m >= n >= 1");
    exit(1);
  }
  printf("Testing mmqr with %s\n", cqr_version());
  float* A = (float*)malloc((size_t)m * n * sizeof(float));
  float* RV = (float*)malloc((size_t)m * n * sizeof(float));
  int rowPanels, colPanels;
  getPanelDims(m, n, &rowPanels, &colPanels);
  float* tau = (float*)malloc((size_t)rowPanels * colPanels * PC * sizeof(float));
  if (!A || !RV || !tau) {
    puts("out of host memory");
    exit(1);
  }
  srand(12);
  for (size_t i = 0; i < (size_t)m * n; i++) A[i] = (float)rand() / RAND_MAX;
  memcpy(RV, A, (size_t)m * n * sizeof(float));
  double elapsed = 0;
  struct timeval cur, next;
  gettimeofday(&cur, NULL);
  for (int i = 0; i < TRIALS; i++) {
    mmqr(RV, tau, m, n);
    gettimeofday(&next, NULL);
    elapsed += (next.tv_sec + 1e-6 * next.tv_usec) - (cur.tv_sec + 1e-6 * cur.tv_usec);
    if (i != TRIALS - 1) memcpy(RV, A, (size_t)m * n * sizeof(float)); /* not part of the algorithm, not timed */
    gettimeofday(&cur, NULL);
  }
  printf(" MMQR ran QR on %dx%d matrix in %f s (avg over %d)\n", m, n, elapsed / TRIALS, TRIALS);
  const double flops = 2.0 * m * (double)n * n - 2.0 * (double)n * n * n / 3.0;
  printf("%.1f GFLOP/s (2mn^2 - 2n^3/3, host buffers: transfers included)\n", flops / (elapsed / TRIALS) / 1e9);
  if (compare) { /* qr.cu:790-806 with cuSOLVER in MAGMA's place: same loop, same timing, transfers included */
    float* CV = (float*)malloc((size_t)m * n * sizeof(float));
    float* ctau = (float*)malloc((size_t)n * sizeof(float));
    memcpy(CV, A, (size_t)m * n * sizeof(float));
    double cmpElapsed = 0;
    int rc = cqr_compare_cusolver_sgeqrf(CV, ctau, m, n); /* warm-up: library load, handle, workspace */
    memcpy(CV, A, (size_t)m * n * sizeof(float));
    gettimeofday(&cur, NULL);
    for (int i = 0; i < TRIALS && rc == 0; i++) {
      rc = cqr_compare_cusolver_sgeqrf(CV, ctau, m, n);
      gettimeofday(&next, NULL);
      cmpElapsed += (next.tv_sec + 1e-6 * next.tv_usec) - (cur.tv_sec + 1e-6 * cur.tv_usec);
      if (i != TRIALS - 1) memcpy(CV, A, (size_t)m * n * sizeof(float));
      gettimeofday(&cur, NULL);
    }
    if (rc == 0) printf("cuSOLVER ran QR on %dx%d matrix in %f s (avg over %d)\n", m, n, cmpElapsed / TRIALS, TRIALS);
    else printf("cuSOLVER comparator unavailable (status %d)\n", rc);
    free(CV); free(ctau);
  }
  if (check) {
    if ((double)m * m * 4 > 2e9) {
      puts("check: Q is m x m; use m <= 20000");
      exit(1);
    }
    float* Q = (float*)malloc((size_t)m * m * sizeof(float));
    float* R = (float*)malloc((size_t)m * n * sizeof(float));
    float* QR = (float*)malloc((size_t)m * n * sizeof(float));
    explicitQR(RV, tau, Q, R, m, n);
    dgemm(Q, R, QR, m, m, n);
    double err = 0, nrm = 0;
    for (size_t i = 0; i < (size_t)m * n; i++) {
      const double d = (double)QR[i] - A[i];
      err += d * d;
      nrm += (double)A[i] * A[i];
    }
    printf("L2 norm of residual QR-A: %.9g\n", sqrt(err));
    printf("relative: %.3g  (in units of n*eps: %.3f)\n", sqrt(err / nrm), sqrt(err / nrm) / (n * 1.1920929e-07));
    free(Q); free(R); free(QR);
  }
  free(A); free(RV); free(tau);
  return 0;
}
