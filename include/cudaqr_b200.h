/*
 * cudaqr_b200.h -- C ABI of libcudaqr_b200.so: a B200 (sm_100a) drop-in for the QR hot
 * path of brian-kelley/CUDA-QR (qr.c / qr.cu).
 *
 * Part 1 re-exports the reference's own entry points with identical signatures and data
 * contract (host pointers, column-major, lda = m, fp32, m >= n, blocking, errors print
 * and exit(1) as qr.cu:467-471 does).  Part 2 is the device-resident API those wrappers
 * are built on (what benchmarks time; no host copies, explicit stream, int status).
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef CUDAQR_B200_H
#define CUDAQR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * Part 1 -- legacy entry points (replace the same-named functions of the reference)
 * ---------------------------------------------------------------------------------- */

/* Replaces getPanelDims, qr.cu:49-55 (qr.c:47-53 with PR=64, PC=4, qr.cu:21-23).
 * Callers size tau as rowPanels*colPanels*4 floats (qr.cu:764); that is always >= n,
 * which is what this library's tau (one value per column) needs. */
void getPanelDims(int m, int n, int* rowPanels, int* colPanels);

/* Replaces mmqr, qr.cu:475-553.  In place: on return `mat` holds R on and above the
 * diagonal and the Householder vectors below it (unit diagonal implicit); tau[0..n)
 * holds the reflector scalars, the rest of the caller's rowPanels*colPanels*4 buffer is
 * zero-filled (qr.c:62).  tau's layout is this library's own: only the
 * mmqr -> explicitQR round trip is contractual.  Any m >= n >= 1 is legal (the
 * reference silently mis-factors shapes off its window grid). */
void mmqr(float* mat, float* tau, int m, int n);

/* Replaces the CPU variant mmqr, qr.c:55-313: callee malloc()s *tau, caller free()s. */
void mmqr_alloc(float* mat, float** tau, int m, int n);

/* Double precision: the pair a `#define Scalar double` build of the reference would export (qr.c:9 "can be float or
 * double", qr.cu:747-754).  Same contract as mmqr / explicitQR above with double buffers. */
void mmqr_f64(double* mat, double* tau, int m, int n);
void explicitQR_f64(double* A, double* tau, double* Q, double* R, int m, int n);

/* mmqr with the REFERENCE's storage on output: its window sweep with PR = 64, PC = 4 (qr.cu:21-23) run on
 * the device, leaving the per-window reflector segments in place (qr.c:109-167) and
 * tau[(rowPanels*pcCount + prCount)*4 + col] (qr.c:300-304), so that the reference's own explicitQR (qr.c:330-438)
 * consumes the result unchanged.  tau holds rowPanels*colPanels*4 floats (getPanelDims).  Only the shapes the
 * reference factors correctly: m = 64 + 60 k, n a multiple of 4, n <= m -- prints and exit(1)s otherwise. */
void mmqr_reference_format(float* mat, float* tau, int m, int n);

/* Replaces explicitQR, qr.c:330-438 / qr.cu:582-686.  Q is m x m, R is m x n (zero
 * below the diagonal), A = Q*R; all column-major host buffers owned by the caller. */
void explicitQR(float* A, float* tau, float* Q, float* R, int m, int n);

/* Replaces dgemm, qr.c:443-459 / qr.cu:691-707: C(k x n) = A(k x m) * B(m x n). */
void dgemm(float* A, float* B, float* C, int k, int m, int n);

/* Replaces identity, qr.c:316-324 / qr.cu:568-576. */
void identity(float* A, int m);

/* Replaces printMat, qr.c:21-33 / qr.cu:35-47 (same text format). */
void printMat(float* mat, int m, int n);

/* ------------------------------------------------------------------------------------
 * Part 2 -- device-resident API.  All matrix pointers are DEVICE pointers, column-major.
 * Every function returns 0 on success, a negative CQR_E* on bad arguments, or a
 * positive cudaError_t.  Work is enqueued on the context's stream; nothing synchronises
 * unless stated.
 * ---------------------------------------------------------------------------------- */
typedef struct cqr_context cqr_context;

enum {
  CQR_OK = 0,
  CQR_EINVAL = -1,      /* bad shape / pointer / leading dimension            */
  CQR_ENOMEM = -2,      /* workspace allocation failed                        */
  CQR_ESTATE = -3,      /* call order (e.g. tsqr_form_q without tsqr_factor)  */
  CQR_EUNSUPPORTED = -4,
  CQR_ESINGULAR = -5    /* cqr_solve_ls: R has an exactly zero diagonal entry */
};

/* Options for cqr_set_option. */
enum {
  CQR_OPT_GEMM = 1,         /* 0 = fp32 SIMT GEMMs, 1 = tcgen05 3xTF32 (default when shapes allow) */
  CQR_OPT_OUTER_BLOCK = 2,  /* aggregated block width for the trailing update: 64..512 (default 256) */
  CQR_OPT_TILE_ROWS = 3,    /* TSQR leaf height: 128 or 256 (default 256)                            */
  CQR_OPT_SPLITK = 4,       /* 0 = automatic                                                        */
  CQR_OPT_LOOKAHEAD = 5,    /* 0: none; 1: next block's panels overlap the trailing update on a side stream; 2 (default): as 1, and in the
                             * panel-bound phase each finished panel is applied to the block after next's columns at once (panel-wise slices) */
  CQR_OPT_PANEL = 6,        /* 1 (default): one-launch multi-CTA Householder panel; 0: TSQR tree + Householder reconstruction */
  CQR_OPT_FLAT_TSQR = 7,    /* R-only cqr_tsqr_r on >= 16384 rows: 4 (default) Gram leaf, see below; 1 SIMT flat-tree Householder leaf, 2 tensor-pipe flat-tree leaf (tsqr_mma.cu),
                             * 3 SIMT leaf with two pivot columns per reduction, 0 256-row tile leaves; 4 = Gram leaf on tcgen05
                             * (gram_umma.cu: error-free bf16 slicing, fp64 Cholesky, R with a positive diagonal) with the
                             * Householder leaf (1) behind a device-side gate for ill-conditioned / badly scaled input */
  CQR_OPT_PARTITION = 8     /* read-only (cqr_get_option): 1 when the look-ahead streams own disjoint SM partitions (CUDA green contexts),
                             * 0 when they share the device (CQR_PARTITION=0, an injecting profiler, or no driver support) */
};

int cqr_create(cqr_context** ctx, int device);
int cqr_destroy(cqr_context* ctx);
int cqr_set_stream(cqr_context* ctx, void* cuda_stream);   /* cudaStream_t; NULL = default stream */
int cqr_set_option(cqr_context* ctx, int option, int value);
int cqr_get_option(cqr_context* ctx, int option, int* value);
int cqr_synchronize(cqr_context* ctx);
const char* cqr_error_string(int status);
/* Kernels this context has launched since creation (bench.py's gpu_launches). */
long long cqr_launch_count(cqr_context* ctx);
/* Per-kernel-class timing for roofline reports: between begin and end every launch group is
 * bracketed by CUDA events on the context's stream; end synchronises and returns, per class,
 * the summed device milliseconds, algorithmic flops, algorithmic bytes and launch counts. */
enum { CQR_PROF_PANEL = 0, CQR_PROF_GEMM_TN = 1, CQR_PROF_GEMM_NN = 2, CQR_PROF_MISC = 3,
       /* the same three classes when launched on the panel-chain stream (inner updates of a block) */
       CQR_PROF_CHAIN_TN = 4, CQR_PROF_CHAIN_NN = 5, CQR_PROF_CHAIN_MISC = 6, CQR_PROF_NCAT = 7 };
int cqr_profile_begin(cqr_context* ctx);
int cqr_profile_end(cqr_context* ctx, double* ms, double* flops, double* bytes, long long* launches, int ncat);
/* Timeline of the bracketed launch groups since cqr_profile_begin (call BEFORE cqr_profile_end): start/end in ms
 * relative to the first bracket, class per bracket.  Returns the number of brackets (<= cap written), < 0 on error. */
int cqr_profile_timeline(cqr_context* ctx, double* t0_ms, double* t1_ms, int* cls, int cap);
/* Pre-size the internal workspace (bytes) so no allocation happens in a timed region. */
int cqr_reserve(cqr_context* ctx, size_t bytes);

/* Blocked Householder QR (device core of mmqr).  A: m x n, lda >= m.  tau: n floats. */
int cqr_geqrf(cqr_context* ctx, float* dA, int lda, int m, int n, float* dtau);
/* Partial factorisation: Householder QR of the first nfact columns (m >= nfact) with Q^T applied to all n columns
 * (n may exceed m): what one outer step of a blocked / communication-avoiding driver needs, without rebuilding (V, T)
 * for a separate cqr_apply_q.  With one outer block to factor the finished 64-column panels are applied to the
 * trailing columns while the panel chain is still running. */
int cqr_geqrf_partial(cqr_context* ctx, float* dA, int lda, int m, int n, int nfact, float* dtau);

/* R = triu(A) into dR (r_rows x n, r_rows = m reproduces explicitQR's m x n R; r_rows = n
 * gives the square factor). */
int cqr_extract_r(cqr_context* ctx, const float* dA, int lda, int m, int n, float* dR, int ldr, int r_rows);

/* Q = H_0 H_1 ... H_{n-1} restricted to its first q_cols columns (q_cols = m: the
 * reference's explicit m x m Q; q_cols = n: thin Q).  dQ must not alias dA. */
int cqr_form_q(cqr_context* ctx, const float* dA, int lda, int m, int n, const float* dtau,
               float* dQ, int ldq, int q_cols);

/* C <- Q*C (trans = 0) or Q^T*C (trans = 1); C is m x nc. */
int cqr_apply_q(cqr_context* ctx, int trans, const float* dA, int lda, int m, int n, const float* dtau,
                float* dC, int ldc, int nc);

/* Least squares through the factorisation: B (m x nrhs, ldb >= m) <- Q^T B, then R X = B(0:n, :) by blocked back
 * substitution; X is returned in the first n rows of B (rows n..m-1 hold the residual's Q^T components).
 * Blocking (reads a singularity flag back).  The natural consumer of mmqr's output (the reference only forms a dense
 * Q, qr.c:330-438). */
int cqr_solve_ls(cqr_context* ctx, const float* dA, int lda, int m, int n, const float* dtau, float* dB, int ldb, int nrhs);

/* Communication-avoiding tall-skinny QR (n <= 64).  R-only: A is read once, never written. */
int cqr_tsqr_r(cqr_context* ctx, const float* dA, int lda, long long m, int n, float* dR, int ldr);
/* Verdict of the Gram leaf (CQR_OPT_FLAT_TSQR = 4) of the last cqr_tsqr_r / cqr_tsqr_dist_r on this context; synchronises
 * the stream.  *bound = n ||Rs^-1||_F^2 >= cond_2 of the unit-diagonal Gram matrix (Rs = R with columns scaled to unit
 * norm; -1: Cholesky broke down or a column's scale was out of range), *householder = 1 when the Householder leaf
 * produced R instead (bound above CQR_GRAM_BOUND, default 32768).  CQR_ESTATE if the last call did not use the Gram leaf. */
int cqr_tsqr_gram_info(cqr_context* ctx, double* bound, int* householder);
/* Factor keeping the implicit Q: leaf reflectors overwrite A, tree levels live in the context. */
int cqr_tsqr_factor(cqr_context* ctx, float* dA, int lda, long long m, int n, float* dR, int ldr);
/* Thin Q (m x n) of the last cqr_tsqr_factor on this context times the n x n seed dX
 * (NULL = identity): dQ = Q * [X; 0].  Used by the multi-GPU R-tree to push the tree's
 * own Q blocks down into each rank's local Q. */
int cqr_tsqr_form_q(cqr_context* ctx, const float* dX, int ldx, float* dQ, int ldq);
/* QR of nblk stacked n x n upper-triangular blocks (dRs: (nblk*n) x n, ld = ldrs); keeps
 * the reflectors so cqr_stack_form_q can expand it.  The combine step of the R-tree. */
int cqr_stack_qr(cqr_context* ctx, float* dRs, int ldrs, int nblk, int n, float* dtau, float* dR, int ldr);
int cqr_stack_form_q(cqr_context* ctx, const float* dRs, int ldrs, int nblk, int n, const float* dtau,
                     const float* dX, int ldx, float* dQs, int ldqs);

/* Device core of mmqr_reference_format: in place on dA, dtau_grid = rowPanels*colPanels*4 floats (zero-filled here).
 * CQR_EUNSUPPORTED for shapes off the reference's window grid. */
int cqr_mmqr_reference_format(cqr_context* ctx, float* dA, int lda, int m, int n, float* dtau_grid);

/* Double-precision blocked Householder QR (fp64 SIMT kernels, 32-column panels factored by one cooperative launch each;
 * LAPACK storage, tau[0..n)); form / apply Q and R extraction as in the fp32 API.  Device pointers, column-major. */
int cqr_dgeqrf(cqr_context* ctx, double* dA, int lda, int m, int n, double* dtau);
int cqr_dform_q(cqr_context* ctx, const double* dA, int lda, int m, int n, const double* dtau, double* dQ, int ldq, int q_cols);
int cqr_dapply_q(cqr_context* ctx, int trans, const double* dA, int lda, int m, int n, const double* dtau, double* dC, int ldc, int nc);
int cqr_dextract_r(cqr_context* ctx, const double* dA, int lda, int m, int n, double* dR, int ldr, int r_rows);

/* Row-partitioned TSQR across the GPUs of one box (BASELINE config 3), one process per GPU, R tree over peer memory:
 * cqr_dist_export allocates this rank's exchange slab and returns its 64-byte cudaIpc handle; the launcher hands every
 * rank all `world` handles (rank order, 64 bytes each) for cqr_dist_attach; cqr_tsqr_dist_r is then the local R-only
 * TSQR of this rank's m_loc x n rows plus one kernel per rank that combines the R factors, the n x n blocks moving as
 * NVLink stores into the receiver's slab (no NCCL, no host synchronisation between calls): up to 8 ranks every rank
 * stores into rank 0's slab and rank 0 factors the stacked (64 world) x n matrix in one go, beyond that (or with
 * CQR_RTREE=tree) a binary reduction tree.  The combined R is on rank 0.  Every rank makes the same sequence of calls.  The reference has no multi-GPU path (qr.cu:737). */
int cqr_dist_export(cqr_context* ctx, void* handle_out_64_bytes);
int cqr_dist_attach(cqr_context* ctx, int rank, int world, const void* handles);
int cqr_dist_detach(cqr_context* ctx);
int cqr_tsqr_dist_r(cqr_context* ctx, const float* dA, int lda, long long m_loc, int n, float* dR, int ldr);

/* `batch` independent m x n matrices (m <= 256, n <= 64, m >= n), one CTA each; matrix i
 * starts at dA + i*stride, lda >= m.  tau: batch x n. */
int cqr_geqrf_batched(cqr_context* ctx, float* dA, int lda, long long stride, int m, int n, int batch, float* dtau);

/* D = alpha * op(A) * B + beta * D, fp32, column-major (device core of dgemm).
 * transA = 0: A is M x K; transA = 1: A is K x M. */
int cqr_gemm(cqr_context* ctx, int transA, int M, int N, int K, float alpha, const float* dA, int lda,
             const float* dB, int ldb, float beta, float* dD, int ldd);

/* D = alpha * op(A) * B + beta * D on the tcgen05 tensor cores with the 3xTF32 operand split
 * (fp32-faithful to ~2^-21).  transA = 1 supports alpha = 1, beta = 0 only (split-K partials).
 * CQR_EUNSUPPORTED if pointers are not 16-byte aligned or leading dimensions not multiples of 4. */
int cqr_gemm_tf32x3(cqr_context* ctx, int transA, int M, int N, int K, float alpha, const float* dA, int lda,
                    const float* dB, int ldb, float beta, float* dD, int ldd);

int cqr_set_identity(cqr_context* ctx, float* dA, int lda, int m, int n);

/* Library build info: "sm_100a;<date>;<features>" */
const char* cqr_version(void);

/* ---- Part 3: comparator slot of the reference's command line (never on the hot path) ----------------------
 * The reference can time MAGMA's magma_sgeqrf2_gpu next to mmqr (qr.cu:555-565 magmaQR, qr.cu:790-806; timing only,
 * compiled out by default).  Same call shape with cuSOLVER's geqrf, dlopen'ed at run time: host matrix up, library
 * factorisation, matrix and tau[0..n) back.  Returns 0, CQR_EUNSUPPORTED when libcusolver is not on the box. */
int cqr_compare_cusolver_sgeqrf(float* mat, float* tau, int m, int n);

#ifdef __cplusplus
}
#endif
#endif /* CUDAQR_B200_H */
