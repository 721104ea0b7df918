// common.cuh -- shared device helpers and kernel launch declarations (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#ifndef CQR_SLOT
#define CQR_SLOT 64            // panel width / height of one R slot in the TSQR tree
#endif

namespace cqr {

constexpr unsigned kFull = 0xffffffffu;

// One-time per-DEVICE setup guard (cudaFuncSetAttribute and the device queries apply to the current device, not the
// process): `static PerDeviceOnce once; if (once.first()) { ... }`.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
  void retry() { int d = 0; if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) done[d] = false; }
};

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while its predecessor on the
// stream is still running; it must execute pdl_wait() before it reads anything the predecessor wrote (a no-op when the kernel
// was launched the ordinary way).  pdl_trigger() lets the NEXT kernel on the stream be scheduled from this point on.  Used for
// the chain of gated launches behind the Gram leaf (gram_umma.cu): seven kernels that return at once when the gate is 0 cost
// ~2.5 us each as ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename Kernel, typename Params>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Params& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// N independent warp all-reduces, issued stage by stage so the N shuffle chains overlap (a
// branchy per-value loop serialises them: ~150 cycles each).
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&s)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(kFull, s[i], o);
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] += t[i];
  }
}

// x = hi + lo with hi carrying the top 11 significand bits (the TF32 payload, low 13 bits
// zero) and lo the exact remainder: the operand split behind the 3xTF32 tensor-core GEMMs.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }

// ---- tile (TSQR leaf / tree node / batched) Householder kernels: tile_qr.cu ------------
struct TileSrc {
  float* base;          // tile t starts at base + t * tile_stride
  long long tile_stride;
  long long ld;
  long long rows_total; // rows in tile t = clamp(rows_total - t*TH, 0, TH)
};

struct TileQRParams {
  TileSrc a;            // tiles to factor
  int ncols;            // reflectors per tile (<= 64)
  int write_back;       // store V (and the tile's R) over the source tile
  float* tau;           // [tiles][64]
  float* r_out;         // R of tile t -> r_out + (t / fan) * r_tile_stride + (t % fan) * 64, ld = r_ld
  long long r_tile_stride;
  long long r_ld;
  int r_rows;           // rows of the R slot to write (64 inside the tree, ncols for the root)
  int fan;              // tiles stacked per parent tile (TH / 64)
  // batched mode: tau row stride other than 64 and no R output
  int tau_stride;
  const int* gate;      // non-null: the kernel runs only if *gate != 0 (Householder leaf behind the Gram leaf, gram_umma.cu)
};

struct TileApplyParams {
  TileSrc v;            // reflector tiles (as written by tile_qr with write_back)
  const float* tau;     // [tiles][64]
  int nref;             // reflectors per tile
  int nc;               // columns of the block being transformed (<= 64)
  const float* x;       // seed: tile t reads the 64-row slot x + (t / fan)*x_tile_stride + (t % fan)*64, ld x_ld
  long long x_tile_stride;
  long long x_ld;
  int x_rows;           // valid seed rows (<= 64); x == nullptr means identity
  int fan;
  TileSrc out;          // output tiles: rows_total masks ragged stores
};

void launch_tile_qr(const TileQRParams& p, int tiles, int tile_rows, cudaStream_t s);
void launch_tile_apply_q(const TileApplyParams& p, int tiles, int tile_rows, cudaStream_t s);
// batched QR of m x n matrices with m, n <= 64: one 64-thread CTA per matrix, one thread per column (LAPACK storage
// in place, tau[b * n + j])
void launch_batched_qr_col(float* base, long long stride, long long lda, int m, int n, int batch, float* tau, cudaStream_t s);

// ---- warp-resident flat-tree TSQR leaf (R only): tsqr_flat.cu -------------------------------
struct FlatTsqrParams {
  const float* a; long long lda;   // m x n source, never written
  long long m; int n;
  long long rows_per_chain;        // multiple of 64; chain k factors rows [k * rows_per_chain, ...)
  int chains;
  float* r_out;                    // chain k's 64 x 64 R -> r_out + (k / fan) * r_tile_stride + (k % fan) * 64, ld r_ld
  long long r_tile_stride, r_ld;
  int fan;
  // implicit-Q variant only: reflectors are written over the source (a_out == a) and taus to tau_out[block][64]
  float* a_out;
  float* tau_out;
  const int* gate;                 // non-null: the kernel runs only if *gate != 0 (see TileQRParams::gate)
};
struct FlatApplyParams {
  const float* v; long long ldv;   // the factored matrix (V blocks as written by the implicit-Q leaf)
  const float* tau;                // [blocks][64]
  long long m; int n;              // rows, reflectors per block
  int nc;                          // output columns (<= 64)
  long long rows_per_chain; int chains;
  const float* x;                  // seeds: chain k reads the 64-row slot x + (k / fan) * x_tile_stride + (k % fan) * 64, ld x_ld
  long long x_tile_stride, x_ld;   //   (nullptr = identity)
  int x_rows, fan;
  float* q; long long ldq;         // output m x nc
};
int flat_tsqr_max_chains(int sm_count);   // chains resident in one wave
void launch_tsqr_flat_r(const FlatTsqrParams& p, cudaStream_t s, bool pair = false);   // pair: two pivot columns per reduction
void launch_tsqr_flat_keep(const FlatTsqrParams& p, cudaStream_t s);   // after launch_tsqr_flat_first_blocks
void launch_tsqr_flat_first_blocks(float* a, long long lda, long long m, int n, long long rows_per_chain, int chains, float* tau,
                                   cudaStream_t s);
void launch_tsqr_flat_apply(const FlatApplyParams& p, cudaStream_t s);
// batched QR of m x n matrices (m, n <= 64), one warp per matrix, LAPACK storage in place, tau[b * n + j]
void launch_batched_qr_warp(float* base, long long stride, long long lda, int m, int n, int batch, float* tau, cudaStream_t s);

// ---- tensor-pipe flat-tree TSQR (R only): tsqr_mma.cu ------------------------------------------
// One launch turns an m x n matrix into one 64 x 64 R per CTA: warp chains of 64-row blocks, then the CTA's warps are
// combined.  CTA b writes rows [64 b, 64 b + out_rows) of r_out (column-major, ld r_ld): out_rows = 64 for a level that
// feeds the next launch (the stacked R's are its input matrix), n for the last one (a single CTA).
struct MmaTsqrParams {
  const float* a; long long lda;
  long long m; int n;
  long long rows_per_chain;        // multiple of 64
  int chains;
  float* r_out; long long r_ld;
  int out_rows;
};
int mma_tsqr_warps_per_cta();
#ifdef CQR_MMA_TRACE
void mma_tsqr_read_trace(long long* out);   // [loads, sub-panels, trailing] clock cycles and block steps of warp 0 of CTA 0
#endif
void launch_tsqr_mma_r(const MmaTsqrParams& p, cudaStream_t s);

// ---- the chain's K <= 64 inner update in one launch: chain_update.cu ------------------------------
struct ChainUpdParams {
  const float* v; long long ldv;   // explicit V (mp x kb)
  const float* t; long long ldt;   // kb x kb upper-triangular T
  float* c; long long ldc;         // mp x nc, updated in place
  long long mp; int kb, nc, trans; // trans != 0: T^T (Q^T C)
  float* wpart;                    // ctas x 64 x 192 scratch
  float* x;                        // 64 x 192 scratch
  unsigned* bar; unsigned bar_base;
  int* err;
};
bool chain_update_fits(int kb, int nc, const float* v, long long ldv, const float* c, long long ldc);
long long chain_update_part_floats(int nc);
int chain_update_min_ctas(int nc);
void launch_chain_update(const ChainUpdParams& p, int ctas, cudaStream_t s);

// ---- double-precision variant: f64_qr.cu -----------------------------------------------------------
size_t f64_workspace_bytes(long long m, int nc_max);
int f64_geqrf(double* a, long long lda, long long m, int n, double* tau, void* ws, int sm_count, cudaStream_t s);
int f64_apply_q(int trans, const double* a, long long lda, long long m, int n, const double* tau, double* c, long long ldc, int nc,
                bool c_is_identity_start, void* ws, cudaStream_t s);
void f64_set_identity(double* a, long long lda, long long m, int n, cudaStream_t s);
void f64_extract_r(const double* a, long long lda, int n, double* r, long long ldr, int r_rows, cudaStream_t s);

// ---- cross-GPU R tree over peer memory: rtree_peer.cu ------------------------------------------
constexpr int kRtreeMaxLevels = 8, kRtreeMaxWorld = 16;   // slots per epoch parity: tree levels (binary tree) or sender rank - 1 (one-hop gather, world <= 8)
constexpr int kRtreeFlatMaxWorld = 8;
struct RtreeSlab {                               // one per rank, cudaMalloc'ed, mapped into every peer with cudaIpc
  float slot[2][kRtreeMaxLevels][64 * 64];       // [epoch parity][tree level]: the R a peer hands over at that level
  unsigned ready[2][kRtreeMaxLevels];            // epoch of the R in the slot (written by the sender)
  unsigned ack[2][kRtreeMaxLevels];              // last epoch the receiver of MY R at that level consumed (written by the receiver)
};
struct RtreePeerParams {
  float* r; long long ldr;                       // this rank's n x n R (in: local TSQR result; out on rank 0: the combined R)
  int n, rank, world;
  unsigned epoch;                                // same on every rank, starts at 1, +1 per call
  unsigned long long timeout_ns;
  int* err;                                      // set to 1 if a peer never arrived
  RtreeSlab* slabs[kRtreeMaxWorld];              // slabs[rank] is this rank's own
};
void launch_rtree_peer(const RtreePeerParams& p, cudaStream_t s);   // CQR_RTREE=tree: always the binary tree; default: one-hop gather up to 8 ranks

// ---- exporter for the reference's own storage format: legacy_format.cu ------------------------
bool legacy_format_shape_ok(int m, int n);       // m = 64 + 60 k, n a multiple of 4, n <= m (the reference's legal shapes)
void launch_legacy_sweep(float* a, long long lda, int m, int n, float* tau, int rowPanels, int colPanels, float* scratch, cudaStream_t s);

// ---- Householder reconstruction + T builder: reconstruct.cu ---------------------------------
struct HrParams {
  const float* q;  long long ldq;     // thin Q of the panel (mp x b)
  const float* rt; long long ldrt;    // R from TSQR (b x b upper)
  float* a;        long long lda;     // panel of A (mp x b): Y below the diagonal, S*R on/above
  float* tau;                         // b
  float* t;        long long ldt;     // b x b compact-WY T (upper)
  float* uinv;                        // 64 x 64 scratch: inverse of the LU's U factor
  float* vbuf;     long long ldv;     // explicit Y (unit diagonal, zeros above), mp x b
  long long mp; int b;
};
void launch_hr_top(const HrParams& p, cudaStream_t s);
void launch_hr_rows(const HrParams& p, cudaStream_t s);

// T(kb x kb) from the Gram matrix G = V^T V and tau.  If have_diag != 0 the 64 x 64
// diagonal blocks of T are already in place and only the off-diagonal blocks are built.
void launch_build_t(const float* g, long long ldg, const float* tau, float* t, long long ldt, int kb,
                    int have_diag, cudaStream_t s);

// ---- multi-CTA Householder panel: panel_hh.cu -------------------------------------------------
constexpr int kPanelHHMaxCtas = 128;   // slab partials exchanged per step (slots: 2 x 128 x 64 + 2 x 64 uint2)
struct PanelHHParams {
  float* a; long long lda;      // panel (m_p x b), overwritten with R / v (LAPACK geqrf storage)
  long long mp; int b;
  float* tau;                   // b
  float* vbuf; long long ldv;   // explicit V (unit diagonal, zeros above), m_p x b
  float* t; long long ldt;      // b x b compact-WY T (upper, zero below); nullptr = not wanted
  uint2* slots; int pmax;       // exchange area {value bits, tag}
  unsigned epoch;               // unique per launch: tags of earlier launches never match
  int* err;                     // set to 1 if a spin timed out (grid was not co-resident)
};
inline size_t panel_hh_slot_bytes() { return (size_t)(2 * kPanelHHMaxCtas * 64 + 2 * 64) * sizeof(uint2); }
bool panel_hh_plan(long long mp, int max_ctas, int* ri, int* ctas);
void launch_panel_hh(const PanelHHParams& p, int ri, int ctas, cudaStream_t s);
// cluster variant (DSMEM exchange): one cluster for m_p <= 8192, two for m_p <= 16384; rows per thread rr, cluster
// size cs, cluster count ncl; launch returns false if the launch failed
bool panel_hh_cluster_plan(long long mp, int* rr, int* cs, int* ncl);
bool launch_panel_hh_cluster(const PanelHHParams& p, int rr, int cs, int ncl, cudaStream_t s);
// warp-block layout (panel_wb.cu), m_p <= 16384: warps per CTA wpc (1, 2, 4, 8), cluster size cs, clusters ncl (1 or 2)
bool panel_wb_plan(long long mp, int* wpc, int* cs, int* ncl);
bool launch_panel_wb(const PanelHHParams& p, int wpc, int cs, int ncl, cudaStream_t s);
// experimental (panel_wb2.cu, CQR_PANEL_PAIR): same plan, two pivot columns per cluster exchange; b == 64, ncl == 1 only
bool launch_panel_wb2(const PanelHHParams& p, int wpc, int cs, int ncl, int mode, cudaStream_t s);
#ifdef CQR_HH_TRACE
void panel_hh_read_trace(long long* out);
void panel_wb2_read_trace(long long* steps, long long* marks);   // [2][8][64][8] and [2][8][5] clock64 values
#endif

// ---- fp32 SIMT GEMMs and element-wise utilities: gemm_simt.cu -------------------------------
// D[z](M x N) = A(:, z-th K chunk)^T * B(z-th K chunk, :) ; D[z] = d + z * d_split_stride
void launch_gemm_tn_simt(int M, int N, int K, const float* a, long long lda, const float* b, long long ldb,
                         float* d, long long ldd, int splits, long long d_split_stride, cudaStream_t s);
// D = alpha * A * B + beta * D
void launch_gemm_nn_simt(int M, int N, int K, float alpha, const float* a, long long lda, const float* b,
                         long long ldb, float beta, float* d, long long ldd, cudaStream_t s);
// out = sum_z part[z] (M x N, part ld = ldp, stride between parts = stride)
void launch_reduce_splits(int M, int N, const float* part, long long ldp, long long stride, int splits, float* out,
                          long long ldo, cudaStream_t s);
// x(kb x nc) = op(T) * sum_z part[z]: split-K reduction fused with the multiplication by the kb x kb upper-triangular
// T (kb <= 256); trans != 0 -> T^T
void launch_tw_fused(int kb, int nc, const float* part, long long ldp, long long stride, int splits, const float* t,
                     long long ldt, int trans, float* x, long long ldx, cudaStream_t s);
void launch_set_identity(float* a, long long lda, int m, int n, cudaStream_t s);
void launch_fill_zero(float* a, long long lda, long long m, int n, cudaStream_t s);
void launch_copy_matrix(long long m, int n, const float* a, long long lda, float* b, long long ldb, cudaStream_t s);
// R extraction: r(i,j) = (i <= j) ? a(i,j) : 0 for i < r_rows
void launch_extract_r(const float* a, long long lda, int m, int n, float* r, long long ldr, int r_rows,
                      cudaStream_t s);
// V extraction from LAPACK-format storage: v(i,j) = i<j ? 0 : i==j ? 1 : a(i,j), for a panel whose
// diagonal starts at local row d0 of the mp x b block
void launch_extract_v(const float* a, long long lda, long long mp, int b, int d0, float* v, long long ldv,
                      cudaStream_t s);

// x = R^-1 b for one kb x kb (kb <= 64) upper-triangular block and nrhs right-hand sides, in place; *singular = 1 on a zero pivot
void launch_trsm_upper_block(const float* r, long long ldr, int kb, float* b, long long ldb, int nrhs, int* singular, cudaStream_t s);

// ---- tcgen05 3xTF32 GEMMs: gemm_umma.cu -------------------------------------------------------
// Both return false (nothing launched) when shape/alignment rules out the TMA path.
bool umma_available();
int umma_effective_splits(int K, int splits);   // K splits the tensor kernel really uses for a request
// max_ctas: SMs available to the launching stream (persistent grid = min(tiles, max_ctas)).
bool launch_gemm_tn_umma(int M, int N, int K, const float* a, long long lda, const float* b, long long ldb, float* d,
                         long long ldd, int splits, long long d_split_stride, int max_ctas, cudaStream_t s);
bool launch_gemm_nn_umma(int M, int N, int K, float alpha, const float* a, long long lda, const float* b, long long ldb,
                         float beta, float* d, long long ldd, int max_ctas, cudaStream_t s);

// ---- Gram leaf of the R-only TSQR on tcgen05 (gram_umma.cu): error-free bf16 slicing, fp64 Cholesky, gated fallback ----
bool gram_tsqr_eligible(const float* a, long long lda, long long m, int n);
size_t gram_tsqr_workspace_floats(int sm_count);
bool launch_tsqr_gram_r(const float* a, long long lda, long long m, int n, float* r, long long ldr, float* ws, int* flags, int sm_count,
                        int max_ctas, double bound_max, int** gate_out, double** info_out, cudaStream_t s);

// global launch counter (gpu_launches evidence)
extern std::atomic<long long> g_launches;   // host threads may drive contexts on several devices

}  // namespace cqr
