// rtree_peer.cu -- the cross-GPU R tree of the row-partitioned TSQR (BASELINE config 3) over peer memory.
//
// After its local TSQR every rank holds one n x n R.  The binary reduction tree (rank r with bit s set hands its R to
// rank r - s, which stacks [R_mine; R_peer] and re-factors) used to be three serialized ncclSend / ncclRecv hops wrapped
// in copies and one cqr_stack_qr launch each.  Here it is ONE kernel launch per rank, behind the local TSQR on the same
// stream: a sender STORES its 16 KiB R straight into the receiver's slab (a cudaIpc-mapped peer pointer: NVLink stores)
// and releases a flag there; the receiver's CTA spins on its own flag, runs the stacked Householder QR in registers
// (tile_qr_core<4>: 128 x 64) and either forwards the result one level up the same way or, on rank 0, leaves the final R.
// Flags carry the call's epoch; slots are double-buffered by epoch parity and a sender waits for the receiver's
// acknowledgement of epoch - 2 before it reuses a slot, so back-to-back calls need no host synchronisation.
// Up to 8 ranks the tree is flattened to ONE hop (rtree_flat_kernel): every rank r > 0 stores its R into slot r - 1 of rank 0's
// slab, rank 0 waits for all of them and factors the stacked (64 world) x 64 matrix in registers in one go (tile_qr_core<4 / 8 /
// 16>).  The binary tree's log2(world) hops are serial -- store, flag, 128 x 64 QR, ~32 us each: 99 us of the 232 us an 8-GPU
// 8M x 64 TSQR takes with the Gram leaf -- the gather pays one store + flag latency and one larger QR.
// The reference has no multi-GPU path (qr.cu:737); parity is against the single-device R of the same matrix.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tile_qr_core.cuh"

namespace cqr {

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until pred(value at p) holds; false on timeout (the peer never arrived: the caller flags the result void)
template <typename Pred>
__device__ __forceinline__ bool spin_until(const unsigned* p, Pred pred, unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  while (!pred(ld_acquire_sys(p))) {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > timeout_ns) return false;
  }
  return true;
}

__global__ void __launch_bounds__(256, 1) rtree_peer_kernel(RtreePeerParams p) {
  __shared__ float vs[2][128];
  __shared__ int ok_s;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int par = p.epoch & 1;
  const int n = p.n;
  RtreeSlab* mine = p.slabs[p.rank];
  for (int level = 0, s = 1; s < p.world; ++level, s <<= 1) {
    if (p.rank % (2 * s) == s) {                 // hand my R to rank - s and finish
      RtreeSlab* peer = p.slabs[p.rank - s];
      if (tid == 0) ok_s = spin_until(&mine->ack[par][level], [&](unsigned v) { return v + 2 >= p.epoch; }, p.timeout_ns) ? 1 : 0;
      __syncthreads();
      if (!ok_s) { if (tid == 0) *p.err = 1; return; }
      float* dst = peer->slot[par][level];
      for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int r = idx & 63, c = idx >> 6;
        dst[idx] = (r <= c && c < n) ? p.r[r + (long long)c * p.ldr] : 0.f;
      }
      __threadfence_system();
      __syncthreads();
      if (tid == 0) st_release_sys(&peer->ready[par][level], p.epoch);
      return;
    }
    if (p.rank % (2 * s) == 0 && p.rank + s < p.world) {   // merge the R of rank + s into mine
      if (tid == 0) ok_s = spin_until(&mine->ready[par][level], [&](unsigned v) { return v == p.epoch; }, p.timeout_ns) ? 1 : 0;
      __syncthreads();
      if (!ok_s) { if (tid == 0) *p.err = 1; return; }
      const float* src = mine->slot[par][level];
      float a[8][4];
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c = w + 8 * ci;
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
          const int r = l + 32 * ri;
          float v = 0.f;
          if (c < n) {
            if (r < 64) v = (r <= c) ? p.r[r + (long long)c * p.ldr] : 0.f;
            else v = __ldcg(src + (r - 64) + 64 * c);   // written by the peer GPU: read through L2, never a stale L1 line
          }
          a[ci][ri] = v;
        }
      }
      __syncthreads();                            // every thread has read R before anyone overwrites it
      tile_qr_core<4>(a, n, vs, nullptr, w, l);
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c = w + 8 * ci;
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
          const int r = l + 32 * ri;
          if (c < n && r < n) p.r[r + (long long)c * p.ldr] = (r <= c) ? a[ci][ri] : 0.f;
        }
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_sys(&p.slabs[p.rank + s]->ack[par][level], p.epoch);
    }
  }
}


// One-hop gather: slot index = sender rank - 1; same epoch / acknowledgement protocol as the tree.
template <int RI>
__global__ void __launch_bounds__(256, 1) rtree_flat_kernel(RtreePeerParams p) {
  constexpr int TH = 32 * RI;
  __shared__ float vs[2][TH];
  __shared__ int ok_s;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int par = p.epoch & 1;
  const int n = p.n;
  RtreeSlab* mine = p.slabs[p.rank];
  if (p.rank != 0) {
    const int slot = p.rank - 1;
    RtreeSlab* root = p.slabs[0];
    if (tid == 0) ok_s = spin_until(&mine->ack[par][slot], [&](unsigned v) { return v + 2 >= p.epoch; }, p.timeout_ns) ? 1 : 0;
    __syncthreads();
    if (!ok_s) { if (tid == 0) *p.err = 1; return; }
    float* dst = root->slot[par][slot];
    for (int idx = tid; idx < 64 * 64; idx += 256) {
      const int r = idx & 63, c = idx >> 6;
      dst[idx] = (r <= c && c < n) ? p.r[r + (long long)c * p.ldr] : 0.f;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys(&root->ready[par][slot], p.epoch);
    return;
  }
  // rank 0: thread s - 1 waits for rank s
  if (tid == 0) ok_s = 1;
  __syncthreads();
  if (tid < p.world - 1) {
    if (!spin_until(&mine->ready[par][tid], [&](unsigned v) { return v == p.epoch; }, p.timeout_ns)) ok_s = 0;
  }
  __syncthreads();
  if (!ok_s) { if (tid == 0) *p.err = 1; return; }
  float a[8][RI];
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    const int c = w + 8 * ci;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = l + 32 * ri, blk = r >> 6, rr = r & 63;
      float v = 0.f;
      if (c < n && rr <= c && blk < p.world) {
        if (blk == 0) v = p.r[rr + (long long)c * p.ldr];
        else v = __ldcg(mine->slot[par][blk - 1] + rr + 64 * c);   // written by a peer GPU: read through L2, never a stale L1 line
      }
      a[ci][ri] = v;
    }
  }
  __syncthreads();                            // every thread has read R before anyone overwrites it
  tile_qr_core<RI>(a, n, vs, nullptr, w, l);
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    const int c = w + 8 * ci;
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      const int r = l + 32 * ri;
      if (c < n && r < n) p.r[r + (long long)c * p.ldr] = (r <= c) ? a[ci][ri] : 0.f;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid < p.world - 1) st_release_sys(&p.slabs[tid + 1]->ack[par][tid], p.epoch);
}

}  // namespace

void launch_rtree_peer(const RtreePeerParams& p, cudaStream_t s) {
  static int tree_only = -1;
  if (tree_only < 0) { const char* e = getenv("CQR_RTREE"); tree_only = (e && strcmp(e, "tree") == 0) ? 1 : 0; }
  ++g_launches;
  if (tree_only || p.world > kRtreeFlatMaxWorld) rtree_peer_kernel<<<1, 256, 0, s>>>(p);
  else if (p.world <= 2) rtree_flat_kernel<4><<<1, 256, 0, s>>>(p);
  else if (p.world <= 4) rtree_flat_kernel<8><<<1, 256, 0, s>>>(p);
  else rtree_flat_kernel<16><<<1, 256, 0, s>>>(p);
}

}  // namespace cqr
