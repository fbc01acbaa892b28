// Batched greedy NMS, entirely on the device (no D2H round trip, no host scan).
//
// Replaces `_C.nms` (lib/model/csrc/nms.h:10-28 -> nms_cuda, csrc/cuda/nms.cu:70-131) and the
// per-image python loop around it (lib/model/rpn/proposal_layer.py:134-164).
//
//   1. gather   : boxes in descending-score order (the caller supplies the order)
//   2. bitmask  : 64x64 tiles of the IoU > thr relation, upper triangle only (the reference
//                 computes all tiles, nms.cu:28 has the early exit commented out), one warp-pair
//                 per tile, 64 candidate boxes staged in shared memory
//   3. scan     : one CTA per image walks the 64-box blocks in order.  Warp 0 resolves the 64
//                 in-block decisions from the diagonal words held in registers; then every
//                 thread ORs the kept rows into its slice of the suppression vector (coalesced
//                 64-bit loads).  In "proposal" mode the scan stops as soon as max_out boxes are
//                 kept (the caller only consumes keep[:post_nms_topN], proposal_layer.py:156).
//   4. emit     : kept positions, or (mode 1) original indices in ascending order like the
//                 reference's final sort (nms.cu:127-130), plus the zero-padded roi tensor.
//
// IoU arithmetic follows devIoU (nms.cu:13-21) with the legacy +1 convention; every operation
// is an explicitly rounded IEEE op (__fadd_rn/__fmul_rn/__fdiv_rn) so no FMA contraction can
// change a comparison: results are bit-exact against the C oracle (oracle/oracle_ops.c).  The
// IEEE division is only executed when inter / den is within 4e-6 of the threshold (iou_gt).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

typedef unsigned long long u64;

__global__ void nms_gather_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ order,
                                  float4* __restrict__ sorted, int n_total, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t src = order ? order[(size_t)b * n + i] : i;
  sorted[(size_t)b * n + i] = boxes[(size_t)b * n_total + src];
}

// iou_legacy(a, b) > thr, decided without the IEEE division unless the quotient is within 2^-18 of thr:
// disjoint boxes have inter == 0 (quotient 0 or NaN, never > thr >= 0); otherwise compare inter with
// thr * den using a guard band and fall back to the exact division only inside it.
// Pre-test (most pairs end here): the intersection is no larger than the smaller box and the union no smaller than the
// larger one, so iou <= min(sa, sb) / max(sa, sb); with both areas positive, min < 0.98 * thr * max puts the computed
// quotient (three roundings, < 1e-6 relative) strictly below thr -- a conservative reject that never changes a decision.
__device__ __forceinline__ bool iou_gt(const float4 a, float sa, const float4 b, float sb, float thr) {
  if (thr > 0.f && sa > 0.f && sb > 0.f) {
    const float t98 = 0.98f * thr;
    if (sa < t98 * sb || sb < t98 * sa) return false;
  }
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  if (thr >= 0.f && inter == 0.f) return false;
  const float den = __fsub_rn(__fadd_rn(sa, sb), inter);
  const float t = __fmul_rn(thr, den);
  if (thr >= 0.f && den > 0.f) {
    if (inter > __fmul_rn(t, 1.000004f)) return true;
    if (inter < __fmul_rn(t, 0.999996f)) return false;
  }
  return __fdiv_rn(inter, den) > thr;
}
__device__ __forceinline__ float area_legacy(const float4 a) {
  return __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
}

// grid (col_blk, row_blk, B), 64 threads.  mask[b][row][col_blk] bit j = IoU(row, col_blk*64+j) > thr
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ sorted, u64* __restrict__ mask, int n, int nblk, float thr) {
  const int col_blk = blockIdx.x, row_blk = blockIdx.y, b = blockIdx.z;
  if (col_blk < row_blk) return;  // lower triangle is never read by the scan
  const float4* bx = sorted + (size_t)b * n;
  __shared__ float4 cols[64];
  __shared__ float col_area[64];
  const int col_size = min(n - col_blk * 64, 64);
  const int row_size = min(n - row_blk * 64, 64);
  if ((int)threadIdx.x < col_size) {
    const float4 c = bx[col_blk * 64 + threadIdx.x];
    cols[threadIdx.x] = c;
    col_area[threadIdx.x] = area_legacy(c);
  }
  __syncthreads();
  if ((int)threadIdx.x < row_size) {
    const int row = row_blk * 64 + threadIdx.x;
    const float4 me = bx[row];
    const float my_area = area_legacy(me);
    u64 t = 0;
    const int start = (row_blk == col_blk) ? threadIdx.x + 1 : 0;
    for (int j = start; j < col_size; ++j)
      if (iou_gt(me, my_area, cols[j], col_area[j], thr)) t |= 1ULL << j;
    mask[((size_t)b * n + row) * nblk + col_blk] = t;
  }
}

__device__ __forceinline__ u64 warp_or64(u64 v) {
  const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
  const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
  return ((u64)hi << 32) | lo;
}

// Greedy resolve of one 64-box block, warp-collective (all 32 lanes of one warp call it).  Lane l holds the successor
// masks of block rows l and l + 32 (s0 / s1: bit c > row set when IoU(row, c) > thr -- the upper triangle of the diagonal
// tile); `avail` (uniform) = boxes that exist and are not suppressed by a box kept in an earlier block.
// The sequential rule "box j is kept iff no KEPT box i < j overlaps it" (nms.cu:112-123) is evaluated in rounds: an
// undecided box none of whose overlapping predecessors is still undecided is kept (its kept predecessors would already
// have removed it), and what it overlaps becomes suppressed.  The lowest undecided box is never blocked, so every round
// decides at least one box; typical NMS blocks finish in 3-6 rounds of four warp reductions instead of 64 dependent steps.
// Returns the kept mask (uncapped).
__device__ __forceinline__ u64 resolve_block(u64 s0, u64 s1, u64 avail, int lane) {
  u64 und = avail, kept = 0ULL;
  while (und) {
    const u64 m0 = ((und >> lane) & 1ULL) ? s0 : 0ULL, m1 = ((und >> (lane + 32)) & 1ULL) ? s1 : 0ULL;
    const u64 blocked = warp_or64(m0 | m1);            // overlapped by an undecided predecessor
    const u64 now = und & ~blocked;
    kept |= now;
    const u64 k0 = ((now >> lane) & 1ULL) ? s0 : 0ULL, k1 = ((now >> (lane + 32)) & 1ULL) ? s1 : 0ULL;
    const u64 supp = warp_or64(k0 | k1);               // overlapped by a box kept in this round
    und &= ~now & ~supp;
  }
  return kept;
}

// The same decision walked in order, one iteration per KEPT box: the lowest undecided box is kept, its successor mask
// (fetched from the lane that holds it: two shuffles) retires everything it overlaps.  Cost ~ 50 clk x boxes kept.
__device__ __forceinline__ u64 resolve_block_serial(u64 s0, u64 s1, u64 avail, int lane) {
  u64 und = avail, kept = 0ULL;
  while (und) {
    const int j = __ffsll((long long)und) - 1;
    const u64 sel = (j & 32) ? s1 : s0;
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)sel, j & 31);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(sel >> 32), j & 31);
    const u64 bit = 1ULL << j;
    kept |= bit;
    und &= ~((((u64)hi << 32) | lo) | bit);
  }
  return kept;
}

// keep only the first `room` set bits
__device__ __forceinline__ u64 first_bits(u64 kept, int room) {
  if (room <= 0) return 0ULL;
  while (__popcll(kept) > room) kept &= ~(1ULL << (63 - __clzll((long long)kept)));
  return kept;
}

// one CTA per image.  The reference copies the whole mask to the host and scans it there (nms.cu:100-123); here one CTA
// walks the 64-box blocks with everything it needs per block arriving in ONE overlapped round trip:
//   warp 0        resolves block blk (resolve_block) from the diagonal words prefetched during the previous block
//   warps 1-7     gather column blk + 1 of the rows kept so far (positions in shared memory): the suppression word of the
//                 next block -- one strided 8-byte load per kept box, all independent
//   threads 64-127 read word blk + 1 of ALL 64 rows of block blk (before it is known which of them are kept)
//   threads 0-63  prefetch the diagonal words of block blk + 1
// then the kept boxes of blk add their (already loaded) words.  Two barriers per block; no thread ever waits on a chain
// of dependent global loads (the first version: 25 serial loads per thread and a single-thread 64-step resolve, 1.1 ms
// at n = 6000).
static constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const u64* __restrict__ mask, int n, int nblk, int max_out, int stop_early, int rounds,
                int32_t* __restrict__ keep_pos /*[B, n]*/, int32_t* __restrict__ n_keep_all /*[B]*/) {
  extern __shared__ int32_t klist[];   // [n] kept positions (score order)
  __shared__ u64 diag[2][64];
  __shared__ u64 nextw[64];
  __shared__ u64 wpart[32];
  __shared__ u64 s_kept, s_remv;
  __shared__ int s_total;
  const int b = blockIdx.x;
  const u64* m = mask + (size_t)b * n * nblk;
  int32_t* kp = keep_pos + (size_t)b * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) diag[0][tid] = (tid < n) ? m[(size_t)tid * nblk] : 0ULL;
  if (tid == 0) { s_total = 0; s_remv = 0ULL; }
  __syncthreads();

  for (int blk = 0; blk < nblk; ++blk) {
    const int base = blk * 64;
    const int cnt = min(n - base, 64);
    const int total = s_total;
    const bool more = blk + 1 < nblk;
    // ---- loads of this iteration, all issued before anything waits
    u64 pre = 0ULL;
    if (more && tid < 64) {                       // diagonal word of block blk + 1
      const int row = base + 64 + tid;
      if (row < n) pre = m[(size_t)row * nblk + blk + 1];
    } else if (more && tid < 128) {               // word blk + 1 of row (base + tid - 64)
      const int row = base + tid - 64;
      if (row < n) pre = m[(size_t)row * nblk + blk + 1];
    }
    if (warp == 0) {
      u64 avail = ~s_remv;
      if (cnt < 64) avail &= (1ULL << cnt) - 1ULL;            // boxes past the end do not exist
      u64 kept = rounds ? resolve_block(diag[blk & 1][lane], diag[blk & 1][lane + 32], avail, lane)
                        : resolve_block_serial(diag[blk & 1][lane], diag[blk & 1][lane + 32], avail, lane);
      if (stop_early) kept = first_bits(kept, max_out - total);
      if (lane == 0) s_kept = kept;
    } else if (more) {
      u64 acc = 0ULL;
#pragma unroll 4
      for (int i = tid - 32; i < total; i += kScanThreads - 32) acc |= m[(size_t)klist[i] * nblk + blk + 1];
      acc = warp_or64(acc);
      if (lane == 0) wpart[warp] = acc;
    }
    if (more && tid < 64) diag[(blk + 1) & 1][tid] = pre;
    else if (more && tid < 128) nextw[tid - 64] = pre;
    __syncthreads();
    const u64 kept = s_kept;
    if (tid < 64 && ((kept >> tid) & 1ULL)) {
      const int pos = total + __popcll(kept & ((1ULL << tid) - 1ULL));
      klist[pos] = base + tid;
      kp[pos] = base + tid;
    }
    const int new_total = total + __popcll(kept);
    const bool done = stop_early && new_total >= max_out;
    if (warp == 0 && more) {
      u64 v = (((kept >> lane) & 1ULL) ? nextw[lane] : 0ULL) | (((kept >> (lane + 32)) & 1ULL) ? nextw[lane + 32] : 0ULL);
      if (lane >= 1) v |= wpart[lane];
      v = warp_or64(v);
      if (lane == 0) s_remv = v;
    }
    if (tid == 32) s_total = new_total;
    __syncthreads();
    if (done) break;  // uniform: derived from shared values read after the barrier
  }
  if (tid == 0) n_keep_all[b] = s_total;
}

// mode 0 outputs: keep_out[b, i] = kept position (or -1), n_keep, rois_out
__global__ void nms_emit_proposals_kernel(const float4* __restrict__ sorted, const int32_t* __restrict__ keep_pos,
                                          const int32_t* __restrict__ n_keep_all, int n, int max_out,
                                          int64_t* __restrict__ keep_out, int32_t* __restrict__ n_keep,
                                          float* __restrict__ rois_out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk = min(n_keep_all[b], max_out);
  if (i == 0) n_keep[b] = nk;
  if (i >= max_out) return;
  const bool live = i < nk;
  const int pos = live ? keep_pos[(size_t)b * n + i] : -1;
  keep_out[(size_t)b * max_out + i] = pos;
  if (rois_out) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) bx = sorted[(size_t)b * n + pos];
    float* r = rois_out + ((size_t)b * max_out + i) * 5;
    r[0] = (float)b;
    r[1] = bx.x;
    r[2] = bx.y;
    r[3] = bx.z;
    r[4] = bx.w;
  }
}

// mode 1: flag the kept boxes in original index space ...
__global__ void nms_mark_kernel(const int32_t* __restrict__ keep_pos, const int32_t* __restrict__ n_keep_all,
                                const int64_t* __restrict__ order, int n, int n_total, uint8_t* __restrict__ flags) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keep_all[b]) return;
  const int pos = keep_pos[(size_t)b * n + i];
  const int64_t orig = order ? order[(size_t)b * n + pos] : pos;
  flags[(size_t)b * n_total + orig] = 1;
}

// ... and compact them in ascending index order (one CTA per image, ordered block scan)
__global__ void __launch_bounds__(1024)
nms_compact_kernel(const uint8_t* __restrict__ flags, int n_total, int max_out, int64_t* __restrict__ keep_out,
                   int32_t* __restrict__ n_keep) {
  __shared__ int warp_cnt[32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const uint8_t* f = flags + (size_t)b * n_total;
  int64_t* ko = keep_out + (size_t)b * max_out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n_total; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const bool on = i < n_total && f[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (on) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      if (pos < max_out) ko[pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_cnt[w];
      s_base += t;
    }
    __syncthreads();
  }
  const int total = s_base;
  if (threadIdx.x == 0) n_keep[b] = min(total, max_out);
  for (int i = total + threadIdx.x; i < max_out; i += blockDim.x) ko[i] = -1;
}

// ---------------------------------------------------------------------------------------------
// "Proposal mode" fast path (mode 0, max_out <= kLazyMaxOut): greedy NMS only ever compares a
// candidate with boxes that were KEPT before it, and the caller consumes just the first max_out
// survivors (proposal_layer.py:156 takes keep[:post_nms_topN]).  So instead of the N^2/2 bitmask
// (18 M IoU tests at N = 6000) one CTA per image walks the candidates in score order, 64 at a time,
// with the kept boxes in shared memory:
//   1. the 64 candidates against every box kept so far          (64 x kept tests, all threads)
//   2. the 64 x 64 upper triangle inside the block              (bitmask words in shared memory)
//   3. the resolve of the block from those words                 (warp 0, resolve_block: a few reduction rounds)
// and stops as soon as max_out boxes are kept: at most N x max_out tests (1.8 M), typically far
// fewer, no mask in HBM, and the gather by `order` is fused into the candidate load.  Decisions use
// the same iou_gt predicate in the same (kept, candidate) argument order as the bitmask path, so the
// two paths and the C oracle agree bit for bit.
// ---------------------------------------------------------------------------------------------
static constexpr int kLazyMaxOut = 1024;
static constexpr int kLazyThreads = 1024;

__global__ void __launch_bounds__(kLazyThreads)
nms_lazy_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ order, int n_total, int n, float thr,
                int max_out, int rounds, int64_t* __restrict__ keep_out, int32_t* __restrict__ n_keep,
                float* __restrict__ rois_out) {
  extern __shared__ __align__(16) uint8_t lazy_smem[];
  float4* kb = reinterpret_cast<float4*>(lazy_smem);              // [max_out] kept boxes
  float* ka = reinterpret_cast<float*>(kb + max_out);             // [max_out] their areas
  int32_t* kpos = reinterpret_cast<int32_t*>(ka + max_out);       // [max_out] their positions in score order
  __shared__ float4 cand[2][64];     // double buffer: the next block's candidates are fetched while this one is resolved
  __shared__ float cand_area[2][64];
  __shared__ u64 diag[64];
  __shared__ unsigned int supp[2];  // candidates suppressed by earlier kept boxes (bits 0-31, 32-63)
  __shared__ u64 s_kept;
  __shared__ int s_total;
  const int b = blockIdx.x;
  const float4* bx = boxes + (size_t)b * n_total;
  const int64_t* ord = order ? order + (size_t)b * n : nullptr;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_total = 0;
  if (tid < 64) {
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < n) c = bx[ord ? ord[tid] : (int64_t)tid];
    cand[0][tid] = c;
    cand_area[0][tid] = area_legacy(c);
  }
  __syncthreads();

  int buf = 0;
  for (int base = 0; base < n; base += 64, buf ^= 1) {
    const int cnt = min(64, n - base);
    const int nk = s_total;
    const float4* cd = cand[buf];
    const float* ca = cand_area[buf];
    // the next block's candidates: two dependent global loads (order, then box) issued now, consumed after step 3
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid >= 64 && tid < 128) {
      const int j = base + tid;   // = base + 64 + (tid - 64)
      if (j < n) nxt = bx[ord ? ord[j] : (int64_t)j];
    }
    if (tid < 2) supp[tid] = 0u;
    __syncthreads();
    // 1. candidates x kept: pair p -> (kept i = p / 64, candidate j = p % 64); a warp shares one kept box
    for (int p = tid; p < nk * 64; p += kLazyThreads) {
      const int i = p >> 6, j = p & 63;
      const bool hit = j < cnt && iou_gt(kb[i], ka[i], cd[j], ca[j], thr);
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && bal) atomicOr(&supp[j >> 5], bal);
    }
    // 2. upper triangle inside the block: (row r, column c > r), 8 threads per row
    if (tid < 512) {
      const int r = tid >> 3, c0 = tid & 7;
      u64 bits = 0;
      if (r < cnt) {
        const float4 me = cd[r];
        const float my_area = ca[r];
        for (int c = r + 1 + c0; c < cnt; c += 8)
          if (iou_gt(me, my_area, cd[c], ca[c], thr)) bits |= 1ULL << c;
      }
      bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
      if (c0 == 0) diag[r] = bits;
    }
    __syncthreads();
    // 3. resolve the block (warp 0, see resolve_block), capped at max_out survivors in total
    if (tid < 32) {
      u64 avail = ~((u64)supp[0] | ((u64)supp[1] << 32));
      if (cnt < 64) avail &= (1ULL << cnt) - 1ULL;
      u64 kept = rounds ? resolve_block(diag[lane], diag[lane + 32], avail, lane)
                        : resolve_block_serial(diag[lane], diag[lane + 32], avail, lane);
      kept = first_bits(kept, max_out - nk);
      if (lane == 0) s_kept = kept;
    }
    if (tid >= 64 && tid < 128) {   // park the prefetched candidates in the other buffer
      cand[buf ^ 1][tid - 64] = nxt;
      cand_area[buf ^ 1][tid - 64] = area_legacy(nxt);
    }
    __syncthreads();
    const u64 kept = s_kept;
    if (tid < 64 && ((kept >> tid) & 1ULL)) {
      const int pos = nk + __popcll(kept & ((1ULL << tid) - 1ULL));
      kb[pos] = cd[tid];
      ka[pos] = ca[tid];
      kpos[pos] = base + tid;
    }
    const int total = nk + __popcll(kept);
    if (tid == 0) s_total = total;
    __syncthreads();
    if (total >= max_out) break;
  }
  const int nk = s_total;
  if (tid == 0) n_keep[b] = nk;
  for (int i = tid; i < max_out; i += kLazyThreads) {
    const bool live = i < nk;
    keep_out[(size_t)b * max_out + i] = live ? kpos[i] : -1;
    if (rois_out) {
      const float4 k = live ? kb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      float* r = rois_out + ((size_t)b * max_out + i) * 5;
      r[0] = (float)b;
      r[1] = k.x;
      r[2] = k.y;
      r[3] = k.z;
      r[4] = k.w;
    }
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct NmsWs {
  size_t sorted, mask, keep_pos, n_keep_all, flags, total;
};
static NmsWs nms_layout(int B, int n_total, int n) {
  NmsWs w;
  const size_t nblk = (n + 63) / 64;
  size_t off = 0;
  w.sorted = off; off += align256((size_t)B * n * 16);
  w.mask = off; off += align256((size_t)B * n * nblk * 8);
  w.keep_pos = off; off += align256((size_t)B * n * 4);
  w.n_keep_all = off; off += align256((size_t)B * 4);
  w.flags = off; off += align256((size_t)B * n_total);
  w.total = off;
  return w;
}

size_t nms_workspace_bytes(int B, int n_total, int n) { return nms_layout(B, n_total, n).total; }

int nms_run(const float* boxes, const int64_t* order, int B, int n_total, int n, float thr, int max_out,
            int mode, int64_t* keep_out, int32_t* n_keep, float* rois_out, void* ws, size_t ws_bytes,
            cudaStream_t stream) {
  AITB_REQUIRE(B > 0 && n > 0 && n_total >= n, "aitb_nms: bad sizes B=%d n_total=%d n=%d", B, n_total, n);
  AITB_REQUIRE(mode == 0 || mode == 1, "aitb_nms: mode must be 0 or 1");
  AITB_REQUIRE(max_out > 0, "aitb_nms: max_out must be positive");
  AITB_REQUIRE(mode == 0 || max_out >= n, "aitb_nms: mode 1 needs max_out >= n");
  AITB_REQUIRE(order != nullptr || n == n_total, "aitb_nms: order == NULL requires n == n_total");
  AITB_REQUIRE(boxes && keep_out && n_keep && ws, "aitb_nms: null pointer");
  AITB_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)ws & 255) == 0, "aitb_nms: misaligned boxes/workspace");
  static const bool no_lazy = getenv("AITB_NMS_NO_LAZY") != nullptr;   // debug: force the bitmask path
  static const bool resolve_rounds = getenv("AITB_NMS_SERIAL") == nullptr;   // default: round-based block resolve (measured 48.7 vs 71.1 us for 8 images; mask-mode scan 215 vs 315 us); A/B: AITB_NMS_SERIAL=1
  if (mode == 0 && max_out <= kLazyMaxOut && !no_lazy) {
    const size_t smem = (size_t)max_out * (16 + 4 + 4);
    nms_lazy_kernel<<<B, kLazyThreads, smem, stream>>>(reinterpret_cast<const float4*>(boxes), order, n_total, n, thr,
                                                       max_out, resolve_rounds ? 1 : 0, keep_out, n_keep, rois_out);
    return check_launch("nms_lazy_kernel");
  }
  const NmsWs L = nms_layout(B, n_total, n);
  AITB_REQUIRE(ws_bytes >= L.total, "aitb_nms: workspace too small (%zu < %zu)", ws_bytes, L.total);
  const int nblk = (n + 63) / 64;
  AITB_REQUIRE((size_t)n * 4 <= 200 * 1024, "aitb_nms: n=%d too large for the on-chip kept list (51200 boxes)", n);
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  float4* sorted = reinterpret_cast<float4*>(w + L.sorted);
  u64* mask = reinterpret_cast<u64*>(w + L.mask);
  int32_t* keep_pos = reinterpret_cast<int32_t*>(w + L.keep_pos);
  int32_t* n_keep_all = reinterpret_cast<int32_t*>(w + L.n_keep_all);
  uint8_t* flags = w + L.flags;

  nms_gather_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(reinterpret_cast<const float4*>(boxes), order,
                                                                  sorted, n_total, n);
  if (check_launch("nms_gather_kernel")) return 1;
  nms_mask_kernel<<<dim3(nblk, nblk, B), 64, 0, stream>>>(sorted, mask, n, nblk, thr);
  if (check_launch("nms_mask_kernel")) return 1;
  const size_t smem = (size_t)n * 4;   // kept positions
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    AITB_REQUIRE(e == cudaSuccess, "aitb_nms: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  }
  nms_scan_kernel<<<B, kScanThreads, smem, stream>>>(mask, n, nblk, max_out, mode == 0 ? 1 : 0, resolve_rounds ? 1 : 0, keep_pos,
                                                     n_keep_all);
  if (check_launch("nms_scan_kernel")) return 1;
  if (mode == 0) {
    nms_emit_proposals_kernel<<<dim3((max_out + 127) / 128, B), 128, 0, stream>>>(sorted, keep_pos, n_keep_all, n,
                                                                                 max_out, keep_out, n_keep, rois_out);
    if (check_launch("nms_emit_proposals_kernel")) return 1;
  } else {
    cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)B * n_total, stream);
    AITB_REQUIRE(e == cudaSuccess, "aitb_nms: memset failed: %s", cudaGetErrorString(e));
    nms_mark_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(keep_pos, n_keep_all, order, n, n_total, flags);
    if (check_launch("nms_mark_kernel")) return 1;
    nms_compact_kernel<<<B, 1024, 0, stream>>>(flags, n_total, max_out, keep_out, n_keep);
    if (check_launch("nms_compact_kernel")) return 1;
  }
  return 0;
}

}  // namespace aitb
