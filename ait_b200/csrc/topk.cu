// Descending top-n selection of RPN scores per image, stable (ties: lower index first).
// Replaces `torch.sort(scores, 1, True)` + `[:pre_nms_topN]`
// (lib/model/rpn/proposal_layer.py:129,144-145) in front of the NMS.
//
// n_total is ~21.5k (VOC, 9 anchors) / ~28.7k (COCO, 12 anchors) and only n = 6000 / 12000 are
// consumed, so a full sort is wasted work.  Three small kernels, all exact and deterministic:
//   1. 4096-bin histogram of the 12 most significant bits of an order-preserving key, find the
//      threshold bin that contains the n-th largest score;
//   2. compact the candidates (bin >= threshold) -- order of compaction is irrelevant;
//   3. sort the candidates, packed as (key << 32 | ~index) so that one 64-bit comparison is the total
//      order (score desc, index asc): a bitonic network in shared memory (one CTA per image, up to
//      16384 candidates), or -- beyond that -- exact rank by counting.  Either way the output never
//      depends on the (arbitrary) compaction order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

__device__ __forceinline__ uint32_t order_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;  // -0.0 == +0.0 for the comparison, like torch.sort
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // larger float -> larger key
}

__global__ void __launch_bounds__(1024)
topk_hist_kernel(const float* __restrict__ scores, int n_total, int n, int32_t* __restrict__ thr_bin,
                 int32_t* __restrict__ cand_count, const int32_t* __restrict__ fallback) {
  __shared__ int hist[4096];
  __shared__ int chunk_sum[32];
  const int b = blockIdx.x;
  if (fallback && fallback[b] == 0) {   // done by topk_bucket_kernel: the kernels after this one see 0 candidates
    if (threadIdx.x == 0) { thr_bin[b] = 0x7fffffff; cand_count[b] = 0; }
    return;
  }
  const float* s = scores + (size_t)b * n_total;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_total; i += blockDim.x) atomicAdd(&hist[order_key(s[i]) >> 20], 1);
  __syncthreads();
  // 32 chunks of 128 bins, scanned from the top
  if (threadIdx.x < 32) {
    int t = 0;
    const int c = threadIdx.x;
    for (int i = 0; i < 128; ++i) t += hist[c * 128 + i];
    chunk_sum[c] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int above = 0, c = 31;
    while (c > 0 && above + chunk_sum[c] < n) above += chunk_sum[c--];
    int bin = c * 128 + 127;
    while (bin > c * 128 && above + hist[bin] < n) above += hist[bin--];
    thr_bin[b] = bin;
    cand_count[b] = 0;
  }
}

// candidates are packed as (order_key << 32) | ~index: one 64-bit descending order = score desc, index asc
__global__ void __launch_bounds__(256)
topk_compact_kernel(const float* __restrict__ scores, int n_total, const int32_t* __restrict__ thr_bin,
                    int32_t* __restrict__ cand_count, unsigned long long* __restrict__ cand) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const uint32_t key = order_key(scores[(size_t)b * n_total + i]);
  if ((int)(key >> 20) >= thr_bin[b]) {
    const int pos = atomicAdd(&cand_count[b], 1);
    cand[(size_t)b * n_total + pos] = ((unsigned long long)key << 32) | (unsigned long long)(~(uint32_t)i);
  }
}

// nc <= kSortMax: bitonic sort of the candidates in shared memory, one CTA per image.
static constexpr int kSortMax = 16384;

__global__ void __launch_bounds__(1024)
topk_bitonic_kernel(const int32_t* __restrict__ cand_count, const unsigned long long* __restrict__ cand, int n_total,
                    int n, int max_nc, int64_t* __restrict__ order) {
  extern __shared__ unsigned long long sk[];
  const int b = blockIdx.x;
  const int nc = cand_count[b];
  if (nc > max_nc || nc == 0) return;  // handled by topk_rank_kernel / already done by topk_bucket_kernel
  int S = 1024;
  while (S < nc) S <<= 1;
  const unsigned long long* c = cand + (size_t)b * n_total;
  for (int i = threadIdx.x; i < S; i += blockDim.x) sk[i] = i < nc ? c[i] : 0ULL;
  __syncthreads();
  for (int k = 2; k <= S; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (S >> 1); t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int hi = lo | j;
        const bool desc = (lo & k) == 0;                        // overall descending order
        const unsigned long long a = sk[lo], d = sk[hi];
        if ((a < d) == desc) {
          sk[lo] = d;
          sk[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < n; r += blockDim.x) order[(size_t)b * n + r] = (int64_t)(~(uint32_t)sk[r]);
}

// Exact rank by counting: the position of a candidate in the descending order is the number of candidates that compare
// greater (keys are distinct: the index is part of the key).  Quadratic, but embarrassingly parallel -- nc^2 = 36 M
// compare-and-count steps per image at nc = 6000 spread over every SM finish in a fraction of the time one CTA needs to
// walk the 91 barrier-separated stages of the shared-memory bitonic network (measured 80 us -> see DESIGN), and the
// result cannot depend on the (arbitrary) compaction order.  min_nc: images with fewer candidates are left to the bitonic kernel.
__global__ void __launch_bounds__(256)
topk_rank_kernel(const int32_t* __restrict__ cand_count, const unsigned long long* __restrict__ cand, int n_total,
                 int n, int min_nc, int64_t* __restrict__ order) {
  constexpr int kTile = 2048;
  __shared__ __align__(16) unsigned long long tk[kTile];
  const int b = blockIdx.y;
  const int nc = cand_count[b];
  if (nc <= min_nc || (int)(blockIdx.x * blockDim.x) >= nc) return;  // uniform per CTA
  const unsigned long long* c = cand + (size_t)b * n_total;
  const int me = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = me < nc;
  const unsigned long long mk = live ? c[me] : ~0ULL;
  int rank = 0;
  for (int t0 = 0; t0 < nc; t0 += kTile) {
    __syncthreads();
    for (int j = threadIdx.x; j < kTile; j += blockDim.x) tk[j] = t0 + j < nc ? c[t0 + j] : 0ULL;   // padding never counts
    __syncthreads();
    const int lim = min(kTile, (nc - t0 + 1) & ~1);
#pragma unroll 8
    for (int j = 0; j < lim; j += 2) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&tk[j]);
      rank += (v.x > mk) + (v.y > mk);
    }
  }
  if (live && rank < n) order[(size_t)b * n + rank] = (int64_t)(~(uint32_t)mk);
}

// ---------------------------------------------------------------------------------------------
// Fused top-n for the common case (one CTA per image, everything after the score reads in shared memory):
//   1. 4096-bin histogram of the key's 12 most significant bits + the largest key  -> threshold bin, candidate count nc
//   2. candidates (bin >= threshold) are bucketed by a MONOTONE linear map of the key onto 1024 buckets between the
//      threshold bin's lower bound and the largest key: count, descending exclusive scan, scatter into bucket segments
//   3. exact rank inside the bucket by counting (a bucket holds ~nc / 1024 keys for spread-out scores)
// = one pass of counting sort + tiny insertion ranks instead of the 91 barrier-separated stages of an 8192-key bitonic
// network (80 us -> see DESIGN).  The result is the same total order (score desc, index asc), independent of any atomics'
// arrival order.  Degenerate score distributions (all candidates in one bucket) stay exact, just slower (nc^2 / 1024
// compares per thread).  Images with more than kBucketCap candidates keep their segments in global memory (`spill`).
// ---------------------------------------------------------------------------------------------
static constexpr int kBucketCap = 16384;
static constexpr int kBuckets = 1024;

__global__ void __launch_bounds__(1024)
topk_bucket_kernel(const float* __restrict__ scores, int n_total, int n, int64_t* __restrict__ order,
                   unsigned long long* __restrict__ spill) {
  extern __shared__ __align__(16) unsigned long long seg_s[];  // [kBucketCap]
  __shared__ int hist[4096];
  __shared__ int bcnt[kBuckets], bbase[kBuckets], bfill[kBuckets];
  __shared__ int chunk_sum[32], wsum[32];
  __shared__ unsigned int s_kmax;
  __shared__ int s_thr, s_nc;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* s = scores + (size_t)b * n_total;
  for (int i = tid; i < 4096; i += 1024) hist[i] = 0;
  bcnt[tid] = 0;
  bfill[tid] = 0;
  if (tid == 0) s_kmax = 0u;
  __syncthreads();
  unsigned int kmax = 0u;
#pragma unroll 4
  for (int i = tid; i < n_total; i += 1024) {
    const uint32_t key = order_key(s[i]);
    atomicAdd(&hist[key >> 20], 1);
    kmax = max(kmax, key);
  }
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) atomicMax(&s_kmax, kmax);
  __syncthreads();
  if (tid < 32) {                       // 32 chunks of 128 bins, scanned from the top
    int t = 0;
    for (int i = 0; i < 128; ++i) t += hist[tid * 128 + i];
    chunk_sum[tid] = t;
  }
  __syncthreads();
  if (tid == 0) {
    int above = 0, c = 31;
    while (c > 0 && above + chunk_sum[c] < n) above += chunk_sum[c--];
    int bin = c * 128 + 127;
    while (bin > c * 128 && above + hist[bin] < n) above += hist[bin--];
    s_thr = bin;
    s_nc = above + hist[bin];
  }
  __syncthreads();
  const int thr = s_thr, nc = s_nc;
  // the bucket segments live in shared memory; an image with more than kBucketCap candidates (a threshold bin crowded with
  // equal-ish scores) spills them to the caller's workspace instead -- same algorithm, L2 latency on the rank reads
  unsigned long long* seg = nc <= kBucketCap ? seg_s : spill + (size_t)b * n_total;
  const uint32_t kmin = (uint32_t)thr << 20;
  const uint32_t range = s_kmax - kmin;               // >= 0: the largest key is a candidate
  int shift = 0;
  while ((range >> shift) >= (uint32_t)kBuckets) ++shift;
  // ---- count per bucket
#pragma unroll 4
  for (int i = tid; i < n_total; i += 1024) {
    const uint32_t key = order_key(s[i]);
    if ((int)(key >> 20) >= thr) atomicAdd(&bcnt[(key - kmin) >> shift], 1);
  }
  __syncthreads();
  // ---- descending exclusive scan: bbase[k] = number of candidates in buckets above k
  {
    const int k = kBuckets - 1 - tid;                 // thread 0 owns the top bucket
    const int v = bcnt[k];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;                            // exclusive prefix of the warp totals
    }
    __syncthreads();
    bbase[k] = wsum[warp] + incl - v;
  }
  __syncthreads();
  // ---- scatter into the bucket segments (order inside a bucket is arbitrary; the ranks below do not depend on it)
#pragma unroll 4
  for (int i = tid; i < n_total; i += 1024) {
    const uint32_t key = order_key(s[i]);
    if ((int)(key >> 20) >= thr) {
      const int k = (int)((key - kmin) >> shift);
      const int slot = bbase[k] + atomicAdd(&bfill[k], 1);
      seg[slot] = ((unsigned long long)key << 32) | (unsigned long long)(~(uint32_t)i);
    }
  }
  __syncthreads();
  // ---- exact rank inside the bucket
  for (int p = tid; p < nc; p += 1024) {
    const unsigned long long mk = seg[p];
    const int k = (int)(((uint32_t)(mk >> 32) - kmin) >> shift);
    const int lo = bbase[k], hi = lo + bcnt[k];
    int rank = lo;
    for (int q = lo; q < hi; ++q) rank += seg[q] > mk;
    if (rank < n) order[(size_t)b * n + rank] = (int64_t)(~(uint32_t)mk);
  }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t topk_workspace_bytes(int B, int n_total, int n) {
  (void)n;
  return al256((size_t)B * 4) * 3 + al256((size_t)B * n_total * 8);
}

int topk_run(const float* scores, int B, int n_total, int n, int64_t* order, void* ws, size_t ws_bytes,
             cudaStream_t stream) {
  AITB_REQUIRE(B > 0 && n > 0 && n_total >= n, "aitb_topk_desc: bad sizes B=%d n_total=%d n=%d", B, n_total, n);
  AITB_REQUIRE(scores && order && ws, "aitb_topk_desc: null pointer");
  AITB_REQUIRE(ws_bytes >= topk_workspace_bytes(B, n_total, n), "aitb_topk_desc: workspace too small");
  AITB_REQUIRE(B <= 65535, "aitb_topk_desc: B too large");
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  int32_t* thr_bin = reinterpret_cast<int32_t*>(w);
  w += al256((size_t)B * 4);
  int32_t* cand_count = reinterpret_cast<int32_t*>(w);
  w += al256((size_t)B * 4);
  int32_t* fallback = reinterpret_cast<int32_t*>(w);
  w += al256((size_t)B * 4);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(w);
  // the fused bucket kernel (default); AITB_TOPK_NO_BUCKET=1: the round-1 histogram / compact / bitonic(+rank) kernels (A/B)
  static const bool use_bucket = getenv("AITB_TOPK_NO_BUCKET") == nullptr;
  if (use_bucket) {
    static SmemAttrOnce bonce;
    const int bsmem = kBucketCap * 8;
    if (ensure_dyn_smem((const void*)topk_bucket_kernel, bsmem, bonce, "topk_bucket_kernel")) return 1;
    topk_bucket_kernel<<<B, 1024, bsmem, stream>>>(scores, n_total, n, order, cand);
    return check_launch("topk_bucket_kernel");
  }
  (void)fallback;
  topk_hist_kernel<<<B, 1024, 0, stream>>>(scores, n_total, n, thr_bin, cand_count, nullptr);
  if (check_launch("topk_hist_kernel")) return 1;
  topk_compact_kernel<<<dim3((n_total + 255) / 256, B), 256, 0, stream>>>(scores, n_total, thr_bin, cand_count, cand);
  if (check_launch("topk_compact_kernel")) return 1;
  // ordering: rank-by-counting across all SMs (default); the one-CTA-per-image shared-memory bitonic network stays as the
  // A/B path (AITB_TOPK_BITONIC=1: it then takes every image with <= kSortMax candidates, the rank kernel the rest)
  static const bool use_bitonic = getenv("AITB_TOPK_RANK") == nullptr;   // default: bitonic (measured 100 us vs 177 us for rank counting, 8 images x 6000 of 21546)
  const int split_nc = use_bitonic ? kSortMax : 0;   // images with nc <= split_nc -> bitonic, others -> rank
  if (use_bitonic) {
    static SmemAttrOnce once;
    const int smem = kSortMax * 8;
    if (ensure_dyn_smem((const void*)topk_bitonic_kernel, smem, once, "topk_bitonic_kernel")) return 1;
    topk_bitonic_kernel<<<B, 1024, smem, stream>>>(cand_count, cand, n_total, n, split_nc, order);
    if (check_launch("topk_bitonic_kernel")) return 1;
  }
  if (!use_bitonic || n_total > kSortMax) {
    topk_rank_kernel<<<dim3((n_total + 255) / 256, B), 256, 0, stream>>>(cand_count, cand, n_total, n, split_nc, order);
    if (check_launch("topk_rank_kernel")) return 1;
  }
  return 0;
}

}  // namespace aitb
