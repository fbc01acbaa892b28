// Descending top-n selection of RPN scores per image, stable (ties: lower index first).
// Replaces `torch.sort(scores, 1, True)` + `[:pre_nms_topN]`
// (lib/model/rpn/proposal_layer.py:129,144-145) in front of the NMS.
//
// n_total is ~21.5k (VOC, 9 anchors) / ~28.7k (COCO, 12 anchors) and only n = 6000 / 12000 are
// consumed, so a full sort is wasted work.  Three small kernels, all exact and deterministic:
//   1. 4096-bin histogram of the 12 most significant bits of an order-preserving key, find the
//      threshold bin that contains the n-th largest score;
//   2. compact the candidates (bin >= threshold) -- order of compaction is irrelevant;
//   3. exact rank of every candidate by counting (shared-memory tiles); rank < n scatters the
//      original index to order[rank].  Ranks are a total order (key desc, index asc), so the
//      output never depends on atomics ordering.
#include <cuda_runtime.h>
#include <stdint.h>

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
                 int32_t* __restrict__ cand_count) {
  __shared__ int hist[4096];
  __shared__ int chunk_sum[32];
  const int b = blockIdx.x;
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

__global__ void __launch_bounds__(256)
topk_compact_kernel(const float* __restrict__ scores, int n_total, const int32_t* __restrict__ thr_bin,
                    int32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_key, int32_t* __restrict__ cand_idx) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const uint32_t key = order_key(scores[(size_t)b * n_total + i]);
  if ((int)(key >> 20) >= thr_bin[b]) {
    const int pos = atomicAdd(&cand_count[b], 1);
    cand_key[(size_t)b * n_total + pos] = key;
    cand_idx[(size_t)b * n_total + pos] = i;
  }
}

__global__ void __launch_bounds__(256)
topk_rank_kernel(const int32_t* __restrict__ cand_count, const uint32_t* __restrict__ cand_key,
                 const int32_t* __restrict__ cand_idx, int n_total, int n, int64_t* __restrict__ order) {
  __shared__ uint32_t tk[1024];
  __shared__ int32_t ti[1024];
  const int b = blockIdx.y;
  const int nc = cand_count[b];
  if ((int)(blockIdx.x * blockDim.x) >= nc) return;  // uniform per CTA
  const uint32_t* ck = cand_key + (size_t)b * n_total;
  const int32_t* ci = cand_idx + (size_t)b * n_total;
  const int me = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = me < nc;
  const uint32_t mk = live ? ck[me] : 0u;
  const int32_t mi = live ? ci[me] : 0;
  int rank = 0;
  for (int t0 = 0; t0 < nc; t0 += 1024) {
    __syncthreads();
    for (int j = threadIdx.x; j < 1024; j += blockDim.x) {
      const bool in = t0 + j < nc;
      tk[j] = in ? ck[t0 + j] : 0u;
      ti[j] = in ? ci[t0 + j] : 0x7fffffff;
    }
    __syncthreads();
    const int lim = min(1024, nc - t0);
#pragma unroll 8
    for (int j = 0; j < lim; ++j) rank += (tk[j] > mk) || (tk[j] == mk && ti[j] < mi);
  }
  if (live && rank < n) order[(size_t)b * n + rank] = mi;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t topk_workspace_bytes(int B, int n_total, int n) {
  (void)n;
  return al256((size_t)B * 4) * 2 + al256((size_t)B * n_total * 4) * 2;
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
  uint32_t* cand_key = reinterpret_cast<uint32_t*>(w);
  w += al256((size_t)B * n_total * 4);
  int32_t* cand_idx = reinterpret_cast<int32_t*>(w);
  topk_hist_kernel<<<B, 1024, 0, stream>>>(scores, n_total, n, thr_bin, cand_count);
  if (check_launch("topk_hist_kernel")) return 1;
  topk_compact_kernel<<<dim3((n_total + 255) / 256, B), 256, 0, stream>>>(scores, n_total, thr_bin, cand_count,
                                                                         cand_key, cand_idx);
  if (check_launch("topk_compact_kernel")) return 1;
  topk_rank_kernel<<<dim3((n_total + 255) / 256, B), 256, 0, stream>>>(cand_count, cand_key, cand_idx, n_total, n,
                                                                      order);
  return check_launch("topk_rank_kernel");
}

}  // namespace aitb
