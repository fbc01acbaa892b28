// Training-only samplers and losses around the head, on the device:
//
//   anchor target layer    lib/model/rpn/anchor_target_layer.py:49-199 (_AnchorTargetLayer.forward)
//   proposal target layer  lib/model/rpn/proposal_target_layer_cascade.py:33-220 (_ProposalTargetLayer.forward)
//   RPN losses             lib/model/rpn/rpn.py:99-126 (cross entropy over the sampled anchors, smooth L1 sigma 3)
//   detection losses       lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:340-361 (cross entropy,
//                          3 * MarginRankingLoss over the pairwise |p_i - p_j| map, smooth L1) + their gradients
//
// Both samplers are two-phase: an ASSIGN phase (IoU of every anchor / roi with every ground-truth box, max / argmax,
// candidate classes and per-image candidate counts) and a FINISH phase that applies a random selection expressed in
// RANKS within the image's ascending candidate lists -- exactly what the reference's `fg_inds[rand_num[...]]` indexing
// means.  The host draws the ranks (from numpy's stream, call for call like the reference, when bit parity with it
// is wanted); the candidate lists themselves never leave the device: the finish kernels recover rank -> index with a
// block scan.  IoU / threshold arithmetic uses explicitly rounded fp32 operations in the reference's order
// (bbox_transform.py:167-257), so every label decision is bit-identical; only log() in the regression targets
// differs from the host libm by an ulp.
//
// HBM-bound integer/byte work: one thread per anchor, ground-truth boxes in shared memory, coalesced NCHW stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kMaxGt = 128;   // ground-truth boxes per image held in shared memory

struct GtBox { float x1, y1, x2, y2, gx, gy, area; int zero; };

__device__ __forceinline__ void load_gt(GtBox* s, const float* __restrict__ gt, int K) {
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    GtBox g;
    g.x1 = gt[k * 5 + 0]; g.y1 = gt[k * 5 + 1]; g.x2 = gt[k * 5 + 2]; g.y2 = gt[k * 5 + 3];
    g.gx = __fadd_rn(__fsub_rn(g.x2, g.x1), 1.f);        // bbox_transform.py:185-186
    g.gy = __fadd_rn(__fsub_rn(g.y2, g.y1), 1.f);
    g.area = __fmul_rn(g.gx, g.gy);
    g.zero = (g.gx == 1.f) && (g.gy == 1.f);             // :193
    s[k] = g;
  }
}

// one entry of bbox_overlaps_batch (bbox_transform.py:196-212)
__device__ __forceinline__ float overlap(float4 a, float a_area, bool a_zero, const GtBox& g) {
  float iw = __fadd_rn(__fsub_rn(fminf(a.z, g.x2), fmaxf(a.x, g.x1)), 1.f);
  float ih = __fadd_rn(__fsub_rn(fminf(a.w, g.y2), fmaxf(a.y, g.y1)), 1.f);
  if (iw < 0.f) iw = 0.f;
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  const float ua = __fsub_rn(__fadd_rn(a_area, g.area), inter);
  float ov = __fdiv_rn(inter, ua);
  if (g.zero) ov = 0.f;
  if (a_zero) ov = -1.f;
  return ov;
}

// order-preserving float -> uint32 (for atomicMax)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bbox_transform_batch (bbox_transform.py:38-75) for one (ex, gt) pair
__device__ __forceinline__ float4 box_targets(float4 ex, float gx1, float gy1, float gx2, float gy2) {
  const float ew = __fadd_rn(__fsub_rn(ex.z, ex.x), 1.f), eh = __fadd_rn(__fsub_rn(ex.w, ex.y), 1.f);
  const float ecx = __fadd_rn(ex.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(ex.y, __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(gx2, gx1), 1.f), gh = __fadd_rn(__fsub_rn(gy2, gy1), 1.f);
  const float gcx = __fadd_rn(gx1, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(gy1, __fmul_rn(0.5f, gh));
  float4 t;
  t.x = __fdiv_rn(__fsub_rn(gcx, ecx), ew);
  t.y = __fdiv_rn(__fsub_rn(gcy, ecy), eh);
  t.z = logf(__fdiv_rn(gw, ew));
  t.w = logf(__fdiv_rn(gh, eh));
  return t;
}

__device__ __forceinline__ float4 anchor_at(const float* __restrict__ base, int i, int A, int W, float stride) {
  const int cell = i / A, a = i - cell * A;
  const int y = cell / W, x = cell - y * W;
  const float sx = (float)x * stride, sy = (float)y * stride;
  return make_float4(__fadd_rn(base[a * 4 + 0], sx), __fadd_rn(base[a * 4 + 1], sy), __fadd_rn(base[a * 4 + 2], sx),
                     __fadd_rn(base[a * 4 + 3], sy));
}
// anchor_target_layer.py:85-88 with _allowed_border = 0; the limits come from im_info[0] for every image
__device__ __forceinline__ bool anchor_inside(float4 an, const float* __restrict__ im_info) {
  const float lim_w = (float)(long long)im_info[1], lim_h = (float)(long long)im_info[0];
  return an.x >= 0.f && an.y >= 0.f && an.z < lim_w && an.w < lim_h;
}

// ---------------------------------------------------------------------------------------------
// anchor target, phase 1a: per inside anchor max / argmax over the gt boxes (first maximum, like torch.max on the
// host); per gt box the maximum over the inside anchors (atomicMax on order-preserving keys)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
anchor_overlap_kernel(const float* __restrict__ base, const float* __restrict__ gt_boxes, const float* __restrict__ im_info,
                      int A, int H, int W, int K, float stride, float* __restrict__ max_ov, int32_t* __restrict__ argmax,
                      unsigned* __restrict__ gt_max) {
  __shared__ GtBox sg[kMaxGt];
  __shared__ unsigned smax[kMaxGt];
  const int b = blockIdx.y;
  load_gt(sg, gt_boxes + (size_t)b * K * 5, K);
  for (int k = threadIdx.x; k < K; k += blockDim.x) smax[k] = 0u;
  __syncthreads();
  const int total = A * H * W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    const float4 an = anchor_at(base, i, A, W, stride);
    float best = -INFINITY;
    int arg = -1;
    if (anchor_inside(an, im_info)) {
      const float ax = __fadd_rn(__fsub_rn(an.z, an.x), 1.f), ay = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
      const float a_area = __fmul_rn(ax, ay);
      const bool a_zero = (ax == 1.f) && (ay == 1.f);
      for (int k = 0; k < K; ++k) {
        const float ov = overlap(an, a_area, a_zero, sg[k]);
        if (ov > best) { best = ov; arg = k; }
        atomicMax(&smax[k], f2ord(ov));
      }
    }
    max_ov[(size_t)b * total + i] = best;
    argmax[(size_t)b * total + i] = arg;      // -1: outside the image
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (smax[k]) atomicMax(&gt_max[(size_t)b * K + k], smax[k]);
}

// phase 1b: labels before sub-sampling (anchor_target_layer.py:110-123) + per-image fg / bg counts
__global__ void __launch_bounds__(256)
anchor_label_kernel(const float* __restrict__ base, const float* __restrict__ gt_boxes, int A, int H, int W, int K,
                    float stride, float neg_thr, float pos_thr, int clobber, const float* __restrict__ max_ov,
                    const int32_t* __restrict__ argmax, const unsigned* __restrict__ gt_max, int8_t* __restrict__ labels,
                    int32_t* __restrict__ counts) {
  __shared__ GtBox sg[kMaxGt];
  __shared__ float sgm[kMaxGt];
  __shared__ int s_cnt[2];
  const int b = blockIdx.y;
  load_gt(sg, gt_boxes + (size_t)b * K * 5, K);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const unsigned u = gt_max[(size_t)b * K + k];
    float m = u ? ord2f(u) : 0.f;
    if (m == 0.f) m = 1e-5f;                               // :113
    sgm[k] = m;
  }
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int total = A * H * W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lab = -1;
  if (i < total && argmax[(size_t)b * total + i] >= 0) {
    const float4 an = anchor_at(base, i, A, W, stride);
    const float ax = __fadd_rn(__fsub_rn(an.z, an.x), 1.f), ay = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
    const float a_area = __fmul_rn(ax, ay);
    const bool a_zero = (ax == 1.f) && (ay == 1.f);
    const float mo = max_ov[(size_t)b * total + i];
    if (!clobber && mo < neg_thr) lab = 0;
    bool hit = false;
    for (int k = 0; k < K; ++k) hit |= (overlap(an, a_area, a_zero, sg[k]) == sgm[k]);
    if (hit) lab = 1;
    if (mo >= pos_thr) lab = 1;
    if (clobber && mo < neg_thr) lab = 0;
  }
  if (i < total) labels[(size_t)b * total + i] = (int8_t)lab;
  const unsigned fg = __ballot_sync(0xffffffffu, lab == 1), bg = __ballot_sync(0xffffffffu, lab == 0);
  if ((threadIdx.x & 31) == 0) {
    if (fg) atomicAdd(&s_cnt[0], __popc(fg));
    if (bg) atomicAdd(&s_cnt[1], __popc(bg));
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[b * 2 + threadIdx.x], s_cnt[threadIdx.x]);
}

// block-wide exclusive scan of one int per thread (1024 threads max); returns the exclusive prefix, total in *sum
__device__ __forceinline__ int block_exscan(int v, int* s_warp, int* sum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int before = warp ? s_warp[warp - 1] : 0;
  *sum = s_warp[nw - 1];
  __syncthreads();
  return before + inc - v;
}

// phase 2a: apply the sub-sampling.  drop [B, 2, ld] bytes: drop[b][0][r] != 0 disables the r-th foreground anchor of
// image b (r = rank in ascending anchor order = position in the reference's `fg_inds`), drop[b][1][r] the r-th
// background anchor.  One CTA per image; n_examples[b] = anchors left with label >= 0.
__global__ void __launch_bounds__(1024)
anchor_subsample_kernel(int8_t* __restrict__ labels, int total, const uint8_t* __restrict__ drop, int ld,
                        int32_t* __restrict__ n_examples) {
  __shared__ int s_warp[32];
  __shared__ int s_left;
  const int b = blockIdx.x;
  int8_t* lab = labels + (size_t)b * total;
  const uint8_t* dfg = drop ? drop + (size_t)b * 2 * ld : nullptr;
  const uint8_t* dbg = drop ? dfg + ld : nullptr;
  if (threadIdx.x == 0) s_left = 0;
  const int per = (total + blockDim.x - 1) / blockDim.x;
  const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
  int nf = 0, nb = 0;
  for (int i = lo; i < hi; ++i) { nf += lab[i] == 1; nb += lab[i] == 0; }
  int tot_f, tot_b;
  int rf = block_exscan(nf, s_warp, &tot_f);
  int rb = block_exscan(nb, s_warp, &tot_b);
  int left = 0;
  for (int i = lo; i < hi; ++i) {
    const int l = lab[i];
    if (l == 1) {
      if (dfg && rf < ld && dfg[rf]) lab[i] = -1; else ++left;
      ++rf;
    } else if (l == 0) {
      if (dbg && rb < ld && dbg[rb]) lab[i] = -1; else ++left;
      ++rb;
    }
  }
  if (left) atomicAdd(&s_left, left);
  __syncthreads();
  if (threadIdx.x == 0) n_examples[b] = s_left;
}

// ---------------------------------------------------------------------------------------------
// Device-side sampling (no host round trip): the same selections drawn from a counter-based generator instead of
// numpy's stream.  Philox4x32-10 keyed by (seed, image, list, candidate index) gives every candidate an independent
// uniform 32-bit key; "a uniformly random subset of size k" = the k smallest keys (radix select in shared memory),
// "rand * n with replacement" = a key per output slot.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cand_key(uint64_t seed, int image, int list, int idx) {
  return philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)image, (uint32_t)list, 0x5eedu),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x;
}

// keep the `keep` candidates (lab[i] == L) with the smallest keys, set the others to -1.  Whole CTA; s_hist[256], s_sel[3].
__device__ void keep_smallest_keys(int8_t* __restrict__ lab, int total, int L, int n_cand, int keep, uint64_t seed, int image,
                                   int* s_hist, int* s_sel) {
  if (n_cand <= keep) return;                       // CTA-uniform
  uint32_t prefix = 0;
  int remaining = keep;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x)
      if (lab[i] == L) {
        const uint32_t k = cand_key(seed, image, L, i);
        if (pass == 3 || (k >> (8 * (pass + 1))) == prefix) atomicAdd(&s_hist[(k >> (8 * pass)) & 255u], 1);
      }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, b = 0;
      for (; b < 256; ++b) {
        if (cum + s_hist[b] >= remaining) break;
        cum += s_hist[b];
      }
      s_sel[0] = b;
      s_sel[1] = remaining - cum;                  // how many to take from bin b
    }
    __syncthreads();
    prefix = (prefix << 8) | (uint32_t)s_sel[0];
    remaining = s_sel[1];
    __syncthreads();
  }
  // keys < prefix stay; keys == prefix: `remaining` of them (32-bit ties are ~n^2 / 2^33 rare; any of them is a valid draw)
  if (threadIdx.x == 0) s_sel[2] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    if (lab[i] == L) {
      const uint32_t k = cand_key(seed, image, L, i);
      if (k > prefix || (k == prefix && atomicAdd(&s_sel[2], 1) >= remaining)) lab[i] = -1;
    }
  __syncthreads();
}

// anchor target sub-sampling on the device (anchor_target_layer.py:130-152): at most num_fg foreground anchors, then
// rpn_batch - (#fg kept) background anchors.  One CTA per image; counts [B, 2] from the assign phase.
__global__ void __launch_bounds__(1024)
anchor_subsample_device_kernel(int8_t* __restrict__ labels, int total, const int32_t* __restrict__ counts, int num_fg, int rpn_batch,
                               uint64_t seed) {
  __shared__ int s_hist[256];
  __shared__ int s_sel[3];
  const int b = blockIdx.x;
  int8_t* lab = labels + (size_t)b * total;
  const int n_f = counts[b * 2], n_b = counts[b * 2 + 1];
  keep_smallest_keys(lab, total, 1, n_f, num_fg, seed, b, s_hist, s_sel);
  const int fg_left = min(n_f, num_fg);
  keep_smallest_keys(lab, total, 0, n_b, rpn_batch - fg_left, seed, b, s_hist, s_sel);
}

// proposal target picks on the device (proposal_target_layer_cascade.py:154-199): fg_this = min(fg_per_image, nf)
// distinct foreground ranks in random order (rank of the rank's key among all nf keys), the other slots background
// ranks with replacement (floor(u * nb)); foreground-only / background-only images fill all S slots with replacement.
// picks [B, S], n_fg_pick [B]; *bad = 1 for an image with no candidate at all (the reference raises there).
__global__ void __launch_bounds__(1024)
proposal_picks_device_kernel(const int32_t* __restrict__ counts, int S, int fg_per_image, uint64_t seed, int32_t* __restrict__ picks,
                             int32_t* __restrict__ n_fg_pick, int32_t* __restrict__ bad) {
  const int b = blockIdx.x;
  const int nf = counts[b * 2], nb = counts[b * 2 + 1];
  int fg_this;
  if (nf > 0 && nb > 0) fg_this = min(fg_per_image, nf);
  else if (nf > 0) fg_this = S;
  else fg_this = 0;
  if (threadIdx.x == 0) {
    n_fg_pick[b] = fg_this;
    if (nf == 0 && nb == 0) *bad = 1;
  }
  int32_t* pk = picks + (size_t)b * S;
  if (nf == 0 && nb == 0) {
    for (int j = threadIdx.x; j < S; j += blockDim.x) pk[j] = -1;
    return;
  }
  const bool distinct_fg = nf > 0 && nb > 0;
  __shared__ uint32_t s_key[4096];
  if (distinct_fg) {
    // position of rank r in the key order; the first fg_this positions are the sample (a uniform permutation prefix)
    const bool cached = nf <= 4096;
    if (cached)
      for (int r = threadIdx.x; r < nf; r += blockDim.x) s_key[r] = cand_key(seed, b, 2, r);
    __syncthreads();
    for (int r = threadIdx.x; r < nf; r += blockDim.x) {
      const uint32_t kr = cached ? s_key[r] : cand_key(seed, b, 2, r);
      int pos = 0;
      for (int q = 0; q < nf; ++q) {
        const uint32_t kq = cached ? s_key[q] : cand_key(seed, b, 2, q);
        pos += (kq < kr) || (kq == kr && q < r);
      }
      if (pos < fg_this) pk[pos] = r;
    }
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const bool is_fg = j < fg_this;
    if (is_fg && distinct_fg) continue;
    const int n = is_fg ? nf : nb;
    const uint32_t u = cand_key(seed, b, 3, j);
    pk[j] = (int)(((uint64_t)u * (uint64_t)n) >> 32);         // floor(u / 2^32 * n)
  }
}

// phase 2b: the four outputs in the reference's layouts (anchor_target_layer.py:155-197):
//   labels [B, 1, A*H, W] (index a*H*W + cell), bbox_targets / inside / outside weights [B, 4A, H, W] (channel a*4 + j)
// outside weight = 1 / (examples of the LAST image) for every sampled anchor (:161-168, RPN_POSITIVE_WEIGHT < 0)
__global__ void __launch_bounds__(256)
anchor_emit_kernel(const float* __restrict__ base, const float* __restrict__ gt_boxes, int A, int H, int W, int K, float stride,
                   const int8_t* __restrict__ labels, const int32_t* __restrict__ argmax, const int32_t* __restrict__ n_examples,
                   int B, float inside_w, float* __restrict__ labels_out, float* __restrict__ targets, float* __restrict__ inside,
                   float* __restrict__ outside) {
  const int b = blockIdx.y;
  const int cells = H * W, total = A * cells;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;     // = a * cells + cell: consecutive threads, consecutive cells
  if (o >= total) return;
  const int a = o / cells, cell = o - a * cells;
  const int i = cell * A + a;
  const int lab = labels[(size_t)b * total + i];
  const int arg = argmax[(size_t)b * total + i];
  labels_out[(size_t)b * total + o] = (float)lab;
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (arg >= 0) {
    const float* g = gt_boxes + ((size_t)b * K + arg) * 5;
    t = box_targets(anchor_at(base, i, A, W, stride), g[0], g[1], g[2], g[3]);
  }
  const float w_in = lab == 1 ? inside_w : 0.f;
  const float w_out = lab >= 0 ? (float)(1.0 / (double)n_examples[B - 1]) : 0.f;   // python double division, then float32
  const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t q = ((size_t)b * 4 * A + a * 4 + j) * cells + cell;
    targets[q] = tv[j];
    inside[q] = w_in;
    outside[q] = w_out;
  }
}

// ---------------------------------------------------------------------------------------------
// proposal target, phase 1: candidates = rois ++ gt boxes (proposal_target_layer_cascade.py:41-45); per candidate
// max / argmax overlap and its class: 1 foreground (>= fg_thr), 0 background ([bg_lo, bg_hi)), -1 neither
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 candidate_box(const float* __restrict__ rois, const float* __restrict__ gt, int R, int i) {
  if (i < R) return make_float4(rois[i * 5 + 1], rois[i * 5 + 2], rois[i * 5 + 3], rois[i * 5 + 4]);
  const float* g = gt + (size_t)(i - R) * 5;
  return make_float4(g[0], g[1], g[2], g[3]);
}

__global__ void __launch_bounds__(256)
proposal_assign_kernel(const float* __restrict__ rois, const float* __restrict__ gt_boxes, int R, int K, float fg_thr,
                       float bg_hi, float bg_lo, float* __restrict__ max_ov, int32_t* __restrict__ assign,
                       int8_t* __restrict__ cls, int32_t* __restrict__ counts) {
  __shared__ GtBox sg[kMaxGt];
  __shared__ int s_cnt[2];
  const int b = blockIdx.y;
  const float* gt = gt_boxes + (size_t)b * K * 5;
  load_gt(sg, gt, K);
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int N = R + K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = -1;
  if (i < N) {
    const float4 bx = candidate_box(rois + (size_t)b * R * 5, gt, R, i);
    const float ax = __fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), ay = __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f);
    const float a_area = __fmul_rn(ax, ay);
    const bool a_zero = (ax == 1.f) && (ay == 1.f);
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      const float ov = overlap(bx, a_area, a_zero, sg[k]);
      if (ov > best) { best = ov; arg = k; }
    }
    if (best >= fg_thr) c = 1;
    else if (best < bg_hi && best >= bg_lo) c = 0;
    max_ov[(size_t)b * N + i] = best;
    assign[(size_t)b * N + i] = arg;
    cls[(size_t)b * N + i] = (int8_t)c;
  }
  const unsigned fg = __ballot_sync(0xffffffffu, c == 1), bg = __ballot_sync(0xffffffffu, c == 0);
  if ((threadIdx.x & 31) == 0) {
    if (fg) atomicAdd(&s_cnt[0], __popc(fg));
    if (bg) atomicAdd(&s_cnt[1], __popc(bg));
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[b * 2 + threadIdx.x], s_cnt[threadIdx.x]);
}

// phase 2: picks [B, S] = ranks within the image's ascending foreground list for the first n_fg_pick[b] slots and
// within its background list for the rest (the reference's `fg_inds[rand_num[:n]]` ++ `bg_inds[rand_num]`,
// :154-214); outputs rois [B,S,5], labels [B,S], bbox_targets / inside / outside weights [B,S,4] (:84-121).
// lists: workspace [B, 2, N] int32 (rank -> candidate index).  One CTA per image.
__global__ void __launch_bounds__(1024)
proposal_sample_kernel(const float* __restrict__ rois, const float* __restrict__ gt_boxes, int R, int K,
                       const int8_t* __restrict__ cls, const int32_t* __restrict__ assign, const int32_t* __restrict__ picks,
                       const int32_t* __restrict__ n_fg_pick, int S, float4 means, float4 stds, float4 inw,
                       int32_t* __restrict__ lists, float* __restrict__ rois_out, float* __restrict__ labels_out,
                       float* __restrict__ targets, float* __restrict__ inside, float* __restrict__ outside,
                       int32_t* __restrict__ bad) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x;
  const int N = R + K;
  const int8_t* c = cls + (size_t)b * N;
  int32_t* lf = lists + (size_t)b * 2 * N;
  int32_t* lb = lf + N;
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
  int nf = 0, nb = 0;
  for (int i = lo; i < hi; ++i) { nf += c[i] == 1; nb += c[i] == 0; }
  int tot_f, tot_b;
  int rf = block_exscan(nf, s_warp, &tot_f);
  int rb = block_exscan(nb, s_warp, &tot_b);
  for (int i = lo; i < hi; ++i) {
    if (c[i] == 1) lf[rf++] = i;
    else if (c[i] == 0) lb[rb++] = i;
  }
  __syncthreads();
  const float* gt = gt_boxes + (size_t)b * K * 5;
  const int nfp = n_fg_pick[b];
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const bool is_fg = j < nfp;
    const int r = picks[(size_t)b * S + j];
    const int n_list = is_fg ? tot_f : tot_b;
    if (r < 0 || r >= n_list) { atomicExch(bad, 1); continue; }   // inconsistent picks: flagged, row left untouched
    const int i = is_fg ? lf[r] : lb[r];
    const float4 bx = candidate_box(rois + (size_t)b * R * 5, gt, R, i);
    const float* g = gt + (size_t)assign[(size_t)b * N + i] * 5;
    const float lab = is_fg ? g[4] : 0.f;                         // labels_batch[i][fg_rois_per_this_image:] = 0
    float* ro = rois_out + ((size_t)b * S + j) * 5;
    ro[0] = (float)b; ro[1] = bx.x; ro[2] = bx.y; ro[3] = bx.z; ro[4] = bx.w;
    labels_out[(size_t)b * S + j] = lab;
    float4 t = box_targets(bx, g[0], g[1], g[2], g[3]);
    t.x = __fdiv_rn(__fsub_rn(t.x, means.x), stds.x); t.y = __fdiv_rn(__fsub_rn(t.y, means.y), stds.y);
    t.z = __fdiv_rn(__fsub_rn(t.z, means.z), stds.z); t.w = __fdiv_rn(__fsub_rn(t.w, means.w), stds.w);
    const bool on = lab > 0.f;
    const size_t q = ((size_t)b * S + j) * 4;
    reinterpret_cast<float4*>(targets + q)[0] = on ? t : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 wi = on ? inw : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(inside + q)[0] = wi;
    reinterpret_cast<float4*>(outside + q)[0] = make_float4(wi.x > 0.f ? 1.f : 0.f, wi.y > 0.f ? 1.f : 0.f,
                                                            wi.z > 0.f ? 1.f : 0.f, wi.w > 0.f ? 1.f : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// losses
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < nw ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

// _smooth_l1_loss element (net_utils.py:75-85): value and d(value)/d(pred)
__device__ __forceinline__ float smooth_l1_elem(float pred, float tgt, float w_in, float w_out, float s2, float* dpred) {
  const float d = w_in * (pred - tgt);
  const float ad = fabsf(d);
  float v, g;
  if (ad < 1.f / s2) { v = d * d * (s2 * 0.5f); g = s2 * d; }
  else { v = ad - 0.5f / s2; g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
  *dpred = w_out * g * w_in;
  return w_out * v;
}

// two-class cross entropy of one row: value, p1 = softmax[1]
__device__ __forceinline__ float ce2(float s0, float s1, int label, float* p1) {
  const float m = fmaxf(s0, s1);
  const float e0 = expf(s0 - m), e1 = expf(s1 - m);
  const float den = e0 + e1;
  *p1 = e1 / den;
  return (m + logf(den)) - (label ? s1 : s0);
}

// RPN losses, pass 1 (rpn.py:103-124): acc[0] = sum of cross entropies over anchors with label != -1, acc[1] = their
// number, acc[2] = sum of the weighted smooth-L1 terms.  score [B, 2A, H, W]: bg of anchor (a, cell) in channel a, fg in
// channel A + a (the `reshape(x, 2)` view, rpn.py:70); labels [B, A*H*W] indexed a*H*W + cell; bbox tensors [B, 4A, H, W].
__global__ void __launch_bounds__(256)
rpn_loss_reduce_kernel(const float* __restrict__ score, const float* __restrict__ bbox_pred, const float* __restrict__ labels,
                       const float* __restrict__ targets, const float* __restrict__ inside, const float* __restrict__ outside,
                       int B, int A, int cells, float s2, double* __restrict__ acc) {
  __shared__ double s_red[32];
  const size_t n = (size_t)B * A * cells;
  double ce = 0.0, cnt = 0.0, box = 0.0;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(o / ((size_t)A * cells));
    const int r = (int)(o - (size_t)b * A * cells);
    const int a = r / cells, cell = r - a * cells;
    const float lab = labels[o];
    if (lab != -1.f) {
      float p1;
      ce += ce2(score[((size_t)b * 2 * A + a) * cells + cell], score[((size_t)b * 2 * A + A + a) * cells + cell], lab != 0.f, &p1);
      cnt += 1.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t q = ((size_t)b * 4 * A + a * 4 + j) * cells + cell;
      const float wo = outside[q];
      if (wo != 0.f) { float g; box += smooth_l1_elem(bbox_pred[q], targets[q], inside[q], wo, s2, &g); }
    }
  }
  ce = block_sum(ce, s_red);
  cnt = block_sum(cnt, s_red);
  box = block_sum(box, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, ce);
    atomicAdd(acc + 1, cnt);
    atomicAdd(acc + 2, box);
  }
}

// pass 2: losses[0] = cross entropy (mean over the sampled anchors), losses[1] = smooth L1 (sum per image, mean over
// the batch); gradients of gscale[0]*losses[0] + gscale[1]*losses[1] w.r.t. score and bbox_pred (gscale NULL = 1, 1)
__global__ void __launch_bounds__(256)
rpn_loss_grad_kernel(const float* __restrict__ score, const float* __restrict__ bbox_pred, const float* __restrict__ labels,
                     const float* __restrict__ targets, const float* __restrict__ inside, const float* __restrict__ outside,
                     int B, int A, int cells, float s2, const double* __restrict__ acc, const float* __restrict__ gscale,
                     float* __restrict__ losses, float* __restrict__ d_score, float* __restrict__ d_bbox) {
  const double cnt = acc[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    losses[0] = (float)(acc[0] / cnt);
    losses[1] = (float)(acc[2] / (double)B);
  }
  if (!d_score && !d_bbox) return;
  const float g0 = (gscale ? gscale[0] : 1.f) / (float)cnt, g1 = (gscale ? gscale[1] : 1.f) / (float)B;
  const size_t n = (size_t)B * A * cells;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(o / ((size_t)A * cells));
    const int r = (int)(o - (size_t)b * A * cells);
    const int a = r / cells, cell = r - a * cells;
    if (d_score) {
      const size_t q0 = ((size_t)b * 2 * A + a) * cells + cell, q1 = ((size_t)b * 2 * A + A + a) * cells + cell;
      const float lab = labels[o];
      float d0 = 0.f, d1 = 0.f;
      if (lab != -1.f) {
        float p1;
        ce2(score[q0], score[q1], lab != 0.f, &p1);
        d1 = g0 * (p1 - (lab != 0.f ? 1.f : 0.f));
        d0 = -d1;
      }
      d_score[q0] = d0;
      d_score[q1] = d1;
    }
    if (d_bbox) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t q = ((size_t)b * 4 * A + a * 4 + j) * cells + cell;
        const float wo = outside[q];
        float g = 0.f;
        if (wo != 0.f) smooth_l1_elem(bbox_pred[q], targets[q], inside[q], wo, s2, &g);
        d_bbox[q] = g1 * g;
      }
    }
  }
}

// Detection losses (faster_rcnn_coatt_transformer_sk.py:334-361), one CTA per image of P rois (P <= 1024):
//   acc[0] += sum CE(score, label), acc[1] += sum over the P x P map of max(0, -t (|p_i-p_j| - |l_i-l_j|) + margin),
//   acc[2] += sum smooth L1;  cls_prob[n] = softmax(score)[1];  gradients w.r.t. score / bbox_pred of
//   gscale[0] * mean CE + gscale[1] * margin_scale * mean(map) + gscale[2] * mean_rows(sum smooth L1)
__global__ void __launch_bounds__(1024)
rcnn_loss_kernel(const float* __restrict__ score, const float* __restrict__ bbox_pred, const float* __restrict__ labels,
                 const float* __restrict__ targets, const float* __restrict__ inside, const float* __restrict__ outside,
                 int bs, int P, float margin, float margin_scale, const float* __restrict__ gscale, double* __restrict__ acc,
                 float* __restrict__ cls_prob, float* __restrict__ d_score, float* __restrict__ d_bbox) {
  __shared__ float sp[1024], sl[1024];
  __shared__ double s_red[32];
  const int b = blockIdx.x, i = threadIdx.x;
  const size_t row = (size_t)b * P + i;
  const float n_rows = (float)bs * (float)P;
  const float g0 = (gscale ? gscale[0] : 1.f) / n_rows;
  const float g1 = (gscale ? gscale[1] : 1.f) * margin_scale / (n_rows * (float)P);
  const float g2 = (gscale ? gscale[2] : 1.f) / n_rows;
  double ce = 0.0, box = 0.0, mar = 0.0;
  float p = 0.f, lab = 0.f, ds1 = 0.f;
  if (i < P) {
    lab = labels[row];
    const float s0 = score[row * 2], s1 = score[row * 2 + 1];
    ce = ce2(s0, s1, lab != 0.f, &p);
    ds1 = g0 * (p - (lab != 0.f ? 1.f : 0.f));
    if (cls_prob) cls_prob[row] = p;
    float gb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      box += smooth_l1_elem(bbox_pred[row * 4 + j], targets[row * 4 + j], inside[row * 4 + j], outside[row * 4 + j], 1.f, &gb[j]);
    if (d_bbox) reinterpret_cast<float4*>(d_bbox)[row] = make_float4(g2 * gb[0], g2 * gb[1], g2 * gb[2], g2 * gb[3]);
    sp[i] = p;
    sl[i] = lab;
  }
  __syncthreads();
  if (i < P) {
    float dp = 0.f;
    for (int j = 0; j < P; ++j) {
      const float diff = p - sp[j];
      const float pr = fabsf(diff), gm = fabsf(lab - sl[j]);
      const float t = -((gm - 1.f) * (gm - 1.f)) + gm;
      const float v = -t * (pr - gm) + margin;
      if (v > 0.f) {
        mar += v;
        dp += -t * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
      }
    }
    // the map is symmetric: entries (i, j) and (j, i) both depend on p_i
    ds1 += g1 * 2.f * dp * p * (1.f - p);
    if (d_score) { d_score[row * 2] = -ds1; d_score[row * 2 + 1] = ds1; }
  }
  ce = block_sum(ce, s_red);
  mar = block_sum(mar, s_red);
  box = block_sum(box, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, ce);
    atomicAdd(acc + 1, mar);
    atomicAdd(acc + 2, box);
  }
}

__global__ void rcnn_loss_final_kernel(const double* __restrict__ acc, int bs, int P, float margin_scale, float* __restrict__ losses) {
  const double n = (double)bs * P;
  losses[0] = (float)(acc[0] / n);
  losses[1] = (float)(margin_scale * (float)(acc[1] / (n * P)));
  losses[2] = (float)(acc[2] / n);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace aitb

using namespace aitb;

extern "C" {

size_t aitb_anchor_target_workspace_bytes(int B, int A, int H, int W, int K) {
  const size_t total = (size_t)A * H * W;
  return align_up((size_t)B * total * 4, 256) + align_up((size_t)B * K * 4, 256);
}

int aitb_anchor_target_assign(const float* base_anchors, const float* gt_boxes, const float* im_info, int B, int A, int H,
                              int W, int K, float feat_stride, float neg_thr, float pos_thr, int clobber_positives,
                              int8_t* labels, int32_t* argmax, int32_t* counts, void* workspace, size_t workspace_bytes,
                              aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(base_anchors && gt_boxes && im_info && labels && argmax && counts && workspace, "aitb_anchor_target_assign: null pointer");
  AITB_REQUIRE(B > 0 && B <= 65535 && A > 0 && H > 0 && W > 0 && K > 0 && K <= kMaxGt,
               "aitb_anchor_target_assign: bad sizes (at most %d ground-truth boxes per image)", kMaxGt);
  AITB_REQUIRE(workspace_bytes >= aitb_anchor_target_workspace_bytes(B, A, H, W, K), "aitb_anchor_target_assign: workspace too small");
  const int total = A * H * W;
  float* max_ov = reinterpret_cast<float*>(workspace);
  unsigned* gt_max = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + align_up((size_t)B * total * 4, 256));
  cudaError_t e = cudaMemsetAsync(gt_max, 0, (size_t)B * K * 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, (size_t)B * 2 * 4, st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_anchor_target_assign: memset failed: %s", cudaGetErrorString(e));
  const dim3 grid((total + 255) / 256, B);
  anchor_overlap_kernel<<<grid, 256, 0, st>>>(base_anchors, gt_boxes, im_info, A, H, W, K, feat_stride, max_ov, argmax, gt_max);
  if (check_launch("anchor_overlap_kernel")) return 1;
  anchor_label_kernel<<<grid, 256, 0, st>>>(base_anchors, gt_boxes, A, H, W, K, feat_stride, neg_thr, pos_thr,
                                            clobber_positives, max_ov, argmax, gt_max, labels, counts);
  return check_launch("anchor_label_kernel");
}

int aitb_anchor_target_finish(const float* base_anchors, const float* gt_boxes, int B, int A, int H, int W, int K,
                              float feat_stride, int8_t* labels, const int32_t* argmax, const uint8_t* drop, int ld_drop,
                              float inside_weight, int32_t* n_examples, float* labels_out, float* bbox_targets,
                              float* inside_w, float* outside_w, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(base_anchors && gt_boxes && labels && argmax && n_examples && labels_out && bbox_targets && inside_w && outside_w,
               "aitb_anchor_target_finish: null pointer");
  AITB_REQUIRE(B > 0 && B <= 65535 && A > 0 && H > 0 && W > 0 && K > 0 && (!drop || ld_drop > 0), "aitb_anchor_target_finish: bad sizes");
  const int total = A * H * W;
  anchor_subsample_kernel<<<B, 1024, 0, st>>>(labels, total, drop, ld_drop, n_examples);
  if (check_launch("anchor_subsample_kernel")) return 1;
  anchor_emit_kernel<<<dim3((total + 255) / 256, B), 256, 0, st>>>(base_anchors, gt_boxes, A, H, W, K, feat_stride, labels, argmax,
                                                                   n_examples, B, inside_weight, labels_out, bbox_targets,
                                                                   inside_w, outside_w);
  return check_launch("anchor_emit_kernel");
}

int aitb_anchor_target_subsample_device(int8_t* labels, const int32_t* counts, int B, int total, int num_fg, int rpn_batch,
                                        uint64_t seed, aitb_stream_t stream) {
  AITB_REQUIRE(labels && counts && B > 0 && total > 0 && num_fg >= 0 && rpn_batch >= num_fg, "aitb_anchor_target_subsample_device: bad arguments");
  anchor_subsample_device_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(labels, total, counts, num_fg, rpn_batch, seed);
  return check_launch("anchor_subsample_device_kernel");
}

int aitb_proposal_target_picks_device(const int32_t* counts, int B, int S, int fg_per_image, uint64_t seed, int32_t* picks,
                                      int32_t* n_fg_pick, int32_t* bad_flag, aitb_stream_t stream) {
  AITB_REQUIRE(counts && picks && n_fg_pick && bad_flag && B > 0 && S > 0 && fg_per_image > 0, "aitb_proposal_target_picks_device: bad arguments");
  proposal_picks_device_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(counts, S, fg_per_image, seed, picks, n_fg_pick, bad_flag);
  return check_launch("proposal_picks_device_kernel");
}

int aitb_proposal_target_assign(const float* rois, const float* gt_boxes, int B, int R, int K, float fg_thr, float bg_hi,
                                float bg_lo, float* max_overlaps, int32_t* assignment, int8_t* cls, int32_t* counts,
                                aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(rois && gt_boxes && max_overlaps && assignment && cls && counts, "aitb_proposal_target_assign: null pointer");
  AITB_REQUIRE(B > 0 && B <= 65535 && R >= 0 && K > 0 && K <= kMaxGt,
               "aitb_proposal_target_assign: bad sizes (at most %d ground-truth boxes per image)", kMaxGt);
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)B * 2 * 4, st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_proposal_target_assign: memset failed: %s", cudaGetErrorString(e));
  proposal_assign_kernel<<<dim3((R + K + 255) / 256, B), 256, 0, st>>>(rois, gt_boxes, R, K, fg_thr, bg_hi, bg_lo, max_overlaps,
                                                                       assignment, cls, counts);
  return check_launch("proposal_assign_kernel");
}

int aitb_proposal_target_sample(const float* rois, const float* gt_boxes, int B, int R, int K, const int8_t* cls,
                                const int32_t* assignment, const int32_t* picks, const int32_t* n_fg_pick, int S,
                                const float* h_means, const float* h_stds, const float* h_inside_w, int32_t* lists,
                                float* rois_out, float* labels_out, float* bbox_targets, float* inside_w, float* outside_w,
                                int32_t* bad_flag, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(rois && gt_boxes && cls && assignment && picks && n_fg_pick && h_means && h_stds && h_inside_w && lists && rois_out &&
                   labels_out && bbox_targets && inside_w && outside_w && bad_flag, "aitb_proposal_target_sample: null pointer");
  AITB_REQUIRE(B > 0 && R >= 0 && K > 0 && S > 0, "aitb_proposal_target_sample: bad sizes");
  AITB_REQUIRE((((uintptr_t)bbox_targets | (uintptr_t)inside_w | (uintptr_t)outside_w) & 15) == 0,
               "aitb_proposal_target_sample: outputs must be 16-byte aligned");
  proposal_sample_kernel<<<B, 1024, 0, st>>>(rois, gt_boxes, R, K, cls, assignment, picks, n_fg_pick, S,
                                             make_float4(h_means[0], h_means[1], h_means[2], h_means[3]),
                                             make_float4(h_stds[0], h_stds[1], h_stds[2], h_stds[3]),
                                             make_float4(h_inside_w[0], h_inside_w[1], h_inside_w[2], h_inside_w[3]), lists,
                                             rois_out, labels_out, bbox_targets, inside_w, outside_w, bad_flag);
  return check_launch("proposal_sample_kernel");
}

int aitb_rpn_loss(const float* rpn_cls_score, const float* rpn_bbox_pred, const float* labels, const float* bbox_targets,
                  const float* inside_w, const float* outside_w, int B, int A, int H, int W, float sigma,
                  const float* gscale, float* losses, float* d_cls_score, float* d_bbox_pred, double* acc,
                  aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(rpn_cls_score && rpn_bbox_pred && labels && bbox_targets && inside_w && outside_w && losses && acc,
               "aitb_rpn_loss: null pointer");
  AITB_REQUIRE(B > 0 && A > 0 && H > 0 && W > 0 && sigma > 0.f, "aitb_rpn_loss: bad sizes");
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_rpn_loss: memset failed: %s", cudaGetErrorString(e));
  const size_t n = (size_t)B * A * H * W;
  const int blocks = (int)((n + 255) / 256 < (size_t)(4 * current_sm_count()) ? (n + 255) / 256 : 4 * current_sm_count());
  rpn_loss_reduce_kernel<<<blocks, 256, 0, st>>>(rpn_cls_score, rpn_bbox_pred, labels, bbox_targets, inside_w, outside_w, B, A,
                                                 H * W, sigma * sigma, acc);
  if (check_launch("rpn_loss_reduce_kernel")) return 1;
  rpn_loss_grad_kernel<<<blocks, 256, 0, st>>>(rpn_cls_score, rpn_bbox_pred, labels, bbox_targets, inside_w, outside_w, B, A,
                                               H * W, sigma * sigma, acc, gscale, losses, d_cls_score, d_bbox_pred);
  return check_launch("rpn_loss_grad_kernel");
}

int aitb_rcnn_loss(const float* score, const float* bbox_pred, const float* labels, const float* bbox_targets,
                   const float* inside_w, const float* outside_w, int bs, int P, float margin, float margin_scale,
                   const float* gscale, float* losses, float* cls_prob, float* d_score, float* d_bbox_pred, double* acc,
                   aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(score && bbox_pred && labels && bbox_targets && inside_w && outside_w && losses && acc, "aitb_rcnn_loss: null pointer");
  AITB_REQUIRE(bs > 0 && P > 0 && P <= 1024, "aitb_rcnn_loss: bad sizes (1 <= rois per image <= 1024)");
  AITB_REQUIRE(!d_bbox_pred || ((uintptr_t)d_bbox_pred & 15) == 0, "aitb_rcnn_loss: d_bbox_pred must be 16-byte aligned");
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_rcnn_loss: memset failed: %s", cudaGetErrorString(e));
  const int threads = (P + 31) / 32 * 32;
  rcnn_loss_kernel<<<bs, threads, 0, st>>>(score, bbox_pred, labels, bbox_targets, inside_w, outside_w, bs, P, margin, margin_scale,
                                           gscale, acc, cls_prob, d_score, d_bbox_pred);
  if (check_launch("rcnn_loss_kernel")) return 1;
  rcnn_loss_final_kernel<<<1, 1, 0, st>>>(acc, bs, P, margin_scale, losses);
  return check_launch("rcnn_loss_final_kernel");
}

}  // extern "C"
