// Box arithmetic either side of the head, on the device:
//
//   rpn_decode_kernel    the front of _ProposalLayer.forward (lib/model/rpn/proposal_layer.py:66-118):
//                        anchors = base anchors + cell shifts, bbox_transform_inv, clip_boxes, and the
//                        NCHW -> [B, K*A] re-ordering of the RPN outputs (`permute(0,2,3,1)`), in one pass.
//                        Feeds aitb_topk_desc / aitb_nms_batched: the whole proposal layer runs without
//                        a python per-image loop or a host round trip.
//   box_decode_kernel    the detection decode of test_net_voc.py:380-407: de-normalise bbox_pred with the
//                        precomputed stds / means, bbox_transform_inv on the rois, clip to the image,
//                        divide by the image scale; also emits the `score > thresh` keys for the sort.
//   det_assemble_kernel  test_net_voc.py:421-446 after the final NMS: [x1,y1,x2,y2,score] rows in
//                        descending score order, limited to max_per_image (ties at the cut kept, like
//                        the reference's `>= image_thresh`).
//
// bbox_transform_inv / clip_boxes follow lib/model/rpn/bbox_transform.py:77-133 with explicitly rounded
// fp32 operations in the reference's order; the only non-bit-identical step is exp (expf vs the host libm).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

// bbox_transform.py:77-106 for one box, then clip (:125-133) to [0, w-1] x [0, h-1]
__device__ __forceinline__ float4 decode_clip(float4 b, float dx, float dy, float dw, float dh, float im_h, float im_w) {
  const float w = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f);
  const float h = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
  const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w));
  const float cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);
  const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w);
  const float ph = __fmul_rn(expf(dh), h);
  float4 o;
  o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  const float mx = __fsub_rn(im_w, 1.0f), my = __fsub_rn(im_h, 1.0f);
  o.x = fminf(fmaxf(o.x, 0.f), mx);   // clamp_(0, im_shape[i, 1] - 1)
  o.y = fminf(fmaxf(o.y, 0.f), my);
  o.z = fminf(fmaxf(o.z, 0.f), mx);
  o.w = fminf(fmaxf(o.w, 0.f), my);
  return o;
}

// scores_nchw [B, 2A, H, W] (fg = channels A..2A), deltas_nchw [B, 4A, H, W], base [A, 4], im_info [B, 3]
// -> proposals [B, H*W*A, 4], fg [B, H*W*A]; candidate index = (y*W + x)*A + a (proposal_layer.py:92-106)
__global__ void __launch_bounds__(256)
rpn_decode_kernel(const float* __restrict__ scores, const float* __restrict__ deltas, const float* __restrict__ base,
                  const float* __restrict__ im_info, int A, int H, int W, float stride, float4* __restrict__ proposals,
                  float* __restrict__ fg) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // = a * H*W + cell: consecutive threads read consecutive cells
  const int cells = H * W;
  if (i >= A * cells) return;
  const int a = i / cells, cell = i - a * cells;
  const int y = cell / W, x = cell - y * W;
  const float sx = (float)x * stride, sy = (float)y * stride;
  float4 an;
  an.x = __fadd_rn(base[a * 4 + 0], sx);
  an.y = __fadd_rn(base[a * 4 + 1], sy);
  an.z = __fadd_rn(base[a * 4 + 2], sx);
  an.w = __fadd_rn(base[a * 4 + 3], sy);
  const float* d = deltas + ((size_t)b * 4 * A + 4 * a) * cells + cell;
  const float4 o = decode_clip(an, d[0], d[cells], d[2 * (size_t)cells], d[3 * (size_t)cells], im_info[b * 3 + 0],
                               im_info[b * 3 + 1]);
  const size_t out = (size_t)b * cells * A + (size_t)cell * A + a;
  proposals[out] = o;
  fg[out] = scores[((size_t)b * 2 * A + A + a) * cells + cell];
}

// boxes [B, N, 4] (rois without the batch column: `boxes_stride` floats per row, box at column `boxes_off`),
// deltas [B, N, 4], cls [B, N] -> pred [B, N, 4], key [B, N] = score > thresh ? score : -inf, n_valid [B]
__global__ void __launch_bounds__(256)
box_decode_kernel(const float* __restrict__ boxes, int boxes_stride, int boxes_off, const float* __restrict__ deltas,
                  const float* __restrict__ cls, const float* __restrict__ im_info, int N, float4 stds, float4 means,
                  float thresh, int divide_by_scale, float4* __restrict__ pred, float* __restrict__ key,
                  int* __restrict__ n_valid) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  if (i < N) {
    const float* bp = boxes + ((size_t)b * N + i) * boxes_stride + boxes_off;
    const float4 bx = make_float4(bp[0], bp[1], bp[2], bp[3]);
    const float4 d = reinterpret_cast<const float4*>(deltas)[(size_t)b * N + i];
    // box_deltas * BBOX_NORMALIZE_STDS + BBOX_NORMALIZE_MEANS  (test_net_voc.py:393-396)
    float4 o = decode_clip(bx, __fadd_rn(__fmul_rn(d.x, stds.x), means.x), __fadd_rn(__fmul_rn(d.y, stds.y), means.y),
                           __fadd_rn(__fmul_rn(d.z, stds.z), means.z), __fadd_rn(__fmul_rn(d.w, stds.w), means.w),
                           im_info[b * 3 + 0], im_info[b * 3 + 1]);
    if (divide_by_scale) {  // pred_boxes /= im_scale  (:410)
      const float s = im_info[b * 3 + 2];
      o.x = __fdiv_rn(o.x, s); o.y = __fdiv_rn(o.y, s); o.z = __fdiv_rn(o.z, s); o.w = __fdiv_rn(o.w, s);
    }
    pred[(size_t)b * N + i] = o;
    const float sc = cls[(size_t)b * N + i];
    valid = sc > thresh;
    key[(size_t)b * N + i] = valid ? sc : -INFINITY;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_valid + b, __popc(bal));
}

// order [B, N] (descending key), keep_pos [B, N] (kept positions in that order, ascending), n_keep [B], n_valid [B]
// -> dets [B, N, 5] zero-padded, n_det [B]
__global__ void __launch_bounds__(256)
det_assemble_kernel(const float4* __restrict__ pred, const float* __restrict__ cls, const int64_t* __restrict__ order,
                    const int64_t* __restrict__ keep_pos, const int32_t* __restrict__ n_keep,
                    const int32_t* __restrict__ n_valid, int N, int max_per_image, float* __restrict__ dets,
                    int32_t* __restrict__ n_det) {
  const int b = blockIdx.x;
  __shared__ int s_n;
  const int64_t* ord = order + (size_t)b * N;
  const int64_t* kp = keep_pos + (size_t)b * N;
  if (threadIdx.x == 0) {
    // kept candidates that passed the score threshold: positions < n_valid (the failed ones sort last and,
    // being lower-scored, can never have suppressed a valid box)
    int nk = n_keep[b];
    const int nv = n_valid[b];
    while (nk > 0 && kp[nk - 1] >= nv) --nk;
    int n = nk;
    if (max_per_image > 0 && nk > max_per_image) {
      const float cut = cls[(size_t)b * N + ord[kp[max_per_image - 1]]];   // np.sort(scores)[-max_per_image]
      n = max_per_image;
      while (n < nk && cls[(size_t)b * N + ord[kp[n]]] >= cut) ++n;        // ties at the cut are kept (>=)
    }
    s_n = n;
    n_det[b] = n;
  }
  __syncthreads();
  const int n = s_n;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float* o = dets + ((size_t)b * N + j) * 5;
    if (j < n) {
      const int64_t src = ord[kp[j]];
      const float4 bx = pred[(size_t)b * N + src];
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
      o[4] = cls[(size_t)b * N + src];
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
    }
  }
}

// RPN head outputs in token-major layout (the [rows, ld] output of the fused RPN_cls_score | RPN_bbox_pred GEMM:
// columns 0..A-1 background scores, A..2A-1 foreground scores, 2A..6A-1 box deltas; lib/model/rpn/rpn.py:70-85)
// -> pair softmax (rpn.py:76-78: reshape(x, 2) + softmax over the bg / fg pair of every anchor), anchors,
// bbox_transform_inv, clip: proposals [B, H*W*A, 4], fg [B, H*W*A]; optionally also the reference's NCHW tensors
// rpn_cls_prob [B, 2A, H, W] and rpn_bbox_pred [B, 4A, H, W].
// MODE: AITB_F32 (fp32 rows), AITB_BF16, AITB_F32S (two bf16 planes per row, lo plane `ld` elements further)
template <int MODE>
__device__ __forceinline__ float tok_load(const void* base, size_t row, int ld, int col) {
  if constexpr (MODE == AITB_F32) {
    return reinterpret_cast<const float*>(base)[row * ld + col];
  } else if constexpr (MODE == AITB_BF16) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[row * ld + col]);
  } else {
    const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(base) + row * 2 * (size_t)ld;
    return __bfloat162float(r[col]) + __bfloat162float(r[ld + col]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
rpn_head_decode_kernel(const void* __restrict__ heads, int ld, const float* __restrict__ base, const float* __restrict__ im_info,
                       int A, int H, int W, float stride, float4* __restrict__ proposals, float* __restrict__ fg,
                       float* __restrict__ cls_prob_nchw, float* __restrict__ bbox_pred_nchw) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // = cell * A + a: the candidate index of the proposal layer
  const int cells = H * W;
  if (i >= A * cells) return;
  const int cell = i / A, a = i - cell * A;
  const int y = cell / W, x = cell - y * W;
  const size_t row = (size_t)b * cells + cell;
  const float bg = tok_load<MODE>(heads, row, ld, a), fgs = tok_load<MODE>(heads, row, ld, A + a);
  const float m = fmaxf(bg, fgs);
  const float e0 = expf(__fsub_rn(bg, m)), e1 = expf(__fsub_rn(fgs, m));
  const float den = __fadd_rn(e0, e1);
  const float p1 = __fdiv_rn(e1, den);
  float d[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = tok_load<MODE>(heads, row, ld, 2 * A + 4 * a + j);
  const float sx = (float)x * stride, sy = (float)y * stride;
  float4 an;
  an.x = __fadd_rn(base[a * 4 + 0], sx);
  an.y = __fadd_rn(base[a * 4 + 1], sy);
  an.z = __fadd_rn(base[a * 4 + 2], sx);
  an.w = __fadd_rn(base[a * 4 + 3], sy);
  const size_t out = (size_t)b * cells * A + i;
  proposals[out] = decode_clip(an, d[0], d[1], d[2], d[3], im_info[b * 3 + 0], im_info[b * 3 + 1]);
  fg[out] = p1;
  if (cls_prob_nchw) {
    cls_prob_nchw[((size_t)b * 2 * A + a) * cells + cell] = __fdiv_rn(e0, den);
    cls_prob_nchw[((size_t)b * 2 * A + A + a) * cells + cell] = p1;
  }
  if (bbox_pred_nchw) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bbox_pred_nchw[((size_t)b * 4 * A + 4 * a + j) * cells + cell] = d[j];
  }
}

int rpn_head_decode_run(const void* heads, int dtype, int ld, const float* base_anchors, const float* im_info, int B, int A,
                        int H, int W, float feat_stride, float* proposals, float* fg_scores, float* cls_prob_nchw,
                        float* bbox_pred_nchw, cudaStream_t st) {
  AITB_REQUIRE(heads && base_anchors && im_info && proposals && fg_scores, "aitb_rpn_forward: null pointer");
  AITB_REQUIRE(B > 0 && A > 0 && H > 0 && W > 0 && 6 * A <= ld, "aitb_rpn_forward: bad sizes");
  const int n = A * H * W;
  dim3 grid((n + 255) / 256, B);
  float4* pr = reinterpret_cast<float4*>(proposals);
  if (dtype == AITB_F32)
    rpn_head_decode_kernel<AITB_F32><<<grid, 256, 0, st>>>(heads, ld, base_anchors, im_info, A, H, W, feat_stride, pr,
                                                           fg_scores, cls_prob_nchw, bbox_pred_nchw);
  else if (dtype == AITB_BF16)
    rpn_head_decode_kernel<AITB_BF16><<<grid, 256, 0, st>>>(heads, ld, base_anchors, im_info, A, H, W, feat_stride, pr,
                                                            fg_scores, cls_prob_nchw, bbox_pred_nchw);
  else
    rpn_head_decode_kernel<AITB_F32S><<<grid, 256, 0, st>>>(heads, ld, base_anchors, im_info, A, H, W, feat_stride, pr,
                                                            fg_scores, cls_prob_nchw, bbox_pred_nchw);
  return check_launch("rpn_head_decode_kernel");
}

int rpn_decode_run(const float* scores_nchw, const float* deltas_nchw, const float* base_anchors, const float* im_info,
                   int B, int A, int H, int W, float feat_stride, float* proposals, float* fg_scores, cudaStream_t st) {
  AITB_REQUIRE(scores_nchw && deltas_nchw && base_anchors && im_info && proposals && fg_scores, "aitb_rpn_decode: null pointer");
  AITB_REQUIRE(B > 0 && A > 0 && H > 0 && W > 0 && B <= 65535, "aitb_rpn_decode: bad sizes");
  AITB_REQUIRE(((uintptr_t)proposals & 15) == 0, "aitb_rpn_decode: proposals must be 16-byte aligned");
  const int n = A * H * W;
  rpn_decode_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(scores_nchw, deltas_nchw, base_anchors, im_info, A, H, W,
                                                              feat_stride, reinterpret_cast<float4*>(proposals), fg_scores);
  return check_launch("rpn_decode_kernel");
}

int box_decode_run(const float* boxes, int boxes_stride, int boxes_off, const float* deltas, const float* cls,
                   const float* im_info, int B, int N, const float* stds, const float* means, float thresh,
                   int divide_by_scale, float* pred, float* key, int32_t* n_valid, cudaStream_t st) {
  AITB_REQUIRE(boxes && deltas && cls && im_info && pred && key && n_valid && stds && means, "aitb_box_decode: null pointer");
  AITB_REQUIRE(B > 0 && N > 0 && B <= 65535 && boxes_stride >= 4 && boxes_off >= 0, "aitb_box_decode: bad sizes");
  AITB_REQUIRE(((uintptr_t)deltas & 15) == 0 && ((uintptr_t)pred & 15) == 0, "aitb_box_decode: deltas / pred must be 16-byte aligned");
  cudaError_t e = cudaMemsetAsync(n_valid, 0, (size_t)B * 4, st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_box_decode: memset failed: %s", cudaGetErrorString(e));
  box_decode_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(
      boxes, boxes_stride, boxes_off, deltas, cls, im_info, N, make_float4(stds[0], stds[1], stds[2], stds[3]),
      make_float4(means[0], means[1], means[2], means[3]), thresh, divide_by_scale, reinterpret_cast<float4*>(pred), key,
      n_valid);
  return check_launch("box_decode_kernel");
}

int det_assemble_run(const float* pred, const float* cls, const int64_t* order, const int64_t* keep_pos,
                     const int32_t* n_keep, const int32_t* n_valid, int B, int N, int max_per_image, float* dets,
                     int32_t* n_det, cudaStream_t st) {
  AITB_REQUIRE(pred && cls && order && keep_pos && n_keep && n_valid && dets && n_det, "aitb_det_assemble: null pointer");
  AITB_REQUIRE(B > 0 && N > 0, "aitb_det_assemble: bad sizes");
  det_assemble_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const float4*>(pred), cls, order, keep_pos, n_keep, n_valid, N,
                                         max_per_image, dets, n_det);
  return check_launch("det_assemble_kernel");
}

}  // namespace aitb
