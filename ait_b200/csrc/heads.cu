// Tail of the detection head: spatial mean over the 4x4 layer-4 map (`_head_to_tail`,
// lib/model/faster_rcnn/resnet_coatt_transformer_sk.py:476-485), RCNN_bbox_pred (Linear 2048->4),
// RCNN_cls_score (Linear 4096->8 -> Linear 8->2) on cat(props_feat, query_feat) and
// softmax(score)[:, 1]  (lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:318-337).
//
// One CTA per proposal-query pair: the 16 x 2048 map is read once (coalesced 128-bit loads), the
// pooled vector lives in shared memory, the 4 + 8 dot products are warp-shuffle reductions.
// The query half of the 4096-wide first layer (W1[:, 2048:] . qfeat) is identical for all P
// proposals of a unit; each CTA recomputes it from the unit's pooled query feature (8 x 2048 MACs).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kFeat = 2048;
static constexpr int kPos = 16;

// SPLIT: `top` is a split matrix [G*16, hi 2048 | lo 2048] (AITB_F32S)
template <typename T, bool SPLIT = false>
__global__ void __launch_bounds__(256)
pool_heads_kernel(const T* __restrict__ top, int P, const float* __restrict__ qfeat, const float* __restrict__ w_bbox,
                  const float* __restrict__ b_bbox, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ feat_out,
                  float* __restrict__ bbox_out, float* __restrict__ cls_out) {
  __shared__ float feat[kFeat];
  __shared__ float dots[12];
  const int gidx = blockIdx.x;
  constexpr int kPitch = SPLIT ? 2 * kFeat : kFeat;
  const T* tg = top + (size_t)gidx * kPos * kPitch;
  // thread owns 8 consecutive channels
  {
    const int c = threadIdx.x * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int p = 0; p < kPos; ++p) {
      float v[8];
      ld8(tg + (size_t)p * kPitch + c, v);
      if constexpr (SPLIT) {
        float u[8];
        ld8(tg + (size_t)p * kPitch + kFeat + c, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += u[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    // .mean(3).mean(2) on a 4x4 map: two successive means of 4 == sum / 16 up to rounding
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] *= (1.f / 16.f);
      feat[c + j] = acc[j];
    }
    if (feat_out) {
      float* fo = feat_out + (size_t)gidx * kFeat + c;
      *reinterpret_cast<float4*>(fo) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(fo + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  if (w_bbox == nullptr) return;  // pooling only (query branch)
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qf = qfeat + (size_t)(gidx / P) * kFeat;
  for (int d = warp; d < 12; d += 8) {
    float acc = 0.f;
    if (d < 4) {
      const float* w = w_bbox + (size_t)d * kFeat;
      for (int c = lane * 4; c < kFeat; c += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
        acc += w4.x * feat[c] + w4.y * feat[c + 1] + w4.z * feat[c + 2] + w4.w * feat[c + 3];
      }
    } else {
      const float* w = w1 + (size_t)(d - 4) * 2 * kFeat;
      for (int c = lane * 4; c < kFeat; c += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
        acc += w4.x * feat[c] + w4.y * feat[c + 1] + w4.z * feat[c + 2] + w4.w * feat[c + 3];
        const float4 u4 = __ldg(reinterpret_cast<const float4*>(w + kFeat + c));
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(qf + c));
        acc += u4.x * q4.x + u4.y * q4.y + u4.z * q4.z + u4.w * q4.w;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dots[d] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 4) bbox_out[(size_t)gidx * 4 + threadIdx.x] = dots[threadIdx.x] + b_bbox[threadIdx.x];
  if (threadIdx.x == 32) {
    float hdn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hdn[j] = dots[4 + j] + b1[j];
    float s0 = b2[0], s1 = b2[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s0 += w2[j] * hdn[j];
      s1 += w2[8 + j] * hdn[j];
    }
    const float m = fmaxf(s0, s1);
    const float e0 = expf(s0 - m), e1 = expf(s1 - m);
    cls_out[gidx] = e1 / (e0 + e1);
  }
}

int pool_heads_run(const void* top, int dtype, int G, int P, const float* qfeat, const float* w_bbox,
                   const float* b_bbox, const float* w1, const float* b1, const float* w2, const float* b2,
                   float* feat_out, float* bbox_out, float* cls_out, cudaStream_t stream) {
  AITB_REQUIRE(G > 0 && P > 0, "aitb_pool_heads: bad sizes G=%d P=%d", G, P);
  AITB_REQUIRE(top != nullptr, "aitb_pool_heads: null input");
  if (w_bbox) {
    AITB_REQUIRE(qfeat && b_bbox && w1 && b1 && w2 && b2 && bbox_out && cls_out, "aitb_pool_heads: null head pointer");
    AITB_REQUIRE(G % P == 0, "aitb_pool_heads: G=%d is not a multiple of P=%d", G, P);
  } else {
    AITB_REQUIRE(feat_out != nullptr, "aitb_pool_heads: pooling-only call needs feat_out");
  }
  if (dtype == AITB_F32)
    pool_heads_kernel<float><<<G, 256, 0, stream>>>((const float*)top, P, qfeat, w_bbox, b_bbox, w1, b1, w2, b2,
                                                    feat_out, bbox_out, cls_out);
  else if (dtype == AITB_BF16)
    pool_heads_kernel<__nv_bfloat16><<<G, 256, 0, stream>>>((const __nv_bfloat16*)top, P, qfeat, w_bbox, b_bbox, w1,
                                                            b1, w2, b2, feat_out, bbox_out, cls_out);
  else if (dtype == AITB_F32S)
    pool_heads_kernel<__nv_bfloat16, true><<<G, 256, 0, stream>>>((const __nv_bfloat16*)top, P, qfeat, w_bbox, b_bbox,
                                                                  w1, b1, w2, b2, feat_out, bbox_out, cls_out);
  else {
    set_error("aitb_pool_heads: bad dtype %d", dtype);
    return 1;
  }
  return check_launch("pool_heads_kernel");
}

}  // namespace aitb
