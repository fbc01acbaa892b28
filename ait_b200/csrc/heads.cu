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
//
// HEADS = false: pooling only (the query maps): one CTA per map.
// HEADS = true (round 2): the first version read the 12 x 2048 proposal-side weights AND the 8 x 2048 query-side weights + the
// unit's query feature from L2 for every pair (168 KB per pair, 403 MB per step -- more than the 315 MB map it pools) in a
// serial phase behind the pooling loads.  Now a CTA owns a CONTIGUOUS range of pairs: the 12 proposal-side weight rows stay in
// shared memory (96 KB, two CTAs per SM), the query half W1[:, 2048:] . qfeat is computed once per unit the range touches,
// the pooled vector never leaves registers (thread = 8 channels: 24 LDS.128 + 96 FMA per pair for the 12 partial dots), and
// while one CTA of the SM reduces its dots the other streams its next map.
template <typename T, bool SPLIT, bool HEADS>
__global__ void __launch_bounds__(256, HEADS ? 2 : 4)
pool_heads_kernel(const T* __restrict__ top, int G, int P, const float* __restrict__ qfeat, const float* __restrict__ w_bbox,
                  const float* __restrict__ b_bbox, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ feat_out,
                  float* __restrict__ bbox_out, float* __restrict__ cls_out) {
  extern __shared__ __align__(16) float hsm[];
  float* Wsm = hsm;                       // HEADS: [12][2048]: rows 0-3 w_bbox, 4-11 W1[:, :2048]
  float* part = hsm + (HEADS ? 12 * kFeat : 0);   // [8 warps][12]
  float* qdot = part + 8 * 12;            // [8]: W1[:, 2048:] . qfeat[unit]
  constexpr int kPitch = SPLIT ? 2 * kFeat : kFeat;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 8;                  // this thread's 8 consecutive channels
  int g0 = blockIdx.x, g1 = blockIdx.x + 1;
  if constexpr (HEADS) {
    const int per = (G + gridDim.x - 1) / gridDim.x;
    g0 = blockIdx.x * per;
    g1 = g0 + per < G ? g0 + per : G;
    for (int i = tid; i < 12 * kFeat / 4; i += 256) {
      const int d = i / (kFeat / 4), cc = (i % (kFeat / 4)) * 4;
      const float* src = d < 4 ? w_bbox + (size_t)d * kFeat + cc : w1 + (size_t)(d - 4) * 2 * kFeat + cc;
      *reinterpret_cast<float4*>(Wsm + d * kFeat + cc) = __ldg(reinterpret_cast<const float4*>(src));
    }
    __syncthreads();
  }
  int unit = -1;
  for (int gidx = g0; gidx < g1; ++gidx) {
    const T* tg = top + (size_t)gidx * kPos * kPitch;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
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
    for (int j = 0; j < 8; ++j) acc[j] *= (1.f / 16.f);
    if (feat_out) {
      float* fo = feat_out + (size_t)gidx * kFeat + c;
      *reinterpret_cast<float4*>(fo) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(fo + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if constexpr (HEADS) {
      if (gidx / P != unit) {   // the query half of the first score layer: once per unit of this CTA's range (block-uniform)
        unit = gidx / P;
        const float* qf = qfeat + (size_t)unit * kFeat;
        float qd[8];
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(qf + c)), q1 = __ldg(reinterpret_cast<const float4*>(qf + c + 4));
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const float* w = w1 + (size_t)d * 2 * kFeat + kFeat + c;
          const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w + 4));
          qd[d] = warp_sum(a.x * q0.x + a.y * q0.y + a.z * q0.z + a.w * q0.w + b.x * q1.x + b.y * q1.y + b.z * q1.z + b.w * q1.w);
        }
        __syncthreads();        // the previous pair's readers of part / qdot are done
        if (lane == 0) {
#pragma unroll
          for (int d = 0; d < 8; ++d) part[warp * 12 + d] = qd[d];
        }
        __syncthreads();
        if (tid < 8) {
          float t = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) t += part[w8 * 12 + tid];
          qdot[tid] = t;
        }
      }
      float dsum[12];
#pragma unroll
      for (int d = 0; d < 12; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(Wsm + d * kFeat + c), b = *reinterpret_cast<const float4*>(Wsm + d * kFeat + c + 4);
        dsum[d] = warp_sum(a.x * acc[0] + a.y * acc[1] + a.z * acc[2] + a.w * acc[3] + b.x * acc[4] + b.y * acc[5] + b.z * acc[6] + b.w * acc[7]);
      }
      __syncthreads();          // the previous pair's final reads of part are done (and qdot is published)
      if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 12; ++d) part[warp * 12 + d] = dsum[d];
      }
      __syncthreads();
      if (tid < 32) {
        float dot = 0.f;
        if (tid < 12) {
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) dot += part[w8 * 12 + tid];
        }
        if (tid < 4) bbox_out[(size_t)gidx * 4 + tid] = dot + b_bbox[tid];
        // hidden layer of RCNN_cls_score: lanes 4..11 hold the proposal halves
        float hdn = (tid >= 4 && tid < 12) ? dot + qdot[tid - 4] + b1[tid - 4] : 0.f;
        float s0 = 0.f, s1 = 0.f;
        if (tid >= 4 && tid < 12) { s0 = w2[tid - 4] * hdn; s1 = w2[8 + tid - 4] * hdn; }
        s0 = warp_sum(s0) + b2[0];
        s1 = warp_sum(s1) + b2[1];
        if (tid == 0) {
          const float m = fmaxf(s0, s1);
          const float e0 = expf(s0 - m), e1 = expf(s1 - m);
          cls_out[gidx] = e1 / (e0 + e1);
        }
      }
    }
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
  const bool heads = w_bbox != nullptr;
  const int smem = ((heads ? 12 * kFeat : 0) + 8 * 12 + 8) * (int)sizeof(float);
  int grid = G;
  if (heads) {   // contiguous ranges of pairs, two CTAs per SM
    grid = 2 * current_sm_count();
    if (grid > G) grid = G;
  }
#define AITB_POOL_LAUNCH(TT, SP)                                                                                          \
  do {                                                                                                                    \
    if (heads) {                                                                                                          \
      static SmemAttrOnce once;                                                                                           \
      if (ensure_dyn_smem((const void*)pool_heads_kernel<TT, SP, true>, smem, once, "pool_heads_kernel")) return 1;       \
      pool_heads_kernel<TT, SP, true><<<grid, 256, smem, stream>>>((const TT*)top, G, P, qfeat, w_bbox, b_bbox, w1, b1, w2, \
                                                                  b2, feat_out, bbox_out, cls_out);                        \
    } else {                                                                                                              \
      pool_heads_kernel<TT, SP, false><<<grid, 256, smem, stream>>>((const TT*)top, G, P, qfeat, w_bbox, b_bbox, w1, b1,  \
                                                                   w2, b2, feat_out, bbox_out, cls_out);                   \
    }                                                                                                                     \
  } while (0)
  if (dtype == AITB_F32)
    AITB_POOL_LAUNCH(float, false);
  else if (dtype == AITB_BF16)
    AITB_POOL_LAUNCH(__nv_bfloat16, false);
  else if (dtype == AITB_F32S)
    AITB_POOL_LAUNCH(__nv_bfloat16, true);
  else {
    set_error("aitb_pool_heads: bad dtype %d", dtype);
    return 1;
  }
#undef AITB_POOL_LAUNCH
  return check_launch("pool_heads_kernel");
}

}  // namespace aitb
