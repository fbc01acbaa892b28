// Training of the detection head's last layers on the pooled layer-4 features (rows a12 + f4):
//   bbox_pred = RCNN_bbox_pred(feat)                       Linear(2048 -> 4)
//   score     = RCNN_cls_score(cat(feat, qfeat[unit]))     Linear(4096 -> 8) -> Linear(8 -> 2)
// (lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:318-337, resnet_coatt_transformer_sk.py:419-427)
// forward keeping what the backward needs (the 8 hidden values and the LOGITS -- the losses of
// faster_rcnn_coatt_transformer_sk.py:340-361 are taken on `score`, not on the probability the inference kernel
// emits), and the backward: gradients of both feature inputs and of the six parameter tensors, plus the adjoint of
// the 4x4 spatial mean (`_head_to_tail`, resnet_coatt_transformer_sk.py:476-485).  The reference gets all of this
// from torch autograd.  HBM-bound row work: one CTA per pair for the row passes, (channel block x row chunk) CTAs
// with per-thread channels for the weight gradients (feature rows read coalesced, the 12 upstream values of a row
// broadcast from shared memory).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kF = 2048;

// feat [G, 2048], qfeat [G / P, 2048] -> bbox [G, 4], hidden [G, 8], score [G, 2]
__global__ void __launch_bounds__(256)
heads_fwd_train_kernel(const float* __restrict__ feat, const float* __restrict__ qfeat, int P, const float* __restrict__ w_bbox,
                       const float* __restrict__ b_bbox, const float* __restrict__ w1, const float* __restrict__ b1,
                       const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ bbox,
                       float* __restrict__ hidden, float* __restrict__ score) {
  __shared__ float dots[12];
  const int g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* f = feat + (size_t)g * kF;
  const float* qf = qfeat + (size_t)(g / P) * kF;
  for (int d = warp; d < 12; d += 8) {
    float acc = 0.f;
    const float* w = d < 4 ? w_bbox + (size_t)d * kF : w1 + (size_t)(d - 4) * 2 * kF;
    for (int c = lane * 4; c < kF; c += 128) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(f + c));
      acc += w4.x * x4.x + w4.y * x4.y + w4.z * x4.z + w4.w * x4.w;
      if (d >= 4) {
        const float4 u4 = __ldg(reinterpret_cast<const float4*>(w + kF + c));
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(qf + c));
        acc += u4.x * q4.x + u4.y * q4.y + u4.z * q4.z + u4.w * q4.w;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dots[d] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 4) bbox[(size_t)g * 4 + threadIdx.x] = dots[threadIdx.x] + b_bbox[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 40) hidden[(size_t)g * 8 + threadIdx.x - 32] = dots[4 + threadIdx.x - 32] + b1[threadIdx.x - 32];
  if (threadIdx.x >= 64 && threadIdx.x < 66) {
    const int i = threadIdx.x - 64;
    float s = b2[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += w2[i * 8 + j] * (dots[4 + j] + b1[j]);
    score[(size_t)g * 2 + i] = s;
  }
}

// row pass of the backward: d_hidden [G, 8] = d_score W2, d_feat [G, 2048] = d_hidden W1[:, :2048] + d_bbox W_bbox
__global__ void __launch_bounds__(256)
heads_bwd_rows_kernel(const float* __restrict__ d_score, const float* __restrict__ d_bbox, const float* __restrict__ w_bbox,
                      const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ d_hidden,
                      float* __restrict__ d_feat) {
  __shared__ float up[12];   // 0..3 d_bbox, 4..11 d_hidden
  const int g = blockIdx.x;
  if (threadIdx.x < 4) up[threadIdx.x] = d_bbox[(size_t)g * 4 + threadIdx.x];
  else if (threadIdx.x < 12) {
    const int j = threadIdx.x - 4;
    const float v = d_score[(size_t)g * 2] * w2[j] + d_score[(size_t)g * 2 + 1] * w2[8 + j];
    up[threadIdx.x] = v;
    d_hidden[(size_t)g * 8 + j] = v;
  }
  __syncthreads();
  const int c = threadIdx.x * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int d = 0; d < 12; ++d) {
    const float* w = (d < 4 ? w_bbox + (size_t)d * kF : w1 + (size_t)(d - 4) * 2 * kF) + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w + 4));
    const float u = up[d];
    acc[0] += u * a.x; acc[1] += u * a.y; acc[2] += u * a.z; acc[3] += u * a.w;
    acc[4] += u * b.x; acc[5] += u * b.y; acc[6] += u * b.z; acc[7] += u * b.w;
  }
  float* o = d_feat + (size_t)g * kF + c;
  *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// per unit: D [B, 8] = sum over its P pairs of d_hidden; d_qfeat [B, 2048] = D W1[:, 2048:]
__global__ void __launch_bounds__(256)
heads_bwd_query_kernel(const float* __restrict__ d_hidden, int P, const float* __restrict__ w1, float* __restrict__ d_unit,
                       float* __restrict__ d_qfeat) {
  __shared__ float part[8][8];
  __shared__ float D[8];
  const int u = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // warp w sums rows w, w + 8, ... ; lanes = (row sub-index 4) x (hidden 8)
    const int j = lane & 7;
    float s = 0.f;
    for (int p = warp * 4 + (lane >> 3); p < P; p += 32) s += d_hidden[((size_t)u * P + p) * 8 + j];
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (lane < 8) part[warp][lane] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
    D[threadIdx.x] = s;
    d_unit[(size_t)u * 8 + threadIdx.x] = s;
  }
  __syncthreads();
  const int c = threadIdx.x * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* w = w1 + (size_t)j * 2 * kF + kF + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w + 4));
    const float d = D[j];
    acc[0] += d * a.x; acc[1] += d * a.y; acc[2] += d * a.z; acc[3] += d * a.w;
    acc[4] += d * b.x; acc[5] += d * b.y; acc[6] += d * b.z; acc[7] += d * b.w;
  }
  float* o = d_qfeat + (size_t)u * kF + c;
  *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// weight gradients: out[d][c] += sum over rows r of up[r][d] * x[r][c], d < ND.  grid (2048 / 256, row chunks);
// thread = channel (feature rows read coalesced), the ND upstream values of a row broadcast from shared memory.
//   up row r = (upA[r][0..NA), upB[r][0..NB))  (NA + NB = ND <= 12);  out row d = d < NA ? outA[d] : outB[d - NA], pitch ldo
template <int ND>
__global__ void __launch_bounds__(256)
heads_wgrad_kernel(const float* __restrict__ x, int rows, int rows_per_cta, const float* __restrict__ upA, int NA,
                   const float* __restrict__ upB, int NB, float* __restrict__ outA, int ldoA, float* __restrict__ outB, int ldoB) {
  __shared__ float s_up[64][ND];
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) acc[d] = 0.f;
  for (int rb = r0; rb < r1; rb += 64) {
    const int n = min(64, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < n * ND; i += 256) {
      const int r = i / ND, d = i - r * ND;
      s_up[r][d] = d < NA ? upA[(size_t)(rb + r) * NA + d] : upB[(size_t)(rb + r) * NB + (d - NA)];
    }
    __syncthreads();
    for (int r = 0; r < n; ++r) {
      const float v = x[(size_t)(rb + r) * kF + c];
#pragma unroll
      for (int d = 0; d < ND; ++d) acc[d] += s_up[r][d] * v;
    }
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    float* o = d < NA ? outA + (size_t)d * ldoA + c : outB + (size_t)(d - NA) * ldoB + c;
    atomicAdd(o, acc[d]);
  }
}

// the small ones: db_bbox [4] += sum d_bbox, db1 [8] += sum d_hidden, db2 [2] += sum d_score,
// dw2 [2, 8] += d_score^T hidden.  One CTA; thread t < 30 owns one output.
__global__ void __launch_bounds__(1024)
heads_small_grads_kernel(const float* __restrict__ d_bbox, const float* __restrict__ d_hidden, const float* __restrict__ d_score,
                         const float* __restrict__ hidden, int G, float* __restrict__ db_bbox, float* __restrict__ db1,
                         float* __restrict__ db2, float* __restrict__ dw2) {
  __shared__ float red[32][30];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[30];
#pragma unroll
  for (int i = 0; i < 30; ++i) acc[i] = 0.f;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float ds[2] = {d_score[(size_t)g * 2], d_score[(size_t)g * 2 + 1]};
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += d_bbox[(size_t)g * 4 + k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[4 + j] += d_hidden[(size_t)g * 8 + j];
      const float h = hidden[(size_t)g * 8 + j];
      acc[14 + j] += ds[0] * h;
      acc[22 + j] += ds[1] * h;
    }
    acc[12] += ds[0];
    acc[13] += ds[1];
  }
#pragma unroll
  for (int i = 0; i < 30; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 30) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    const int i = threadIdx.x;
    float* o = i < 4 ? db_bbox + i : i < 12 ? db1 + (i - 4) : i < 14 ? db2 + (i - 12) : dw2 + (i - 14);
    *o += s;
  }
}

// adjoint of the 4x4 spatial mean: d_top [G, 16, 2048] = d_feat [G, 2048] / 16
__global__ void __launch_bounds__(256)
mean_pool_bwd_kernel(const float* __restrict__ d_feat, float* __restrict__ d_top) {
  const int g = blockIdx.x, c = threadIdx.x * 8;
  const float4 a = *reinterpret_cast<const float4*>(d_feat + (size_t)g * kF + c);
  const float4 b = *reinterpret_cast<const float4*>(d_feat + (size_t)g * kF + c + 4);
  const float s = 1.f / 16.f;
  const float4 a2 = make_float4(a.x * s, a.y * s, a.z * s, a.w * s), b2 = make_float4(b.x * s, b.y * s, b.z * s, b.w * s);
#pragma unroll 4
  for (int p = 0; p < 16; ++p) {
    float* o = d_top + ((size_t)g * 16 + p) * kF + c;
    *reinterpret_cast<float4*>(o) = a2;
    *reinterpret_cast<float4*>(o + 4) = b2;
  }
}

}  // namespace aitb

using namespace aitb;

extern "C" {

int aitb_heads_forward_train(const float* feat, const float* qfeat, int G, int P, const float* w_bbox, const float* b_bbox,
                             const float* w1, const float* b1, const float* w2, const float* b2, float* bbox, float* hidden,
                             float* score, aitb_stream_t stream) {
  AITB_REQUIRE(feat && qfeat && w_bbox && b_bbox && w1 && b1 && w2 && b2 && bbox && hidden && score, "aitb_heads_forward_train: null pointer");
  AITB_REQUIRE(G > 0 && P > 0 && G % P == 0, "aitb_heads_forward_train: G=%d must be a positive multiple of P=%d", G, P);
  heads_fwd_train_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(feat, qfeat, P, w_bbox, b_bbox, w1, b1, w2, b2, bbox, hidden, score);
  return check_launch("heads_fwd_train_kernel");
}

size_t aitb_heads_backward_workspace_bytes(int G, int P) { return ((size_t)G * 8 + (size_t)(G / (P > 0 ? P : 1)) * 8) * sizeof(float); }

int aitb_heads_backward(const float* feat, const float* qfeat, const float* hidden, const float* d_score, const float* d_bbox,
                        int G, int P, const float* w_bbox, const float* w1, const float* w2, float* d_feat, float* d_qfeat,
                        float* dw_bbox, float* db_bbox, float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                        size_t workspace_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(feat && qfeat && hidden && d_score && d_bbox && w_bbox && w1 && w2 && d_feat && d_qfeat && dw_bbox && db_bbox && dw1 &&
                   db1 && dw2 && db2 && workspace, "aitb_heads_backward: null pointer");
  AITB_REQUIRE(G > 0 && P > 0 && G % P == 0, "aitb_heads_backward: G=%d must be a positive multiple of P=%d", G, P);
  AITB_REQUIRE(workspace_bytes >= aitb_heads_backward_workspace_bytes(G, P), "aitb_heads_backward: workspace too small");
  const int B = G / P;
  float* d_hidden = reinterpret_cast<float*>(workspace);
  float* d_unit = d_hidden + (size_t)G * 8;
  heads_bwd_rows_kernel<<<G, 256, 0, st>>>(d_score, d_bbox, w_bbox, w1, w2, d_hidden, d_feat);
  if (check_launch("heads_bwd_rows_kernel")) return 1;
  heads_bwd_query_kernel<<<B, 256, 0, st>>>(d_hidden, P, w1, d_unit, d_qfeat);
  if (check_launch("heads_bwd_query_kernel")) return 1;
  // dW_bbox [4, 2048] and dW1[:, :2048] from the pair rows; dW1[:, 2048:] from the unit rows
  const int chunk = 64;
  heads_wgrad_kernel<12><<<dim3(kF / 256, (G + chunk - 1) / chunk), 256, 0, st>>>(feat, G, chunk, d_bbox, 4, d_hidden, 8, dw_bbox, kF,
                                                                                 dw1, 2 * kF);
  if (check_launch("heads_wgrad_kernel")) return 1;
  heads_wgrad_kernel<8><<<dim3(kF / 256, (B + chunk - 1) / chunk), 256, 0, st>>>(qfeat, B, chunk, d_unit, 8, nullptr, 0, dw1 + kF, 2 * kF,
                                                                                nullptr, 0);
  if (check_launch("heads_wgrad_kernel")) return 1;
  heads_small_grads_kernel<<<1, 1024, 0, st>>>(d_bbox, d_hidden, d_score, hidden, G, db_bbox, db1, db2, dw2);
  return check_launch("heads_small_grads_kernel");
}

int aitb_mean_pool_backward(const float* d_feat, int G, float* d_top, aitb_stream_t stream) {
  AITB_REQUIRE(d_feat && d_top && G > 0, "aitb_mean_pool_backward: bad arguments");
  mean_pool_bwd_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(d_feat, d_top);
  return check_launch("mean_pool_bwd_kernel");
}

}  // extern "C"
