// Training-mode dropout of the AIT module (nn.Dropout sites of lib/model/system/Models.py:104,166,
// SubLayers.py:97 (`q = self.dropout(self.fc(q))`), :182 (`x = self.dropout(x)` in the FFN) -- all of them sit between a
// projection and the residual-add + LayerNorm that follows; the attention-probability dropout of Modules.py:24 lives in the
// attention kernels, attn.cu / bwd.cu).
//
// With dropout the fused GEMM + LayerNorm epilogue of the inference / dropout-free training path is split: the GEMM writes
// z = x W^T (+ bias) and ONE streaming kernel does  y = LayerNorm( dropout(z [+ pos]) + residual )  in place, one warp per
// 512-feature row.  The masks are never stored: a counter-based generator (Philox4x32-10 keyed by the step's seed and
// the site, counter = (row, column quad)) regenerates them in the backward (`drop_bwd_kernel`: the gradient of the
// projection output = dx * mask / (1 - p); the residual path takes dx unmasked).
// The reference's Philox stream cannot be reproduced (torch's generator state, one draw per element in launch order):
// parity is checked with OUR masks injected into the oracle (aitb_dropout_mask / aitb_attn_dropout_mask materialise them).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

// z [rows, 512] in place -> y; rows are groups of `grp` (64) tokens, z of rows >= `valid` of a group reads as 0; pos row = row % grp; residual row =
// ((row / res_div) / res_rep) * res_div + row % res_div  (res_rep = P: the unit's residual broadcast to its P pairs)
template <typename T>
__global__ void __launch_bounds__(256)
drop_res_ln_kernel(T* __restrict__ z, const float* __restrict__ pos, int grp, int valid, const T* __restrict__ res, int res_div,
                   int res_rep, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, DropCfg dc,
                   int rows, int round_tf, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    // rows t >= valid of a group carry no projection output (the encoder's 15 zero-padded rows, Models.py:268-270): z = 0, so
    // y = LayerNorm(dropout(pos_t)) -- NOT dead: they are queries of the encoder self-attention and enter its head gate
    const bool pad = row % grp >= valid;
    float v[16];
    T* zr = z + (size_t)row * 512;
    const float* pr = pos ? pos + (size_t)(row % grp) * 512 : nullptr;
    const T* rr = res ? res + (size_t)(((row / res_div) / res_rep) * res_div + row % res_div) * 512 : nullptr;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j * 128 + lane * 4;
      float4 x = pad ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(zr + c);
      if (pr) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c));
        x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
      }
      if (dc.thr) {
        const uint4 d = drop_row_draw(dc, (uint32_t)row, (uint32_t)(j * 32 + lane));
        x.x *= drop_mul(dc, d.x); x.y *= drop_mul(dc, d.y); x.z *= drop_mul(dc, d.z); x.w *= drop_mul(dc, d.w);
      }
      if (rr) {
        const float4 q = ld4(rr + c);
        x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
      }
      v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
      sum += x.x + x.y + x.z + x.w;
    }
    const float mean = warp_sum(sum) * (1.f / 512.f);
    float ssq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) ssq += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(ssq) * (1.f / 512.f) + eps);
    if (rstd_out && lane == 0) rstd_out[row] = rstd;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j * 128 + lane * 4;
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[4 * j] - mean) * rstd * g4.x + b4.x;
      o.y = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
      o.z = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
      o.w = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
      if (sizeof(T) == 2 || round_tf) st4r(zr + c, o);
      else *reinterpret_cast<float4*>(zr + c) = o;
    }
  }
}

// dz[r', :] = dx[r', :] * mask(row, :) / (1 - p); dx / dz hold `valid` of every `grp` rows (the encoder's 64 -> 49 compaction):
// r' = (row / grp) * valid + row % grp for row % grp < valid.  dz may alias dx.
template <typename T>
__global__ void __launch_bounds__(256)
drop_bwd_kernel(const T* __restrict__ dx, T* __restrict__ dz, DropCfg dc, int rows, int grp, int valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const int t = row % grp;
    if (t >= valid) continue;
    const size_t o = ((size_t)(row / grp) * valid + t) * 512;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j * 128 + lane * 4;
      float4 x = ld4(dx + o + c);
      const uint4 d = drop_row_draw(dc, (uint32_t)row, (uint32_t)(j * 32 + lane));
      x.x *= drop_mul(dc, d.x); x.y *= drop_mul(dc, d.y); x.z *= drop_mul(dc, d.z); x.w *= drop_mul(dc, d.w);
      st4r(dz + o + c, x);
    }
  }
}

// ---- mask materialisation (tests): the multipliers 0 | 1 / (1 - p) the kernels above / the attention kernels apply
__global__ void __launch_bounds__(256)
drop_mask_kernel(DropCfg dc, int rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 d = drop_row_draw(dc, (uint32_t)row, (uint32_t)(j * 32 + lane));
      *reinterpret_cast<float4*>(out + (size_t)row * 512 + j * 128 + lane * 4) =
          make_float4(drop_mul(dc, d.x), drop_mul(dc, d.y), drop_mul(dc, d.z), drop_mul(dc, d.w));
    }
  }
}
// out [G, 8, 64, 64]
__global__ void __launch_bounds__(256)
attn_drop_mask_kernel(DropCfg dc, int G, float* __restrict__ out) {
  const size_t n = (size_t)G * 8 * 64 * 32;   // one thread per (grp, h, row, column pair)
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const uint32_t cp = (uint32_t)(i & 31), row = (uint32_t)((i >> 5) & 63), h = (uint32_t)((i >> 11) & 7), grp = (uint32_t)(i >> 14);
    const uint4 d = drop_attn_draw(dc, grp, h, row & ~8u, cp);
    const uint32_t a = (row & 8u) ? d.z : d.x, b = (row & 8u) ? d.w : d.y;
    float* o = out + (((size_t)grp * 8 + h) * 64 + row) * 64 + 2 * cp;
    o[0] = drop_mul(dc, a);
    o[1] = drop_mul(dc, b);
  }
}

int check_launch(const char* what);

template <typename T>
static int drop_res_ln_run_t(T* z, const float* pos, int grp, int valid, const T* res, int res_div, int res_rep, const float* gamma,
                             const float* beta, float eps, float p, unsigned long long seed, int site, int rows, int round_tf,
                             float* rstd_out, cudaStream_t st) {
  AITB_REQUIRE(z && gamma && beta && rows > 0 && grp > 0 && valid > 0 && valid <= grp && res_div > 0 && res_rep > 0 && p >= 0.f &&
                   p < 1.f, "drop_res_ln: bad arguments");
  int grid = (rows + 7) / 8;
  if (grid > 8 * current_sm_count()) grid = 8 * current_sm_count();
  drop_res_ln_kernel<T><<<grid, 256, 0, st>>>(z, pos, grp, valid, res, res_div, res_rep, gamma, beta, eps, make_drop(p, seed, site),
                                              rows, round_tf, rstd_out);
  return check_launch("drop_res_ln_kernel");
}
int drop_res_ln_run(float* z, const float* pos, int grp, int valid, const float* res, int res_div, int res_rep, const float* gamma,
                    const float* beta, float eps, float p, unsigned long long seed, int site, int rows, int round_tf,
                    float* rstd_out, cudaStream_t st) {
  return drop_res_ln_run_t<float>(z, pos, grp, valid, res, res_div, res_rep, gamma, beta, eps, p, seed, site, rows, round_tf, rstd_out, st);
}
int drop_res_ln_run(__nv_bfloat16* z, const float* pos, int grp, int valid, const __nv_bfloat16* res, int res_div, int res_rep,
                    const float* gamma, const float* beta, float eps, float p, unsigned long long seed, int site, int rows,
                    int round_tf, float* rstd_out, cudaStream_t st) {
  return drop_res_ln_run_t<__nv_bfloat16>(z, pos, grp, valid, res, res_div, res_rep, gamma, beta, eps, p, seed, site, rows, round_tf,
                                          rstd_out, st);
}

template <typename T>
static int drop_bwd_run_t(const T* dx, T* dz, float p, unsigned long long seed, int site, int rows, int grp, int valid,
                          cudaStream_t st) {
  AITB_REQUIRE(dx && dz && rows > 0 && grp > 0 && valid > 0 && valid <= grp && rows % grp == 0 && p > 0.f && p < 1.f,
               "drop_bwd: bad arguments");
  int grid = (rows + 7) / 8;
  if (grid > 8 * current_sm_count()) grid = 8 * current_sm_count();
  drop_bwd_kernel<T><<<grid, 256, 0, st>>>(dx, dz, make_drop(p, seed, site), rows, grp, valid);
  return check_launch("drop_bwd_kernel");
}
int drop_bwd_run(const float* dx, float* dz, float p, unsigned long long seed, int site, int rows, int grp, int valid,
                 cudaStream_t st) {
  return drop_bwd_run_t<float>(dx, dz, p, seed, site, rows, grp, valid, st);
}
int drop_bwd_run(const __nv_bfloat16* dx, __nv_bfloat16* dz, float p, unsigned long long seed, int site, int rows, int grp,
                 int valid, cudaStream_t st) {
  return drop_bwd_run_t<__nv_bfloat16>(dx, dz, p, seed, site, rows, grp, valid, st);
}

}  // namespace aitb

using namespace aitb;

extern "C" int aitb_dropout_mask(float p, unsigned long long seed, int site, int rows, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(out && rows > 0 && p >= 0.f && p < 1.f, "aitb_dropout_mask: bad arguments");
  int grid = (rows + 7) / 8;
  if (grid > 8 * current_sm_count()) grid = 8 * current_sm_count();
  drop_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(make_drop(p, seed, site), rows, out);
  return check_launch("drop_mask_kernel");
}

extern "C" int aitb_attn_dropout_mask(float p, unsigned long long seed, int site, int G, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(out && G > 0 && p >= 0.f && p < 1.f, "aitb_attn_dropout_mask: bad arguments");
  attn_drop_mask_kernel<<<8 * current_sm_count(), 256, 0, (cudaStream_t)stream>>>(make_drop(p, seed, site), G, out);
  return check_launch("attn_drop_mask_kernel");
}
