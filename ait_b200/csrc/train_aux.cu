// Small data-movement kernels of the layer-4 (`RCNN_top`) training path (lib/model/faster_rcnn/
// resnet_coatt_transformer_sk.py:73-109, 476-485; the reference gets the backward from torch autograd):
//   relu_bwd        g = dy (rounded to tf32, nearest: it only feeds tensor-core GEMMs) where the saved activation y > 0, else 0
//   im2col3x3       [G, s, s, C] -> [G*s*s, 9*C] tap-major rows (zero padding): the 3x3 weight gradient then is ONE
//                   wgrad GEMM dW[out, 9*C] += dY^T * cols, in the same tap-major layout the forward packs its weights in
//   map_subsample   [G, S, S, C] -> [G, s, s, C] every `stride`-th position (the stride-2 1x1 convolutions of the first
//                   bottleneck read exactly these rows)
//   map_upsample    its adjoint: scatter to the strided positions, zeros elsewhere
// and of the SKNet training path (lib/model/modules/blocks_coatt_transformer_sk.py:960-998: v = relu(conv1x1_g8(x))^2 +
// relu(conv3x3_g8(x))^2, the selective-kernel attention is computed and discarded by the reference):
//   sk_combine      out = r1^2 + r3^2 from the two saved post-ReLU branch maps
//   sk_combine_bwd  d1 = 2 dv r1, d3 = 2 dv r3 (relu(z)^2 is C1: no mask needed), rounded to tf32 (RN) for the GEMMs
//   im2col3x3 with group_c < C: columns ordered (group, tap, channel-in-group), so the weight gradient of a grouped
//                   3x3 convolution is one wgrad per group over a contiguous [rows, 9*group_c] slice
// fp32, HBM-bound, 128-bit accesses (C % 4 == 0).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 d = dy[i], a = y[i];
    out[i] = make_float4(a.x > 0.f ? rn_tf32(d.x) : 0.f, a.y > 0.f ? rn_tf32(d.y) : 0.f, a.z > 0.f ? rn_tf32(d.z) : 0.f,
                         a.w > 0.f ? rn_tf32(d.w) : 0.f);
  }
}

// one CTA per output row (g, y, x); thread -> (group, tap, channel quad); gc4 = group_c / 4 (= C / 4: ungrouped)
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, int s, int C, int gc4, float* __restrict__ out) {
  const int row = blockIdx.x;
  const int g = row / (s * s), p = row - g * s * s, py = p / s, px = p - py * s;
  const int c4 = C / 4;
  float4* o = reinterpret_cast<float4*>(out + (size_t)row * 9 * C);
  for (int i = threadIdx.x; i < 9 * c4; i += blockDim.x) {
    const int grp = i / (9 * gc4), r = i - grp * 9 * gc4;
    const int tap = r / gc4, c = grp * gc4 + (r - tap * gc4);
    const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < s && xx >= 0 && xx < s)
      v = reinterpret_cast<const float4*>(x + ((size_t)(g * s + yy) * s + xx) * C)[c];
    o[i] = v;
  }
}

__global__ void __launch_bounds__(256)
sk_combine_kernel(const float4* __restrict__ r1, const float4* __restrict__ r3, float4* __restrict__ out, size_t n4, int round_tf) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = r1[i], b = r3[i];
    float4 v = make_float4(a.x * a.x + b.x * b.x, a.y * a.y + b.y * b.y, a.z * a.z + b.z * b.z, a.w * a.w + b.w * b.w);
    if (round_tf) v = make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
sk_combine_bwd_kernel(const float4* __restrict__ dv, const float4* __restrict__ r1, const float4* __restrict__ r3,
                      float4* __restrict__ d1, float4* __restrict__ d3, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 g = dv[i], a = r1[i], b = r3[i];
    const float gx = 2.f * g.x, gy = 2.f * g.y, gz = 2.f * g.z, gw = 2.f * g.w;
    d1[i] = make_float4(rn_tf32(gx * a.x), rn_tf32(gy * a.y), rn_tf32(gz * a.z), rn_tf32(gw * a.w));
    d3[i] = make_float4(rn_tf32(gx * b.x), rn_tf32(gy * b.y), rn_tf32(gz * b.z), rn_tf32(gw * b.w));
  }
}

// gather = 1: out [G, s, s, C] = x [G, S, S, C] at (stride*y, stride*x); gather = 0: out [G, S, S, C] = scatter of x [G, s, s, C]
__global__ void __launch_bounds__(256)
map_resample_kernel(const float* __restrict__ x, int S, int s, int stride, int C, int gather, float* __restrict__ out) {
  const int c4 = C / 4;
  if (gather) {
    const int row = blockIdx.x;                       // (g, y, x) of the small map
    const int g = row / (s * s), p = row - g * s * s, py = p / s, px = p - py * s;
    const float4* src = reinterpret_cast<const float4*>(x + ((size_t)(g * S + py * stride) * S + px * stride) * C);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * C);
    for (int i = threadIdx.x; i < c4; i += blockDim.x) dst[i] = src[i];
  } else {
    const int row = blockIdx.x;                       // (g, y, x) of the large map
    const int g = row / (S * S), p = row - g * S * S, py = p / S, px = p - py * S;
    float4* dst = reinterpret_cast<float4*>(out + (size_t)row * C);
    const bool hit = py % stride == 0 && px % stride == 0 && py / stride < s && px / stride < s;
    const float4* src = reinterpret_cast<const float4*>(x + ((size_t)(g * s + py / stride) * s + px / stride) * C);
    for (int i = threadIdx.x; i < c4; i += blockDim.x) dst[i] = hit ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace aitb

using namespace aitb;

// channels-last map [B, H, W, C] -> zero-bordered copy [B, H + 2, W + 2, C] (one CTA per destination position): the 3x3 weight
// gradient on an arbitrary H x W map then is nine plain row-contraction GEMMs on row-shifted views of the padded copies
// (ait_b200/rpn_train.py)
__global__ void __launch_bounds__(256)
map_pad_kernel(const float* __restrict__ x, int H, int W, int C, float* __restrict__ out) {
  const int Wp = W + 2, Hp = H + 2;
  const int pos = blockIdx.x;                     // (b, y', x') of the padded map
  const int xp = pos % Wp, yp = (pos / Wp) % Hp, b = pos / (Wp * Hp);
  float4* o = reinterpret_cast<float4*>(out + (size_t)pos * C);
  if (xp == 0 || xp == Wp - 1 || yp == 0 || yp == Hp - 1) {
    for (int c = threadIdx.x; c < C / 4; c += 256) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* s = reinterpret_cast<const float4*>(x + ((size_t)(b * H + yp - 1) * W + xp - 1) * C);
  for (int c = threadIdx.x; c < C / 4; c += 256) o[c] = __ldg(s + c);
}

extern "C" {

int aitb_relu_bwd(const float* dy, const float* y, float* out, size_t n, aitb_stream_t stream) {
  AITB_REQUIRE(dy && y && out && n > 0 && n % 4 == 0, "aitb_relu_bwd: bad arguments (n must be a positive multiple of 4)");
  AITB_REQUIRE((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)out) & 15) == 0, "aitb_relu_bwd: pointers must be 16-byte aligned");
  const size_t n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < (size_t)(8 * current_sm_count()) ? (n4 + 255) / 256 : 8 * current_sm_count());
  relu_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y),
                                                            reinterpret_cast<float4*>(out), n4);
  return check_launch("relu_bwd_kernel");
}

int aitb_im2col3x3_grouped(const float* x, int G, int s, int C, int group_c, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(x && out && G > 0 && s > 0 && C > 0 && C % 4 == 0, "aitb_im2col3x3: bad arguments");
  AITB_REQUIRE(group_c > 0 && group_c % 4 == 0 && C % group_c == 0, "aitb_im2col3x3: group_c=%d must divide C=%d and be a multiple of 4", group_c, C);
  AITB_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "aitb_im2col3x3: pointers must be 16-byte aligned");
  im2col3x3_kernel<<<G * s * s, 256, 0, (cudaStream_t)stream>>>(x, s, C, group_c / 4, out);
  return check_launch("im2col3x3_kernel");
}

int aitb_im2col3x3(const float* x, int G, int s, int C, float* out, aitb_stream_t stream) {
  return aitb_im2col3x3_grouped(x, G, s, C, C, out, stream);
}

static int ew_blocks(size_t n4) {
  const size_t want = (n4 + 255) / 256, cap = (size_t)(8 * current_sm_count());
  return (int)(want < cap ? want : cap);
}

int aitb_sk_combine(const float* r1, const float* r3, float* out, size_t n, int round_tf32, aitb_stream_t stream) {
  AITB_REQUIRE(r1 && r3 && out && n > 0 && n % 4 == 0, "aitb_sk_combine: bad arguments (n must be a positive multiple of 4)");
  AITB_REQUIRE((((uintptr_t)r1 | (uintptr_t)r3 | (uintptr_t)out) & 15) == 0, "aitb_sk_combine: pointers must be 16-byte aligned");
  sk_combine_kernel<<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(r1), reinterpret_cast<const float4*>(r3),
                                                                       reinterpret_cast<float4*>(out), n / 4, round_tf32);
  return check_launch("sk_combine_kernel");
}

int aitb_sk_combine_bwd(const float* dv, const float* r1, const float* r3, float* d1, float* d3, size_t n, aitb_stream_t stream) {
  AITB_REQUIRE(dv && r1 && r3 && d1 && d3 && n > 0 && n % 4 == 0, "aitb_sk_combine_bwd: bad arguments (n must be a positive multiple of 4)");
  AITB_REQUIRE((((uintptr_t)dv | (uintptr_t)r1 | (uintptr_t)r3 | (uintptr_t)d1 | (uintptr_t)d3) & 15) == 0,
               "aitb_sk_combine_bwd: pointers must be 16-byte aligned");
  sk_combine_bwd_kernel<<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dv), reinterpret_cast<const float4*>(r1), reinterpret_cast<const float4*>(r3),
      reinterpret_cast<float4*>(d1), reinterpret_cast<float4*>(d3), n / 4);
  return check_launch("sk_combine_bwd_kernel");
}

int aitb_map_subsample(const float* x, int G, int S, int s, int stride, int C, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(x && out && G > 0 && S > 0 && s > 0 && stride > 0 && (s - 1) * stride < S && C % 4 == 0, "aitb_map_subsample: bad arguments");
  map_resample_kernel<<<G * s * s, 256, 0, (cudaStream_t)stream>>>(x, S, s, stride, C, 1, out);
  return check_launch("map_resample_kernel");
}

int aitb_map_upsample(const float* x, int G, int S, int s, int stride, int C, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(x && out && G > 0 && S > 0 && s > 0 && stride > 0 && (s - 1) * stride < S && C % 4 == 0, "aitb_map_upsample: bad arguments");
  map_resample_kernel<<<G * S * S, 256, 0, (cudaStream_t)stream>>>(x, S, s, stride, C, 0, out);
  return check_launch("map_resample_kernel");
}

int aitb_map_pad(const float* x, int B, int H, int W, int C, float* out, aitb_stream_t stream) {
  AITB_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "aitb_map_pad: bad arguments");
  AITB_REQUIRE((long long)B * (H + 2) * (W + 2) < (1ll << 31), "aitb_map_pad: too many positions");
  map_pad_kernel<<<B * (H + 2) * (W + 2), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, out);
  return check_launch("map_pad_kernel");
}

}  // extern "C"
