// GroupNorm for the co-attention block (lib/model/modules/blocks_coatt_transformer_sk.py:31-42: `gn(32, 1024)` after
// the theta / omega 1x1 convolutions) on token-major [B, N, C] fp32 activations: one pass of per-(image, group)
// sums in double precision, one pass that normalises, applies the affine parameters and adds the non-local
// residual (`non_img + identity_img`, :108-109).  HBM-bound elementwise work; the GEMMs around it are in capi.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

// sums[b][g] = (sum, sum of squares) over N positions x (C / groups) channels.  grid (row chunks, B), block 256:
// thread t owns channels 4t..4t+3 (C == 1024), `cpg / 4` consecutive threads share a group.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, int N, int rows_per_cta, int cpg, double* __restrict__ sums) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  const float* xb = x + (size_t)b * N * 1024 + threadIdx.x * 4;
  float s = 0.f, ss = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * 1024));
    s += v.x + v.y + v.z + v.w;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const int tpg = cpg / 4;  // threads per group (power of two <= 32)
  for (int o = tpg >> 1; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x % tpg) == 0) {
    const int g = threadIdx.x / tpg;
    atomicAdd(&sums[((size_t)b * (1024 / cpg) + g) * 2], (double)s);
    atomicAdd(&sums[((size_t)b * (1024 / cpg) + g) * 2 + 1], (double)ss);
  }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ identity, const double* __restrict__ sums,
                const float* __restrict__ gamma, const float* __restrict__ beta, int N, int cpg, float eps,
                float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = threadIdx.x * 4;
  const int g = c / cpg;
  const double cnt = (double)N * cpg;
  const double mean_d = sums[((size_t)b * (1024 / cpg) + g) * 2] / cnt;
  const double var_d = sums[((size_t)b * (1024 / cpg) + g) * 2 + 1] / cnt - mean_d * mean_d;
  const float mean = (float)mean_d;
  const float rstd = rsqrtf(fmaxf((float)var_d, 0.f) + eps);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  for (int r = blockIdx.x; r < N; r += gridDim.x) {
    const size_t off = ((size_t)b * N + r) * 1024 + c;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    const float4 id = __ldg(reinterpret_cast<const float4*>(identity + off));
    float4 o;
    o.x = (v.x - mean) * rstd * ga.x + be.x + id.x;
    o.y = (v.y - mean) * rstd * ga.y + be.y + id.y;
    o.z = (v.z - mean) * rstd * ga.z + be.z + id.z;
    o.w = (v.w - mean) * rstd * ga.w + be.w + id.w;
    *reinterpret_cast<float4*>(out + off) = o;
  }
}

// out = GroupNorm(x; groups of `cpg` channels over all N positions of an image) * gamma + beta + identity
int group_norm_residual_run(const float* x, const float* identity, const float* gamma, const float* beta, int B, int N,
                            int groups, float eps, double* sums /*[B, groups, 2] scratch*/, float* out, cudaStream_t st) {
  AITB_REQUIRE(x && identity && gamma && beta && sums && out, "aitb_group_norm: null pointer");
  AITB_REQUIRE(B > 0 && B <= 65535 && N > 0 && groups > 0 && 1024 % groups == 0, "aitb_group_norm: bad sizes");
  const int cpg = 1024 / groups;
  AITB_REQUIRE(cpg % 4 == 0 && cpg <= 128 && (cpg & (cpg - 1)) == 0, "aitb_group_norm: channels per group must be 4..128, power of 2");
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * groups * 2 * sizeof(double), st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_group_norm: memset failed: %s", cudaGetErrorString(e));
  int chunks = (N + 31) / 32;
  if (chunks > 128) chunks = 128;
  const int rpc = (N + chunks - 1) / chunks;
  chunks = (N + rpc - 1) / rpc;
  gn_stats_kernel<<<dim3(chunks, B), 256, 0, st>>>(x, N, rpc, cpg, sums);
  if (check_launch("gn_stats_kernel")) return 1;
  gn_apply_kernel<<<dim3(N < 128 ? N : 128, B), 256, 0, st>>>(x, identity, sums, gamma, beta, N, cpg, eps, out);
  return check_launch("gn_apply_kernel");
}

// ---------------------------------------------------------------------------------------------
// GroupNorm backward (training of the co-attention block; the reference gets it from torch autograd over nn.GroupNorm).
//   y = x_hat * gamma + beta (+ identity),  x_hat = (x - mean_bg) * rstd_bg,  group = (image b, cpg consecutive channels)
//   a = dy * gamma;  dx = rstd * (a - mean_group(a) - x_hat * mean_group(a * x_hat));  dgamma_c += dy * x_hat;  dbeta_c += dy
// Pass 1 (gn_bwd_stats_kernel): per-(b, g) sums of a and a * x_hat in double precision + the per-channel parameter
// gradients (registers over the CTA's rows, one atomic per channel and CTA).  Pass 2 (gn_bwd_apply_kernel): dx.
// `sums` are the forward's (sum, sum of squares) per (b, g).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_mean_rstd(const double* __restrict__ sums, int b, int g, int groups, double cnt, float eps,
                                             float& mean, float& rstd) {
  const double mean_d = sums[((size_t)b * groups + g) * 2] / cnt;
  const double var_d = sums[((size_t)b * groups + g) * 2 + 1] / cnt - mean_d * mean_d;
  mean = (float)mean_d;
  rstd = rsqrtf(fmaxf((float)var_d, 0.f) + eps);
}

__global__ void __launch_bounds__(256)
gn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const double* __restrict__ sums,
                    const float* __restrict__ gamma, int N, int rows_per_cta, int cpg, float eps, double* __restrict__ bsums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int b = blockIdx.y, c = threadIdx.x * 4, g = c / cpg, groups = 1024 / cpg;
  float mean, rstd;
  gn_mean_rstd(sums, b, g, groups, (double)N * cpg, eps, mean, rstd);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  float s1 = 0.f, s2 = 0.f;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; ++r) {
    const size_t off = ((size_t)b * N + r) * 1024 + c;
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy + off));
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    const float h0 = (v.x - mean) * rstd, h1 = (v.y - mean) * rstd, h2 = (v.z - mean) * rstd, h3 = (v.w - mean) * rstd;
    dg.x += d.x * h0; dg.y += d.y * h1; dg.z += d.z * h2; dg.w += d.w * h3;
    db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
    const float a0 = d.x * ga.x, a1 = d.y * ga.y, a2 = d.z * ga.z, a3 = d.w * ga.w;
    s1 += a0 + a1 + a2 + a3;
    s2 += a0 * h0 + a1 * h1 + a2 * h2 + a3 * h3;
  }
  const int tpg = cpg / 4;
  for (int o = tpg >> 1; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x % tpg) == 0) {
    atomicAdd(&bsums[((size_t)b * groups + g) * 2], (double)s1);
    atomicAdd(&bsums[((size_t)b * groups + g) * 2 + 1], (double)s2);
  }
  atomicAdd(dgamma + c, dg.x); atomicAdd(dgamma + c + 1, dg.y); atomicAdd(dgamma + c + 2, dg.z); atomicAdd(dgamma + c + 3, dg.w);
  atomicAdd(dbeta + c, db.x); atomicAdd(dbeta + c + 1, db.y); atomicAdd(dbeta + c + 2, db.z); atomicAdd(dbeta + c + 3, db.w);
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const double* __restrict__ sums,
                    const double* __restrict__ bsums, const float* __restrict__ gamma, int N, int cpg, float eps, int round_tf,
                    float* __restrict__ dx) {
  const int b = blockIdx.y, c = threadIdx.x * 4, g = c / cpg, groups = 1024 / cpg;
  const double cnt = (double)N * cpg;
  float mean, rstd;
  gn_mean_rstd(sums, b, g, groups, cnt, eps, mean, rstd);
  const float m1 = (float)(bsums[((size_t)b * groups + g) * 2] / cnt), m2 = (float)(bsums[((size_t)b * groups + g) * 2 + 1] / cnt);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  for (int r = blockIdx.x; r < N; r += gridDim.x) {
    const size_t off = ((size_t)b * N + r) * 1024 + c;
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy + off));
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    float4 o;
    o.x = rstd * (d.x * ga.x - m1 - (v.x - mean) * rstd * m2);
    o.y = rstd * (d.y * ga.y - m1 - (v.y - mean) * rstd * m2);
    o.z = rstd * (d.z * ga.z - m1 - (v.z - mean) * rstd * m2);
    o.w = rstd * (d.w * ga.w - m1 - (v.w - mean) * rstd * m2);
    if (round_tf) st4r(dx + off, o);   // the consumer is a tf32 GEMM
    else *reinterpret_cast<float4*>(dx + off) = o;
  }
}

int group_norm_backward_run(const float* dy, const float* x, const double* sums, const float* gamma, int B, int N, int groups,
                            float eps, int round_tf, double* bsums /*[B, groups, 2] scratch*/, float* dx, float* dgamma,
                            float* dbeta, cudaStream_t st) {
  AITB_REQUIRE(dy && x && sums && gamma && bsums && dx && dgamma && dbeta, "aitb_group_norm_backward: null pointer");
  AITB_REQUIRE(B > 0 && B <= 65535 && N > 0 && groups > 0 && 1024 % groups == 0, "aitb_group_norm_backward: bad sizes");
  const int cpg = 1024 / groups;
  AITB_REQUIRE(cpg % 4 == 0 && cpg <= 128 && (cpg & (cpg - 1)) == 0,
               "aitb_group_norm_backward: channels per group must be 4..128, power of 2");
  cudaError_t e = cudaMemsetAsync(bsums, 0, (size_t)B * groups * 2 * sizeof(double), st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_group_norm_backward: memset failed: %s", cudaGetErrorString(e));
  int chunks = (N + 31) / 32;
  if (chunks > 128) chunks = 128;
  const int rpc = (N + chunks - 1) / chunks;
  chunks = (N + rpc - 1) / rpc;
  gn_bwd_stats_kernel<<<dim3(chunks, B), 256, 0, st>>>(dy, x, sums, gamma, N, rpc, cpg, eps, bsums, dgamma, dbeta);
  if (check_launch("gn_bwd_stats_kernel")) return 1;
  gn_bwd_apply_kernel<<<dim3(N < 128 ? N : 128, B), 256, 0, st>>>(dy, x, sums, bsums, gamma, N, cpg, eps, round_tf, dx);
  return check_launch("gn_bwd_apply_kernel");
}

}  // namespace aitb

using namespace aitb;

/* GroupNorm(groups, 1024) (+ identity) on token-major [B, N, 1024] fp32, forward keeping the per-(image, group) sums, and its
 * backward -- the training path of the co-attention block (ait_b200/coatt_train.py). */
extern "C" int aitb_group_norm_forward(const float* x, const float* identity, const float* gamma, const float* beta, int B, int N,
                                       int groups, float eps, double* sums, float* out, aitb_stream_t stream) {
  return group_norm_residual_run(x, identity, gamma, beta, B, N, groups, eps, sums, out, (cudaStream_t)stream);
}
extern "C" int aitb_group_norm_backward(const float* dy, const float* x, const double* sums, const float* gamma, int B, int N,
                                        int groups, float eps, int round_tf32, double* bsums, float* dx, float* dgamma,
                                        float* dbeta, aitb_stream_t stream) {
  return group_norm_backward_run(dy, x, sums, gamma, B, N, groups, eps, round_tf32, bsums, dx, dgamma, dbeta,
                                 (cudaStream_t)stream);
}
