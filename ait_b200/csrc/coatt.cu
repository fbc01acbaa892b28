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

}  // namespace aitb
