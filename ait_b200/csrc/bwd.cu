// Backward building blocks of the AIT training step (BASELINE config 4: forward + backward of
// Transformer.forward, system/Models.py:231-280, with the ROIAlign gradient of roi_align.cu in front).
// The reference gets all of these from torch autograd (cuBLAS / ATen kernels); here:
//
//   wgrad_tcgen05_kernel   dW[N, K] += dY[M, N]^T * X[M, K]   -- the contraction runs over the ROWS of
//                          two row-major activations, so both operands are fed to tcgen05.mma as
//                          MN-major tiles (TMA boxes of 32 rows x 128 B land in shared memory exactly in
//                          the canonical MN-major SWIZZLE_128B layout: no transposed copy of any
//                          activation exists), split over row chunks across CTAs, accumulated into the
//                          fp32 gradient with vector reductions (red.global.add.v4.f32).
//   ln_bwd_kernel          LayerNorm backward from the saved OUTPUT y and 1/sigma (x_hat = (y - beta) / gamma),
//                          with the encoder's 64 -> 49 row compaction and d(gamma), d(beta).
//   colsum_kernel          bias gradients.
//   bsum_kernel            gradient of the unit -> proposal broadcast (sum over the P pairs of a unit).
//   attn_bwd_kernel        selective-head attention backward (softmax, gate, head sum), one CTA per pair.
//
// (dgrad GEMMs dX = dY * W reuse the forward kernel of gemm.cu with a transposed weight copy.)
// Two storage configurations: fp32 storage with tf32 tensor-core math (T = float), and bf16 storage with bf16 tensor-core
// math (T = __nv_bfloat16); fp32 accumulation and fp32 parameter gradients in both.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

int encode_map_f32_mn(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const char* what);

int encode_map_bf16(CUtensorMap* tm, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* what);

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
static constexpr int kWgStages = 4;
static constexpr int kWgRows = 32;                      // tf32: contraction rows per pipeline stage (4 MMAs of K = 8)
static constexpr int kWgABytes = 4 * kWgRows * 128;     // 16 KB: the 128 dY columns (MMA M = 128) of one stage, both configurations
static constexpr int kWgThreads = 192;
// per storage type: a stage holds kRows contraction rows as column groups of kGW elements (= 128 bytes, one TMA box each);
// one MMA consumes kKRows rows (K = 8 tf32 / K = 16 bf16), i.e. kKRows * 128 bytes of every group
template <bool BF16> struct WgCfg {
  static constexpr int kRows = BF16 ? 64 : 32;
  static constexpr int kGW = BF16 ? 64 : 32;
  static constexpr int kKRows = BF16 ? 16 : 8;
  static constexpr int kEs = BF16 ? 2 : 4;
};

struct WgradParams {
  int M;          // rows (contraction length)
  int N, K;       // dW is [N, K]
  int bn;         // K-columns per tile (MMA N): 64 | 128 | 256
  int n_tiles, k_tiles, splits;
  int rows_per_split;   // multiple of kWgRows
  float* dw;
  int ldw;
  // X addressing.  conv_S = 0: X is a row-major matrix, columns [nt * x_group_stride + k, ...).  conv_S = S > 0: X is a
  // channels-last S x S map [G, S, S, C] read through a 4-D TMA view, dW column k = tap * conv_cg + c (tap-major 3x3):
  // the rows of a stage are the map positions shifted by the tap, out-of-map positions zero-filled by TMA (= the padding)
  // -- the weight gradient of a 3x3 convolution without an im2col buffer.  x_group_stride: grouped convolution, n-tile
  // nt (128 output channels = one group) reads input channels [nt * x_group_stride, + conv_cg).
  int conv_S, conv_cg, x_group_stride;
};

// MN-major tf32 operand.  tcgen05 accepts exactly one shared-memory layout for it: "128-byte swizzle with a
// 32-byte atom" (layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- rows of 128 contiguous bytes
// along MN (32 fp32), FOUR contraction rows per 512-byte atom, the 32-byte chunks of a row XOR-ed with
// (row & 3).  Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units: atoms `lbo_bytes` apart along
// MN and 512 bytes apart along K, which is exactly what a TMA box of {32 fp32, R rows} leaves behind.
// MN-major bf16 operand: the ordinary 128-byte swizzle (layout type 2; TMA: CU_TENSOR_MAP_SWIZZLE_128B) -- rows of 128
// contiguous bytes along MN (64 bf16), EIGHT contraction rows per 1024-byte atom, 16-byte chunks XOR-ed with (row & 7).
template <bool BF16>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((BF16 ? 1024 : 512) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(BF16 ? 2 : 1) << 61;   // SWIZZLE_128B | SWIZZLE_128B_BASE32B
  return d;
}

template <bool BF16>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                     const WgradParams p) {
  using W = WgCfg<BF16>;
  constexpr int kWgRows = W::kRows, kGW = W::kGW;    // shadow the tf32 constants of the namespace
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int b_bytes = (p.bn / kGW) * kWgRows * 128;
  const int stage_bytes = kWgABytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // tile / split of this CTA
  const int tile = blockIdx.x % (p.n_tiles * p.k_tiles);
  const int split = blockIdx.x / (p.n_tiles * p.k_tiles);
  const int nt = tile / p.k_tiles, kt = tile - nt * p.k_tiles;
  const int row0 = split * p.rows_per_split;
  int rows = p.M - row0;
  if (rows > p.rows_per_split) rows = p.rows_per_split;
  const int iters = rows > 0 ? (rows + kWgRows - 1) / kWgRows : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        // one {128 bytes of columns, kWgRows rows} box per column group (= one MN-major atom column); rows beyond M
        // are zero-filled by TMA
        const int r = row0 + it * kWgRows;
        for (int g = 0; g < 128 / kGW; ++g) tma_load_2d(sa + g * (kWgRows * 128), &tmY, &full_bar[s], nt * 128 + g * kGW, r);
        const int xc = kt * p.bn;
        if (p.conv_S == 0) {
          for (int g = 0; g < p.bn / kGW; ++g)
            tma_load_2d(sa + kWgABytes + g * (kWgRows * 128), &tmX, &full_bar[s], nt * p.x_group_stride + xc + g * kGW, r);
        } else {
          const int tap = xc / p.conv_cg, cc = xc - tap * p.conv_cg + nt * p.x_group_stride;
          const int ss = p.conv_S * p.conv_S;
          const int gi = r / ss, y0 = (r - gi * ss) / p.conv_S;
          for (int g = 0; g < p.bn / kGW; ++g)
            tma_load_4d(sa + kWgABytes + g * (kWgRows * 128), &tmX, &full_bar[s], cc + g * kGW, tap % 3 - 1, y0 + tap / 3 - 1, gi);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      // tf32 / bf16 operands, fp32 accumulate, A and B both MN-major (bits 15 / 16)
      const uint32_t idesc = make_idesc(BF16 ? 1u : 2u, 128u, (uint32_t)p.bn) | (1u << 15) | (1u << 16);
      for (int it = 0; it < iters; ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t adesc = make_mnmajor_desc<BF16>(sa, kWgRows * 128);
        const uint64_t bdesc = make_mnmajor_desc<BF16>(sa + kWgABytes, kWgRows * 128);
        constexpr int kStep = W::kKRows * 128 / 16;   // 8 rows (two 512-byte atoms) per K = 8 tf32 MMA; 16 rows (two 1024-byte atoms) per K = 16 bf16 MMA
#pragma unroll
        for (int k = 0; k < kWgRows / W::kKRows; ++k)
          umma_ss<W::kEs>(tmem_base, adesc + (uint64_t)(k * kStep), bdesc + (uint64_t)(k * kStep), idesc, (it | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(acc_full);
    }
  } else if (iters > 0) {
    const int q = warp & 3;
    const int n = nt * 128 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float* drow = p.dw + (size_t)n * p.ldw + kt * p.bn;
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t raw[32];
      tmem_ld32(t_row + c0, raw);
      tmem_ld_wait();
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(raw[j])),
                       "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                       "f"(__uint_as_float(raw[j + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ---------------------------------------------------------------------------------------------
// Selective-head attention backward, one CTA (4 warps) per proposal-query pair.
//
// Forward (attn.cu; system/Modules.py:16-29, SubLayers.py:22-39,89-92):
//   P_h = softmax(mask(Q_h K_h^T / 8)),  O_h = P_h V_h,  s = mean_T(sum_h O_h),
//   z = W_sk s + b_sk,  g = softmax over h of z,  out = sum_h O_h * g_h
// Backward, given dOut [64, 64]:
//   dg_h[c] = sum_t dOut[t,c] O_h[t,c];  dz_h = g_h * (dg_h - sum_h' g_h' dg_h');  ds = W_sk^T dz
//   dO_h[t,c] = dOut[t,c] g_h[c] + ds[c] / T
//   dV_h = P_h^T dO_h;  dP = dO_h V_h^T;  dS = P * (dP - rowsum(dP * P));  dQ_h = dS K_h / 8;  dK_h = dS^T Q_h / 8
// Pass A recomputes P_h / O_h of every head for dg and s; pass B recomputes P_h again and forms the five
// products.  64^3 products run on mma.sync.m16n8k8 tf32 with fp32 tiles in shared memory; (dz, s) are
// written per pair so that d(W_sk) = dz^T s and d(b_sk) = colsum(dz) become one small wgrad / column sum.
// Masked keys have P = 0, hence dS = 0 and zero dK / dV rows: no special casing.
// ---------------------------------------------------------------------------------------------
static constexpr int kBT = 64;        // tokens == head dim == 64
static constexpr int kBS = 68;        // fp32 tile row stride (conflict-free A-fragment reads)
static constexpr int kBH = 8;
static constexpr int kAttnBwdThreads = 128;
static constexpr int kAttnBwdSmem = (6 * kBT * kBS + 4 * kBT /*col partials*/ + 2 * kBH * kBT /*gate, dg -> dz*/ + 2 * kBT) * 4;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// C[16 x 64] (this warp's rows r0..r0+15) = A * B over a contraction of 64.
//   A(i, k) = TA ? a[k * kBS + i] : a[i * kBS + k]        (i = output row)
//   B(k, n) = BKN ? b[k * kBS + n] : b[n * kBS + k]       (n = output column)
template <bool TA, bool BKN>
__device__ __forceinline__ void mm64(const float* __restrict__ a, const float* __restrict__ b, int r0, float (&c)[8][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < kBT; k0 += 8) {
    uint32_t af[4];
    auto A = [&](int i, int k) { return __float_as_uint(TA ? a[k * kBS + i] : a[i * kBS + k]); };
    af[0] = A(r0 + g, k0 + t);
    af[1] = A(r0 + g + 8, k0 + t);
    af[2] = A(r0 + g, k0 + t + 4);
    af[3] = A(r0 + g + 8, k0 + t + 4);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      auto B = [&](int k, int n) { return __float_as_uint(BKN ? b[k * kBS + n] : b[n * kBS + k]); };
      mma_tf32(c[nt], af, B(k0 + t, nt * 8 + g), B(k0 + t + 4, nt * 8 + g));
    }
  }
}

// global [64 rows, ld] (64 columns from `g`) -> shared fp32 tile, rounded to tf32 (bf16 values are tf32-exact)
template <typename T>
__device__ __forceinline__ void load_tile_f32(const T* __restrict__ g, int ld, float* __restrict__ s) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = threadIdx.x + kAttnBwdThreads * i;
    const int r = f >> 4, c4 = (f & 15) * 4;
    const float4 v = ld4(g + (size_t)r * ld + c4);
    *reinterpret_cast<float4*>(s + r * kBS + c4) = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
  }
}

// this warp's C fragment -> global rows r0.., 64 columns (rounded to the storage's operand precision)
template <typename T>
__device__ __forceinline__ void store_frag(const float (&c)[8][4], T* __restrict__ gp, int ld, int r0, float scale) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    st2r(gp + (size_t)(r0 + g) * ld + nt * 8 + 2 * t, c[nt][0] * scale, c[nt][1] * scale);
    st2r(gp + (size_t)(r0 + g + 8) * ld + nt * 8 + 2 * t, c[nt][2] * scale, c[nt][3] * scale);
  }
}

// mask + softmax of this warp's 16 score rows, in the mma C layout (as in attn.cu)
__device__ __forceinline__ void softmax_frag(float (&p)[8][4], int row0, int mask_mode, int n_keys) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r_lo = row0 + g, r_hi = row0 + g + 8;
  float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + 2 * t + (e & 1);
      const int row = (e < 2) ? r_lo : r_hi;
      const bool masked = mask_mode == 0 ? (col >= n_keys) : (col > row);
      const float v = masked ? -1e9f : p[nt][e] * 0.125f;
      p[nt][e] = v;
      if (e < 2) m_lo = fmaxf(m_lo, v); else m_hi = fmaxf(m_hi, v);
    }
  }
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
  float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    p[nt][0] = expf(p[nt][0] - m_lo); p[nt][1] = expf(p[nt][1] - m_lo);
    p[nt][2] = expf(p[nt][2] - m_hi); p[nt][3] = expf(p[nt][3] - m_hi);
    s_lo += p[nt][0] + p[nt][1];
    s_hi += p[nt][2] + p[nt][3];
  }
  s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1);
  s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
  s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1);
  s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
  const float i_lo = 1.f / s_lo, i_hi = 1.f / s_hi;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    p[nt][0] *= i_lo; p[nt][1] *= i_lo; p[nt][2] *= i_hi; p[nt][3] *= i_hi;
  }
}

// fragment -> shared tile rows r0.. (rounded to tf32: it is an MMA operand next)
__device__ __forceinline__ void frag_to_smem(const float (&c)[8][4], float* __restrict__ s, int r0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<float2*>(s + (r0 + g) * kBS + nt * 8 + 2 * t) = make_float2(tf32r(c[nt][0]), tf32r(c[nt][1]));
    *reinterpret_cast<float2*>(s + (r0 + g + 8) * kBS + nt * 8 + 2 * t) = make_float2(tf32r(c[nt][2]), tf32r(c[nt][3]));
  }
}

// column sums over this warp's 16 rows of a fragment -> part[warp][64]
__device__ __forceinline__ void frag_colsum(const float (&c)[8][4], float* __restrict__ part) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float c0 = c[nt][0] + c[nt][2], c1 = c[nt][1] + c[nt][3];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if (g == 0) {
      part[warp * kBT + nt * 8 + 2 * t] = c0;
      part[warp * kBT + nt * 8 + 2 * t + 1] = c1;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kAttnBwdThreads, 2)
attn_bwd_kernel(const T* __restrict__ q, int ldq, int q_rep, const T* __restrict__ k, const T* __restrict__ v,
                int ldkv, const float* __restrict__ w_sk, const float* __restrict__ b_sk, const T* __restrict__ dout,
                int mask_mode, int n_keys, T* __restrict__ dq, int lddq, T* __restrict__ dk,
                T* __restrict__ dv, int lddkv, T* __restrict__ dz_out, T* __restrict__ s_out, DropCfg dc) {
  extern __shared__ __align__(16) float bsm[];
  float* sQ = bsm;
  float* sK = sQ + kBT * kBS;
  float* sV = sK + kBT * kBS;
  float* sDOut = sV + kBT * kBS;   // dOut of the pair (all heads)
  float* sP = sDOut + kBT * kBS;   // P_h, later reused for dO_h^T-free products
  float* sX = sP + kBT * kBS;      // dO_h (pass B), then dS
  float* part = sX + kBT * kBS;    // [4][64]
  float* gate = part + 4 * kBT;    // [8][64]
  float* dgv = gate + kBH * kBT;   // [8][64] dg, then dz
  float* svec = dgv + kBH * kBT;   // [64] s
  float* dsv = svec + kBT;         // [64] ds / T

  const int grp = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  const T* qg = q + (size_t)(grp / q_rep) * kBT * ldq;
  const T* kg = k + (size_t)grp * kBT * ldkv;
  const T* vg = v + (size_t)grp * kBT * ldkv;
  {  // dOut tile (kept in fp32: it is multiplied elementwise, rounded where it becomes an MMA operand)
    const T* dg_ = dout + (size_t)grp * kBT * kBT;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = tid + kAttnBwdThreads * i;
      const int r = f >> 4, c4 = (f & 15) * 4;
      *reinterpret_cast<float4*>(sDOut + r * kBS + c4) = ld4(dg_ + (size_t)r * kBT + c4);
    }
  }
  float s_acc = 0.f;  // thread c < 64: sum_h sum_t O_h[t, c]

  // ---------------- pass A: dg_h and s
  for (int h = 0; h < kBH; ++h) {
    __syncthreads();
    load_tile_f32(qg + h * kBT, ldq, sQ);
    load_tile_f32(kg + h * kBT, ldkv, sK);
    load_tile_f32(vg + h * kBT, ldkv, sV);
    __syncthreads();
    float p[8][4], o[8][4];
    mm64<false, false>(sQ, sK, row0, p);            // S = Q K^T
    softmax_frag(p, row0, mask_mode, n_keys);
    if (dc.thr) drop_frag(dc, grp, h, row0, p);     // P' = dropout(P): the forward's draws (attn.cu)
    frag_to_smem(p, sP, row0);
    __syncwarp();
    mm64<false, true>(sP, sV, row0, o);             // O_h = P' V   (only this warp's rows of P' are read)
    frag_colsum(o, part);
    __syncthreads();
    if (tid < kBT) s_acc += part[tid] + part[kBT + tid] + part[2 * kBT + tid] + part[3 * kBT + tid];
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {                // o *= dOut  (elementwise), then column sums = dg_h
      const float2 a = *reinterpret_cast<const float2*>(sDOut + (row0 + g) * kBS + nt * 8 + 2 * t);
      const float2 b = *reinterpret_cast<const float2*>(sDOut + (row0 + g + 8) * kBS + nt * 8 + 2 * t);
      o[nt][0] *= a.x; o[nt][1] *= a.y; o[nt][2] *= b.x; o[nt][3] *= b.y;
    }
    frag_colsum(o, part);
    __syncthreads();
    if (tid < kBT) dgv[h * kBT + tid] = part[tid] + part[kBT + tid] + part[2 * kBT + tid] + part[3 * kBT + tid];
  }
  if (tid < kBT) svec[tid] = s_acc * (1.f / kBT);
  __syncthreads();
  // ---------------- gate forward + backward
  for (int o = tid; o < kBH * kBT; o += kAttnBwdThreads) {
    const float* wr = w_sk + (size_t)o * kBT;
    float acc = __ldg(b_sk + o);
#pragma unroll 8
    for (int c = 0; c < kBT; c += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + c));
      acc += w4.x * svec[c] + w4.y * svec[c + 1] + w4.z * svec[c + 2] + w4.w * svec[c + 3];
    }
    gate[o] = acc;
  }
  __syncthreads();
  if (tid < kBT) {
    const int c = tid;
    float m = -INFINITY;
#pragma unroll
    for (int h = 0; h < kBH; ++h) m = fmaxf(m, gate[h * kBT + c]);
    float e[kBH], sum = 0.f;
#pragma unroll
    for (int h = 0; h < kBH; ++h) { e[h] = expf(gate[h * kBT + c] - m); sum += e[h]; }
    const float inv = 1.f / sum;
    float dot = 0.f;
#pragma unroll
    for (int h = 0; h < kBH; ++h) { e[h] *= inv; dot += e[h] * dgv[h * kBT + c]; }
#pragma unroll
    for (int h = 0; h < kBH; ++h) {
      gate[h * kBT + c] = e[h];
      const float dzv = e[h] * (dgv[h * kBT + c] - dot);
      dgv[h * kBT + c] = dzv;                                   // dz
      st1(dz_out + (size_t)grp * kBH * kBT + h * kBT + c, dzv);
    }
    st1(s_out + (size_t)grp * kBT + c, svec[c]);
  }
  __syncthreads();
  if (tid < kBT) {   // ds[c'] = sum_o W_sk[o, c'] dz[o]; stored as ds / T (the mean over T rows)
    float acc = 0.f;
    for (int o = 0; o < kBH * kBT; ++o) acc += __ldg(w_sk + (size_t)o * kBT + tid) * dgv[o];
    dsv[tid] = acc * (1.f / kBT);
  }

  // ---------------- pass B
  for (int h = 0; h < kBH; ++h) {
    __syncthreads();
    load_tile_f32(qg + h * kBT, ldq, sQ);
    load_tile_f32(kg + h * kBT, ldkv, sK);
    load_tile_f32(vg + h * kBT, ldkv, sV);
    // dO_h = dOut * g_h + ds / T  -> sX (tf32)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = tid + kAttnBwdThreads * i;
      const int r = f >> 4, c4 = (f & 15) * 4;
      const float4 d4 = *reinterpret_cast<const float4*>(sDOut + r * kBS + c4);
      const float4 g4 = *reinterpret_cast<const float4*>(gate + h * kBT + c4);
      const float4 s4 = *reinterpret_cast<const float4*>(dsv + c4);
      *reinterpret_cast<float4*>(sX + r * kBS + c4) = make_float4(tf32r(d4.x * g4.x + s4.x), tf32r(d4.y * g4.y + s4.y),
                                                                  tf32r(d4.z * g4.z + s4.z), tf32r(d4.w * g4.w + s4.w));
    }
    __syncthreads();
    float p[8][4], pm[8][4], dp[8][4];
    mm64<false, false>(sQ, sK, row0, p);
    softmax_frag(p, row0, mask_mode, n_keys);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { pm[nt][0] = p[nt][0]; pm[nt][1] = p[nt][1]; pm[nt][2] = p[nt][2]; pm[nt][3] = p[nt][3]; }
    if (dc.thr) drop_frag(dc, grp, h, row0, pm);    // P' = P * M (M = mask / (1 - p_drop)); pm == p without dropout
    frag_to_smem(pm, sP, row0);
    mm64<false, false>(sX, sV, row0, dp);           // dP' = dO_h V^T;  dP = dP' * M
    // dS = P * (dP - rowsum(dP * P)) = P' * dP' - P * rowsum(dP' * P')
    float r_lo = 0.f, r_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      r_lo += dp[nt][0] * pm[nt][0] + dp[nt][1] * pm[nt][1];
      r_hi += dp[nt][2] * pm[nt][2] + dp[nt][3] * pm[nt][3];
    }
    r_lo += __shfl_xor_sync(0xffffffffu, r_lo, 1);
    r_lo += __shfl_xor_sync(0xffffffffu, r_lo, 2);
    r_hi += __shfl_xor_sync(0xffffffffu, r_hi, 1);
    r_hi += __shfl_xor_sync(0xffffffffu, r_hi, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      dp[nt][0] = pm[nt][0] * dp[nt][0] - p[nt][0] * r_lo; dp[nt][1] = pm[nt][1] * dp[nt][1] - p[nt][1] * r_lo;
      dp[nt][2] = pm[nt][2] * dp[nt][2] - p[nt][2] * r_hi; dp[nt][3] = pm[nt][3] * dp[nt][3] - p[nt][3] * r_hi;
    }
    __syncthreads();                                // every warp is done reading sX (dO_h) as the A operand of dP ...
    // dV_h = P^T dO_h needs sX (dO_h) and sP from all warps; dS goes to a separate tile: reuse sQ? no -- Q is
    // still needed for dK.  Write dS over sDS = sX only after dV is done; so: dV first.
    {
      float dvf[8][4];
      mm64<true, true>(sP, sX, row0, dvf);          // [j, c] = sum_t P'[t, j] dO[t, c]
      store_frag(dvf, dv + (size_t)grp * kBT * lddkv + h * kBT, lddkv, row0, 1.f);
    }
    __syncthreads();
    frag_to_smem(dp, sX, row0);                     // dS (this warp's rows)
    __syncwarp();
    {
      float dqf[8][4];
      mm64<false, true>(sX, sK, row0, dqf);         // dQ = dS K / 8  (own rows of dS only)
      store_frag(dqf, dq + (size_t)grp * kBT * lddq + h * kBT, lddq, row0, 0.125f);
    }
    __syncthreads();
    {
      float dkf[8][4];
      mm64<true, true>(sX, sQ, row0, dkf);          // [j, d] = sum_t dS[t, j] Q[t, d] / 8
      store_frag(dkf, dk + (size_t)grp * kBT * lddkv + h * kBT, lddkv, row0, 0.125f);
    }
  }
}

template <typename T>
static int attn_bwd_run_t(const T* q, int ldq, int q_rep, const T* k, const T* v, int ldkv, const float* w_sk,
                          const float* b_sk, const T* dout, int G, int mask_mode, int n_keys, T* dq, int lddq, T* dk,
                          T* dv, int lddkv, T* dz, T* s_out, cudaStream_t stream, const DropCfg* drop) {
  DropCfg dc;
  dc.scale = 1.f; dc.thr = dc.k0 = dc.k1 = 0u;
  if (drop && drop->thr) dc = *drop;
  AITB_REQUIRE(G > 0 && q && k && v && w_sk && b_sk && dout && dq && dk && dv && dz && s_out, "aitb_attn_bwd: bad arguments");
  AITB_REQUIRE(q_rep >= 1 && (mask_mode == 0 || mask_mode == 1) && n_keys >= 1 && n_keys <= kBT, "aitb_attn_bwd: bad mode");
  AITB_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0 && lddq % 2 == 0 && lddkv % 2 == 0, "aitb_attn_bwd: bad leading dimensions");
  static SmemAttrOnce once;
  if (ensure_dyn_smem((const void*)attn_bwd_kernel<T>, kAttnBwdSmem, once, "attn_bwd_kernel")) return 1;
  attn_bwd_kernel<T><<<G, kAttnBwdThreads, kAttnBwdSmem, stream>>>(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, dout, mask_mode,
                                                                  n_keys, dq, lddq, dk, dv, lddkv, dz, s_out, dc);
  return check_launch("attn_bwd_kernel");
}
int attn_bwd_run(const float* q, int ldq, int q_rep, const float* k, const float* v, int ldkv, const float* w_sk,
                 const float* b_sk, const float* dout, int G, int mask_mode, int n_keys, float* dq, int lddq, float* dk,
                 float* dv, int lddkv, float* dz, float* s_out, cudaStream_t stream, const DropCfg* drop) {
  return attn_bwd_run_t<float>(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, dout, G, mask_mode, n_keys, dq, lddq, dk, dv, lddkv, dz,
                               s_out, stream, drop);
}
int attn_bwd_run(const __nv_bfloat16* q, int ldq, int q_rep, const __nv_bfloat16* k, const __nv_bfloat16* v, int ldkv,
                 const float* w_sk, const float* b_sk, const __nv_bfloat16* dout, int G, int mask_mode, int n_keys,
                 __nv_bfloat16* dq, int lddq, __nv_bfloat16* dk, __nv_bfloat16* dv, int lddkv, __nv_bfloat16* dz,
                 __nv_bfloat16* s_out, cudaStream_t stream, const DropCfg* drop) {
  return attn_bwd_run_t<__nv_bfloat16>(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, dout, G, mask_mode, n_keys, dq, lddq, dk, dv,
                                       lddkv, dz, s_out, stream, drop);
}

static int sms_bwd() { return current_sm_count(); }


// ---------------------------------------------------------------------------------------------
// LayerNorm backward from the saved output:  y = x_hat * gamma + beta,  x_hat = (x - mean) * rstd
//   a = g * gamma;  dx = rstd * (a - mean(a) - x_hat * mean(a * x_hat));  dgamma += g * x_hat;  dbeta += g
// x_hat is recovered as (y - beta) / gamma (the forward keeps y and 1/sigma only; a channel whose gamma is
// exactly 0 contributes x_hat = 0).  One warp per row (512 channels, 16 per lane), persistent over rows so the
// parameter gradients are reduced in registers and leave the CTA as 2 x 512 atomics.
// Row compaction: rows are grouped by `grp` (64 tokens); only the first `valid` rows of a group store dx, at
// row (group * valid + t) -- the encoder's 64 -> 49 un-padding (system/Models.py:268-270); the pad rows still
// feed dgamma / dbeta (their x_hat is LayerNorm(pos_table[t])).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ g, const T* __restrict__ y, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ rstd, int rows, int grp, int valid,
              T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8][512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[16], bt[16], ig[16], dg_acc[16], db_acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = (i >> 2) * 128 + lane * 4 + (i & 3);
    gm[i] = gamma[c];
    bt[i] = beta[c];
    ig[i] = gm[i] != 0.f ? 1.f / gm[i] : 0.f;
    dg_acc[i] = db_acc[i] = 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float gv[16], xh[16];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 g4 = ld4(g + (size_t)row * 512 + j * 128 + lane * 4);
      const float4 y4 = ld4(y + (size_t)row * 512 + j * 128 + lane * 4);
      gv[4 * j] = g4.x; gv[4 * j + 1] = g4.y; gv[4 * j + 2] = g4.z; gv[4 * j + 3] = g4.w;
      xh[4 * j] = y4.x; xh[4 * j + 1] = y4.y; xh[4 * j + 2] = y4.z; xh[4 * j + 3] = y4.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      xh[i] = (xh[i] - bt[i]) * ig[i];
      dg_acc[i] += gv[i] * xh[i];
      db_acc[i] += gv[i];
      gv[i] *= gm[i];           // a
      m1 += gv[i];
      m2 += gv[i] * xh[i];
    }
    m1 = warp_sum(m1) * (1.f / 512.f);
    m2 = warp_sum(m2) * (1.f / 512.f);
    const int t = row % grp;
    if (t < valid) {
      const float rs = rstd[row];
      T* o = dx + ((size_t)(row / grp) * valid + t) * 512;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 r;
        r.x = rs * (gv[4 * j] - m1 - xh[4 * j] * m2);          // stored rounded (RN): dx feeds tensor-core MMAs next
        r.y = rs * (gv[4 * j + 1] - m1 - xh[4 * j + 1] * m2);
        r.z = rs * (gv[4 * j + 2] - m1 - xh[4 * j + 2] * m2);
        r.w = rs * (gv[4 * j + 3] - m1 - xh[4 * j + 3] * m2);
        st4r(o + j * 128 + lane * 4, r);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = (i >> 2) * 128 + lane * 4 + (i & 3);
    red[0][warp][c] = dg_acc[i];
    red[1][warp][c] = db_acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 512; c += 256) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

template <typename T>
static int ln_bwd_run_t(const T* g, const T* y, const float* gamma, const float* beta, const float* rstd, int rows,
                        int grp, int valid, T* dx, float* dgamma, float* dbeta, cudaStream_t stream) {
  AITB_REQUIRE(g && y && gamma && beta && rstd && dx && dgamma && dbeta, "aitb_ln_bwd: null pointer");
  AITB_REQUIRE(rows > 0 && grp > 0 && valid > 0 && valid <= grp && rows % grp == 0, "aitb_ln_bwd: bad row grouping");
  int grid = (rows + 7) / 8;
  if (grid > 4 * sms_bwd()) grid = 4 * sms_bwd();
  ln_bwd_kernel<T><<<grid, 256, 0, stream>>>(g, y, gamma, beta, rstd, rows, grp, valid, dx, dgamma, dbeta);
  return check_launch("ln_bwd_kernel");
}
int ln_bwd_run(const float* g, const float* y, const float* gamma, const float* beta, const float* rstd, int rows,
               int grp, int valid, float* dx, float* dgamma, float* dbeta, cudaStream_t stream) {
  return ln_bwd_run_t<float>(g, y, gamma, beta, rstd, rows, grp, valid, dx, dgamma, dbeta, stream);
}
int ln_bwd_run(const __nv_bfloat16* g, const __nv_bfloat16* y, const float* gamma, const float* beta, const float* rstd,
               int rows, int grp, int valid, __nv_bfloat16* dx, float* dgamma, float* dbeta, cudaStream_t stream) {
  return ln_bwd_run_t<__nv_bfloat16>(g, y, gamma, beta, rstd, rows, grp, valid, dx, dgamma, dbeta, stream);
}

// out[c] += sum over rows of x[row, c]   (bias gradients), x row-major [rows, ld], c < cols (cols % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int ld, int rows, int cols, int rows_per_cta, float* __restrict__ out) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; ++r) {
    const float4 v = ld4(x + (size_t)r * ld + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  atomicAdd(out + c, acc.x);
  atomicAdd(out + c + 1, acc.y);
  atomicAdd(out + c + 2, acc.z);
  atomicAdd(out + c + 3, acc.w);
}

template <typename T>
static int colsum_run_t(const T* x, int ld, int rows, int cols, float* out, cudaStream_t stream) {
  AITB_REQUIRE(x && out && rows > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0, "aitb_colsum: bad arguments");
  const int gx = (cols / 4 + 255) / 256;
  int gy = (4 * sms_bwd() + gx - 1) / gx;
  if (gy > (rows + 31) / 32) gy = (rows + 31) / 32;
  if (gy < 1) gy = 1;
  const int rpc = (rows + gy - 1) / gy;
  gy = (rows + rpc - 1) / rpc;
  colsum_kernel<T><<<dim3(gx, gy), 256, 0, stream>>>(x, ld, rows, cols, rpc, out);
  return check_launch("colsum_kernel");
}
int colsum_run(const float* x, int ld, int rows, int cols, float* out, cudaStream_t stream) {
  return colsum_run_t<float>(x, ld, rows, cols, out, stream);
}
int colsum_run(const __nv_bfloat16* x, int ld, int rows, int cols, float* out, cudaStream_t stream) {
  return colsum_run_t<__nv_bfloat16>(x, ld, rows, cols, out, stream);
}

// out[b, l] = sum_p x[b, p, l]   (gradient of the unit -> proposal broadcast), l < L (L % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256)
bsum_kernel(const T* __restrict__ x, int P, int L, T* __restrict__ out) {
  const int l = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (l >= L) return;
  const T* xb = x + (size_t)blockIdx.y * P * L + l;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < P; ++p) {
    const float4 v = ld4(xb + (size_t)p * L);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(out + (size_t)blockIdx.y * L + l) = acc;   // fp32: unrounded, as before
  else st4r(out + (size_t)blockIdx.y * L + l, acc);
}

template <typename T>
static int bsum_run_t(const T* x, int B, int P, int L, T* out, cudaStream_t stream) {
  AITB_REQUIRE(x && out && B > 0 && P > 0 && L > 0 && L % 4 == 0, "aitb_bsum: bad arguments");
  bsum_kernel<T><<<dim3((L / 4 + 255) / 256, B), 256, 0, stream>>>(x, P, L, out);
  return check_launch("bsum_kernel");
}
int bsum_run(const float* x, int B, int P, int L, float* out, cudaStream_t stream) { return bsum_run_t<float>(x, B, P, L, out, stream); }
int bsum_run(const __nv_bfloat16* x, int B, int P, int L, __nv_bfloat16* out, cudaStream_t stream) {
  return bsum_run_t<__nv_bfloat16>(x, B, P, L, out, stream);
}

// dw [N, ldw] += dy[M, ldy (cols n_off .. n_off + N)]^T * x[M, ldx (cols 0 .. K)]
// shared tail of the two entry points: tiling, split over row chunks, launch
template <bool BF16>
static int wgrad_launch(const void* dy, int ldy, const CUtensorMap& tmX, int M, int N, int K, int bn, float* dw, int ldw,
                        int conv_S, int conv_cg, int x_group_stride, cudaStream_t stream) {
  constexpr int kWgRows = WgCfg<BF16>::kRows, kGW = WgCfg<BF16>::kGW;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.bn = bn;
  p.n_tiles = N / 128;
  p.k_tiles = K / bn;
  p.conv_S = conv_S;
  p.conv_cg = conv_cg;
  p.x_group_stride = x_group_stride;
  const int tiles = p.n_tiles * p.k_tiles;
  const int chunks = (M + kWgRows - 1) / kWgRows;
  int splits = (2 * sms_bwd() + tiles - 1) / tiles;           // ~2 waves of CTAs
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  p.rows_per_split = ((chunks + splits - 1) / splits) * kWgRows;
  p.splits = (M + p.rows_per_split - 1) / p.rows_per_split;
  p.dw = dw;
  p.ldw = ldw;
  CUtensorMap tmY;
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldy * WgCfg<BF16>::kEs};
    const uint32_t box[2] = {(uint32_t)kGW, (uint32_t)kWgRows};
    if (BF16 ? encode_map_bf16(&tmY, dy, dims, str, box, "wgrad dY") : encode_map_f32_mn(&tmY, dy, 2, dims, str, box, "wgrad dY"))
      return 1;
  }
  const int smem = kWgStages * (kWgABytes + (bn / kGW) * kWgRows * 128) + 1024 + 256;   // bn * 128 bytes of X per stage either way
  static SmemAttrOnce once;
  if (ensure_dyn_smem((const void*)wgrad_tcgen05_kernel<BF16>, kWgStages * (kWgABytes + 256 * 128) + 1024 + 256, once,
                      "wgrad_tcgen05_kernel"))
    return 1;
  wgrad_tcgen05_kernel<BF16><<<tiles * p.splits, kWgThreads, smem, stream>>>(tmY, tmX, p);
  return check_launch("wgrad_tcgen05_kernel");
}

int wgrad_run(const float* dy, int ldy, const float* x, int ldx, int M, int N, int K, float* dw, int ldw,
              cudaStream_t stream) {
  AITB_REQUIRE(dy && x && dw, "aitb_wgrad: null pointer");
  AITB_REQUIRE(M > 0 && N > 0 && K > 0, "aitb_wgrad: empty problem");
  AITB_REQUIRE(N % 128 == 0, "aitb_wgrad: N=%d must be a multiple of 128", N);
  AITB_REQUIRE(K % 64 == 0, "aitb_wgrad: K=%d must be a multiple of 64", K);
  AITB_REQUIRE(ldy % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, "aitb_wgrad: leading dimensions must be multiples of 4");
  AITB_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dw & 15) == 0,
               "aitb_wgrad: pointers must be 16-byte aligned");
  const int bn = K % 256 == 0 ? 256 : (K % 128 == 0 ? 128 : 64);
  CUtensorMap tmX;
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  const uint64_t str[1] = {(uint64_t)ldx * 4};
  const uint32_t box[2] = {32, (uint32_t)kWgRows};
  if (encode_map_f32_mn(&tmX, x, 2, dims, str, box, "wgrad X")) return 1;
  return wgrad_launch<false>(dy, ldy, tmX, M, N, K, bn, dw, ldw, 0, 0, 0, stream);
}

// bf16 storage: dY and X are bf16 row-major activations, dW stays fp32 (accumulated)
int wgrad_run(const __nv_bfloat16* dy, int ldy, const __nv_bfloat16* x, int ldx, int M, int N, int K, float* dw, int ldw,
              cudaStream_t stream) {
  AITB_REQUIRE(dy && x && dw, "aitb_wgrad: null pointer");
  AITB_REQUIRE(M > 0 && N > 0 && K > 0, "aitb_wgrad: empty problem");
  AITB_REQUIRE(N % 128 == 0, "aitb_wgrad: N=%d must be a multiple of 128", N);
  AITB_REQUIRE(K % 64 == 0, "aitb_wgrad: K=%d must be a multiple of 64", K);
  AITB_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0 && ldw % 4 == 0, "aitb_wgrad: bf16 leading dimensions must be multiples of 8 (dW: 4)");
  AITB_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dw & 15) == 0,
               "aitb_wgrad: pointers must be 16-byte aligned");
  const int bn = K % 256 == 0 ? 256 : (K % 128 == 0 ? 128 : 64);
  CUtensorMap tmX;
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  const uint64_t str[1] = {(uint64_t)ldx * 2};
  const uint32_t box[2] = {64, (uint32_t)WgCfg<true>::kRows};
  if (encode_map_bf16(&tmX, x, dims, str, box, "wgrad X")) return 1;
  return wgrad_launch<true>(dy, ldy, tmX, M, N, K, bn, dw, ldw, 0, 0, 0, stream);
}

// Weight gradient of a (grouped) 1x1 or 3x3 stride-1 "same" convolution on a channels-last S x S map, no im2col:
//   dW[n, tap * Cg + c] += sum over (g, y, x) of dY[(g, y, x), n] * X[g, y + ky - 1, x + kx - 1, group(n) * Cg + c]
// (tap = ky * 3 + kx; Cg = C / groups; taps = 1: the 1x1 case, dW [N, Cg]).  groups > 1 needs N / groups == 128 (one
// n-tile per group).  dY is [G * S * S, ldy] row-major; X [G, S, S, C]; dW [N, taps * Cg] with leading dimension ldw.
int wgrad_conv_run(const float* dy, int ldy, const float* x, int G, int S, int C, int N, int groups, int taps, float* dw,
                   int ldw, cudaStream_t stream) {
  AITB_REQUIRE(dy && x && dw, "aitb_wgrad_conv: null pointer");
  AITB_REQUIRE(G > 0 && C > 0 && N > 0 && groups > 0 && C % groups == 0, "aitb_wgrad_conv: bad sizes");
  AITB_REQUIRE(taps == 1 || taps == 9, "aitb_wgrad_conv: taps must be 1 (1x1) or 9 (3x3)");
  AITB_REQUIRE(S == 4 || S == 8, "aitb_wgrad_conv: S=%d (4x4 and 8x8 maps: a 32-row stage must be whole map rows)", S);
  AITB_REQUIRE(N % 128 == 0 && (groups == 1 || N / groups == 128),
               "aitb_wgrad_conv: N=%d must be a multiple of 128, and 128 per group for a grouped convolution", N);
  const int cg = C / groups;
  AITB_REQUIRE(cg % 64 == 0, "aitb_wgrad_conv: %d channels per group must be a multiple of 64", cg);
  AITB_REQUIRE(ldy % 4 == 0 && ldw % 4 == 0, "aitb_wgrad_conv: leading dimensions must be multiples of 4");
  AITB_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dw & 15) == 0,
               "aitb_wgrad_conv: pointers must be 16-byte aligned");
  AITB_REQUIRE((long long)G * S * S < (1ll << 31), "aitb_wgrad_conv: too many rows");
  const int M = G * S * S, K = taps * cg;
  const int bn = cg % 256 == 0 ? 256 : (cg % 128 == 0 ? 128 : 64);   // a k-tile never straddles two taps
  CUtensorMap tmX;
  if (taps == 1) {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {32, (uint32_t)kWgRows};
    if (encode_map_f32_mn(&tmX, x, 2, dims, str, box, "wgrad_conv X")) return 1;
  } else {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)S, (uint64_t)S, (uint64_t)G};
    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)S * C * 4, (uint64_t)S * S * C * 4};
    const uint32_t box[4] = {32, (uint32_t)S, (uint32_t)(S == 8 ? 4 : 4), (uint32_t)(S == 8 ? 1 : 2)};   // 32 rows per stage
    if (encode_map_f32_mn(&tmX, x, 4, dims, str, box, "wgrad_conv X map")) return 1;
  }
  return wgrad_launch<false>(dy, ldy, tmX, M, N, K, bn, dw, ldw, taps == 9 ? S : 0, cg, groups > 1 ? cg : 0, stream);
}

}  // namespace aitb
