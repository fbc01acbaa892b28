// tcgen05 attention core for the bf16 and split ("fp32") configurations: the batched proposal-query score
// products Q K^T and P V on the 5th-generation tensor cores with TMEM accumulators, operands staged by TMA.
//
// Same math as attn.cu (system/Modules.py:16-29, SubLayers.py:22-39,89-92):
//   P_h = softmax(mask(Q_h K_h^T / 8)),  O_h = P_h V_h,  s = mean_T(sum_h O_h),  g = softmax_h(W_sk s + b_sk),
//   out = sum_h O_h * g_h
//
// A CTA works on a COUPLE of proposal-query pairs (2 x 64 query rows = one M = 128 MMA):
//   MMA 1  S[128, 128] = [Q0; Q1] . [K0; K1]^T     one N = 128 instruction group; thread row r reads only its
//                                                   own pair's 64 columns (the off-diagonal blocks are unused)
//   MMA 2  O[:, 0:64]  = [P0; P1] . V0,  O[:, 64:128] = [P0; P1] . V1     the P tile is dense (no zero blocks);
//                                                   row r < 64 keeps columns 0..63, row r >= 64 columns 64..127
//   V is consumed as an MN-major B operand straight from its [keys, d] layout (no transposed copy).
// Split configuration: every operand is two bf16 planes and every product is three MMAs (hi*hi + hi*lo + lo*hi).
// Accumulators live in TMEM (S double-buffered: the next head's Q K^T runs under the current head's softmax); one
// thread owns one query row (tcgen05.ld 32x32b), so softmax needs no shuffles.  The selective-head gate needs
// mean_T sum_h O_h before any head can be weighted and 8 x 64 fp32 per row do not fit in registers, so the heads
// are walked twice: pass A accumulates sum_h O_h per row (one column reduction per couple), pass B re-runs the
// products and accumulates O_h * g_h.  Roles: warps 0-3 softmax / epilogue, warp 4 TMA producer, warp 5 MMA issuer.
//
// STATUS: bit-for-bit within the same tolerances as attn.cu (tests/test_gpu_gemm_attn.py, test_gpu_split.py run both),
// but MEASURED SLOWER than the mma.sync one-pass kernel on the benchmark shape (split: 770 vs 376 us per 2400 pairs;
// bf16: +0.5 ms per step): the gate forces the second pass, and with one query row per thread the softmax is a
// serial 64-element chain on only four warps.  An L2 prefetch of the next tiles changed nothing (not load-bound).
// It is therefore OPT-IN (AITB_ATTN_TC=1); the default stays attn.cu, whose fragment layout keeps all eight O_h
// tiles in registers across eight warps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

int encode_map_bf16(CUtensorMap* tm, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* what);

static constexpr int kTcThreads = 192;
static constexpr int kTile = 128 * 128;      // one operand plane of a couple: 128 rows x 128 B
static constexpr int kVTile = 2 * 64 * 128;  // V plane: two pairs x 64 keys x 128 B

template <bool SPLIT>
struct TcCfg {
  static constexpr int kP = SPLIT ? 2 : 1;
  static constexpr int kQkStage = 2 * kP * kTile;          // Q planes then K planes
  static constexpr int kOffV = 2 * kQkStage;               // after the two Q/K stages
  static constexpr int kVBytes = kP * kVTile;
  static constexpr int kOffP = kOffV + kVBytes;
  static constexpr int kPBytes = kP * kTile;
  static constexpr int kOffMisc = kOffP + kPBytes;
  // misc: barriers (256 B) + column partials [4][64] + s [2][64] + gate [2][8][64] floats
  static constexpr int kSmem = kOffMisc + 256 + (4 * 64 + 2 * 64 + 2 * 8 * 64) * 4 + 1024 /*alignment*/;
};

struct TcParams {
  int G;            // pairs
  int q_rep;        // pairs sharing one Q block
  int mask_mode, n_keys;
  int q_lo, kv_lo;  // split: element offset of the lo plane inside a row (0 otherwise)
  int kv_rows;      // rows per pair in the K / V buffers (64, or 49: compact encoder rows)
  const float* w_sk;
  const float* b_sk;
  __nv_bfloat16* out;
};

// MN-major bf16 B operand, SWIZZLE_128B: rows of 128 contiguous bytes along N (64 bf16), 8 contraction rows per
// 1024-byte atom ((8,n),(8,k)):((1,LBO),(8,SBO)); a single N atom here, so LBO is unused.
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void named_bar_128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const TcParams p) {
  using Cfg = TcCfg<SPLIT>;
  constexpr int kP = Cfg::kP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffMisc);
  uint64_t* qk_full = bars;        // [2]
  uint64_t* qk_empty = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;     // [2]
  uint64_t* s_empty = bars + 8;    // [2]
  uint64_t* p_full = bars + 10;
  uint64_t* p_empty = bars + 11;
  uint64_t* o_full = bars + 12;
  uint64_t* o_empty = bars + 13;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);
  float* part = reinterpret_cast<float*>(smem + Cfg::kOffMisc + 256);  // [4 warps][64]
  float* svec = part + 4 * 64;                                          // [2 pairs][64]
  float* gate = svec + 2 * 64;                                          // [2][8][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256;   // S buffers at columns 0 / 128, O at 256

  const int n_couples = (p.G + 1) >> 1;
  const int my_couples = ((int)blockIdx.x < n_couples) ? (n_couples - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_iters = my_couples * 16;   // (pass, head) steps of this CTA

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int n = 0; n < n_iters; ++n) {
        const int couple = (int)blockIdx.x + (n >> 4) * (int)gridDim.x;
        const int h = n & 7;
        const int pair0 = 2 * couple, pair1 = 2 * couple + 1;
        const int qrow0 = (pair0 / p.q_rep) * 64, qrow1 = (pair1 / p.q_rep) * 64;
        const uint32_t buf = n & 1, ph = (n >> 1) & 1;
        mbar_wait(&qk_empty[buf], ph ^ 1);
        uint8_t* sq = smem + buf * Cfg::kQkStage;
        uint8_t* sk = sq + kP * kTile;
        mbar_arrive_expect_tx(&qk_full[buf], Cfg::kQkStage);
#pragma unroll
        for (int pl = 0; pl < kP; ++pl) {
          const int qc = h * 64 + pl * p.q_lo, kc = h * 64 + pl * p.kv_lo;
          tma_load_2d(sq + pl * kTile, &tmQ, &qk_full[buf], qc, qrow0);
          tma_load_2d(sq + pl * kTile + 64 * 128, &tmQ, &qk_full[buf], qc, qrow1);
          tma_load_2d(sk + pl * kTile, &tmK, &qk_full[buf], kc, pair0 * p.kv_rows);
          tma_load_2d(sk + pl * kTile + 64 * 128, &tmK, &qk_full[buf], kc, pair1 * p.kv_rows);
        }
        mbar_wait(v_empty, (n & 1) ^ 1);
        uint8_t* sv = smem + Cfg::kOffV;
        mbar_arrive_expect_tx(v_full, Cfg::kVBytes);
#pragma unroll
        for (int pl = 0; pl < kP; ++pl) {
          const int vc = h * 64 + pl * p.kv_lo;
          tma_load_2d(sv + pl * kVTile, &tmV, v_full, vc, pair0 * p.kv_rows);
          tma_load_2d(sv + pl * kVTile + 64 * 128, &tmV, v_full, vc, pair1 * p.kv_rows);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_iters > 0) {
      constexpr uint32_t idesc1 = make_idesc(1u, 128u, 128u);                  // bf16, K-major A and B
      constexpr uint32_t idesc2 = make_idesc(1u, 128u, 64u) | (1u << 16);      // B (= V) MN-major
      auto issue_mma1 = [&](int n) {
        const uint32_t buf = n & 1, ph = (n >> 1) & 1;
        mbar_wait(&s_empty[buf], ph ^ 1);
        mbar_wait(&qk_full[buf], ph);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + buf * Cfg::kQkStage);
        const uint32_t sk = sq + kP * kTile;
        const uint64_t qh = make_sw128_kmajor_desc(sq), kh = make_sw128_kmajor_desc(sk);
        const uint32_t d = tS0 + buf * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_ss<2>(d, qh + (uint64_t)(k * 2), kh + (uint64_t)(k * 2), idesc1, k != 0 ? 1u : 0u);
          if constexpr (SPLIT) {
            umma_ss<2>(d, qh + (uint64_t)(k * 2), kh + (uint64_t)(kTile / 16 + k * 2), idesc1, 1u);
            umma_ss<2>(d, qh + (uint64_t)(kTile / 16 + k * 2), kh + (uint64_t)(k * 2), idesc1, 1u);
          }
        }
        tc_commit(&qk_empty[buf]);
        tc_commit(&s_full[buf]);
      };
      issue_mma1(0);
      for (int n = 0; n < n_iters; ++n) {
        if (n + 1 < n_iters) issue_mma1(n + 1);
        mbar_wait(o_empty, (n & 1) ^ 1);
        mbar_wait(v_full, n & 1);
        mbar_wait(p_full, n & 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + Cfg::kOffP), sv = smem_u32(smem + Cfg::kOffV);
        const uint64_t phd = make_sw128_kmajor_desc(sp);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint64_t vh = make_mn_sw128_desc(sv + blk * (64 * 128));
          const uint32_t d = tO + blk * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 16 keys per MMA: 32 bytes along P's rows, two 8-key atoms (2048 B) of V
            umma_ss<2>(d, phd + (uint64_t)(k * 2), vh + (uint64_t)(k * 128), idesc2, k != 0 ? 1u : 0u);
            if constexpr (SPLIT) {
              umma_ss<2>(d, phd + (uint64_t)(k * 2), vh + (uint64_t)(kVTile / 16 + k * 128), idesc2, 1u);
              umma_ss<2>(d, phd + (uint64_t)(kTile / 16 + k * 2), vh + (uint64_t)(k * 128), idesc2, 1u);
            }
          }
        }
        tc_commit(v_empty);
        tc_commit(p_empty);
        tc_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue: thread = query row
    const int r = threadIdx.x;            // 0..127
    const int blk = r >> 6, t = r & 63;   // pair of the couple, token
    const uint32_t lane_off = ((uint32_t)(warp * 32) << 16);
    uint8_t* sp = smem + Cfg::kOffP;
    float acc[64];
    for (int n = 0; n < n_iters; ++n) {
      const int couple = (int)blockIdx.x + (n >> 4) * (int)gridDim.x;
      const int pass = (n >> 3) & 1, h = n & 7;
      const uint32_t buf = n & 1, ph = (n >> 1) & 1;
      if ((n & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      }
      // ---- S row -> P row
      mbar_wait(&s_full[buf], ph);
      tc_fence_after();
      uint32_t raw[64];
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[32]);
        tmem_ld32(tS0 + buf * 128 + lane_off + blk * 64, lo);
        tmem_ld32(tS0 + buf * 128 + lane_off + blk * 64 + 32, hi);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const bool masked = p.mask_mode == 0 ? (j >= p.n_keys) : (j > t);
        const float x = masked ? -1e9f : __uint_as_float(raw[j]) * 0.125f;   // masked_fill(mask == 0, -1e9)
        raw[j] = __float_as_uint(x);
        mx = fmaxf(mx, x);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float e = expf(__uint_as_float(raw[j]) - mx);
        raw[j] = __float_as_uint(e);
        sum += e;
      }
      const float inv = 1.f / sum;
      mbar_wait(p_empty, (n & 1) ^ 1);    // the previous head's P V has finished reading the P tile
#pragma unroll
      for (int c = 0; c < 8; ++c) {       // 8 keys (16 bytes) per chunk, 128-byte swizzle: chunk ^ (row & 7)
        uint4 hi4, lo4;
        uint32_t* hp = reinterpret_cast<uint32_t*>(&hi4);
        uint32_t* lp = reinterpret_cast<uint32_t*>(&lo4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = __uint_as_float(raw[8 * c + 2 * e]) * inv, b2 = __uint_as_float(raw[8 * c + 2 * e + 1]) * inv;
          const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b2);
          hp[e] = *reinterpret_cast<const uint32_t*>(&hb);
          if constexpr (SPLIT) {
            const float2 hf = __bfloat1622float2(hb);
            const __nv_bfloat162 lb = __floats2bfloat162_rn(a - hf.x, b2 - hf.y);
            lp[e] = *reinterpret_cast<const uint32_t*>(&lb);
          }
        }
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sp + off) = hi4;
        if constexpr (SPLIT) *reinterpret_cast<uint4*>(sp + kTile + off) = lo4;
      }
      fence_proxy_async();                // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(p_full);
      // ---- O row
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[32]);
        tmem_ld32(tO + lane_off + blk * 64, lo);
        tmem_ld32(tO + lane_off + blk * 64 + 32, hi);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      if (pass == 0) {
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] += __uint_as_float(raw[j]);
        if (h == 7) {
          // s = mean over the 64 rows of each pair of sum_h O_h: one column reduction per couple
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            float v = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == (j & 31)) part[warp * 64 + j] = v;
            acc[j] = 0.f;
          }
          named_bar_128(1);
          svec[blk * 64 + t] = (part[(2 * blk) * 64 + t] + part[(2 * blk + 1) * 64 + t]) * (1.f / 64.f);
          named_bar_128(1);
          // gate logits z = W_sk s + b_sk: 512 per pair, 8 per thread (o = hh * 64 + t)
          float z[8];
#pragma unroll
          for (int hh = 0; hh < 8; ++hh) {
            const int o = hh * 64 + t;
            const float* wr = p.w_sk + (size_t)o * 64;
            float a = __ldg(p.b_sk + o);
#pragma unroll 8
            for (int c = 0; c < 64; c += 4) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + c));
              a += w4.x * svec[blk * 64 + c] + w4.y * svec[blk * 64 + c + 1] + w4.z * svec[blk * 64 + c + 2] +
                   w4.w * svec[blk * 64 + c + 3];
            }
            z[hh] = a;
          }
          float m = z[0];
#pragma unroll
          for (int hh = 1; hh < 8; ++hh) m = fmaxf(m, z[hh]);
          float zs = 0.f;
#pragma unroll
          for (int hh = 0; hh < 8; ++hh) { z[hh] = expf(z[hh] - m); zs += z[hh]; }
          const float zi = 1.f / zs;
#pragma unroll
          for (int hh = 0; hh < 8; ++hh) gate[(blk * 8 + hh) * 64 + t] = z[hh] * zi;
          named_bar_128(1);
        }
      } else {
        const float* gp = gate + (blk * 8 + h) * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gp + j);
          acc[j] += __uint_as_float(raw[j]) * g4.x;
          acc[j + 1] += __uint_as_float(raw[j + 1]) * g4.y;
          acc[j + 2] += __uint_as_float(raw[j + 2]) * g4.z;
          acc[j + 3] += __uint_as_float(raw[j + 3]) * g4.w;
        }
        if (h == 7) {
          const int pair = 2 * couple + blk;
          if (pair < p.G) {
            constexpr int kOut = SPLIT ? 128 : 64;
            __nv_bfloat16* og = p.out + ((size_t)pair * 64 + t) * kOut;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint4 hi4, lo4;
              uint32_t* hp = reinterpret_cast<uint32_t*>(&hi4);
              uint32_t* lp = reinterpret_cast<uint32_t*>(&lo4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = acc[8 * c + 2 * e], b2 = acc[8 * c + 2 * e + 1];
                const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b2);
                hp[e] = *reinterpret_cast<const uint32_t*>(&hb);
                if constexpr (SPLIT) {
                  const float2 hf = __bfloat1622float2(hb);
                  const __nv_bfloat162 lb = __floats2bfloat162_rn(a - hf.x, b2 - hf.y);
                  lp[e] = *reinterpret_cast<const uint32_t*>(&lb);
                }
              }
              *reinterpret_cast<uint4*>(og + 8 * c) = hi4;
              if constexpr (SPLIT) *reinterpret_cast<uint4*>(og + 64 + 8 * c) = lo4;
            }
          }
          named_bar_128(1);   // gate / svec of this couple are dead before the next couple rewrites them
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// q / k / v: bf16 matrices (split: two planes per row, the lo plane `ld` LOGICAL elements further); ldq / ldkv are
// LOGICAL row widths like in attn_core_run.  Returns -1 when this kernel does not apply (caller falls back).
int attn_tc_run(const void* q, int ldq, int q_rep, const void* k, const void* v, int ldkv, const float* w_sk,
                const float* b_sk, int G, int mask_mode, int n_keys, int dtype, void* out, cudaStream_t stream,
                int kv_rows) {
  const bool split = dtype == AITB_F32S;
  if (dtype != AITB_BF16 && !split) return -1;
  if (ldq % 8 != 0 || ldkv % 8 != 0) return -1;
  if ((((uintptr_t)q) | ((uintptr_t)k) | ((uintptr_t)v)) & 15) return -1;
  const int pl = split ? 2 : 1;
  const int q_units = (G + q_rep - 1) / q_rep;
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[2] = {64, 64};
  {
    const uint64_t dims[2] = {(uint64_t)64 * 8 * pl + (split ? (uint64_t)(ldq - 512) : 0), (uint64_t)q_units * 64};
    const uint64_t str[1] = {(uint64_t)ldq * pl * 2};
    if (encode_map_bf16(&tmQ, q, dims, str, box, "attention Q")) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)64 * 8 * pl + (split ? (uint64_t)(ldkv - 512) : 0), (uint64_t)G * kv_rows};
    const uint64_t str[1] = {(uint64_t)ldkv * pl * 2};
    if (encode_map_bf16(&tmK, k, dims, str, box, "attention K")) return 1;
    if (encode_map_bf16(&tmV, v, dims, str, box, "attention V")) return 1;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.G = G;
  p.q_rep = q_rep;
  p.mask_mode = mask_mode;
  p.n_keys = n_keys;
  p.q_lo = split ? ldq : 0;
  p.kv_lo = split ? ldkv : 0;
  p.kv_rows = kv_rows;
  p.w_sk = w_sk;
  p.b_sk = b_sk;
  p.out = (__nv_bfloat16*)out;
  const int couples = (G + 1) / 2;
  const int sms = current_sm_count();
  const int grid = couples < sms ? couples : sms;
  if (split) {
    static SmemAttrOnce once;
    if (ensure_dyn_smem((const void*)attn_tc_kernel<true>, TcCfg<true>::kSmem, once, "attn_tc_kernel<split>")) return 1;
    attn_tc_kernel<true><<<grid, kTcThreads, TcCfg<true>::kSmem, stream>>>(tmQ, tmK, tmV, p);
  } else {
    static SmemAttrOnce once;
    if (ensure_dyn_smem((const void*)attn_tc_kernel<false>, TcCfg<false>::kSmem, once, "attn_tc_kernel<bf16>")) return 1;
    attn_tc_kernel<false><<<grid, kTcThreads, TcCfg<false>::kSmem, stream>>>(tmQ, tmK, tmV, p);
  }
  return check_launch("attn_tc_kernel");
}

}  // namespace aitb
