// tcgen05 GEMM for the AIT head:  out[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
//   * A and W tiles arrive by TMA (128-byte swizzle) into a shared-memory mbarrier ring.
//     A is addressed through a 4-D strided view, so a 3x3 convolution on an 8x8 / 4x4 map is a
//     loop over 9 shifted boxes (out-of-bounds rows are zero-filled by TMA = the conv padding)
//     and a stride-2 1x1 convolution is just a strided view.  No im2col buffer exists.
//   * one elected thread issues tcgen05.mma (K = 32 bytes per instruction: 8 x tf32 or 16 x bf16);
//     accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the main loop of
//     tile i+1; persistent CTAs walk the tile list.
//   * the epilogue warps read TMEM with tcgen05.ld (one output row per thread) and apply the fused
//     epilogue: bias, ReLU / ReLU^2, dual-accumulator sum, residual, positional table, LayerNorm.
//
// Three kernels share that machinery (and `epilogue_tile`):
//   gemm2_tcgen05_kernel          2-CTA pairs (cta_group::2): 256x256 tile, each CTA stages its 128 rows of A
//                                 and HALF of the W tile, 6-stage ring, two epilogue warpgroups (setmaxnreg).
//                                 All wide plain / conv GEMMs (QKV, FFN w_1, K/V, dec_trans, layer4).
//   gemm_tcgen05_kernel<.., CL>   "cluster LayerNorm": the N = 512 row is split over a 2-CTA cluster
//                                 (2 x 256 columns); partial row statistics travel through DSMEM.
//                                 enc_emb / dec_emb, attention fc, FFN w_2.
//   gemm_tcgen05_kernel<.., !CL>  single-CTA (cta_group::1, M = 128): grouped SKBlock convolutions with the
//                                 dual accumulator (BLOCK_N = 128) and the tiny query-branch GEMMs.
//
// Reference ops this replaces (all plain torch calls into cuBLAS/cuDNN in the reference):
//   nn.Conv2d 1x1 (system/Models.py:188-209), nn.Linear in MultiHeadAttention / FFN
//   (system/SubLayers.py:51-58,172-187), grouped Conv2d in SKBlock
//   (modules/blocks_coatt_transformer_sk.py:929-935), ResNet layer4 convs + frozen BN
//   (faster_rcnn/resnet_coatt_transformer_sk.py:73-109).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kBlockM = 128;
static constexpr int kABytes = kBlockM * 128;  // one A stage: 128 rows x 128 B
static constexpr int kThreads = 192;           // warp0 TMA, warp1 MMA, warps 2..5 epilogue

struct GemmKParams {
  int M, N;
  int k_chunks;  // 128-byte K chunks per tap
  int taps;
  int a_m_dim, a_m_step, a_group_c;
  int simple_rows_ok;   // host: AITB_NO_SIMPLE_ROWS unset (A/B of the strided-row epilogue addressing)
  int bias_smem;   // 2-byte-output plain epilogue: per-warp shared-memory copy of the tile's bias slice (A/B: AITB_BIAS_GLOBAL=1)
  // a_m_dim == 2 ("map" mode: a W x H map tiled by boxes of map_bx x map_by = 128 positions, one image per
  // coordinate 3): m-tile t -> image t / map_tpg, first map row (t % map_tpg) * map_by; tile row r -> position
  // (x = r % map_bx, y = y0 + r / map_bx), valid while x < map_w and y < map_h; output row = image * map_w * map_h + y * map_w + x
  int map_bx, map_by, map_tpg, map_w, map_h;
  int ke;  // elements per 128-byte chunk
  int8_t tap_dx[9], tap_dy[9];
  int m_tiles, n_tiles;
  // epilogue
  int flags;
  void* out;
  int ldo;
  int rows_in, rows_out;
  const float* bias;
  const void* res;
  int ldr, res_div, res_rep;
  const float* pos;
  int pos_rows;
  const float* gamma;
  const float* beta;
  float eps;
  int round_tf32;
  int dual;  // extra centre tap into a second accumulator (BLOCK_N = 128)
  const float* bias2;
  float* ln_rstd;  // optional [rows_out space]: 1 / sigma of every LayerNorm row (saved for the backward pass)
  // split (bf16 hi/lo planes) mode: plane distances in bf16 elements
  int a_lo, w_lo, o_lo, r_lo;
  // split mode: compensation of the tensor core's round-toward-zero accumulation (see gemm_run)
  float acc_scale, acc_scale2;
  // split mode, element format of the two planes (0 = bf16, 1 = fp16; see aitb_gemm_desc.in_f16): MMA operands, the
  // output planes this launch writes, the residual planes it reads
  int in_f16, out_f16, res_f16;
};

__device__ __forceinline__ uint4 ld_global_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// GEMM row m (tile-major) -> output row; *ok = the row exists.  Plain: optional rows_in -> rows_out regrouping.
// Map mode: see GemmKParams.
__device__ __forceinline__ int out_row_of(const GemmKParams& p, int m, bool* ok) {
  if (p.a_m_dim == 2) {
    const int t = m >> 7, r = m & 127;
    const int g = t / p.map_tpg, y = (t - g * p.map_tpg) * p.map_by + r / p.map_bx, x = r % p.map_bx;
    *ok = m < p.M && x < p.map_w && y < p.map_h;
    return *ok ? (g * p.map_h + y) * p.map_w + x : 0;
  }
  *ok = m < p.M;
  const int mm = *ok ? m : 0;
  // no regrouping (every GEMM but the 49 <-> 64 row remaps): skip the integer division -- ncu showed 408 instructions of
  // per-tile set-up per epilogue warp (nine division sequences) against 4 x 226 for the tile's four chunks
  if (p.rows_in == p.rows_out) return mm;
  const int grp = mm / p.rows_in, r = mm - grp * p.rows_in;
  if (r >= p.rows_out) *ok = false;   // rows_out < rows_in: only the first rows_out rows of every group are stored
  return *ok ? grp * p.rows_out + r : 0;
}

// staging-tile accesses with shared-space instructions: through a generic pointer they compile to LD.E / ST.E
// (ncu: the first staged read of every chunk was the top long-scoreboard stall of the bf16 epilogue)
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}

__device__ __forceinline__ uint4 __float4_as_uint4(float4 f) {
  return make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// HALF (the split cluster-LayerNorm kernel): a stage holds 64-byte operand rows (32 bf16 of K, 64-byte swizzle)
// instead of 128-byte ones, so the four planes of a stage are 48 KB and the ring stays four deep.
template <int BLOCK_N, bool SPLIT, bool HALF = false>
struct GemmCfg {
  static constexpr int kRowBytes = HALF ? 64 : 128;   // bytes of K per operand row and stage
  static constexpr int kKSlices = kRowBytes / 32;     // 32-byte K slices (one tcgen05.mma each) per stage
  static constexpr int kABytes = kBlockM * kRowBytes;
  static constexpr int kBBytes = BLOCK_N * kRowBytes;
  static constexpr int kPlanes = SPLIT ? 2 : 1;  // split: [A_hi][A_lo][B_hi][B_lo] per stage
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStages = (!SPLIT || HALF) ? 4 : (BLOCK_N >= 256 ? 2 : (BLOCK_N >= 128 ? 3 : 4));
  static constexpr int kAccStages = 2;
  static constexpr int kUmmaN = BLOCK_N;
  // accumulator stage stride in TMEM columns; BLOCK_N = 128 reserves room for the dual accumulator
  static constexpr int kAccStride = (BLOCK_N == 128) ? 256 : BLOCK_N;
  static constexpr int kTmemCols = (kAccStride * kAccStages < 32) ? 32 : kAccStride * kAccStages;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    2048 /*LN row statistics*/ + 4 * 4096 /*epilogue staging, one 32x128 B tile per warp*/;
};

// ---------------------------------------------------------------------------------------------
// Epilogue of one 128 x BLOCK_N accumulator tile, executed by the 4 epilogue warps of a CTA.
// Each thread owns one accumulator row (tcgen05.ld 32x32b).  Global traffic never uses that
// row-per-thread shape: 32-column chunks go through a per-warp XOR-swizzled staging tile so that
// every global load/store instruction covers whole contiguous row segments (4 rows x 128 B for
// fp32, 8 rows x 64 B for bf16) -- fully coalesced, conflict-free in shared memory.
//   mt : 128-row tile index (rows mt*128 ...), n0 : first output column, t_row : TMEM address of this
//   warp's lane quadrant in the accumulator stage, stg : this warp's 4 KB staging tile.
// ---------------------------------------------------------------------------------------------
// PARTS (CL LayerNorm): partial row statistics combined per row -- 2 = one per CTA of the cluster (each thread
// drains all BLOCK_N columns of its CTA), 4 = two epilogue warpgroups per CTA, each draining COLS = BLOCK_N / 2.
// SIMPLE (compile-time, chosen per launch by the 2-CTA kernel): output rows are the GEMM rows (no 49 <-> 64 regrouping, no
// map mode), so the rows a lane stores are a fixed stride apart -- one base pointer + it * stride instead of kIt live
// 64-bit pointers that the compiler otherwise re-derives from the kernel parameters inside the chunk loop.
template <typename T, int BLOCK_N, bool CL, bool SPLIT, int COLS = BLOCK_N, int PARTS = 2, bool SIMPLE = false>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, uint8_t* stg, float2* stats,
                                              uint64_t* stats_full_bar, uint64_t* acc_full_bar, uint32_t t_row,
                                              int q, int lane, int mt, int n0, uint32_t as, uint32_t aph,
                                              uint32_t cta_rank, int c_begin = 0, int c_end = BLOCK_N) {
  constexpr int kRowBytes = 32 * (int)sizeof(T);       // one staged row: 32 columns
  constexpr int kCh = kRowBytes / 16;                  // 16-byte pieces per row (8 | 4)
  constexpr int kRpi = 32 / kCh;                       // rows covered by one warp instruction (4 | 8)
  constexpr int kIt = kCh;                             // instructions per 32-row chunk (8 | 4)
  constexpr uint32_t kFull = 0xffffffffu;
  const int row_in_tile = q * 32 + lane;
  const int piece = lane % kCh;
  const int srow0 = lane / kCh;
  auto phys = [](int r, int j) { return j ^ ((r / (8 / kCh)) % kCh); };
  const uint32_t stg_s = smem_u32(stg);
  const T* res = reinterpret_cast<const T*>(p.res);

  // rows this lane serves in the coalesced phases
  T* optr[SIMPLE ? 1 : kIt];
  const T* rptr[kIt];
  uint32_t vmask = 0;
  // SIMPLE: row of it = 0; rows at or beyond M are masked by vmask, so their (never dereferenced) addresses do not matter
  T* const optr0 = reinterpret_cast<T*>(p.out) + (size_t)(mt * kBlockM + q * 32 + srow0) * p.ldo + n0 + piece * (16 / (int)sizeof(T));
  const size_t ostep = (size_t)kRpi * p.ldo;
  auto optr_of = [&](int it) -> T* {
    if constexpr (SIMPLE) return optr0 + (size_t)it * ostep;
    else return optr[it];
  };
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int m = mt * kBlockM + q * 32 + it * kRpi + srow0;
    bool ok;
    const int orow = out_row_of(p, m, &ok);
    if constexpr (!SIMPLE) optr[it] = reinterpret_cast<T*>(p.out) + (size_t)orow * p.ldo + n0 + piece * (16 / (int)sizeof(T));
    int rrow = 0;
    if (p.flags & (AITB_EPI_RES | AITB_EPI_RELU_MASK)) {
      const int rb = (p.flags & AITB_EPI_RES_ROW_M) ? (ok ? m : 0) : orow;   // residual indexed by GEMM row or by output row
      rrow = ((rb / p.res_div) / p.res_rep) * p.res_div + (rb % p.res_div);
    }
    rptr[it] = res + (size_t)rrow * p.ldr + n0 + piece * (16 / (int)sizeof(T));
    vmask |= (ok ? 1u : 0u) << it;
  }
  // this thread's own row (for the fp32 positional table)
  const int m_own = mt * kBlockM + row_in_tile;
  bool ok_own;
  const int orow_own = out_row_of(p, m_own, &ok_own);
  const float* prow = p.pos + (size_t)((p.flags & AITB_EPI_POS) ? (orow_own % p.pos_rows) : 0) * p.N + n0;

  // issue this lane's share of a coalesced 32-row x 32-column global read (no dependent use yet)
  auto issue_loads = [&](const T* const* ptrs, int c0, uint4 (&val)[kIt]) {
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      val[it] = make_uint4(0u, 0u, 0u, 0u);
      if ((vmask >> it) & 1u) val[it] = ld_global_v4(ptrs[it] + c0);
    }
  };
  // the same from the output rows themselves (ACCUM epilogue)
  auto issue_out_loads = [&](int c0, uint4 (&val)[kIt]) {
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      val[it] = make_uint4(0u, 0u, 0u, 0u);
      if ((vmask >> it) & 1u) val[it] = ld_global_v4(optr_of(it) + c0);
    }
  };
  // ... and turn it into this thread's 32 values of its own row through the swizzled staging tile
  const bool of16 = SPLIT && p.out_f16 != 0, rf16 = SPLIT && p.res_f16 != 0;
  auto exchange = [&](const uint4 (&val)[kIt], float (&r)[32], bool f16 = false) {
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int sr = it * kRpi + srow0;
      sts128(stg_s + sr * kRowBytes + phys(sr, piece) * 16, val[it]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kCh; ++j) {
      const uint4 x = lds128(stg_s + lane * kRowBytes + phys(lane, j) * 16);
      if constexpr (sizeof(T) == 4) {
        r[4 * j + 0] = __uint_as_float(x.x); r[4 * j + 1] = __uint_as_float(x.y);
        r[4 * j + 2] = __uint_as_float(x.z); r[4 * j + 3] = __uint_as_float(x.w);
      } else if (f16) {
        const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          r[8 * j + 2 * e] = f.x;
          r[8 * j + 2 * e + 1] = f.y;
        }
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          r[8 * j + 2 * e] = f.x;
          r[8 * j + 2 * e + 1] = f.y;
        }
      }
    }
    __syncwarp();
  };
  // Per-column parameters (bias / gamma / beta) of the columns this warp drains, [c_begin, c_end), are
  // fetched ONCE per tile -- before the accumulator wait, so their latency hides behind the main loop --
  // as one 128-bit load per lane and 128-column slice, and handed to the row-owning threads by shuffles.
  // COLS = number of columns this warp drains (c_end - c_begin)
  struct LaneVec { float4 s[COLS > 128 ? 2 : 1]; };
  auto load_lane_vec = [&](const float* base) {  // base = parameter array + n0 + c_begin
    LaneVec lv;
    lv.s[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane * 4 < COLS) lv.s[0] = __ldg(reinterpret_cast<const float4*>(base) + lane);
    if constexpr (COLS > 128) lv.s[1] = __ldg(reinterpret_cast<const float4*>(base + 128) + lane);
    return lv;
  };
  // v[j] (+)= param[c_rel + j], c_rel = chunk offset from c_begin (multiple of 32)
  auto lane_vec_chunk = [&](const LaneVec& lv, int c_rel, float (&o)[32]) {
    float4 b = lv.s[0];
    if constexpr (COLS > 128) {
      if (c_rel & 128) b = lv.s[1];
    }
    const int sl = (c_rel & 127) >> 2;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      o[4 * jj + 0] = __shfl_sync(kFull, b.x, sl + jj);
      o[4 * jj + 1] = __shfl_sync(kFull, b.y, sl + jj);
      o[4 * jj + 2] = __shfl_sync(kFull, b.z, sl + jj);
      o[4 * jj + 3] = __shfl_sync(kFull, b.w, sl + jj);
    }
  };
  auto add_lane_vec = [&](const LaneVec& lv, int c_rel, float (&v)[32]) {
    float o[32];
    lane_vec_chunk(lv, c_rel, o);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += o[j];
  };
  // fp32 positional table rows differ per thread: direct vector loads
  auto add_vec = [&](const float* src, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  };
  // this thread's 32 values -> coalesced global store at column `c0` of the rows this warp covers
  auto stage_store_raw = [&](int c0, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < kCh; ++j) {
      uint4 x;
      if constexpr (sizeof(T) == 4) {
        x = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                       __float_as_uint(v[4 * j + 3]));
      } else if (of16) {
        __half2* h = reinterpret_cast<__half2*>(&x);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
      } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
      }
      sts128(stg_s + lane * kRowBytes + phys(lane, j) * 16, x);
    }
    __syncwarp();
    uint4 val[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int sr = it * kRpi + srow0;
      val[it] = lds128(stg_s + sr * kRowBytes + phys(sr, piece) * 16);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it)
      if ((vmask >> it) & 1u) *reinterpret_cast<uint4*>(optr_of(it) + c0) = val[it];
    __syncwarp();
  };
  // split mode: hi = bf16(v) into the hi plane, lo = bf16(v - hi) into the lo plane (o_lo columns further)
  auto stage_store = [&](int c0, const float (&vin)[32]) {
    if constexpr (SPLIT) {
      float v[32], lo[32];
      if (of16) {   // fp16 planes: 11 + 11 bits; the hi plane saturates at the fp16 range instead of overflowing to inf
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = fminf(fmaxf(vin[j], -65504.f), 65504.f);
          lo[j] = v[j] - __half2float(__float2half_rn(v[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = vin[j];
          lo[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
        }
      }
      stage_store_raw(c0, v);
      stage_store_raw(c0 + p.o_lo, lo);
    } else {
      stage_store_raw(c0, vin);
    }
  };
  // residual chunk (columns c0..c0+31 of this thread's row) from prefetched coalesced loads
  auto add_residual = [&](const uint4 (&hi)[kIt], const uint4 (&lo)[kIt], float (&v)[32]) {
    float r[32];
    exchange(hi, r, rf16);
    if (p.flags & AITB_EPI_RELU_MASK) {   // backward of ReLU: the "residual" stream is the saved activation
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = r[j] > 0.f ? v[j] : 0.f;
      return;
    }
    if constexpr (SPLIT) {
      float r2[32];
      exchange(lo, r2, rf16);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r[j] + r2[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r[j];
    }
  };
  // pull this warp's residual rows towards L2 while the main loop of the tile is still running
  auto prefetch_residual = [&]() {
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      if (!((vmask >> it) & 1u)) continue;
      // the kCh lanes that share a row split its lines among themselves
      for (int l = piece; l < (COLS * (int)sizeof(T) + 127) / 128; l += kCh) {
        const T* a = rptr[it] - piece * (16 / (int)sizeof(T)) + c_begin + l * (128 / (int)sizeof(T));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        if constexpr (SPLIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + p.r_lo));
      }
    }
  };

  const bool aux_res = (p.flags & (AITB_EPI_RES | AITB_EPI_RELU_MASK)) != 0;
  LaneVec lv_bias, lv_gamma, lv_beta;  // CL (LayerNorm path) only
  if constexpr (CL) {   // the LayerNorm path hands bias / gamma / beta to the row owners by shuffles; the plain path loads them
    if ((p.flags & AITB_EPI_BIAS) && (p.flags & AITB_EPI_LN)) lv_bias = load_lane_vec(p.bias + n0 + c_begin);
  }
  if constexpr (CL) {
    if (p.flags & AITB_EPI_LN) {
      lv_gamma = load_lane_vec(p.gamma + n0 + c_begin);
      lv_beta = load_lane_vec(p.beta + n0 + c_begin);
    }
  }
  if (aux_res) prefetch_residual();
  // 2-byte outputs use only the lower 2 KB of the warp's 4 KB staging tile (32 rows x 64 B): the upper half keeps this
  // warp's private copy of the tile's bias slice, fetched here -- before the accumulator wait -- so the chunk loop reads it
  // with broadcast ld.shared (~25 clk) instead of global loads at the point of use (ncu: those missed L1 next to 225 KB of
  // shared memory and cost an L2 round trip per chunk, ~35 % of the loop's stall samples).  AITB_BIAS_GLOBAL=1: A/B.
  constexpr bool kBiasSmem = sizeof(T) == 2 && !CL && COLS <= 256;
  const uint32_t bias_s = stg_s + 2048;
  const bool bias_smem = kBiasSmem && (p.flags & AITB_EPI_BIAS) != 0 && p.bias_smem;
  if constexpr (kBiasSmem) {
    if (bias_smem) {
      const float4* bsrc = reinterpret_cast<const float4*>(p.bias + n0 + c_begin);
      if (lane * 4 < COLS) sts128(bias_s + lane * 16, __float4_as_uint4(__ldg(bsrc + lane)));
      if constexpr (COLS > 128) sts128(bias_s + 512 + lane * 16, __float4_as_uint4(__ldg(bsrc + 32 + lane)));
      __syncwarp();
    }
  }

  mbar_wait(acc_full_bar, aph);
  tc_fence_after();

  if (!CL || (p.flags & AITB_EPI_LN) == 0) {
    // one auxiliary read stream (residual, or the output itself for ACCUM) is prefetched one chunk ahead;
    // the TMEM read of chunk c+1 is in flight while chunk c is processed
    const bool aux_acc = !aux_res && (p.flags & AITB_EPI_ACCUM) != 0;
    const bool dual = BLOCK_N == 128 && (p.flags & AITB_EPI_DUAL) != 0;
    uint4 pre[kIt], pre2[kIt];
    if (aux_res) issue_loads(rptr, c_begin, pre);
    if (aux_acc) issue_out_loads(c_begin, pre);
    if constexpr (SPLIT) { if (aux_res) issue_loads(rptr, c_begin + p.r_lo, pre2); }
    uint32_t raw[32], raw2[32];
    tmem_ld32(t_row + c_begin, raw);
    if (dual) tmem_ld32(t_row + BLOCK_N + c_begin, raw2);
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      tmem_ld_wait();
      float v[32], u[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if constexpr (BLOCK_N == 128) {
        if (dual) {
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __uint_as_float(raw2[j]);
        }
      }
      if (c0 + 32 < c_end) {  // next chunk's TMEM read overlaps this chunk's math and stores
        tmem_ld32(t_row + c0 + 32, raw);
        if (dual) tmem_ld32(t_row + BLOCK_N + c0 + 32, raw2);
      }
      if (SPLIT || p.acc_scale != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.acc_scale;
      }
      // bias: eight uniform-address 128-bit loads (one L1 wavefront each) instead of 32 shuffles through the MIO pipe
      if (bias_smem) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 b = lds128(bias_s + (uint32_t)(c0 - c_begin) * 4u + (uint32_t)j * 16u);
          v[4 * j] += __uint_as_float(b.x); v[4 * j + 1] += __uint_as_float(b.y);
          v[4 * j + 2] += __uint_as_float(b.z); v[4 * j + 3] += __uint_as_float(b.w);
        }
      } else if (p.flags & AITB_EPI_BIAS) {
        add_vec(p.bias + n0 + c0, v);
      }
      if (p.flags & AITB_EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.flags & AITB_EPI_SQUARE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] * v[j];
      }
      if constexpr (BLOCK_N == 128) {
        if (dual) {  // second accumulator: same activation, then summed
          if constexpr (SPLIT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) u[j] *= p.acc_scale2;
          }
          add_vec(p.bias2 + n0 + c0, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (p.flags & AITB_EPI_RELU) u[j] = fmaxf(u[j], 0.f);
            if (p.flags & AITB_EPI_SQUARE) u[j] = u[j] * u[j];
            v[j] += u[j];
          }
        }
      }
      if (aux_res) {
        add_residual(pre, pre2, v);
        if (c0 + 32 < c_end) {  // next chunk's residual (its lines were prefetched into L2 before the tile)
          issue_loads(rptr, c0 + 32, pre);
          if constexpr (SPLIT) issue_loads(rptr, c0 + 32 + p.r_lo, pre2);
        }
      }
      if (p.flags & AITB_EPI_POS) add_vec(prow + c0, v);
      if (p.flags & AITB_EPI_ACCUM) {
        float r[32];
        if (!aux_acc) issue_out_loads(c0, pre);   // RES and ACCUM together: second stream not prefetched
        exchange(pre, r);
        if (aux_acc && c0 + 32 < c_end) issue_out_loads(c0 + 32, pre);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += r[j];
      }
      if (p.flags & AITB_EPI_RES_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (sizeof(T) == 4 && p.round_tf32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
      }
      stage_store(c0, v);
    }
  } else {
    // ---- full-row LayerNorm.  Plain: the CTA's accumulator holds all N = BLOCK_N columns of the row.
    //      CL: it holds half of the row; partial (sum, M2) are exchanged with the peer CTA via DSMEM.
    // Pass 1 builds x = acc + bias + residual + pos, writes it back to TMEM and accumulates the shifted sums
    // S1 = sum(x - s), S2 = sum((x - s)^2) with s = mean of the row's first 32 columns (so the variance
    // M2 = S2 - S1^2 / n loses nothing to cancellation); pass 2 normalises.
    float shift = 0.f, S1 = 0.f, S2 = 0.f;
    uint4 pre[kIt], pre2[kIt];
    if (aux_res) issue_loads(rptr, c_begin, pre);
    if constexpr (SPLIT) { if (aux_res) issue_loads(rptr, c_begin + p.r_lo, pre2); }
    uint32_t raw[32];
    tmem_ld32(t_row + c_begin, raw);
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (c0 + 32 < c_end) tmem_ld32(t_row + c0 + 32, raw);
      if constexpr (SPLIT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.acc_scale;
      }
      if (p.flags & AITB_EPI_BIAS) add_lane_vec(lv_bias, c0 - c_begin, v);
      if (aux_res) {
        add_residual(pre, pre2, v);
        if (c0 + 32 < c_end) {
          issue_loads(rptr, c0 + 32, pre);
          if constexpr (SPLIT) issue_loads(rptr, c0 + 32 + p.r_lo, pre2);
        }
      }
      if (p.flags & AITB_EPI_POS) add_vec(prow + c0, v);
      if (c0 == c_begin) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) s += v[j];
        shift = s * (1.f / 32.f);
      }
      uint32_t xo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - shift;
        S1 += d;
        S2 += d * d;
        xo[j] = __float_as_uint(v[j]);
      }
      // the next chunk's tcgen05.ld is in flight: a st to a DIFFERENT column range is independent of it
      tmem_st32(t_row + c0, xo);
    }
    tmem_st_wait();
    float mean = shift + S1 * (1.f / COLS);
    float ssq = S2 - S1 * S1 * (1.f / COLS);  // centred M2 of my columns
    float n_cols = (float)COLS;
    if constexpr (CL && PARTS == 4) {
      // four partials per row: (cluster rank, column half).  Every epilogue thread of the cluster publishes its
      // (sum, M2) in BOTH CTAs' statistics tables and arrives on both barriers (count 512), then combines.
      const float sum = mean * (float)COLS;
      const uint32_t peer = cta_rank ^ 1u;
      const int slot = (int)cta_rank * 2 + (c_begin != 0 ? 1 : 0);
      float2* mine = &stats[(as * 4 + slot) * 128 + row_in_tile];
      *mine = make_float2(sum, ssq);
      st_cluster_f32x2(mapa_u32(smem_u32(mine), peer), sum, ssq);
      mbar_arrive(stats_full_bar);
      mbar_arrive_cluster(mapa_u32(smem_u32(stats_full_bar), peer));
      tmem_ld32(t_row + c_begin, raw);  // first chunk of pass 2 travels while we wait
      mbar_wait_cluster(stats_full_bar, aph);
      float2 ps[4];
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ps[i] = stats[(as * 4 + i) * 128 + row_in_tile];
        tot += ps[i].x;
      }
      const float mean_all = tot * (1.f / (4 * COLS));
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dm = ps[i].x * (1.f / COLS) - mean_all;
        m2 += ps[i].y + (float)COLS * dm * dm;
      }
      ssq = m2;
      mean = mean_all;
      n_cols = 4.f * COLS;
    } else if constexpr (CL) {
      // send (sum, M2) of my half-row to the peer, wait for the peer's, combine (Chan et al.)
      const float sum = mean * (float)BLOCK_N;
      const uint32_t peer = cta_rank ^ 1u;
      st_cluster_f32x2(mapa_u32(smem_u32(&stats[as * 128 + row_in_tile]), peer), sum, ssq);
      mbar_arrive_cluster(mapa_u32(smem_u32(stats_full_bar), peer));
      tmem_ld32(t_row, raw);  // first chunk of pass 2 travels while we wait for the peer
      mbar_wait_cluster(stats_full_bar, aph);
      const float2 ps = stats[as * 128 + row_in_tile];
      const float mean_p = ps.x * (1.f / BLOCK_N);
      const float mean_all = (sum + ps.x) * (0.5f / BLOCK_N);
      const float da = mean - mean_all, db = mean_p - mean_all;
      ssq = ssq + ps.y + (float)BLOCK_N * (da * da + db * db);
      mean = mean_all;
      n_cols = 2.f * BLOCK_N;
    } else {
      tmem_ld32(t_row, raw);
    }
    const float rstd = rsqrtf(fmaxf(ssq, 0.f) / n_cols + p.eps);
    if (p.ln_rstd != nullptr && cta_rank == 0 && c_begin == 0 && ok_own) p.ln_rstd[orow_own] = rstd;
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(raw[j]) - mean) * rstd;
      if (c0 + 32 < c_end) tmem_ld32(t_row + c0 + 32, raw);
      float g[32];
      lane_vec_chunk(lv_gamma, c0 - c_begin, g);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= g[j];
      add_lane_vec(lv_beta, c0 - c_begin, v);
      if (sizeof(T) == 4 && p.round_tf32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
      }
      stage_store(c0, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast epilogue of the 2-CTA kernel for 2-byte (bf16) outputs whose rows are the GEMM rows and whose epilogue is at most
// bias + ReLU (QKV / K,V projections, FFN w_1, dec_trans, the layer-4 1x1 / 3x3 convolutions): the single-pass bf16 GEMMs
// were measured EPILOGUE-bound (tensor pipe 58-64 % active; 196 SASS instructions per 32-column chunk and warp, 408 more
// per tile: profiles/r01f_bf16_gemm2_epilogue.txt).  Here a chunk is: tcgen05.ld -> 32 FADD with the bias slice (broadcast
// ld.shared.v4) -> 16 cvt.rn[.relu].bf16x2.f32 (the ReLU rides in the conversion) -> 4 st.shared.v4 into a 32 x 64 B tile in
// the TMA 64-byte-swizzle pattern -> ONE cp.async.bulk.tensor store issued by lane 0.  No per-row global pointers, no
// staged read-back, no bounds masks (TMA clips rows >= M), every chunk of a tile has its own staging tile so nothing waits
// for a store inside a tile.
//   stg_s   : this warp's 4 x 2 KB staging tiles (1024-byte aligned)
//   bias_s  : this warpgroup's bias slice [2][128] floats (double-buffered by accumulator stage)
//   n0      : first output column of this warp's 128 columns;   row0 : first output row of this warp's 32 rows
// ---------------------------------------------------------------------------------------------
// SPLIT_OUT (the one-pass GEMMs of the fp32 configuration's precision plan): the output is two 16-bit planes (bf16 or fp16,
// p.out_f16) hi | lo, N columns apart; per chunk a hi tile and a lo tile, two staging buffers per warp (chunk c reuses the
// buffer of chunk c - 2 after cp.async.bulk.wait_group.read 1).  p.flags & AITB_EPI_HI_ONLY: the lo plane is not written
// (the only consumer reads the hi plane: the FFN hidden tensor between two one-pass GEMMs).
// Residual (plain bf16 only: the layer-4 conv3 launches, bias + residual + ReLU): the four 32 x 32 residual tiles of the warp are
// TMA-LOADED into the same staging tiles at the start of the tile (one mbarrier per warp, `res_bar`), long before the
// accumulator is ready; a thread then reads its own row (4 ld.shared.v4), adds, and writes the result back IN PLACE.
template <bool SPLIT_OUT>
__device__ __forceinline__ void epilogue_fast_tile(const GemmKParams& p, const CUtensorMap* tmO, uint32_t stg_s, uint32_t bias_s,
                                                   int bar_id, uint64_t* acc_full_bar, uint64_t* acc_empty_bar, uint32_t aph,
                                                   uint32_t t_row, int lane, int wg_tid, int row0, int n0, uint32_t as,
                                                   const CUtensorMap* tmR = nullptr, uint64_t* res_bar = nullptr,
                                                   uint32_t res_phase = 0) {
  const bool has_bias = (p.flags & AITB_EPI_BIAS) != 0;
  const bool relu = (p.flags & AITB_EPI_RELU) != 0;
  const bool has_res = !SPLIT_OUT && (p.flags & AITB_EPI_RES) != 0;
  const bool res_relu = (p.flags & AITB_EPI_RES_RELU) != 0;
  const bool of16 = SPLIT_OUT && p.out_f16 != 0;
  const bool hi_only = SPLIT_OUT && (p.flags & AITB_EPI_HI_ONLY) != 0;
  const uint32_t bias_buf = bias_s + as * 512u;
  if (lane == 0) {
    bulk_wait_read_all();                         // the previous tile's stores no longer read the staging tiles
    if (has_res) {                                // residual tiles -> the staging tiles (rows >= M are zero-filled by TMA)
      mbar_arrive_expect_tx(res_bar, 4 * 2048);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_2d(reinterpret_cast<void*>(__cvta_shared_to_generic(stg_s + (uint32_t)c * 2048u)), tmR, res_bar, n0 + c * 32, row0);
    }
  }
  if (has_bias) {
    const float b = __ldg(p.bias + n0 + wg_tid);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_buf + (uint32_t)wg_tid * 4u), "f"(b) : "memory");
    named_bar_sync(bar_id, 128);                  // the slice is complete (and lane 0's wait above is ordered before our writes)
  } else {
    __syncwarp();
  }
  mbar_wait(acc_full_bar, aph);
  tc_fence_after();
  uint32_t raw[32];
  tmem_ld32(t_row, raw);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (c + 1 < 4) {
      tmem_ld32(t_row + (uint32_t)(c + 1) * 32u, raw);
    } else {                                      // the accumulator stage is drained: hand it back before the stores
      tc_fence_before();
      mbar_arrive_remote(acc_empty_bar, 0);
    }
    if (SPLIT_OUT || p.acc_scale != 1.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.acc_scale;
    }
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 b = lds128(bias_buf + (uint32_t)c * 128u + (uint32_t)j * 16u);
        v[4 * j] += __uint_as_float(b.x); v[4 * j + 1] += __uint_as_float(b.y);
        v[4 * j + 2] += __uint_as_float(b.z); v[4 * j + 3] += __uint_as_float(b.w);
      }
    }
    const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
    if constexpr (SPLIT_OUT) {
      if (c >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // chunk c - 2 left its buffer
      __syncwarp();
      const uint32_t tile = stg_s + (uint32_t)(c & 1) * 4096u;    // [hi 2 KB | lo 2 KB]
      const uint32_t rowb = tile + (uint32_t)lane * 64u;
      if (relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo[8];
        uint4 hi, lw;
        hi.x = split_hi2(v[8 * j], v[8 * j + 1], of16, lo[0], lo[1]);
        hi.y = split_hi2(v[8 * j + 2], v[8 * j + 3], of16, lo[2], lo[3]);
        hi.z = split_hi2(v[8 * j + 4], v[8 * j + 5], of16, lo[4], lo[5]);
        hi.w = split_hi2(v[8 * j + 6], v[8 * j + 7], of16, lo[6], lo[7]);
        sts128(rowb + (((uint32_t)j ^ sw) << 4), hi);
        if (!hi_only) {
          lw.x = pack_plane2(lo[0], lo[1], of16); lw.y = pack_plane2(lo[2], lo[3], of16);
          lw.z = pack_plane2(lo[4], lo[5], of16); lw.w = pack_plane2(lo[6], lo[7], of16);
          sts128(rowb + 2048u + (((uint32_t)j ^ sw) << 4), lw);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmO, tile, n0 + c * 32, row0);
        if (!hi_only) tma_store_2d(tmO, tile + 2048u, p.o_lo + n0 + c * 32, row0);
        bulk_commit_group();
      }
    } else {
      const uint32_t tile = stg_s + (uint32_t)c * 2048u;
      const uint32_t rowb = tile + (uint32_t)lane * 64u;
      bool relu_c = relu;
      if (has_res) {
        if (c == 0) mbar_wait(res_bar, res_phase);
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 r = lds128(rowb + (((uint32_t)j ^ sw) << 4));
          const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[8 * j + 2 * e] += __uint_as_float(rw[e] << 16);
            v[8 * j + 2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
          }
        }
        relu_c = res_relu;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 x;
        if (relu_c) {
          x.x = pack_bf16x2_relu(v[8 * j], v[8 * j + 1]); x.y = pack_bf16x2_relu(v[8 * j + 2], v[8 * j + 3]);
          x.z = pack_bf16x2_relu(v[8 * j + 4], v[8 * j + 5]); x.w = pack_bf16x2_relu(v[8 * j + 6], v[8 * j + 7]);
        } else {
          x.x = pack_bf16x2(v[8 * j], v[8 * j + 1]); x.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          x.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); x.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        }
        sts128(rowb + (((uint32_t)j ^ sw) << 4), x);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmO, tile, n0 + c * 32, row0);
        bulk_commit_group();
      }
    }
  }
}

// CL = true: "cluster LayerNorm" variant.  A 2-CTA cluster shares one 128-row m-tile; CTA rank r owns the
// output columns [256 r, 256 r + 256) of the N = 512 row (BLOCK_N = 256 machinery: 4-stage ring, double-
// buffered accumulator).  The LayerNorm row statistics are combined across the two CTAs through
// distributed shared memory (partial sum + centred M2 per row, Chan's parallel-variance formula).
template <typename T, int BLOCK_N, bool CL, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p) {
  using Cfg = GemmCfg<BLOCK_N, SPLIT, CL && SPLIT>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAcc = Cfg::kAccStages;
  constexpr int kABytes = Cfg::kABytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;    // [kAcc]
  uint64_t* acc_empty = acc_full + kAcc;      // [kAcc]
  uint64_t* stats_full = acc_empty + kAcc;    // [kAcc]  (CL only) peer's row statistics have landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(stats_full + kAcc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 128);
      mbar_init(&stats_full[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL) cluster_sync_all();  // peer's mbarriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // tile schedule: plain = tiles round-robin over CTAs (n fastest); CL = m-tiles round-robin over clusters
  const uint32_t cta_rank = CL ? cluster_ctarank() : 0u;
  const int t_first = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_tiles_total = CL ? p.m_tiles : p.m_tiles * p.n_tiles;
  const int n_taps = p.taps + (p.dual ? 1 : 0);
  const int iters_per_tile = n_taps * p.k_chunks;
  const int dual_first_iter = p.dual ? p.taps * p.k_chunks : -1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = t_first; tile < n_tiles_total; tile += t_step) {
        const int mt = CL ? tile : tile / p.n_tiles;
        const int nt = CL ? (int)cta_rank : tile - mt * p.n_tiles;
        const int a_c_base = nt * p.a_group_c;
        for (int tap = 0; tap < n_taps; ++tap) {
          int c1 = tap < p.taps ? p.tap_dx[tap] : 0, c2 = tap < p.taps ? p.tap_dy[tap] : 0, c3 = 0;
          if (p.a_m_dim == 1) c1 += mt * p.a_m_step;
          else if (p.a_m_dim == 3) c3 += mt * p.a_m_step;
          else { c2 += (mt % p.map_tpg) * p.map_by; c3 += mt / p.map_tpg; }
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            const uint32_t s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kPlanes * kABytes;
            mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
            tma_load_4d(sa, &tmA, &full_bar[s], a_c_base + kc * p.ke, c1, c2, c3);
            const int kb = (tap * p.k_chunks + kc) * p.ke;
            tma_load_2d(sb, &tmB, &full_bar[s], kb, nt * BLOCK_N);
            if constexpr (SPLIT) {
              tma_load_4d(sa + kABytes, &tmA, &full_bar[s], p.a_lo + a_c_base + kc * p.ke, c1, c2, c3);
              tma_load_2d(sb + Cfg::kBBytes, &tmB, &full_bar[s], p.w_lo + kb, nt * BLOCK_N);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc((SPLIT && p.in_f16) ? 0u : Act<T>::kFmt, kBlockM, Cfg::kUmmaN);
      uint32_t it = 0;
      uint32_t lt = 0;  // local tile counter
      for (int tile = t_first; tile < n_tiles_total; tile += t_step, ++lt) {
        const uint32_t as = lt % kAcc;
        const uint32_t aph = (lt / kAcc) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem0 = tmem_base + as * Cfg::kAccStride;
        for (int i = 0; i < iters_per_tile; ++i, ++it) {
          const bool second = p.dual && i >= dual_first_iter;
          const uint32_t d_tmem = d_tmem0 + (second ? BLOCK_N : 0);
          const bool fresh = (i == 0) || (i == dual_first_iter);
          const uint32_t s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = make_kmajor_desc<Cfg::kRowBytes>(sa);
          const uint64_t bdesc = make_kmajor_desc<Cfg::kRowBytes>(sa + Cfg::kPlanes * kABytes);
#pragma unroll
          for (int k = 0; k < Cfg::kKSlices; ++k) {  // 32-byte K slices of the stage's operand rows
            umma_ss<Act<T>::kBytes>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                    (fresh && k == 0) ? 0u : 1u);
            if constexpr (SPLIT) {  // + A_hi * W_lo + A_lo * W_hi
              umma_ss<2>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(Cfg::kBBytes / 16 + k * 2), idesc, 1u);
              umma_ss<2>(d_tmem, adesc + (uint64_t)(kABytes / 16 + k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            }
          }
          tc_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
        tc_commit(&acc_full[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    float2* stats = reinterpret_cast<float2*>(smem + kStages * Cfg::kStageBytes + 256);   // [kAcc][128] (CL)
    uint8_t* stg = smem + kStages * Cfg::kStageBytes + 256 + 2048 + q * 4096;
    uint32_t lt = 0;
    for (int tile = t_first; tile < n_tiles_total; tile += t_step, ++lt) {
      const int mt = CL ? tile : tile / p.n_tiles;
      const int nt = CL ? (int)cta_rank : tile - mt * p.n_tiles;
      const uint32_t as = lt % kAcc;
      const uint32_t aph = (lt / kAcc) & 1;
      const uint32_t t_row = tmem_base + as * Cfg::kAccStride + ((uint32_t)(q * 32) << 16);
      epilogue_tile<T, BLOCK_N, CL, SPLIT>(p, stg, stats, &stats_full[as], &acc_full[as], t_row, q, lane, mt,
                                           nt * BLOCK_N, as, aph, cta_rank);
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL) cluster_sync_all();  // the peer may still be writing row statistics into my smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Cluster LayerNorm, two epilogue warpgroups (the production LN kernel).  Same tile schedule and main loop
// as gemm_tcgen05_kernel<.., CL = true>: a 2-CTA cluster shares a 128-row m-tile, CTA rank r owns columns
// [256 r, 256 r + 256) of the N = 512 row.  The LayerNorm epilogue is instruction/latency-bound (two passes over
// TMEM per row, residual exchange, staged stores), so it gets EIGHT warps: warpgroup 1 drains columns 0..127 of
// the CTA's accumulator, warpgroup 2 columns 128..255; the row statistics are four partials per row combined
// through shared memory + DSMEM.  Measured on the K = 64 attention-fc GEMM (pure epilogue): see DESIGN.md.
//   threads: warp 0 TMA, warp 1 MMA (warps 2, 3 idle, 40 registers), warps 4..11 epilogue (232 registers)
// ---------------------------------------------------------------------------------------------
static constexpr int kLnThreads = 384;

// SPLIT_IN: the stages hold both planes of A and W and every K slice issues three MMAs.  A split-OUTPUT kernel with
// SPLIT_IN = false is the one-pass variant (aitb_gemm_desc.passes == 1): hi planes only, the main loop of the plain kernel.
template <typename T, bool SPLIT_IN>
struct LnCfg {
  using Base = GemmCfg<256, SPLIT_IN, SPLIT_IN>;
  static constexpr int kStages = (sizeof(T) == 4) ? 3 : 4;   // fp32 staging tiles are twice as large
  static constexpr int kStageBytes = Base::kStageBytes;
  static constexpr int kStatsBytes = 2 * 4 * 128 * 8;         // [acc stage][part][row] float2
  static constexpr int kStgBytes = 8 * 32 * 32 * (int)sizeof(T);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kStatsBytes + kStgBytes;
};

template <typename T, bool SPLIT, bool ONEPASS = false>
__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmKParams p) {
  constexpr int BLOCK_N = 256;
  constexpr bool SPLIT_IN = SPLIT && !ONEPASS;
  using Cfg = GemmCfg<BLOCK_N, SPLIT_IN, SPLIT_IN>;
  using L = LnCfg<T, SPLIT_IN>;
  constexpr int kStages = L::kStages;
  constexpr int kAcc = 2;
  constexpr int kABytes = Cfg::kABytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = acc_full + kAcc;
  uint64_t* stats_full = acc_empty + kAcc;    // [kAcc][4]: one barrier per accumulator stage and 32-row quarter
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(stats_full + kAcc * 4);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 256);
      // the four partial statistics of a row come from the four warps (2 CTAs x 2 column halves) that own the same 32-row
      // quarter: a warp waits for those three partners only, not for all sixteen epilogue warps of the cluster (ncu r02o:
      // 18 % of the epilogue warps' samples sat in this wait with one barrier of count 512 per stage)
      for (int qq = 0; qq < 4; ++qq) mbar_init(&stats_full[s * 4 + qq], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t cta_rank = cluster_ctarank();
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);
  const int iters_per_tile = p.taps * p.k_chunks;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {
        uint32_t it = 0;
        for (int tile = t_first; tile < p.m_tiles; tile += t_step) {
          const int nt = (int)cta_rank;
          for (int tap = 0; tap < p.taps; ++tap) {
            int c1 = p.tap_dx[tap], c2 = p.tap_dy[tap], c3 = 0;
            if (p.a_m_dim == 1) c1 += tile * p.a_m_step;
          else if (p.a_m_dim == 3) c3 += tile * p.a_m_step;
          else { c2 += (tile % p.map_tpg) * p.map_by; c3 += tile / p.map_tpg; }
            for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
              const uint32_t s = it % kStages;
              const uint32_t ph = (it / kStages) & 1;
              mbar_wait(&empty_bar[s], ph ^ 1);
              uint8_t* sa = smem + s * Cfg::kStageBytes;
              uint8_t* sb = sa + Cfg::kPlanes * kABytes;
              mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
              tma_load_4d(sa, &tmA, &full_bar[s], kc * p.ke, c1, c2, c3);
              const int kb = (tap * p.k_chunks + kc) * p.ke;
              tma_load_2d(sb, &tmB, &full_bar[s], kb, nt * BLOCK_N);
              if constexpr (SPLIT_IN) {
                tma_load_4d(sa + kABytes, &tmA, &full_bar[s], p.a_lo + kc * p.ke, c1, c2, c3);
                tma_load_2d(sb + Cfg::kBBytes, &tmB, &full_bar[s], p.w_lo + kb, nt * BLOCK_N);
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc((SPLIT && p.in_f16) ? 0u : Act<T>::kFmt, kBlockM, BLOCK_N);
        uint32_t it = 0, lt = 0;
        for (int tile = t_first; tile < p.m_tiles; tile += t_step, ++lt) {
          const uint32_t as = lt % kAcc;
          const uint32_t aph = (lt / kAcc) & 1;
          mbar_wait(&acc_empty[as], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BLOCK_N;
          for (int i = 0; i < iters_per_tile; ++i, ++it) {
            const uint32_t s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
            const uint64_t adesc = make_kmajor_desc<Cfg::kRowBytes>(sa);
            const uint64_t bdesc = make_kmajor_desc<Cfg::kRowBytes>(sa + Cfg::kPlanes * kABytes);
#pragma unroll
            for (int k = 0; k < Cfg::kKSlices; ++k) {
              umma_ss<Act<T>::kBytes>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                      (i | k) != 0 ? 1u : 0u);
              if constexpr (SPLIT_IN) {
                umma_ss<2>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(Cfg::kBBytes / 16 + k * 2), idesc, 1u);
                umma_ss<2>(d_tmem, adesc + (uint64_t)(kABytes / 16 + k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
              }
            }
            tc_commit(&empty_bar[s]);
          }
          tc_commit(&acc_full[as]);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float2* stats = reinterpret_cast<float2*>(smem + kStages * Cfg::kStageBytes + 256);
    uint8_t* stg = smem + kStages * Cfg::kStageBytes + 256 + L::kStatsBytes + (warp - 4) * (32 * 32 * (int)sizeof(T));
    uint32_t lt = 0;
    for (int tile = t_first; tile < p.m_tiles; tile += t_step, ++lt) {
      const uint32_t as = lt % kAcc;
      const uint32_t aph = (lt / kAcc) & 1;
      const uint32_t t_row = tmem_base + as * BLOCK_N + ((uint32_t)(q * 32) << 16);
      epilogue_tile<T, BLOCK_N, true, SPLIT, 128, 4>(p, stg, stats, &stats_full[as * 4 + q], &acc_full[as], t_row, q, lane, tile,
                                                      (int)cta_rank * BLOCK_N, as, aph, cta_rank, half * 128,
                                                      half * 128 + 128);
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2) for the wide plain GEMMs: a cluster of two CTAs computes a 256 x 256
// tile.  Each CTA stages its own 128 rows of A and HALF of the weight tile (128 of the 256 W rows);
// the leader's single tcgen05.mma (M = 256, N = 256) reads both halves, so every CTA moves
// 32 KB instead of 48 KB per K chunk from L2 (-33 % bytes per MAC) and the ring is 6 deep.
//   full barrier  : in the leader, count 2 (one arrive per CTA's producer) + 64 KB of TMA bytes
//   empty / acc_full : tcgen05.commit multicast to both CTAs
//   acc_empty     : in the leader, count 256 (both CTAs' epilogue threads)
// ---------------------------------------------------------------------------------------------
static constexpr int k2HalfB = 128 * 128;  // this CTA's half of one W plane: 128 rows x 128 B
static constexpr int k2Threads = 384;  // warpgroup 0: TMA / MMA (+2 idle warps); warpgroups 1, 2: epilogue
static constexpr int k2SmemBytes = 6 * (kABytes + k2HalfB) + 1024 + 256 + 8 * 4096;

// SPLIT: a stage holds [A_hi][A_lo][W_hi half][W_lo half] (64 KB, 3 stages) and every K slice issues three MMAs.
// SIMPLE: see epilogue_tile -- a separate kernel instantiation per addressing mode, chosen on the host (two epilogue
// copies inside one kernel were measured slower than either alone).
// FAST (bf16, not split): the bias + ReLU fast epilogue with TMA stores (epilogue_fast_tile).  Shared memory: 5 stages (160 KB)
// | 8 warps x 4 staging tiles (64 KB) | barriers | 2 warpgroups x 2 x 512 B bias slices.
static constexpr int k2FastStages = 5;
// (no alignment slack: the dynamic shared memory of this kernel is declared __align__(1024); checked at run time)
static constexpr int k2FastSmemBytes = k2FastStages * (kABytes + k2HalfB) + 8 * 8192 + 256 + 2048;
static_assert(k2FastSmemBytes <= 232448, "fast 2-CTA kernel exceeds the 227 KB shared-memory limit");

template <typename T, bool SPLIT, bool SIMPLE = false, bool FAST = false, bool ONEPASS = false>
__global__ void __launch_bounds__(k2Threads, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmKParams p) {
  static_assert(!FAST || (!SIMPLE && sizeof(T) == 2 && (!SPLIT || ONEPASS)), "FAST epilogue: plain bf16, or the one-pass split variant");
  static_assert(!ONEPASS || SPLIT, "ONEPASS: the one-pass variant of the split (two-plane) configuration");
  constexpr bool SPLIT_IN = SPLIT && !ONEPASS;   // stages hold both planes, three MMAs per K slice
  constexpr int BLOCK_N = 256;
  constexpr int kPlanes = SPLIT_IN ? 2 : 1;
  constexpr int k2Stages = FAST ? k2FastStages : (SPLIT_IN ? 3 : 6);
  constexpr int k2StageBytes = kPlanes * (kABytes + k2HalfB);  // A (own 128 rows) + half of B, per plane
  constexpr int kAcc = 2;
  extern __shared__ __align__(1024) uint8_t smem2_raw[];
  uint8_t* smem = FAST ? smem2_raw
                       : reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem2_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  if constexpr (FAST) {
    if ((smem_u32(smem2_raw) & 1023u) != 0u) __trap();   // the layout below has no slack to re-align
  }
  // FAST: the staging tiles follow the stages (1024-byte aligned: they are TMA-store sources in the 64-byte swizzle pattern)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes + (FAST ? 8 * 8192 : 0));
  uint64_t* full_bar = bars;                 // [k2Stages]  (used in the leader)
  uint64_t* empty_bar = bars + k2Stages;     // [k2Stages]  (local in each CTA)
  uint64_t* acc_full = bars + 2 * k2Stages;  // [kAcc]      (local in each CTA)
  uint64_t* acc_empty = acc_full + kAcc;     // [kAcc]      (used in the leader)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + kAcc);
  uint64_t* res_bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bars) + 128);   // FAST: [8] one per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (FAST) {
      tma_prefetch_desc(&tmO);
      if (p.flags & AITB_EPI_RES) tma_prefetch_desc(&tmR);
      for (int s = 0; s < 8; ++s) mbar_init(&res_bars[s], 1);
    }
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 512);  // 2 CTAs x 8 epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_holder, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int n_tiles_total = m_pairs * p.n_tiles;
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);
  const int iters_per_tile = p.taps * p.k_chunks;

  // register re-balancing: the epilogue is instruction-bound (one accumulator row per thread, ~150
  // instructions per 32-column chunk), so it gets two warpgroups and most of the register file
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = t_first; tile < n_tiles_total; tile += t_step) {
        const int mp = tile / p.n_tiles;
        const int nt = tile - mp * p.n_tiles;
        const int mt = mp * 2 + (int)rank;
        for (int tap = 0; tap < p.taps; ++tap) {
          int c1 = p.tap_dx[tap], c2 = p.tap_dy[tap], c3 = 0;
          if (p.a_m_dim == 1) c1 += mt * p.a_m_step;
          else if (p.a_m_dim == 3) c3 += mt * p.a_m_step;
          else { c2 += (mt % p.map_tpg) * p.map_by; c3 += mt / p.map_tpg; }
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            const uint32_t s = it % k2Stages;
            const uint32_t ph = (it / k2Stages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * k2StageBytes;
            uint8_t* sb = sa + kPlanes * kABytes;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * k2StageBytes);
            else mbar_arrive_remote(&full_bar[s], 0);
            tma2_load_4d(sa, &tmA, &full_bar[s], kc * p.ke, c1, c2, c3);
            const int kb = (tap * p.k_chunks + kc) * p.ke;
            tma2_load_2d(sb, &tmB, &full_bar[s], kb, nt * BLOCK_N + (int)rank * 128);
            if constexpr (SPLIT_IN) {
              tma2_load_4d(sa + kABytes, &tmA, &full_bar[s], p.a_lo + kc * p.ke, c1, c2, c3);
              tma2_load_2d(sb + k2HalfB, &tmB, &full_bar[s], p.w_lo + kb, nt * BLOCK_N + (int)rank * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc((SPLIT && p.in_f16) ? 0u : Act<T>::kFmt, 256, BLOCK_N);
      uint32_t it = 0, lt = 0;
      for (int tile = t_first; tile < n_tiles_total; tile += t_step, ++lt) {
        const uint32_t as = lt % kAcc;
        const uint32_t aph = (lt / kAcc) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);  // plain (cta-scope) wait: a cluster-scope acquire costs ~500 clk per poll
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int i = 0; i < iters_per_tile; ++i, ++it) {
          const uint32_t s = it % k2Stages;
          const uint32_t ph = (it / k2Stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * k2StageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + kPlanes * kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma2_ss<Act<T>::kBytes>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                     (i | k) != 0 ? 1u : 0u);
            if constexpr (SPLIT_IN) {
              umma2_ss<2>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k2HalfB / 16 + k * 2), idesc, 1u);
              umma2_ss<2>(d_tmem, adesc + (uint64_t)(kABytes / 16 + k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            }
          }
          tc_commit2(&empty_bar[s]);
        }
        tc_commit2(&acc_full[as]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;            // TMEM lane quadrant
    const int half = (warp - 4) >> 2;  // which 128 columns of the 256-wide tile this warp drains
    uint8_t* stg = smem + k2Stages * k2StageBytes + (FAST ? (warp - 4) * 8192 : 256 + (warp - 4) * 4096);
    uint32_t lt = 0;
    for (int tile = t_first; tile < n_tiles_total; tile += t_step, ++lt) {
      const int mp = tile / p.n_tiles;
      const int nt = tile - mp * p.n_tiles;
      const uint32_t as = lt % kAcc;
      const uint32_t aph = (lt / kAcc) & 1;
      const uint32_t t_row = tmem_base + as * BLOCK_N + ((uint32_t)(q * 32) << 16);
      if constexpr (FAST) {
        const uint32_t bias_s = smem_u32(smem + k2Stages * k2StageBytes + 8 * 8192 + 256) + (uint32_t)half * 1024u;
        epilogue_fast_tile<SPLIT>(p, &tmO, smem_u32(stg), bias_s, 1 + half, &acc_full[as], &acc_empty[as], aph,
                           t_row + (uint32_t)(half * 128), lane, q * 32 + lane, (mp * 2 + (int)rank) * kBlockM + q * 32,
                           nt * BLOCK_N + half * 128, as, &tmR, &res_bars[warp - 4], lt & 1u);
      } else {
        epilogue_tile<T, BLOCK_N, false, SPLIT, 128, 2, SIMPLE>(p, stg, nullptr, nullptr, &acc_full[as], t_row, q, lane,
                                                                mp * 2 + (int)rank, nt * BLOCK_N, as, aph, rank, half * 128,
                                                                half * 128 + 128);
        tc_fence_before();
        mbar_arrive_remote(&acc_empty[as], 0);
      }
    }
    if constexpr (FAST) {
      if (lane == 0) bulk_wait_all();   // this warp's last stores have completed before the CTA may exit
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

static int encode_map(CUtensorMap* tm, int dtype, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const char* what, bool swizzle64 = false,
                      bool swizzle128_atom32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t gdims[4];
  cuuint64_t gstr[3];
  cuuint32_t gbox[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(tm, dtype == AITB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(ptr), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                    : (swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
              what, (int)r, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1;
  }
  return 0;
}

// bf16 2-D map, 128-byte swizzle (attention operand tiles)
int encode_map_bf16(CUtensorMap* tm, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* what) {
  return encode_map(tm, AITB_BF16, ptr, 2, dims, strides_bytes, box, what);
}

// fp32 map with the "128-byte swizzle, 32-byte atom" pattern (the only layout tcgen05 accepts for MN-major tf32 operands)
int encode_map_f32_mn(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const char* what) {
  return encode_map(tm, AITB_F32, ptr, rank, dims, strides_bytes, box, what, false, true);
}

static int num_sms() { return current_sm_count(); }

template <typename T, int BLOCK_N, bool CL, bool SPLIT>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, SPLIT, CL && SPLIT>;
  static SmemAttrOnce once;
  auto kern = gemm_tcgen05_kernel<T, BLOCK_N, CL, SPLIT>;
  if (ensure_dyn_smem((const void*)kern, Cfg::kSmemBytes, once, "gemm_tcgen05_kernel")) return 1;
  if constexpr (CL) {
    const int clusters = kp.m_tiles < num_sms() / 2 ? kp.m_tiles : num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, kp);
    if (e != cudaSuccess) {
      set_error("gemm_tcgen05_kernel<cluster LN>: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return 1;
    }
    return check_launch("gemm_tcgen05_kernel<cluster LN>");
  } else {
    const int tiles = kp.m_tiles * kp.n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, kp);
    return check_launch("gemm_tcgen05_kernel");
  }
}

template <typename T, bool SPLIT, bool SIMPLE, bool FAST = false, bool ONEPASS = false>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmKParams& kp,
                        cudaStream_t stream, const CUtensorMap* tmR = nullptr) {
  static SmemAttrOnce once;
  auto kern = gemm2_tcgen05_kernel<T, SPLIT, SIMPLE, FAST, ONEPASS>;
  constexpr int k2SmemBytes = FAST ? k2FastSmemBytes : aitb::k2SmemBytes;
  if (ensure_dyn_smem((const void*)kern, k2SmemBytes, once, "gemm2_tcgen05_kernel")) return 1;
  const int tiles = ((kp.m_tiles + 1) / 2) * kp.n_tiles;
  const int clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(k2Threads);
  cfg.dynamicSmemBytes = k2SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, tmR ? *tmR : tmA, kp);
  if (e != cudaSuccess) {
    set_error("gemm2_tcgen05_kernel: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  return check_launch("gemm2_tcgen05_kernel");
}

template <typename T, bool SPLIT, bool ONEPASS = false>
static int launch_gemm_ln2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp, cudaStream_t stream) {
  using L = LnCfg<T, SPLIT && !ONEPASS>;
  static SmemAttrOnce once;
  auto kern = gemm_ln2_tcgen05_kernel<T, SPLIT, ONEPASS>;
  if (ensure_dyn_smem((const void*)kern, L::kSmemBytes, once, "gemm_ln2_tcgen05_kernel")) return 1;
  const int clusters = kp.m_tiles < num_sms() / 2 ? kp.m_tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kLnThreads);
  cfg.dynamicSmemBytes = L::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, kp);
  if (e != cudaSuccess) {
    set_error("gemm_ln2_tcgen05_kernel: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  return check_launch("gemm_ln2_tcgen05_kernel");
}

int gemm_run(const aitb_gemm_desc* d, cudaStream_t stream) {
  AITB_REQUIRE(d != nullptr, "aitb_gemm: null descriptor");
  AITB_REQUIRE(d->dtype == AITB_F32 || d->dtype == AITB_BF16 || d->dtype == AITB_F32S, "aitb_gemm: bad dtype %d",
               d->dtype);
  const bool split = d->dtype == AITB_F32S;
  AITB_REQUIRE(d->passes == 0 || d->passes == 1 || d->passes == 3, "aitb_gemm: passes must be 0 (default), 1 or 3");
  AITB_REQUIRE(split || (d->passes != 1 && !d->in_f16 && !d->out_f16 && !d->res_f16),
               "aitb_gemm: passes / in_f16 / out_f16 / res_f16 belong to the split (AITB_F32S) configuration");
  const int eb = d->dtype == AITB_F32 ? 4 : 2;  // bytes per TMA element (split: bf16 planes)
  const int ke = 128 / eb;
  if (split) {
    AITB_REQUIRE(d->a_lo_off > 0 && d->a_lo_off % 8 == 0, "aitb_gemm: split mode needs a_lo_off (multiple of 8)");
    AITB_REQUIRE((d->flags & AITB_EPI_ACCUM) == 0, "aitb_gemm: ACCUM epilogue is not available in split mode");
  }
  AITB_REQUIRE(d->M > 0 && d->N > 0, "aitb_gemm: empty problem M=%d N=%d", d->M, d->N);
  AITB_REQUIRE(d->block_n == 128 || d->block_n == 256 || d->block_n == 512 || d->block_n == 64,
               "aitb_gemm: block_n %d unsupported", d->block_n);
  AITB_REQUIRE(d->N % d->block_n == 0, "aitb_gemm: N=%d not a multiple of block_n=%d", d->N, d->block_n);
  AITB_REQUIRE(d->k_per_tap % ke == 0 && d->k_per_tap > 0, "aitb_gemm: K per tap %d not a multiple of %d",
               d->k_per_tap, ke);
  AITB_REQUIRE(d->taps >= 1 && d->taps <= 9, "aitb_gemm: taps=%d", d->taps);
  AITB_REQUIRE(d->a.box[0] == (uint32_t)ke, "aitb_gemm: A box[0] must span 128 bytes");
  AITB_REQUIRE(d->a.box[1] * d->a.box[2] * d->a.box[3] == 128, "aitb_gemm: A box must cover 128 rows");
  AITB_REQUIRE(d->a_m_dim == 1 || d->a_m_dim == 2 || d->a_m_dim == 3, "aitb_gemm: a_m_dim must be 1, 2 or 3");
  if (d->a_m_dim == 2)
    AITB_REQUIRE(d->map_w > 0 && d->map_h > 0 && d->a.box[1] * d->a.box[2] == 128 && d->a.box[3] == 1 &&
                     d->map_w <= (int)d->a.box[1] && d->M % 128 == 0 && d->rows_in == d->rows_out,
                 "aitb_gemm: bad map-mode view (box x * y must be 128 positions of one image, width <= box x)");
  AITB_REQUIRE((d->flags & AITB_EPI_LN) == 0 || (d->block_n == 512 && d->N == 512),
               "aitb_gemm: the LayerNorm epilogue needs block_n == N == 512");
  AITB_REQUIRE((d->flags & AITB_EPI_LN) != 0 || d->block_n != 512, "aitb_gemm: block_n 512 is the LayerNorm variant");
  AITB_REQUIRE(d->rows_in > 0 && d->rows_out > 0, "aitb_gemm: bad row remap %d->%d", d->rows_in, d->rows_out);
  AITB_REQUIRE(d->ldo % 8 == 0, "aitb_gemm: ldo must be a multiple of 8 elements");
  AITB_REQUIRE(((uintptr_t)d->out & 31) == 0 && ((uintptr_t)d->a.ptr & 15) == 0 && ((uintptr_t)d->w & 15) == 0,
               "aitb_gemm: pointers must be 16/32-byte aligned");
  AITB_REQUIRE(d->out_scale == 0.f || d->out_scale == 1.f || (d->flags & AITB_EPI_LN) == 0 || split,
               "aitb_gemm: out_scale is not applied by the LayerNorm epilogue of the non-split configurations");
  AITB_REQUIRE((d->flags & AITB_EPI_HI_ONLY) == 0 || split, "aitb_gemm: HI_ONLY belongs to the split configuration");
  AITB_REQUIRE((d->flags & AITB_EPI_RELU_MASK) == 0 || (!split && (d->flags & (AITB_EPI_RES | AITB_EPI_LN)) == 0),
               "aitb_gemm: RELU_MASK excludes RES / LN and the split configuration");
  if (d->flags & (AITB_EPI_RES | AITB_EPI_RELU_MASK))
    AITB_REQUIRE(d->res != nullptr && d->res_div > 0 && d->res_rep > 0 && d->ldr % 8 == 0 &&
                     ((uintptr_t)d->res & 31) == 0,
                 "aitb_gemm: bad residual spec");
  if (d->flags & AITB_EPI_POS) AITB_REQUIRE(d->pos != nullptr && d->pos_rows > 0, "aitb_gemm: bad pos spec");
  if (d->flags & AITB_EPI_BIAS) AITB_REQUIRE(d->bias != nullptr, "aitb_gemm: bias flag without bias");
  if (d->flags & AITB_EPI_LN) AITB_REQUIRE(d->gamma && d->beta, "aitb_gemm: LN flag without gamma/beta");

  if (d->dual || (d->flags & AITB_EPI_DUAL))
    AITB_REQUIRE(d->dual && (d->flags & AITB_EPI_DUAL) && d->block_n == 128 && d->bias2 != nullptr &&
                     (d->flags & (AITB_EPI_LN | AITB_EPI_RES | AITB_EPI_ACCUM)) == 0,
                 "aitb_gemm: dual accumulator needs block_n 128, bias2 and the DUAL epilogue");
  const int w_taps = d->taps + (d->dual ? 1 : 0);
  const bool cluster_ln = (d->flags & AITB_EPI_LN) != 0;   // N = 512 split over a 2-CTA cluster
  // the split cluster-LayerNorm kernel stages 64-byte operand rows (32 bf16 of K per stage, 64-byte swizzle)
  static const bool ln_one_wg = getenv("AITB_LN_ONE_WG") != nullptr;   // A/B: the single-warpgroup LayerNorm epilogue
  const bool two_cta_shape = !cluster_ln && d->block_n == 256 && d->a_group_c == 0 && !d->dual && (d->M + kBlockM - 1) / kBlockM >= 2;
  // one MMA pass on the hi planes instead of three (the caller's precision plan): honoured by the 2-CTA kernel and the
  // two-warpgroup cluster-LayerNorm kernel -- the launches that matter; every other kernel runs the three passes
  static const bool no_2cta = getenv("AITB_NO_2CTA") != nullptr;
  const bool onepass = split && d->passes == 1 &&
                       ((two_cta_shape && !no_2cta) || (cluster_ln && !ln_one_wg && d->a_group_c == 0));
  const bool split_in = split && !onepass;
  const bool half_rows = cluster_ln && split_in;
  const int kes = half_rows ? ke / 2 : ke;  // K elements per pipeline stage
  CUtensorMap tmA, tmB;
  uint32_t abox[4] = {(uint32_t)kes, d->a.box[1], d->a.box[2], d->a.box[3]};
  if (encode_map(&tmA, d->dtype, d->a.ptr, 4, d->a.dims, d->a.strides, abox, "A", half_rows)) return 1;
  const uint64_t w_k = (uint64_t)w_taps * d->k_per_tap;  // logical K of the weight matrix
  const uint64_t wdims[2] = {w_k * (split ? 2 : 1), (uint64_t)d->N};
  const uint64_t wstr[1] = {w_k * (split ? 4 : eb)};
  const bool two_cta = two_cta_shape && !no_2cta;
  const uint32_t wbox[2] = {(uint32_t)kes, (uint32_t)(two_cta ? 128 : (d->block_n > 256 ? 256 : d->block_n))};
  if (encode_map(&tmB, d->dtype, d->w, 2, wdims, wstr, wbox, "W", half_rows)) return 1;
  const int pl = split ? 2 : 1;  // physical elements per logical element in out / res rows

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = d->M;
  kp.N = d->N;
  kp.k_chunks = d->k_per_tap / kes;
  kp.taps = d->taps;
  kp.a_m_dim = d->a_m_dim;
  kp.a_m_step = d->a_m_step;
  if (d->a_m_dim == 2) {
    kp.map_bx = (int)d->a.box[1];
    kp.map_by = (int)d->a.box[2];
    kp.map_w = d->map_w;
    kp.map_h = d->map_h;
    kp.map_tpg = (d->map_h + kp.map_by - 1) / kp.map_by;
  }
  kp.a_group_c = d->a_group_c;
  // A/B switches of the epilogue variants: read once per process, not per launch
  static const bool bias_smem_on = getenv("AITB_BIAS_GLOBAL") == nullptr;
  static const bool simple_rows_on = getenv("AITB_NO_SIMPLE_ROWS") == nullptr;
  kp.bias_smem = bias_smem_on;
  kp.simple_rows_ok = simple_rows_on;
  kp.ke = kes;
  for (int i = 0; i < 9; ++i) {
    kp.tap_dx[i] = d->tap_dx[i];
    kp.tap_dy[i] = d->tap_dy[i];
  }
  kp.m_tiles = (d->M + kBlockM - 1) / kBlockM;
  kp.n_tiles = cluster_ln ? 2 : d->N / d->block_n;
  kp.flags = d->flags;
  kp.out = d->out;
  kp.ldo = d->ldo * pl;
  kp.o_lo = d->ldo;
  kp.r_lo = d->ldr;
  kp.a_lo = d->a_lo_off;
  kp.w_lo = (int)w_k;
  kp.rows_in = d->rows_in;
  kp.rows_out = d->rows_out;
  kp.bias = d->bias;
  kp.res = d->res;
  kp.ldr = d->ldr * pl;
  kp.res_div = d->res_div > 0 ? d->res_div : 1;
  kp.res_rep = d->res_rep > 0 ? d->res_rep : 1;
  kp.pos = d->pos;
  kp.pos_rows = d->pos_rows > 0 ? d->pos_rows : 1;
  kp.gamma = d->gamma;
  kp.beta = d->beta;
  kp.eps = d->eps;
  kp.round_tf32 = d->round_tf32;
  kp.dual = d->dual ? 1 : 0;
  kp.bias2 = d->bias2;
  kp.ln_rstd = d->ln_rstd;
  kp.in_f16 = d->in_f16 ? 1 : 0;
  kp.out_f16 = d->out_f16 ? 1 : 0;
  kp.res_f16 = d->res_f16 ? 1 : 0;
  // tcgen05.mma adds each K=16 slice into the fp32 accumulator with round-toward-ZERO (measured, tools/acc_bias.py:
  // mean signed error -1.6e-8 .. -2.0e-8 of the result per accumulate step, growing linearly with the step
  // count; the expectation for RZ on a growing partial sum is 0.18 * 2^-23 = 2.1e-8).  In the tf32 / bf16
  // configurations that bias hides below the operand rounding; in the split configuration it would be the
  // largest error left (1.5e-5 at K = 4096), so the epilogue scales the accumulator by 1 + c * steps.
  {
    static float comp = -1.f;
    if (comp < 0.f) {
      const char* e = getenv("AITB_ACC_COMP");
      comp = e ? (float)atof(e) : 1.5e-8f;
    }
    const float np = onepass ? 1.f : 3.f;
    const float steps = split ? np * (float)(d->taps * d->k_per_tap / 16) : 0.f;
    const float steps2 = split ? np * (float)(d->k_per_tap / 16) : 0.f;
    const float user_scale = d->out_scale != 0.f ? d->out_scale : 1.f;   // 0 = unset
    kp.acc_scale = (1.f + comp * steps) * user_scale;
    kp.acc_scale2 = (1.f + comp * steps2) * user_scale;
  }

#define AITB_DISPATCH(BN)                                                                          \
  (d->dtype == AITB_F32 ? launch_gemm<float, BN, false, false>(tmA, tmB, kp, stream)               \
   : split              ? launch_gemm<__nv_bfloat16, BN, false, true>(tmA, tmB, kp, stream)        \
                        : launch_gemm<__nv_bfloat16, BN, false, false>(tmA, tmB, kp, stream))
  if (cluster_ln && !ln_one_wg && d->a_group_c == 0)
    return d->dtype == AITB_F32 ? launch_gemm_ln2<float, false>(tmA, tmB, kp, stream)
           : onepass            ? launch_gemm_ln2<__nv_bfloat16, true, true>(tmA, tmB, kp, stream)
           : split              ? launch_gemm_ln2<__nv_bfloat16, true>(tmA, tmB, kp, stream)
                                : launch_gemm_ln2<__nv_bfloat16, false>(tmA, tmB, kp, stream);
  if (cluster_ln)
    return d->dtype == AITB_F32 ? launch_gemm<float, 256, true, false>(tmA, tmB, kp, stream)
           : split              ? launch_gemm<__nv_bfloat16, 256, true, true>(tmA, tmB, kp, stream)
                                : launch_gemm<__nv_bfloat16, 256, true, false>(tmA, tmB, kp, stream);
  if (two_cta)
  {
    // output rows = GEMM rows (no regrouping, no large-map mode): the strided-row epilogue instantiation -- for 4-byte
    // outputs only.  Measured on the FFN w_1 shape (ncu, same box): fp32 storage 920.8 k -> 855.5 k cycles (eight live
    // 64-bit row pointers become one), bf16 435.3 k -> 468.7 k (four pointers: the array form is faster there).
    const bool simple = kp.a_m_dim != 2 && kp.rows_in == kp.rows_out && kp.simple_rows_ok && d->dtype == AITB_F32;
    if (simple) return launch_gemm2<float, false, true>(tmA, tmB, tmA, kp, stream);
    // bf16 outputs, rows = GEMM rows, epilogue at most bias + ReLU: the TMA-store fast epilogue (A/B: AITB_NO_FAST_EPI=1)
    static const bool fast_on = getenv("AITB_NO_FAST_EPI") == nullptr;
    const bool fast_base = fast_on && kp.a_m_dim != 2 && kp.rows_in == kp.rows_out &&
                           (((uintptr_t)d->out) & 15) == 0 && ((size_t)d->ldo * 2) % 16 == 0;
    const bool fast_shape = fast_base && (d->flags & ~(AITB_EPI_BIAS | AITB_EPI_RELU | AITB_EPI_HI_ONLY)) == 0;
    // plain bf16 also takes a (row-for-row) residual with the optional ReLU after it: the layer-4 conv3 launches
    const bool fast_res = fast_base && d->dtype == AITB_BF16 && (d->flags & AITB_EPI_RES) != 0 &&
                          (d->flags & ~(AITB_EPI_BIAS | AITB_EPI_RELU | AITB_EPI_RES | AITB_EPI_RES_RELU)) == 0 &&
                          kp.res_div == 1 && kp.res_rep == 1 && (((uintptr_t)d->res) & 15) == 0 && ((size_t)d->ldr * 2) % 16 == 0;
    if ((fast_shape || fast_res) && d->dtype == AITB_BF16 && kp.acc_scale == 1.f) {
      CUtensorMap tmO, tmR;
      const uint64_t odims[2] = {(uint64_t)d->N, (uint64_t)d->M};
      const uint64_t ostr[1] = {(uint64_t)d->ldo * 2};
      const uint32_t obox[2] = {32u, 32u};
      if (encode_map(&tmO, AITB_BF16, d->out, 2, odims, ostr, obox, "O", true)) return 1;
      if (fast_res) {
        const uint64_t rstr[1] = {(uint64_t)d->ldr * 2};
        if (encode_map(&tmR, AITB_BF16, d->res, 2, odims, rstr, obox, "R", true)) return 1;
      }
      return launch_gemm2<__nv_bfloat16, false, false, true>(tmA, tmB, tmO, kp, stream, fast_res ? &tmR : nullptr);
    }
    if (fast_shape && onepass) {   // two 16-bit planes per row: hi at column 0, lo at column ldo
      CUtensorMap tmO;
      const uint64_t odims[2] = {(uint64_t)d->ldo + (uint64_t)d->N, (uint64_t)d->M};
      const uint64_t ostr[1] = {(uint64_t)d->ldo * 4};
      const uint32_t obox[2] = {32u, 32u};
      if (encode_map(&tmO, AITB_BF16, d->out, 2, odims, ostr, obox, "O", true)) return 1;
      return launch_gemm2<__nv_bfloat16, true, false, true, true>(tmA, tmB, tmO, kp, stream);
    }
    return d->dtype == AITB_F32 ? launch_gemm2<float, false, false>(tmA, tmB, tmA, kp, stream)
           : onepass            ? launch_gemm2<__nv_bfloat16, true, false, false, true>(tmA, tmB, tmA, kp, stream)
           : split              ? launch_gemm2<__nv_bfloat16, true, false>(tmA, tmB, tmA, kp, stream)
                                : launch_gemm2<__nv_bfloat16, false, false>(tmA, tmB, tmA, kp, stream);
  }
  switch (d->block_n) {
    case 64: return AITB_DISPATCH(64);
    case 128: return AITB_DISPATCH(128);
    case 256: return AITB_DISPATCH(256);
  }
#undef AITB_DISPATCH
  set_error("aitb_gemm: unreachable");
  return 1;
}

}  // namespace aitb
