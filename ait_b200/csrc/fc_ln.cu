// Attention output projection + residual + LayerNorm (lib/model/system/SubLayers.py:97-100:
// `q = self.layer_norm(self.fc(q) + residual)`, fc = Linear(64 -> 512, no bias)) as ONE memory-bound kernel.
//
// Why not the tcgen05 GEMM: with K = 64 the contraction is 0.4 % of the GEMM kernel's time; what is left is its
// LayerNorm epilogue -- one CTA per SM, eight epilogue warps, each walking one accumulator tile through
// TMEM -> registers -> DSMEM statistics exchange -> staging -> global, a serial dependency chain per tile
// (measured 270 us per launch against 100 us of HBM time, profiles/r01b_*).  This kernel instead treats the op as
// what it is, a streaming row operation:
//   * 16 warps per SM; a group of 4 warps owns 16 token rows, each warp one quarter (128) of the 512 features;
//   * the 64-deep contraction runs on mma.sync.m16n8k16 (bf16, fp32 accumulate; three passes hi*hi + hi*lo + lo*hi
//     in the split fp32-class configuration) with W_fc resident in shared memory in FRAGMENT ORDER (one conflict-free
//     LDS.128 feeds two k-steps);
//   * the MMA column <-> feature map is permuted so that every lane owns 8 CONTIGUOUS features per 32-feature
//     sub-block: residual loads and output stores are 128-bit, 64 B per row per warp instruction (256-bit LDG/STG.E.256
//     with 16 features per lane was measured slower: 200 vs 180 us, register spills);
//   * the un-normalised row never leaves registers; row statistics (two-pass: mean, then centred M2) are combined
//     across the four warps of a group through shared memory and a 128-thread named barrier.
// Same arithmetic as the GEMM path (bf16 operands, fp32 accumulation, fp32 LayerNorm), same storage formats.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kFcThreads = 512;
static constexpr int kFcN = 512, kFcK = 64;

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void unpack8(const uint4& x, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __bfloat1622float2(h[e]);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}
__device__ __forceinline__ void unpack8f(const uint4& x, float (&f)[8], bool f16) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = unpack_plane2(w[e], f16);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 x;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return x;
}

struct FcLnParams {
  const __nv_bfloat16* a;    // [M, PL*64]   (split: hi 64 | lo 64)
  const __nv_bfloat16* w;    // [512, PL*64]
  const __nv_bfloat16* res;  // [*, PL*512]
  __nv_bfloat16* out;        // [*, PL*512]
  const float* gamma;
  const float* beta;
  float eps;
  int M, rows_in, rows_out, res_row_m, res_div, res_rep;
  int res_f16, out_f16;   // split mode: plane element format of the residual read / the output written (0 = bf16, 1 = fp16)
};

// dynamic smem: Wf [64 n-tiles][2 k-pairs][PL][32 lanes] uint4, then gamma[512], beta[512], st1[4][4][16], st2[4][4][16],
// then the A staging tiles As[2 buffers][4 groups][16 rows][PL * 128 B] (16-byte chunks XOR-swizzled by row & 7)
template <bool SPLIT>
__global__ void __launch_bounds__(kFcThreads, 1) fc_ln_kernel(const FcLnParams p) {
  constexpr int PL = SPLIT ? 2 : 1;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* Wf = reinterpret_cast<uint4*>(smem);
  float* s_gamma = reinterpret_cast<float*>(smem + (size_t)64 * 2 * PL * 32 * 16);
  float* s_beta = s_gamma + kFcN;
  float* st1 = s_beta + kFcN;
  float* st2 = st1 + 4 * 4 * 16;
  uint8_t* As = reinterpret_cast<uint8_t*>(st2 + 4 * 4 * 16);
  constexpr int kARow = PL * 128;            // bytes per staged A row
  constexpr int kATile = 16 * kARow;         // one group's 16 rows

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int group = warp >> 2, q = warp & 3;

  // ---- W_fc into fragment order.  n-tile nt = (quarter*4 + sub-block)*4 + j; its MMA column n holds feature
  //      quarter*128 + sub*32 + 8*(n>>1) + 2*j + (n&1)
  for (int i = threadIdx.x; i < 64 * 2 * PL * 32; i += kFcThreads) {
    const int ln = i & 31, pl = (i >> 5) % PL, kp = ((i >> 5) / PL) & 1, nt = (i >> 5) / (2 * PL);
    const int gg = ln >> 2, tt = ln & 3;
    const int feat = (nt >> 2) * 32 + 8 * (gg >> 1) + 2 * (nt & 3) + (gg & 1);
    const uint32_t* wr = reinterpret_cast<const uint32_t*>(p.w + (size_t)feat * (PL * kFcK) + pl * kFcK);
    uint4 v;
    v.x = wr[(2 * kp) * 8 + tt];         // k = (2kp)*16 + 2t
    v.y = wr[(2 * kp) * 8 + 4 + tt];     //     + 8
    v.z = wr[(2 * kp + 1) * 8 + tt];
    v.w = wr[(2 * kp + 1) * 8 + 4 + tt];
    Wf[i] = v;
  }
  for (int i = threadIdx.x; i < kFcN; i += kFcThreads) {
    s_gamma[i] = p.gamma[i];
    s_beta[i] = p.beta[i];
  }
  __syncthreads();

  const int n_blk = (p.M + 63) / 64;
  // this group's 16 A rows of block `b` -> staging buffer `buf`: 16 * PL * 8 chunks of 16 bytes, 128 threads
  const int gthread = threadIdx.x & 127;
  auto stage_a = [&](int b, int buf) {
    uint8_t* tile = As + (buf * 4 + group) * kATile;
#pragma unroll
    for (int c = gthread; c < 16 * PL * 8; c += 128) {
      const int row = c / (PL * 8), ch = c % (PL * 8);
      const int mm = b * 64 + group * 16 + row;
      const uint32_t dst = smem_u32(tile + row * kARow + (ch >> 3) * 128 + (((ch & 7) ^ (row & 7)) << 4));
      const __nv_bfloat16* src = p.a + (size_t)(mm < p.M ? mm : 0) * (PL * kFcK) + ch * 8;
      const int bytes = mm < p.M ? 16 : 0;   // rows past M are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
  };
  if ((int)blockIdx.x < n_blk) stage_a(blockIdx.x, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int it = 0;
  for (int blk = blockIdx.x; blk < n_blk; blk += gridDim.x, ++it) {
    const int row0 = blk * 64 + group * 16;
    // the two token rows of this lane (MMA rows g and g + 8)
    int m[2] = {row0 + g, row0 + g + 8};
    bool ok[2];
    size_t orow[2], rrow[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int mm = m[h];
      const int r_in = mm % p.rows_in;
      ok[h] = mm < p.M && r_in < p.rows_out;
      const int o = (mm / p.rows_in) * p.rows_out + r_in;
      orow[h] = (size_t)(ok[h] ? o : 0);
      const int rb = p.res_row_m ? (ok[h] ? mm : 0) : (int)orow[h];
      rrow[h] = (size_t)(((rb / p.res_div) / p.res_rep) * p.res_div + (rb % p.res_div));
    }
    // ---- A fragments from the staged tile (filled by cp.async one iteration ahead):
    //      a[pl][ks] = {A[g][2t..], A[g+8][2t..], A[g][2t+8..], A[g+8][2t+8..]} at k = ks*16  (one ldmatrix.x4)
    {
      const int nb = blk + gridDim.x;
      if (nb < n_blk) stage_a(nb, (it + 1) & 1);
      // (tried in round 2: prefetch.global.L2 of the next block's residual rows by the t == 0 lanes -- measured 3.5 % SLOWER,
      //  177 vs 171 us: the kernel is bandwidth-, not latency-limited on those loads; dropped)
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
    }
    uint32_t a[PL][4][4];
    {
      const uint8_t* tile = As + ((it & 1) * 4 + group) * kATile;
      const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);     // matrices 0 / 2: rows 0-7, 1 / 3: rows 8-15
      const int kc = lane >> 4;                                // matrices 2, 3: k + 8  (the next 16-byte chunk)
#pragma unroll
      for (int pl = 0; pl < PL; ++pl)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t addr = smem_u32(tile + lrow * kARow + pl * 128 + (((2 * ks + kc) ^ (lrow & 7)) << 4));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a[pl][ks][0]), "=r"(a[pl][ks][1]), "=r"(a[pl][ks][2]), "=r"(a[pl][ks][3]) : "r"(addr));
        }
    }
    float acc[4][4][4];
#pragma unroll
    for (int sb = 0; sb < 4; ++sb)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[sb][j][e] = 0.f;

#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
      const int col = q * 128 + sb * 32 + 8 * t;   // this lane's 8 contiguous features of the sub-block
      uint4 r[2][PL];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int pl = 0; pl < PL; ++pl) {
          r[h][pl] = make_uint4(0u, 0u, 0u, 0u);
          if (ok[h]) r[h][pl] = *reinterpret_cast<const uint4*>(p.res + rrow[h] * (PL * kFcN) + pl * kFcN + col);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nt = (q * 4 + sb) * 4 + j;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          const uint4 wh = Wf[((nt * 2 + kp) * PL + 0) * 32 + lane];
          mma_bf16(acc[sb][j], a[0][2 * kp], wh.x, wh.y);
          mma_bf16(acc[sb][j], a[0][2 * kp + 1], wh.z, wh.w);
          if constexpr (SPLIT) {
            const uint4 wl = Wf[((nt * 2 + kp) * PL + 1) * 32 + lane];
            mma_bf16(acc[sb][j], a[0][2 * kp], wl.x, wl.y);
            mma_bf16(acc[sb][j], a[0][2 * kp + 1], wl.z, wl.w);
            mma_bf16(acc[sb][j], a[1][2 * kp], wh.x, wh.y);
            mma_bf16(acc[sb][j], a[1][2 * kp + 1], wh.z, wh.w);
          }
        }
      }
      // + residual: features col + 2j + e  <->  acc[sb][j][e] (row g), acc[sb][j][2 + e] (row g + 8)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[8];
        unpack8f(r[h][0], f, SPLIT && p.res_f16 != 0);
        if constexpr (SPLIT) {
          float f2[8];
          unpack8f(r[h][1], f2, p.res_f16 != 0);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += f2[e];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[sb][j][2 * h] += f[2 * j];
          acc[sb][j][2 * h + 1] += f[2 * j + 1];
        }
      }
    }

    // ---- LayerNorm statistics over the 512 features of rows g / g + 8: lane -> 4 lanes (t) -> 4 warps (quarters)
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int sb = 0; sb < 4; ++sb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[0] += acc[sb][j][0] + acc[sb][j][1];
        s[1] += acc[sb][j][2] + acc[sb][j][3];
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 1);
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 2);
    }
    if (t == 0) {
      st1[(group * 4 + q) * 16 + g] = s[0];
      st1[(group * 4 + q) * 16 + g + 8] = s[1];
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
    float mean[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float tot = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) tot += st1[(group * 4 + qq) * 16 + g + 8 * h];
      mean[h] = tot * (1.f / kFcN);
    }
    float v2[2] = {0.f, 0.f};
#pragma unroll
    for (int sb = 0; sb < 4; ++sb)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = acc[sb][j][e] - mean[e >> 1];
          v2[e >> 1] += d * d;
        }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      v2[h] += __shfl_xor_sync(0xffffffffu, v2[h], 1);
      v2[h] += __shfl_xor_sync(0xffffffffu, v2[h], 2);
    }
    if (t == 0) {
      st2[(group * 4 + q) * 16 + g] = v2[0];
      st2[(group * 4 + q) * 16 + g + 8] = v2[1];
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
    float rstd[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float tot = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) tot += st2[(group * 4 + qq) * 16 + g + 8 * h];
      rstd[h] = rsqrtf(tot * (1.f / kFcN) + p.eps);
    }

    // ---- normalise, scale / shift, store (hi plane, and the bf16 remainder into the lo plane in split mode)
#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
      const int col = q * 128 + sb * 32 + 8 * t;
      const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + col), g1 = *reinterpret_cast<const float4*>(s_gamma + col + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_beta + col), b1 = *reinterpret_cast<const float4*>(s_beta + col + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            y[2 * j + e] = (acc[sb][j][2 * h + e] - mean[h]) * rstd[h] * gm[2 * j + e] + bt[2 * j + e];
        if (!ok[h]) continue;
        __nv_bfloat16* o = p.out + orow[h] * (PL * kFcN) + col;
        if constexpr (SPLIT) {
          const bool of16 = p.out_f16 != 0;
          float lo[8];
          uint4 hi, lw;
          hi.x = split_hi2(y[0], y[1], of16, lo[0], lo[1]);
          hi.y = split_hi2(y[2], y[3], of16, lo[2], lo[3]);
          hi.z = split_hi2(y[4], y[5], of16, lo[4], lo[5]);
          hi.w = split_hi2(y[6], y[7], of16, lo[6], lo[7]);
          lw.x = pack_plane2(lo[0], lo[1], of16); lw.y = pack_plane2(lo[2], lo[3], of16);
          lw.z = pack_plane2(lo[4], lo[5], of16); lw.w = pack_plane2(lo[6], lo[7], of16);
          *reinterpret_cast<uint4*>(o) = hi;
          *reinterpret_cast<uint4*>(o + kFcN) = lw;
        } else {
          *reinterpret_cast<uint4*>(o) = pack8(y);
        }
      }
    }
  }
}

static SmemAttrOnce g_fc_once[2];

// dtype AITB_BF16 or AITB_F32S; a [M, 64] (planes), w [512, 64] (planes), res / out rows of 512 (planes)
int fc_ln_run(int dtype, const void* a, const void* w, const void* res, const float* gamma, const float* beta, float eps,
              void* out, int M, int rows_in, int rows_out, int res_row_m, int res_div, int res_rep, cudaStream_t st,
              int res_f16, int out_f16) {
  AITB_REQUIRE(dtype == AITB_BF16 || dtype == AITB_F32S, "fc_ln: bf16 / split storage only");
  AITB_REQUIRE(a && w && res && gamma && beta && out && M > 0 && rows_in > 0 && rows_out > 0 && rows_out <= rows_in &&
                   res_div > 0 && res_rep > 0, "fc_ln: bad arguments");
  AITB_REQUIRE(((uintptr_t)w & 3) == 0 && (((uintptr_t)a | (uintptr_t)res | (uintptr_t)out) & 15) == 0, "fc_ln: misaligned pointers");
  const bool split = dtype == AITB_F32S;
  const int pl = split ? 2 : 1;
  const int smem = 64 * 2 * pl * 32 * 16 + 2 * kFcN * 4 + 2 * 4 * 4 * 16 * 4 + 2 * 4 * 16 * pl * 128;
  FcLnParams p;
  p.a = reinterpret_cast<const __nv_bfloat16*>(a);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.M = M;
  p.rows_in = rows_in;
  p.rows_out = rows_out;
  p.res_row_m = res_row_m;
  p.res_div = res_div;
  p.res_rep = res_rep;
  p.res_f16 = split ? res_f16 : 0;
  p.out_f16 = split ? out_f16 : 0;
  const int n_blk = (M + 63) / 64;
  const int grid = n_blk < current_sm_count() ? n_blk : current_sm_count();
  if (split) {
    if (ensure_dyn_smem((const void*)fc_ln_kernel<true>, smem, g_fc_once[0], "fc_ln_kernel<split>")) return 1;
    fc_ln_kernel<true><<<grid, kFcThreads, smem, st>>>(p);
  } else {
    if (ensure_dyn_smem((const void*)fc_ln_kernel<false>, smem, g_fc_once[1], "fc_ln_kernel<bf16>")) return 1;
    fc_ln_kernel<false><<<grid, kFcThreads, smem, st>>>(p);
  }
  return check_launch("fc_ln_kernel");
}

}  // namespace aitb

extern "C" int aitb_fc_ln(int dtype, const void* a, const void* w_fc, const void* res, const float* gamma, const float* beta,
                          float eps, void* out, int M, int rows_in, int rows_out, int res_row_m, int res_div, int res_rep,
                          aitb_stream_t stream) {
  return aitb::fc_ln_run(dtype, a, w_fc, res, gamma, beta, eps, out, M, rows_in, rows_out, res_row_m, res_div, res_rep,
                         (cudaStream_t)stream, 0, 0);
}
