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
// fp32 storage, tf32 tensor-core math, fp32 accumulation.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

int encode_map_f32_mn(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const char* what);

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
static constexpr int kWgStages = 4;
static constexpr int kWgRows = 32;                      // contraction rows per pipeline stage (4 MMAs of K = 8)
static constexpr int kWgABytes = 4 * kWgRows * 128;     // 4 column groups of 32 fp32 (MMA M = 128)
static constexpr int kWgThreads = 192;

struct WgradParams {
  int M;          // rows (contraction length)
  int N, K;       // dW is [N, K]
  int bn;         // K-columns per tile (MMA N): 64 | 128 | 256
  int n_tiles, k_tiles, splits;
  int rows_per_split;   // multiple of kWgRows
  float* dw;
  int ldw;
};

// MN-major tf32 operand.  tcgen05 accepts exactly one shared-memory layout for it: "128-byte swizzle with a
// 32-byte atom" (layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- rows of 128 contiguous bytes
// along MN (32 fp32), FOUR contraction rows per 512-byte atom, the 32-byte chunks of a row XOR-ed with
// (row & 3).  Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units: atoms `lbo_bytes` apart along
// MN and 512 bytes apart along K, which is exactly what a TMA box of {32 fp32, R rows} leaves behind.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                     const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int b_bytes = (p.bn / 32) * kWgRows * 128;
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
        // one {32 cols, 32 rows} box per 128-byte column group (= one MN-major atom column); rows beyond M
        // are zero-filled by TMA
        const int r = row0 + it * kWgRows;
        for (int g = 0; g < 4; ++g) tma_load_2d(sa + g * (kWgRows * 128), &tmY, &full_bar[s], nt * 128 + g * 32, r);
        for (int g = 0; g < p.bn / 32; ++g)
          tma_load_2d(sa + kWgABytes + g * (kWgRows * 128), &tmX, &full_bar[s], kt * p.bn + g * 32, r);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      // tf32, fp32 accumulate, A and B both MN-major (bits 15 / 16)
      const uint32_t idesc = make_idesc(2u, 128u, (uint32_t)p.bn) | (1u << 15) | (1u << 16);
      for (int it = 0; it < iters; ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint64_t adesc = make_mnmajor_desc(sa, kWgRows * 128);
        const uint64_t bdesc = make_mnmajor_desc(sa + kWgABytes, kWgRows * 128);
#pragma unroll
        for (int k = 0; k < kWgRows / 8; ++k)  // 8 contraction rows (two 512-byte atoms) per K = 8 MMA
          umma_ss<4>(tmem_base, adesc + (uint64_t)(k * 64), bdesc + (uint64_t)(k * 64), idesc, (it | k) != 0 ? 1u : 0u);
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

static int g_sms_bwd = 0;
static int sms_bwd() {
  if (g_sms_bwd == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_bwd, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_bwd <= 0) g_sms_bwd = 148;
  }
  return g_sms_bwd;
}

// dw [N, ldw] += dy[M, ldy (cols n_off .. n_off + N)]^T * x[M, ldx (cols 0 .. K)]
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
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.bn = bn;
  p.n_tiles = N / 128;
  p.k_tiles = K / bn;
  const int tiles = p.n_tiles * p.k_tiles;
  const int chunks = (M + kWgRows - 1) / kWgRows;
  int splits = (2 * sms_bwd() + tiles - 1) / tiles;           // ~2 waves of CTAs
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  p.rows_per_split = ((chunks + splits - 1) / splits) * kWgRows;
  p.splits = (M + p.rows_per_split - 1) / p.rows_per_split;
  p.dw = dw;
  p.ldw = ldw;
  CUtensorMap tmY, tmX;
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldy * 4};
    const uint32_t box[2] = {32, (uint32_t)kWgRows};
    if (encode_map_f32_mn(&tmY, dy, 2, dims, str, box, "wgrad dY")) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldx * 4};
    const uint32_t box[2] = {32, (uint32_t)kWgRows};
    if (encode_map_f32_mn(&tmX, x, 2, dims, str, box, "wgrad X")) return 1;
  }
  const int smem = kWgStages * (kWgABytes + (bn / 32) * kWgRows * 128) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kWgStages * (kWgABytes + 8 * kWgRows * 128) + 1024 + 256);
    AITB_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(wgrad) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  wgrad_tcgen05_kernel<<<tiles * p.splits, kWgThreads, smem, stream>>>(tmY, tmX, p);
  return check_launch("wgrad_tcgen05_kernel");
}

}  // namespace aitb
