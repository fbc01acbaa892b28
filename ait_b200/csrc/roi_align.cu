// ROIAlign (legacy maskrcnn-benchmark flavour: no half-pixel shift, adaptive sampling grid)
// forward + backward, and the [G,C,S] <-> [G,S,C] layout kernels around it.
//
// Replaces `_C.roi_align_forward/backward`
// (lib/model/csrc/cuda/ROIAlign_cuda.cu:65-122 / 178-254, semantics of bilinear_interpolate
//  :16-62; CPU twin lib/model/csrc/cpu/ROIAlign_cpu.cpp:17-219).
//
// Design (HBM/L2-bound gather, not GEMM-shaped):
//   * the map is read from a channels-last copy, so each bilinear tap is ONE coalesced,
//     vectorised 128-bit load per lane (lanes = channels) instead of the reference's four
//     scattered 4-byte gathers per output element;
//   * sample positions/weights are channel independent and separable: each CTA folds the adaptive
//     sampling grid of every bin into per-axis cell-weight tables once, so a bin costs
//     (grid_h+1)(grid_w+1) vector loads instead of 4*grid_h*grid_w bilinear taps;
//   * a CTA owns (roi, channel slab); its feature window is small enough to stay in L1, so every
//     map byte is fetched from L2 about once per CTA;
//   * output goes either token-major [K, ph*pw, C] (feeds the enc_emb GEMM K-major, no NCHW
//     round trip) or NCHW [K, C, ph, pw] through a shared-memory transpose (the drop-in layout).
//   Sample coordinates use explicitly rounded fp32 ops in the reference's association order so
//   that floor/clamp decisions match the CPU reference bit for bit; only the accumulation order
//   of the (up to 4*grid) products differs (FMA), well inside the 1e-5 relative tolerance.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

static constexpr int kMaxGrid = 16;   // max adaptive samples per bin per axis held in smem tables
static constexpr int kMaxPooled = 8;  // pooled_h/w up to 8

struct Tap {   // one axis of one bilinear sample
  int lo, hi;  // clamped cell indices
  float wlo, whi;  // (1-l), l ; both 0 when the sample is outside [-1, size]
};

// One axis of bilinear_interpolate (ROIAlign_cuda.cu:22-52): validity, clamp, floor, weights.
__device__ __forceinline__ Tap make_tap(float c, int size) {
  Tap t;
  if (c < -1.0f || c > (float)size) {
    t.lo = t.hi = 0;
    t.wlo = t.whi = 0.f;
    return t;
  }
  if (c <= 0.f) c = 0.f;
  int lo = (int)c;
  int hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    c = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = __fsub_rn(c, (float)lo);
  t.lo = lo;
  t.hi = hi;
  t.whi = l;
  t.wlo = __fsub_rn(1.f, l);
  return t;
}

struct RoiGeom {
  int batch;
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
};

// ROIAlign_cuda.cu:78-100 (same expressions, same order, fp32)
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale, int ph, int pw,
                                               int sampling_ratio) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float end_w = __fmul_rn(roi[3], scale);
  const float end_h = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(end_w, g.start_w), 1.f);
  const float rh = fmaxf(__fsub_rn(end_h, g.start_h), 1.f);
  g.bin_h = __fdiv_rn(rh, (float)ph);
  g.bin_w = __fdiv_rn(rw, (float)pw);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)ph));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)pw));
  return g;
}

// y = roi_start + p*bin + (i + .5f)*bin/grid   (ROIAlign_cuda.cu:109,112)
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  const float a = __fadd_rn(start, __fmul_rn((float)p, bin));
  const float b = __fdiv_rn(__fmul_rn(__fadd_rn((float)i, .5f), bin), (float)grid);
  return __fadd_rn(a, b);
}

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  __device__ static __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  __device__ static __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    uint2 r;
    *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

// 8 bf16 channels (one 128-bit load) per lane: the bf16 map moves half the bytes of the fp32 one, so with 4 channels per
// lane it paid the same instruction count for half the data (measured slower than fp32: 351 vs 333 us)
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float4& a, float4& b) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.z));
  const float2 f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.w));
  a = make_float4(f0.x, f0.y, f1.x, f1.y);
  b = make_float4(f2.x, f2.y, f3.x, f3.y);
}

// acc (8 channels) += sum over ny rows x NX columns of wy[cy] * wx[cx] * feat[...]
template <int NX>
__device__ __forceinline__ void foot_rows8(float4& acc0, float4& acc1, const __nv_bfloat16* __restrict__ fb, uint32_t r, uint32_t rowC,
                                           uint32_t uC, int ny, const float* __restrict__ wy_tab, const float* __restrict__ wx_tab) {
  float wx[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) wx[j] = wx_tab[j];
#pragma unroll 1
  for (int cy = 0; cy < ny; ++cy, r += rowC) {
    const float wy = wy_tab[cy];
    float4 va[NX], vb[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) ld8_bf16(fb + r + (uint32_t)j * uC, va[j], vb[j]);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const float w = wy * wx[j];
      acc0.x += w * va[j].x; acc0.y += w * va[j].y; acc0.z += w * va[j].z; acc0.w += w * va[j].w;
      acc1.x += w * vb[j].x; acc1.y += w * vb[j].y; acc1.z += w * vb[j].z; acc1.w += w * vb[j].w;
    }
  }
}

// Per-axis "footprint" of one output bin: the bilinear samples of a bin touch a contiguous run of cells
// [c0, c0 + n) (sample spacing <= 1 cell), and because bilinear weights are separable the bin value is
//   (1/count) * sum_cy sum_cx Wy[cy] * Wx[cx] * feat[cy][cx],   Wy[c] = sum over the bin's y-samples of
// the weight they give to row c (same for x).  That is (gh+1)(gw+1) loads per bin instead of 4*gh*gw taps,
// the same real-number sum as ROIAlign_cuda.cu:105-118 re-associated (differences ~1e-7 relative).
struct Foot {
  int c0, n;
  float w[kMaxGrid + 1];
};

// returns false if the footprint does not fit the table (fixed sampling_ratio on a huge roi)
__device__ __forceinline__ bool build_foot(Foot& f, float start, int p, float bin, int grid, int size) {
  int c0 = 0, n = 0;
#pragma unroll 1
  for (int i = 0; i < kMaxGrid + 1; ++i) f.w[i] = 0.f;
#pragma unroll 1
  for (int i = 0; i < grid; ++i) {
    const Tap t = make_tap(sample_coord(start, p, bin, i, grid), size);
    if (t.wlo == 0.f && t.whi == 0.f) continue;  // outside [-1, size]: contributes 0 (but counts in `count`)
    if (n == 0) { c0 = t.lo; n = 1; }
    // samples are visited in increasing coordinate order, so lo/hi never fall below c0
    if (t.hi - c0 > kMaxGrid) return false;
    f.w[t.lo - c0] += t.wlo;
    f.w[t.hi - c0] += t.whi;
    n = max(n, t.hi - c0 + 1);
  }
  f.c0 = c0;
  f.n = n;
  return true;
}

// acc += sum over ny rows x NX columns of wy[cy] * wx[cx] * feat[r + cy*rowC + cx*uC]  (this lane's 4 channels)
template <typename T, int NX>
__device__ __forceinline__ void foot_rows(float4& acc, const T* __restrict__ fb, uint32_t r, uint32_t rowC, uint32_t uC, int ny,
                                          const float* __restrict__ wy_tab, const float* __restrict__ wx_tab) {
  float wx[NX];
#pragma unroll
  for (int j = 0; j < NX; ++j) wx[j] = wx_tab[j];
#pragma unroll 1
  for (int cy = 0; cy < ny; ++cy, r += rowC) {
    const float wy = wy_tab[cy];
    float4 v[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) v[j] = Vec4<T>::ld(fb + r + (uint32_t)j * uC);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const float w = wy * wx[j];
      acc.x += w * v[j].x; acc.y += w * v[j].y; acc.z += w * v[j].z; acc.w += w * v[j].w;
    }
  }
}

// grid (ceil(C / 128), K); block 256 = 8 warps; lane owns 4 consecutive channels of the 128-channel slab,
// warps stride over the ph*pw bins.
// OUT_SPLIT (token-major only): fp32 map in, output as two bf16 planes [K, ph*pw, hi C | lo C] (AITB_F32S)
#ifndef AITB_ROI_MINB
#define AITB_ROI_MINB 5
#endif
template <typename T, bool NCHW_OUT, bool OUT_SPLIT = false>
__global__ void __launch_bounds__(256, AITB_ROI_MINB)
roi_align_fwd_kernel(const T* __restrict__ feat, const float* __restrict__ rois, int C, int H, int W, float scale,
                     int ph, int pw, int sampling_ratio, typename std::conditional<OUT_SPLIT, __nv_bfloat16, T>::type* __restrict__ out,
                     int round_tf, int out_f16) {
  __shared__ Foot xf[kMaxPooled];
  __shared__ Foot yf[kMaxPooled];
  extern __shared__ float stage[];  // NCHW_OUT: [128][ph*pw + 1]
  const int k = blockIdx.y;
  const int c0 = blockIdx.x * 128;
  const RoiGeom g = roi_geometry(rois + (size_t)k * 5, scale, ph, pw, sampling_ratio);
  const int gw = g.grid_w, gh = g.grid_h;
  // rois far larger than the map (unclipped inputs) overflow the tables: per-sample taps inline then
  bool fits = true;
  if ((int)threadIdx.x < pw) fits = build_foot(xf[threadIdx.x], g.start_w, threadIdx.x, g.bin_w, gw, W);
  else if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + ph)
    fits = build_foot(yf[threadIdx.x - 32], g.start_h, threadIdx.x - 32, g.bin_h, gh, H);
  const bool tables = __syncthreads_and(fits) != 0;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvalid = min(128, C - c0);          // last slab may be partial (C % 4 == 0)
  const bool lane_on = lane * 4 < cvalid;
  const T* fb = feat + (size_t)g.batch * H * W * C + c0 + (lane_on ? lane * 4 : 0);
  // output_val /= count (:118) as a multiply by 1/count: <= 1 ulp from the division, and the four IEEE divisions per
  // bin were 14 % of the kernel (XU pipe); ncu: 41 executed instructions per tap with a 9-instruction inner loop, so
  // the per-bin bookkeeping is kept division-free and 32-bit (the map of one image is < 2^31 elements, checked on the host)
  const float inv_count = __frcp_rn((float)(g.grid_h * g.grid_w));
  const int nbins = ph * pw;
  const uint32_t uC = (uint32_t)C, rowC = (uint32_t)W * (uint32_t)C;
  int py = warp / pw, px = warp - py * pw;   // bin = warp, then += 8 without dividing again (pw <= 8)
  for (int bin = warp; bin < nbins; bin += 8) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tables) {
      const int ny = yf[py].n, nx = xf[px].n;
      const uint32_t r = (uint32_t)yf[py].c0 * rowC + (uint32_t)xf[px].c0 * uC;
      // footprints are 2-7 cells wide at RPN proposal sizes: one fully unrolled row body per width (x weights in
      // registers, the row's loads issued back to back) instead of a 3-4 trip inner loop -- 26 -> ~8 instructions per tap
      switch (nx) {
        case 1: foot_rows<T, 1>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 2: foot_rows<T, 2>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 3: foot_rows<T, 3>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 4: foot_rows<T, 4>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 5: foot_rows<T, 5>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 6: foot_rows<T, 6>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 7: foot_rows<T, 7>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 8: foot_rows<T, 8>(acc, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        default: {
          uint32_t rr = r;
          for (int cy = 0; cy < ny; ++cy, rr += rowC) {
            const float wy = yf[py].w[cy];
            uint32_t q = rr;
            for (int cx = 0; cx < nx; ++cx, q += uC) {
              const float w = wy * xf[px].w[cx];
              const float4 v = Vec4<T>::ld(fb + q);
              acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
            }
          }
        }
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        const Tap ty = make_tap(sample_coord(g.start_h, py, g.bin_h, iy, gh), H);
        const T* r0 = fb + (size_t)ty.lo * W * C;
        const T* r1 = fb + (size_t)ty.hi * W * C;
        for (int ix = 0; ix < gw; ++ix) {
          const Tap tx = make_tap(sample_coord(g.start_w, px, g.bin_w, ix, gw), W);
          const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
          const float4 v1 = Vec4<T>::ld(r0 + (size_t)tx.lo * C);
          const float4 v2 = Vec4<T>::ld(r0 + (size_t)tx.hi * C);
          const float4 v3 = Vec4<T>::ld(r1 + (size_t)tx.lo * C);
          const float4 v4 = Vec4<T>::ld(r1 + (size_t)tx.hi * C);
          acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
          acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
          acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
          acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
        }
      }
    }
    acc.x *= inv_count; acc.y *= inv_count; acc.z *= inv_count; acc.w *= inv_count;
    px += 8;                                   // advance (py, px) to bin + 8
    while (px >= pw) { px -= pw; ++py; }
    if (!lane_on) continue;
    if (sizeof(T) == 4 && round_tf) {  // engine-internal: the consumer is a tf32 MMA (RN beats HW truncation)
      acc.x = rn_tf32(acc.x); acc.y = rn_tf32(acc.y); acc.z = rn_tf32(acc.z); acc.w = rn_tf32(acc.w);
    }
    if constexpr (NCHW_OUT) {
      const int s = nbins + 1;
      stage[(lane * 4 + 0) * s + bin] = acc.x;
      stage[(lane * 4 + 1) * s + bin] = acc.y;
      stage[(lane * 4 + 2) * s + bin] = acc.z;
      stage[(lane * 4 + 3) * s + bin] = acc.w;
    } else if constexpr (OUT_SPLIT) {
      // hi = bf16(x), lo = bf16(x - hi) with the packed conversion (F2FP.PACK_AB, FMA-class pipe) and the hi values
      // recovered by shifts: the scalar F2F.BF16 conversions of the naive form were 25 % of the kernel's instructions
      // (out_f16: fp16 planes, the A operand of a one-pass enc_emb GEMM -- aitb_head_weights.plan)
      __nv_bfloat16* o = out + ((size_t)k * nbins + bin) * 2 * C + c0 + lane * 4;
      float lx, ly, lz, lw;
      uint2 hi, lo;
      hi.x = split_hi2(acc.x, acc.y, out_f16 != 0, lx, ly);
      hi.y = split_hi2(acc.z, acc.w, out_f16 != 0, lz, lw);
      *reinterpret_cast<uint2*>(o) = hi;
      lo.x = pack_plane2(lx, ly, out_f16 != 0);
      lo.y = pack_plane2(lz, lw, out_f16 != 0);
      *reinterpret_cast<uint2*>(o + C) = lo;
    } else {
      Vec4<T>::st(out + ((size_t)k * nbins + bin) * C + c0 + lane * 4, acc);
    }
  }
  if constexpr (NCHW_OUT) {
    __syncthreads();
    // the slab's channels x nbins outputs are contiguous in NCHW: fully coalesced store
    T* ob = out + ((size_t)k * C + c0) * nbins;
    const int s = nbins + 1;
    for (int i = threadIdx.x; i < cvalid * nbins; i += blockDim.x) {
      const int c = i / nbins, bin = i - c * nbins;
      Act<T>::st(ob + i, stage[c * s + bin]);
    }
  }
}

// bf16 map -> bf16 token-major output, 256-channel slabs (lane = 8 channels): the engine's bf16 configuration.
// Same per-axis footprint tables and bin walk as roi_align_fwd_kernel; grid (ceil(C / 256), K), 256 threads.
__global__ void __launch_bounds__(256, 4)
roi_align_fwd8_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ rois, int C, int H, int W, float scale,
                      int ph, int pw, int sampling_ratio, __nv_bfloat16* __restrict__ out) {
  __shared__ Foot xf[kMaxPooled];
  __shared__ Foot yf[kMaxPooled];
  const int k = blockIdx.y;
  const int c0 = blockIdx.x * 256;
  const RoiGeom g = roi_geometry(rois + (size_t)k * 5, scale, ph, pw, sampling_ratio);
  const int gw = g.grid_w, gh = g.grid_h;
  bool fits = true;
  if ((int)threadIdx.x < pw) fits = build_foot(xf[threadIdx.x], g.start_w, threadIdx.x, g.bin_w, gw, W);
  else if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + ph)
    fits = build_foot(yf[threadIdx.x - 32], g.start_h, threadIdx.x - 32, g.bin_h, gh, H);
  const bool tables = __syncthreads_and(fits) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cvalid = min(256, C - c0);
  const bool lane_on = lane * 8 < cvalid;
  const __nv_bfloat16* fb = feat + (size_t)g.batch * H * W * C + c0 + (lane_on ? lane * 8 : 0);
  const float inv_count = __frcp_rn((float)(g.grid_h * g.grid_w));
  const int nbins = ph * pw;
  const uint32_t uC = (uint32_t)C, rowC = (uint32_t)W * (uint32_t)C;
  int py = warp / pw, px = warp - py * pw;
  for (int bin = warp; bin < nbins; bin += 8) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (tables) {
      const int ny = yf[py].n, nx = xf[px].n;
      const uint32_t r = (uint32_t)yf[py].c0 * rowC + (uint32_t)xf[px].c0 * uC;
      switch (nx) {
        case 1: foot_rows8<1>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 2: foot_rows8<2>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 3: foot_rows8<3>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 4: foot_rows8<4>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 5: foot_rows8<5>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        case 6: foot_rows8<6>(a0, a1, fb, r, rowC, uC, ny, yf[py].w, xf[px].w); break;
        default: {
          uint32_t rr = r;
          for (int cy = 0; cy < ny; ++cy, rr += rowC) {
            const float wy = yf[py].w[cy];
            uint32_t q = rr;
            for (int cx = 0; cx < nx; ++cx, q += uC) {
              const float w = wy * xf[px].w[cx];
              float4 va, vb;
              ld8_bf16(fb + q, va, vb);
              a0.x += w * va.x; a0.y += w * va.y; a0.z += w * va.z; a0.w += w * va.w;
              a1.x += w * vb.x; a1.y += w * vb.y; a1.z += w * vb.z; a1.w += w * vb.w;
            }
          }
        }
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        const Tap ty = make_tap(sample_coord(g.start_h, py, g.bin_h, iy, gh), H);
        for (int ix = 0; ix < gw; ++ix) {
          const Tap tx = make_tap(sample_coord(g.start_w, px, g.bin_w, ix, gw), W);
          const float ws[4] = {ty.wlo * tx.wlo, ty.wlo * tx.whi, ty.whi * tx.wlo, ty.whi * tx.whi};
          const int ys[4] = {ty.lo, ty.lo, ty.hi, ty.hi}, xs[4] = {tx.lo, tx.hi, tx.lo, tx.hi};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float4 va, vb;
            ld8_bf16(fb + ((size_t)ys[t] * W + xs[t]) * C, va, vb);
            a0.x += ws[t] * va.x; a0.y += ws[t] * va.y; a0.z += ws[t] * va.z; a0.w += ws[t] * va.w;
            a1.x += ws[t] * vb.x; a1.y += ws[t] * vb.y; a1.z += ws[t] * vb.z; a1.w += ws[t] * vb.w;
          }
        }
      }
    }
    px += 8;
    while (px >= pw) { px -= pw; ++py; }
    if (!lane_on) continue;
    uint4 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(a0.x * inv_count, a0.y * inv_count);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(a0.z * inv_count, a0.w * inv_count);
    *reinterpret_cast<__nv_bfloat162*>(&o.z) = __floats2bfloat162_rn(a1.x * inv_count, a1.y * inv_count);
    *reinterpret_cast<__nv_bfloat162*>(&o.w) = __floats2bfloat162_rn(a1.z * inv_count, a1.w * inv_count);
    *reinterpret_cast<uint4*>(out + ((size_t)k * nbins + bin) * C + c0 + lane * 8) = o;
  }
}

// Backward (ROIAlign_cuda.cu:178-254): scatter g*w/count to the 4 taps.  Same CTA shape as the
// forward; the NCHW grad slab is staged through shared memory, and the accumulation target is
// channels-last so each atomic warp instruction hits 32 consecutive floats (red.global.add.v4
// is used by the footprint kernel below; this per-tap scalar-atomic form stays for C % 4 != 0).
__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ rois, int C, int H, int W, float scale,
                     int ph, int pw, int sampling_ratio, float* __restrict__ gfeat) {
  __shared__ Tap xtab[kMaxPooled * kMaxGrid];
  __shared__ Tap ytab[kMaxPooled * kMaxGrid];
  extern __shared__ float stage[];  // [32][nbins + 1]
  const int k = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const int nbins = ph * pw;
  const RoiGeom g = roi_geometry(rois + (size_t)k * 5, scale, ph, pw, sampling_ratio);
  const int gw = g.grid_w, gh = g.grid_h;
  const bool xt = gw <= kMaxGrid, yt = gh <= kMaxGrid;
  if (xt)
    for (int i = threadIdx.x; i < pw * gw; i += blockDim.x)
      xtab[i] = make_tap(sample_coord(g.start_w, i / gw, g.bin_w, i % gw, gw), W);
  if (yt)
    for (int i = threadIdx.x; i < ph * gh; i += blockDim.x)
      ytab[i] = make_tap(sample_coord(g.start_h, i / gh, g.bin_h, i % gh, gh), H);
  const float* gb = grad + ((size_t)k * C + c0) * nbins;
  const int s = nbins + 1;
  const int cvalid = min(32, C - c0);
  for (int i = threadIdx.x; i < cvalid * nbins; i += blockDim.x) {
    const int c = i / nbins, bin = i - c * nbins;
    stage[c * s + bin] = gb[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane >= cvalid) return;
  float* fb = gfeat + (size_t)g.batch * H * W * C + c0 + lane;
  const float count = (float)(g.grid_h * g.grid_w);
  for (int bin = warp; bin < nbins; bin += 8) {
    const int py = bin / pw, px = bin - py * pw;
    const float gv = stage[lane * s + bin];
    for (int iy = 0; iy < gh; ++iy) {
      const Tap ty = yt ? ytab[py * gh + iy] : make_tap(sample_coord(g.start_h, py, g.bin_h, iy, gh), H);
      for (int ix = 0; ix < gw; ++ix) {
        const Tap tx = xt ? xtab[px * gw + ix] : make_tap(sample_coord(g.start_w, px, g.bin_w, ix, gw), W);
        if (ty.wlo == 0.f && ty.whi == 0.f) continue;  // sample outside the map (x_low = -1 branch)
        if (tx.wlo == 0.f && tx.whi == 0.f) continue;
        const float g1 = gv * (ty.wlo * tx.wlo) / count, g2 = gv * (ty.wlo * tx.whi) / count;
        const float g3 = gv * (ty.whi * tx.wlo) / count, g4 = gv * (ty.whi * tx.whi) / count;
        atomicAdd(fb + ((size_t)ty.lo * W + tx.lo) * C, g1);
        atomicAdd(fb + ((size_t)ty.lo * W + tx.hi) * C, g2);
        atomicAdd(fb + ((size_t)ty.hi * W + tx.lo) * C, g3);
        atomicAdd(fb + ((size_t)ty.hi * W + tx.hi) * C, g4);
      }
    }
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Footprint backward (C % 4 == 0): the adjoint of the forward kernel above, same CTA shape (roi, 128-channel slab; lane = 4
// channels).  The 4*gh*gw bilinear taps of a bin land on (gh+1)(gw+1) distinct cells; their weights are folded per axis
// (the forward's Foot tables), so a bin issues ONE 128-bit vector reduction (red.global.add.v4.f32, sm_90+) per footprint
// cell into the channels-last gradient map instead of 4*gh*gw scalar atomics per channel: ~2x fewer bytes through the L2
// atomic units and 4x fewer instructions than roi_align_bwd_kernel (kept for C % 4 != 0).  g*w/count (:238-246) becomes
// g*(1/count)*Wy*Wx -- the same real-number sum re-associated; accumulation order is nondeterministic as in the reference.
__global__ void __launch_bounds__(256)
roi_align_bwd_foot_kernel(const float* __restrict__ grad, const float* __restrict__ rois, int C, int H, int W, float scale,
                          int ph, int pw, int sampling_ratio, float* __restrict__ gfeat) {
  __shared__ Foot xf[kMaxPooled];
  __shared__ Foot yf[kMaxPooled];
  extern __shared__ float stage[];  // [nbins][132]: the roi's gradient slab, bin-major so a lane reads its 4 channels as one float4
  constexpr int kS = 132;
  const int k = blockIdx.y;
  const int c0 = blockIdx.x * 128;
  const int nbins = ph * pw;
  const RoiGeom g = roi_geometry(rois + (size_t)k * 5, scale, ph, pw, sampling_ratio);
  const int gw = g.grid_w, gh = g.grid_h;
  bool fits = true;
  if ((int)threadIdx.x < pw) fits = build_foot(xf[threadIdx.x], g.start_w, threadIdx.x, g.bin_w, gw, W);
  else if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + ph)
    fits = build_foot(yf[threadIdx.x - 32], g.start_h, threadIdx.x - 32, g.bin_h, gh, H);
  const int cvalid = min(128, C - c0);
  const float* gb = grad + ((size_t)k * C + c0) * nbins;   // NCHW: the slab's cvalid * nbins values are contiguous
  for (int i = threadIdx.x; i < cvalid * nbins; i += blockDim.x) {
    const int c = i / nbins, bin = i - c * nbins;
    stage[bin * kS + c] = __ldg(gb + i);
  }
  const bool tables = __syncthreads_and(fits) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane * 4 >= cvalid) return;
  float* fb = gfeat + (size_t)g.batch * H * W * C + c0 + lane * 4;
  const float inv_count = __frcp_rn((float)(gh * gw));
  const uint32_t uC = (uint32_t)C, rowC = (uint32_t)W * (uint32_t)C;
  for (int bin = warp; bin < nbins; bin += 8) {
    const int py = bin / pw, px = bin - py * pw;
    float4 gv = *reinterpret_cast<const float4*>(stage + bin * kS + lane * 4);
    gv.x *= inv_count; gv.y *= inv_count; gv.z *= inv_count; gv.w *= inv_count;
    if (tables) {
      const int ny = yf[py].n, nx = xf[px].n;
      uint32_t r = (uint32_t)yf[py].c0 * rowC + (uint32_t)xf[px].c0 * uC;
      for (int cy = 0; cy < ny; ++cy, r += rowC) {
        const float wy = yf[py].w[cy];
        if (wy == 0.f) continue;
        uint32_t q = r;
        for (int cx = 0; cx < nx; ++cx, q += uC) {
          const float w = wy * xf[px].w[cx];
          if (w != 0.f) red_add_v4(fb + q, make_float4(gv.x * w, gv.y * w, gv.z * w, gv.w * w));
        }
      }
    } else {   // footprint wider than the tables (fixed sampling_ratio on a huge roi): per-sample taps
      for (int iy = 0; iy < gh; ++iy) {
        const Tap ty = make_tap(sample_coord(g.start_h, py, g.bin_h, iy, gh), H);
        if (ty.wlo == 0.f && ty.whi == 0.f) continue;
        for (int ix = 0; ix < gw; ++ix) {
          const Tap tx = make_tap(sample_coord(g.start_w, px, g.bin_w, ix, gw), W);
          if (tx.wlo == 0.f && tx.whi == 0.f) continue;
          const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
          red_add_v4(fb + ((size_t)ty.lo * W + tx.lo) * C, make_float4(gv.x * w1, gv.y * w1, gv.z * w1, gv.w * w1));
          red_add_v4(fb + ((size_t)ty.lo * W + tx.hi) * C, make_float4(gv.x * w2, gv.y * w2, gv.z * w2, gv.w * w2));
          red_add_v4(fb + ((size_t)ty.hi * W + tx.lo) * C, make_float4(gv.x * w3, gv.y * w3, gv.z * w3, gv.w * w3));
          red_add_v4(fb + ((size_t)ty.hi * W + tx.hi) * C, make_float4(gv.x * w4, gv.y * w4, gv.z * w4, gv.w * w4));
        }
      }
    }
  }
}

// [G, C, S] -> [G, S, C] (to_cl) or back, 32x32 smem tiles, optional dtype conversion.
// SS / SD: the source / destination is a split (two bf16 planes per row) matrix (AITB_F32S).
template <typename TS, typename TD, bool SS = false, bool SD = false>
__global__ void __launch_bounds__(256)
transpose_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int R, int Cc, int round_tf, int dst_f16 = 0) {
  // src [G, R, Cc] -> dst [G, Cc, R]
  __shared__ float tile[32][33];
  const int g = blockIdx.z;
  constexpr int ps = SS ? 2 : 1, pd = SD ? 2 : 1;
  const TS* s = src + (size_t)g * R * Cc * ps;
  TD* d = dst + (size_t)g * R * Cc * pd;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < R && c < Cc) {
      float v = Act<TS>::ld(s + (size_t)r * Cc * ps + c);
      if constexpr (SS) v += Act<TS>::ld(s + (size_t)r * Cc * ps + Cc + c);
      tile[j][tx] = v;
    }
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < R && c < Cc) {
      float v = tile[tx][j];
      if (sizeof(TD) == 4 && round_tf) v = rn_tf32(v);
      TD* o = d + (size_t)c * R * pd + r;
      if constexpr (SD) {
        if (dst_f16) {   // fp16 planes (precision plan): the 16-bit containers are written through __half
          const float vs = sat_f16(v);
          const __half hi = __float2half_rn(vs);
          reinterpret_cast<__half*>(o)[0] = hi;
          reinterpret_cast<__half*>(o)[R] = __float2half_rn(vs - __half2float(hi));
        } else {
          Act<TD>::st(o, v);
          Act<TD>::st(o + R, v - Act<TD>::ld(o));
        }
      } else {
        Act<TD>::st(o, v);
      }
    }
  }
}

int transpose_run(const void* src, int sdt, void* dst, int ddt, int G, int C, int S, int to_cl, cudaStream_t stream,
                  int round_tf, int dst_f16) {
  AITB_REQUIRE(G > 0 && C > 0 && S > 0, "aitb_transpose_cs: empty tensor");
  AITB_REQUIRE(G <= 65535, "aitb_transpose_cs: G=%d exceeds grid.z", G);
  // to_cl: src [G, C, S] -> dst [G, S, C]  => R = C, Cc = S ; else src [G, S, C] -> dst [G, C, S] => R = S, Cc = C
  const int R = to_cl ? C : S, Cc = to_cl ? S : C;
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, G);
  if (sdt == AITB_F32 && ddt == AITB_F32)
    transpose_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)src, (float*)dst, R, Cc, round_tf);
  else if (sdt == AITB_F32 && ddt == AITB_BF16)
    transpose_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, R, Cc, 0);
  else if (sdt == AITB_BF16 && ddt == AITB_F32)
    transpose_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, R, Cc, 0);
  else if (sdt == AITB_BF16 && ddt == AITB_BF16)
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src,
                                                                              (__nv_bfloat16*)dst, R, Cc, 0);
  else if (sdt == AITB_F32 && ddt == AITB_F32S)
    transpose_kernel<float, __nv_bfloat16, false, true><<<grid, 256, 0, stream>>>((const float*)src,
                                                                                   (__nv_bfloat16*)dst, R, Cc, 0, dst_f16);
  else if (sdt == AITB_F32S && ddt == AITB_F32)
    transpose_kernel<__nv_bfloat16, float, true, false><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src,
                                                                                   (float*)dst, R, Cc, 0);
  else {
    set_error("aitb_transpose_cs: bad dtypes %d -> %d", sdt, ddt);
    return 1;
  }
  return check_launch("transpose_kernel");
}

int roi_align_fwd_run(const void* feat, const float* rois, int B, int C, int H, int W, int K, float scale, int ph,
                      int pw, int sampling_ratio, int dtype, int out_layout, void* out, cudaStream_t stream, int round_tf,
                      int out_f16) {
  AITB_REQUIRE(K >= 0 && B > 0, "aitb_roi_align_forward: bad sizes");
  if (K == 0) return 0;
  AITB_REQUIRE(C % 4 == 0, "aitb_roi_align_forward: C=%d must be a multiple of 4 (128-bit channel vectors)", C);
  AITB_REQUIRE(ph >= 1 && pw >= 1 && ph <= kMaxPooled && pw <= kMaxPooled, "aitb_roi_align_forward: pooled size %dx%d unsupported", ph, pw);
  AITB_REQUIRE(K <= 65535, "aitb_roi_align_forward: K=%d exceeds one launch (chunk the rois)", K);
  AITB_REQUIRE(out_layout == 0 || out_layout == 1, "aitb_roi_align_forward: out_layout must be 0 or 1");
  AITB_REQUIRE((size_t)H * W * C < ((size_t)1 << 31), "aitb_roi_align_forward: one image's map must stay below 2^31 elements");
  dim3 grid((C + 127) / 128, K);
  const size_t smem = out_layout == 0 ? (size_t)128 * (ph * pw + 1) * 4 : 0;
  static const bool no_fwd8 = getenv("AITB_ROI_NO_FWD8") != nullptr;   // A/B: the 4-channels-per-lane bf16 kernel
  if (dtype == AITB_F32) {
    if (out_layout == 0)
      roi_align_fwd_kernel<float, true><<<grid, 256, smem, stream>>>((const float*)feat, rois, C, H, W, scale, ph, pw,
                                                                      sampling_ratio, (float*)out, round_tf, 0);
    else
      roi_align_fwd_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)feat, rois, C, H, W, scale, ph, pw,
                                                                    sampling_ratio, (float*)out, round_tf, 0);
  } else if (dtype == AITB_F32S) {   // fp32 map -> split token-major output (engine-internal)
    AITB_REQUIRE(out_layout == 1, "aitb_roi_align_forward: the split configuration writes token-major output only");
    roi_align_fwd_kernel<float, false, true><<<grid, 256, 0, stream>>>((const float*)feat, rois, C, H, W, scale, ph, pw,
                                                                       sampling_ratio, (__nv_bfloat16*)out, 0, out_f16);
  } else if (dtype == AITB_BF16) {
    if (out_layout == 0)
      roi_align_fwd_kernel<__nv_bfloat16, true><<<grid, 256, smem, stream>>>(
          (const __nv_bfloat16*)feat, rois, C, H, W, scale, ph, pw, sampling_ratio, (__nv_bfloat16*)out, 0, 0);
    else if (C % 8 == 0 && (((uintptr_t)feat | (uintptr_t)out) & 15) == 0 && !no_fwd8)
      roi_align_fwd8_kernel<<<dim3((C + 255) / 256, K), 256, 0, stream>>>((const __nv_bfloat16*)feat, rois, C, H, W, scale, ph,
                                                                          pw, sampling_ratio, (__nv_bfloat16*)out);
    else
      roi_align_fwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(
          (const __nv_bfloat16*)feat, rois, C, H, W, scale, ph, pw, sampling_ratio, (__nv_bfloat16*)out, 0, 0);
  } else {
    set_error("aitb_roi_align_forward: bad dtype %d", dtype);
    return 1;
  }
  return check_launch("roi_align_fwd_kernel");
}

int roi_align_bwd_run(const float* grad, const float* rois, int B, int C, int H, int W, int K, float scale, int ph,
                      int pw, int sampling_ratio, float* gfeat, cudaStream_t stream) {
  AITB_REQUIRE(K >= 0 && B > 0, "aitb_roi_align_backward: bad sizes");
  if (K == 0) return 0;
  AITB_REQUIRE(C >= 1, "aitb_roi_align_backward: bad C");
  AITB_REQUIRE(ph >= 1 && pw >= 1 && ph <= kMaxPooled && pw <= kMaxPooled, "aitb_roi_align_backward: pooled size unsupported");
  AITB_REQUIRE(K <= 65535, "aitb_roi_align_backward: K=%d exceeds one launch", K);
  static const bool scalar_atomics = getenv("AITB_ROI_BWD_SCALAR") != nullptr;   // A/B: the per-tap scalar-atomic kernel
  if (C % 4 == 0 && (((uintptr_t)gfeat) & 15) == 0 && (size_t)H * W * C < ((size_t)1 << 31) && !scalar_atomics) {
    dim3 grid((C + 127) / 128, K);
    const size_t smem = (size_t)ph * pw * 132 * 4;
    roi_align_bwd_foot_kernel<<<grid, 256, smem, stream>>>(grad, rois, C, H, W, scale, ph, pw, sampling_ratio, gfeat);
    return check_launch("roi_align_bwd_foot_kernel");
  }
  dim3 grid((C + 31) / 32, K);
  const size_t smem = (size_t)32 * (ph * pw + 1) * 4;
  roi_align_bwd_kernel<<<grid, 256, smem, stream>>>(grad, rois, C, H, W, scale, ph, pw, sampling_ratio, gfeat);
  return check_launch("roi_align_bwd_kernel");
}

}  // namespace aitb
