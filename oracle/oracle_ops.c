/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C restatement of the reference's integer/byte-exact
 * ops on the head path.  Never linked into libaitb200 or imported by ait_b200/.
 *
 * Pinned (tests/test_oracle_pins.py) against the reference's own CPU build (oracle/_ref, made from
 * lib/model/csrc/cpu/{nms_cpu,ROIAlign_cpu}.cpp) and against committed golden vectors.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared   (no FMA contraction: the comparisons below must
 * see the same IEEE single-precision values as the x86 reference build and the CUDA kernels, which
 * use explicitly rounded __f*_rn intrinsics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* IoU with the legacy "+1" box convention: devIoU, lib/model/csrc/cuda/nms.cu:13-21
 * (same arithmetic as lib/model/csrc/cpu/nms_cpu.cpp:22,52-59). */
static float iou_legacy(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  float inter = width * height;
  float sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
  float sb = (b[2] - b[0] + 1.f) * (b[3] - b[1] + 1.f);
  return inter / (sa + sb - inter);
}

/* Greedy NMS over boxes ALREADY in descending-score order.
 *   boxes [n,4]; keep_pos receives the kept positions (ascending = score order); returns count.
 *   ge == 0: suppress when IoU >  thr  (CUDA path, nms.cu:60 + greedy scan :112-123)
 *   ge == 1: suppress when IoU >= thr  (CPU path, nms_cpu.cpp:60)
 *   max_keep > 0 stops after max_keep survivors (proposal_layer.py:156 consumes keep[:post_nms_topN]). */
int oracle_nms_sorted(const float* boxes, int n, float thr, int ge, int max_keep, int32_t* keep_pos) {
  uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (dead[i]) continue;
    keep_pos[kept++] = i;
    if (max_keep > 0 && kept >= max_keep) break;
    for (int j = i + 1; j < n; ++j) {
      if (dead[j]) continue;
      float o = iou_legacy(boxes + 4 * (size_t)i, boxes + 4 * (size_t)j);
      if (ge ? (o >= thr) : (o > thr)) dead[j] = 1;
    }
  }
  free(dead);
  return kept;
}

/* One axis of bilinear_interpolate (lib/model/csrc/cuda/ROIAlign_cuda.cu:22-52 ==
 * pre_calc_for_bilinear_interpolate, lib/model/csrc/cpu/ROIAlign_cpu.cpp:48-98). */
typedef struct { int lo, hi; float wlo, whi; int valid; } tap_t;

static tap_t make_tap(float c, int size) {
  tap_t t;
  memset(&t, 0, sizeof(t));
  if (c < -1.0f || c > (float)size) return t;
  t.valid = 1;
  if (c <= 0.f) c = 0.f;
  int lo = (int)c, hi;
  if (lo >= size - 1) { hi = lo = size - 1; c = (float)lo; } else { hi = lo + 1; }
  float l = c - (float)lo;
  t.lo = lo; t.hi = hi; t.whi = l; t.wlo = 1.f - l;
  return t;
}

/* ROIAlign forward, NCHW fp32 in -> [K,C,ph,pw] fp32 out
 * (ROIAlignForward_cpu_kernel, ROIAlign_cpu.cpp:113-219; sample grid :145-160, average :205). */
void oracle_roi_align_forward(const float* feat, const float* rois, int C, int H, int W, int K, float scale,
                              int ph, int pw, int sampling_ratio, float* out) {
  for (int k = 0; k < K; ++k) {
    const float* r = rois + 5 * (size_t)k;
    int b = (int)r[0];
    float sw = r[1] * scale, sh = r[2] * scale, ew = r[3] * scale, eh = r[4] * scale;
    float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
    float bin_h = rh / (float)ph, bin_w = rw / (float)pw;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
    float count = (float)(gh * gw);
    tap_t* ty = (tap_t*)malloc(sizeof(tap_t) * (size_t)(ph * gh));
    tap_t* tx = (tap_t*)malloc(sizeof(tap_t) * (size_t)(pw * gw));
    for (int p = 0; p < ph; ++p)
      for (int i = 0; i < gh; ++i)
        ty[p * gh + i] = make_tap(sh + (float)p * bin_h + ((float)i + .5f) * bin_h / (float)gh, H);
    for (int p = 0; p < pw; ++p)
      for (int i = 0; i < gw; ++i)
        tx[p * gw + i] = make_tap(sw + (float)p * bin_w + ((float)i + .5f) * bin_w / (float)gw, W);
    for (int c = 0; c < C; ++c) {
      const float* f = feat + ((size_t)b * C + c) * H * W;
      float* o = out + ((size_t)k * C + c) * ph * pw;
      for (int py = 0; py < ph; ++py)
        for (int px = 0; px < pw; ++px) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            tap_t y = ty[py * gh + iy];
            for (int ix = 0; ix < gw; ++ix) {
              tap_t x = tx[px * gw + ix];
              if (!y.valid || !x.valid) continue; /* contributes 0 (ROIAlign_cuda.cu:22-25) */
              float w1 = y.wlo * x.wlo, w2 = y.wlo * x.whi, w3 = y.whi * x.wlo, w4 = y.whi * x.whi;
              acc += w1 * f[y.lo * W + x.lo] + w2 * f[y.lo * W + x.hi] + w3 * f[y.hi * W + x.lo] +
                     w4 * f[y.hi * W + x.hi];
            }
          }
          o[py * pw + px] = acc / count;
        }
    }
    free(ty);
    free(tx);
  }
}

/* ROIAlign backward (RoIAlignBackwardFeature, ROIAlign_cuda.cu:178-254): grad [K,C,ph,pw] ->
 * gfeat [B,C,H,W] (zeroed by the caller); sequential adds in double to give an order-free target. */
void oracle_roi_align_backward(const float* grad, const float* rois, int C, int H, int W, int K, float scale,
                               int ph, int pw, int sampling_ratio, double* gfeat) {
  for (int k = 0; k < K; ++k) {
    const float* r = rois + 5 * (size_t)k;
    int b = (int)r[0];
    float sw = r[1] * scale, sh = r[2] * scale, ew = r[3] * scale, eh = r[4] * scale;
    float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
    float bin_h = rh / (float)ph, bin_w = rw / (float)pw;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
    float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      double* gf = gfeat + ((size_t)b * C + c) * H * W;
      const float* g = grad + ((size_t)k * C + c) * ph * pw;
      for (int py = 0; py < ph; ++py)
        for (int px = 0; px < pw; ++px) {
          float gv = g[py * pw + px];
          for (int iy = 0; iy < gh; ++iy) {
            tap_t y = make_tap(sh + (float)py * bin_h + ((float)iy + .5f) * bin_h / (float)gh, H);
            for (int ix = 0; ix < gw; ++ix) {
              tap_t x = make_tap(sw + (float)px * bin_w + ((float)ix + .5f) * bin_w / (float)gw, W);
              if (!y.valid || !x.valid) continue;
              gf[y.lo * W + x.lo] += (double)(gv * (y.wlo * x.wlo) / count);
              gf[y.lo * W + x.hi] += (double)(gv * (y.wlo * x.whi) / count);
              gf[y.hi * W + x.lo] += (double)(gv * (y.whi * x.wlo) / count);
              gf[y.hi * W + x.hi] += (double)(gv * (y.whi * x.whi) / count);
            }
          }
        }
    }
  }
}
