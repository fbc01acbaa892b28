// C ABI of libaitb200 (see include/aitb200.h) and the native head engine: the per-step launch
// sequence ROIAlign -> AIT -> SKNet -> RCNN_top -> heads for a batch of (image, query) units.
// Host code only orchestrates: every FLOP/byte of the path runs in the kernels of this library.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/aitb200.h"
#include "common.cuh"

namespace aitb {

// ---- implemented in the other translation units
int gemm_run(const aitb_gemm_desc* d, cudaStream_t stream);
size_t nms_workspace_bytes(int B, int n_total, int n);
int nms_run(const float* boxes, const int64_t* order, int B, int n_total, int n, float thr, int max_out, int mode,
            int64_t* keep_out, int32_t* n_keep, float* rois_out, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t topk_workspace_bytes(int B, int n_total, int n);
int topk_run(const float* scores, int B, int n_total, int n, int64_t* order, void* ws, size_t ws_bytes,
             cudaStream_t stream);
int transpose_run(const void* src, int sdt, void* dst, int ddt, int G, int C, int S, int to_cl, cudaStream_t stream,
                  int round_tf = 0, int dst_f16 = 0);
int roi_align_fwd_run(const void* feat, const float* rois, int B, int C, int H, int W, int K, float scale, int ph,
                      int pw, int sampling_ratio, int dtype, int out_layout, void* out, cudaStream_t stream, int round_tf = 0,
                      int out_f16 = 0);
int roi_align_bwd_run(const float* grad, const float* rois, int B, int C, int H, int W, int K, float scale, int ph,
                      int pw, int sampling_ratio, float* gfeat, cudaStream_t stream);
int attn_core_run(const void* q, int ldq, int q_rep, const void* k, const void* v, int ldkv, const float* w_sk,
                  const float* b_sk, int G, int mask_mode, int n_keys, int dtype, void* out, cudaStream_t stream,
                  int round_tf = 0, int kv_rows = 64, const DropCfg* drop = nullptr);
int pool_heads_run(const void* top, int dtype, int G, int P, const float* qfeat, const float* w_bbox,
                   const float* b_bbox, const float* w1, const float* b1, const float* w2, const float* b2,
                   float* feat_out, float* bbox_out, float* cls_out, cudaStream_t stream);

int wgrad_conv_run(const float* dy, int ldy, const float* x, int G, int S, int C, int N, int groups, int taps, float* dw,
                   int ldw, cudaStream_t stream);
int wgrad_run(const float* dy, int ldy, const float* x, int ldx, int M, int N, int K, float* dw, int ldw,
              cudaStream_t stream);
int ln_bwd_run(const float* g, const float* y, const float* gamma, const float* beta, const float* rstd, int rows,
               int grp, int valid, float* dx, float* dgamma, float* dbeta, cudaStream_t stream);
int colsum_run(const float* x, int ld, int rows, int cols, float* out, cudaStream_t stream);
int bsum_run(const float* x, int B, int P, int L, float* out, cudaStream_t stream);
int attn_bwd_run(const float* q, int ldq, int q_rep, const float* k, const float* v, int ldkv, const float* w_sk,
                 const float* b_sk, const float* dout, int G, int mask_mode, int n_keys, float* dq, int lddq, float* dk,
                 float* dv, int lddkv, float* dz, float* s_out, cudaStream_t stream, const DropCfg* drop = nullptr);
int drop_res_ln_run(float* z, const float* pos, int grp, int valid, const float* res, int res_div, int res_rep, const float* gamma,
                    const float* beta, float eps, float p, unsigned long long seed, int site, int rows, int round_tf,
                    float* rstd_out, cudaStream_t st);
int drop_bwd_run(const float* dx, float* dz, float p, unsigned long long seed, int site, int rows, int grp, int valid,
                 cudaStream_t st);
// bf16-storage overloads of the training kernels (bwd.cu, dropout.cu)
typedef __nv_bfloat16 bf16_t;
int wgrad_run(const bf16_t* dy, int ldy, const bf16_t* x, int ldx, int M, int N, int K, float* dw, int ldw, cudaStream_t stream);
int ln_bwd_run(const bf16_t* g, const bf16_t* y, const float* gamma, const float* beta, const float* rstd, int rows, int grp,
               int valid, bf16_t* dx, float* dgamma, float* dbeta, cudaStream_t stream);
int colsum_run(const bf16_t* x, int ld, int rows, int cols, float* out, cudaStream_t stream);
int bsum_run(const bf16_t* x, int B, int P, int L, bf16_t* out, cudaStream_t stream);
int attn_bwd_run(const bf16_t* q, int ldq, int q_rep, const bf16_t* k, const bf16_t* v, int ldkv, const float* w_sk,
                 const float* b_sk, const bf16_t* dout, int G, int mask_mode, int n_keys, bf16_t* dq, int lddq, bf16_t* dk,
                 bf16_t* dv, int lddkv, bf16_t* dz, bf16_t* s_out, cudaStream_t stream, const DropCfg* drop = nullptr);
int drop_res_ln_run(bf16_t* z, const float* pos, int grp, int valid, const bf16_t* res, int res_div, int res_rep,
                    const float* gamma, const float* beta, float eps, float p, unsigned long long seed, int site, int rows,
                    int round_tf, float* rstd_out, cudaStream_t st);
int drop_bwd_run(const bf16_t* dx, bf16_t* dz, float p, unsigned long long seed, int site, int rows, int grp, int valid,
                 cudaStream_t st);
template <typename T> struct DtOf;
template <> struct DtOf<float> { static constexpr int v = AITB_F32; };
template <> struct DtOf<bf16_t> { static constexpr int v = AITB_BF16; };

int rpn_decode_run(const float* scores_nchw, const float* deltas_nchw, const float* base_anchors, const float* im_info,
                   int B, int A, int H, int W, float feat_stride, float* proposals, float* fg_scores, cudaStream_t st);
int box_decode_run(const float* boxes, int boxes_stride, int boxes_off, const float* deltas, const float* cls,
                   const float* im_info, int B, int N, const float* stds, const float* means, float thresh,
                   int divide_by_scale, float* pred, float* key, int32_t* n_valid, cudaStream_t st);
int rpn_head_decode_run(const void* heads, int dtype, int ld, const float* base_anchors, const float* im_info, int B, int A,
                        int H, int W, float feat_stride, float* proposals, float* fg_scores, float* cls_prob_nchw,
                        float* bbox_pred_nchw, cudaStream_t st);
int group_norm_residual_run(const float* x, const float* identity, const float* gamma, const float* beta, int B, int N,
                            int groups, float eps, double* sums, float* out, cudaStream_t st);
int fc_ln_run(int dtype, const void* a, const void* w, const void* res, const float* gamma, const float* beta, float eps,
              void* out, int M, int rows_in, int rows_out, int res_row_m, int res_div, int res_rep, cudaStream_t st,
              int res_f16 = 0, int out_f16 = 0);
int det_assemble_run(const float* pred, const float* cls, const int64_t* order, const int64_t* keep_pos,
                     const int32_t* n_keep, const int32_t* n_valid, int B, int N, int max_per_image, float* dets,
                     int32_t* n_det, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// error string + launch counter (thread-local)
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_dyn_smem(const void* func, int bytes, SmemAttrOnce& once, const char* what) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("%s: cudaGetDevice failed", what);
    return 1;
  }
  if ((__atomic_load_n(&once.done_mask, __ATOMIC_ACQUIRE) >> dev) & 1ULL) return 0;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%s, %d bytes) failed: %s", what, bytes, cudaGetErrorString(e));
    return 1;
  }
  __atomic_fetch_or(&once.done_mask, 1ULL << dev, __ATOMIC_RELEASE);
  return 0;
}

int current_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    __atomic_store_n(&cache[dev], n, __ATOMIC_RELAXED);
  }
  return n;
}

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// small fused helper: encoder rows t >= n_valid of every pair are zero-padded AFTER enc_emb
// (system/Models.py:268-270), so their encoder input is LayerNorm(pos_table[t]) -- independent of
// the proposal.  One warp per row.
// ---------------------------------------------------------------------------------------------
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256)
ln_pad_rows_kernel(T* __restrict__ x, int n_pairs, int n_valid, const float* __restrict__ pos,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int round_tf, int out_f16) {
  const int n_pad = 64 - n_valid;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_pairs * n_pad) return;
  const int pair = row / n_pad, t = n_valid + row % n_pad;
  const float* pr = pos + (size_t)t * 512;
  float v[16];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = pr[lane + 32 * i];
    sum += v[i];
  }
  const float mean = warp_sum(sum) * (1.f / 512.f);
  float ssq = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) ssq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(ssq) * (1.f / 512.f) + eps);
  T* xr = x + ((size_t)pair * 64 + t) * (SPLIT ? 1024 : 512);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
    if (sizeof(T) == 4 && round_tf) {
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(o));
      o = __uint_as_float(r);
    }
    if constexpr (SPLIT) {
      if (out_f16) {   // fp16 planes (precision plan): 16-bit containers written through __half
        const float os = sat_f16(o);
        const __half hi = __float2half_rn(os);
        reinterpret_cast<__half*>(xr)[c] = hi;
        reinterpret_cast<__half*>(xr)[512 + c] = __float2half_rn(os - __half2float(hi));
      } else {
        Act<T>::st(xr + c, o);
        Act<T>::st(xr + 512 + c, o - Act<T>::ld(xr + c));   // lo plane = x - bf16(x)
      }
    } else {
      Act<T>::st(xr + c, o);
    }
  }
}

static int ln_pad_rows(void* x, int dtype, int n_pairs, int n_valid, const float* pos, const aitb_lnorm& ln,
                       int round_tf, cudaStream_t st, int out_f16 = 0) {
  const int rows = n_pairs * (64 - n_valid);
  if (rows <= 0) return 0;
  const int grid = (rows + 7) / 8;
  if (dtype == AITB_F32)
    ln_pad_rows_kernel<float, false><<<grid, 256, 0, st>>>((float*)x, n_pairs, n_valid, pos, ln.gamma, ln.beta, 1e-6f,
                                                           round_tf, 0);
  else if (dtype == AITB_F32S)
    ln_pad_rows_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((__nv_bfloat16*)x, n_pairs, n_valid, pos, ln.gamma,
                                                                  ln.beta, 1e-6f, 0, out_f16);
  else
    ln_pad_rows_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((__nv_bfloat16*)x, n_pairs, n_valid, pos, ln.gamma,
                                                                   ln.beta, 1e-6f, round_tf, 0);
  return check_launch("ln_pad_rows_kernel");
}

// ---------------------------------------------------------------------------------------------
// GEMM descriptor builders
// ---------------------------------------------------------------------------------------------
// bytes per logical element of an activation / weight row (split mode: two bf16 planes), and bytes per
// COLUMN step inside a row (split mode: the hi plane is bf16, the lo plane sits `ld` elements further)
static inline int esize(int dtype) { return dtype == AITB_BF16 ? 2 : 4; }
static inline int colsize(int dtype) { return dtype == AITB_F32 ? 4 : 2; }

static aitb_gemm_desc gemm_base(int dtype, int M, int N, int K, const void* W, int block_n, void* out, int ldo,
                                int round_tf) {
  aitb_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.dtype = dtype;
  d.M = M;
  d.N = N;
  d.k_per_tap = K;
  d.taps = 1;
  d.w = W;
  d.block_n = block_n;
  d.out = out;
  d.ldo = ldo;
  d.rows_in = d.rows_out = M;
  d.res_div = d.res_rep = 1;
  d.pos_rows = 1;
  d.eps = 1e-6f;
  d.round_tf32 = round_tf;
  return d;
}

// A = row-major [M, K] with leading dimension lda (elements)
static void view_plain(aitb_gemm_desc& d, const void* A, int lda) {
  const int eb = esize(d.dtype);
  d.a.ptr = A;
  d.a.dims[0] = (uint64_t)d.k_per_tap * d.taps;
  if (d.a_group_c) d.a.dims[0] = (uint64_t)lda;  // grouped: the view spans all input channels
  if (d.dtype == AITB_F32S) {                    // TMA elements are bf16; the lo plane starts lda further
    d.a.dims[0] += (uint64_t)lda;
    d.a_lo_off = lda;
  }
  d.a.dims[1] = (uint64_t)d.M;
  d.a.dims[2] = d.a.dims[3] = 1;
  d.a.strides[0] = (uint64_t)lda * eb;
  d.a.strides[1] = d.a.strides[2] = (uint64_t)lda * eb * d.M;
  d.a.box[0] = 128 / colsize(d.dtype);
  d.a.box[1] = 128;
  d.a.box[2] = d.a.box[3] = 1;
  d.a_m_dim = 1;
  d.a_m_step = 128;
}

// A = channels-last map [G, S, S, C] seen as an s x s grid sampled with `stride`; one m-tile = 128 / (s*s) maps
static void view_map(aitb_gemm_desc& d, const void* A, int C, int S, int s, int stride, int G) {
  const int eb = esize(d.dtype);
  d.a.ptr = A;
  d.a.dims[0] = (uint64_t)C;
  if (d.dtype == AITB_F32S) {
    d.a.dims[0] = (uint64_t)2 * C;
    d.a_lo_off = C;
  }
  d.a.dims[1] = d.a.dims[2] = (uint64_t)s;
  d.a.dims[3] = (uint64_t)G;
  d.a.strides[0] = (uint64_t)stride * C * eb;
  d.a.strides[1] = (uint64_t)stride * S * C * eb;
  d.a.strides[2] = (uint64_t)S * S * C * eb;
  d.a.box[0] = 128 / colsize(d.dtype);
  d.a.box[1] = d.a.box[2] = (uint32_t)s;
  d.a.box[3] = (uint32_t)(128 / (s * s));
  d.a_m_dim = 3;
  d.a_m_step = 128 / (s * s);
}

static void taps3x3(aitb_gemm_desc& d) {
  d.taps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      d.tap_dx[ky * 3 + kx] = (int8_t)(kx - 1);
      d.tap_dy[ky * 3 + kx] = (int8_t)(ky - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// workspace bump allocator
// ---------------------------------------------------------------------------------------------
struct Bump {
  uint8_t* base;
  size_t off;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 1023) & ~(size_t)1023;
    return p;
  }
};

struct HeadBufs {
  void *featT, *pooled, *qtok, *X1, *QKV, *AO, *X2, *Hh, *ENC;
  void *T0, *QKVd, *AOd, *T1, *Qc;
  void *KVc, *D1, *DEC, *AIT, *SK;
  void *c1, *c2, *ds, *y0, *y1;
  void *SKq, *qc1, *qc2, *qds, *qy0, *qy1;
  float *feat, *qfeat;
  void* in_tok;  // ait_forward only: token-major copy of x_props
  // training forward: separate buffers where inference reuses one (AOc / H2 alias AO / Hh in inference), and
  // 1 / sigma of every LayerNorm row (NULL in inference)
  void *AOc, *Hh2;
  float *r_X1, *r_X2, *r_ENC, *r_T0, *r_T1, *r_D1, *r_DEC;
};

static void carve(Bump& b, HeadBufs& hb, int B, int P, int HW, int dtype, bool with_roi, bool with_top) {
  const size_t eb = esize(dtype);
  const size_t bp = (size_t)B * P, R = bp * 64, RQ = (size_t)B * 64;
  memset(&hb, 0, sizeof(hb));
  if (with_roi) hb.featT = b.take((size_t)B * HW * 1024 * eb);
  hb.pooled = b.take(bp * 49 * 1024 * eb);
  hb.qtok = b.take(RQ * 1024 * eb);
  hb.X1 = b.take(R * 512 * eb);
  hb.QKV = b.take(R * 1536 * eb);
  hb.AO = b.take(R * 64 * eb);
  hb.X2 = b.take(R * 512 * eb);
  hb.Hh = b.take(R * 2048 * eb);
  hb.ENC = b.take(R * 512 * eb);
  hb.T0 = b.take(RQ * 512 * eb);
  hb.QKVd = b.take(RQ * 1536 * eb);
  hb.AOd = b.take(RQ * 64 * eb);
  hb.T1 = b.take(RQ * 512 * eb);
  hb.Qc = b.take(RQ * 512 * eb);
  hb.KVc = b.take(R * 1024 * eb);
  hb.D1 = b.take(R * 512 * eb);
  hb.DEC = b.take(R * 512 * eb);
  hb.AIT = b.take(R * 1024 * eb);
  if (with_top) {
    // RCNN_top has ONE set of weights for the proposal maps and the query maps (`_head_to_tail` is called on both,
    // faster_rcnn_coatt_transformer_sk.py:299-300): the B query maps sit right behind the bp proposal maps in the SKNet output
    // buffer and layer 4 runs once over bp + B maps -- the query side costs one extra 128-row tile instead of eleven tiny
    // single-tile launches on the side stream (which delayed the persistent main-stream kernels they had to squeeze between)
    const size_t R16 = (bp + B) * 16;
    hb.SK = b.take((R + RQ) * 1024 * eb);
    hb.SKq = (uint8_t*)hb.SK + R * 1024 * eb;
    hb.c1 = b.take(R16 * 512 * eb);
    hb.c2 = b.take(R16 * 512 * eb);
    hb.ds = b.take(R16 * 2048 * eb);
    hb.y0 = b.take(R16 * 2048 * eb);
    hb.y1 = b.take(R16 * 2048 * eb);
    hb.feat = (float*)b.take(bp * 2048 * 4);
    hb.qfeat = (float*)b.take((size_t)B * 2048 * 4);
  }
  hb.AOc = hb.AO;
  hb.Hh2 = hb.Hh;
}

// Side stream for the proposal-independent query branch (a few tiny launches that would otherwise sit
// on the critical path).  Created lazily, one per host thread and device; fork/join with events, so the
// caller still observes plain stream semantics on the stream it passed (also under graph capture).
struct SideStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, q1 = nullptr, q2 = nullptr;
};
static thread_local SideStream g_side[16];

static int side_stream(SideStream** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  AITB_REQUIRE(e == cudaSuccess && dev >= 0 && dev < 16, "side_stream: bad device");
  SideStream& s = g_side[dev];
  if (s.stream == nullptr) {
    e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    AITB_REQUIRE(e == cudaSuccess, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.q1, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.q2, cudaEventDisableTiming);
    s.device = dev;
  }
  *out = &s;
  return 0;
}

// training-mode dropout of aitb_ait_forward_train / aitb_ait_backward (aitb_head_weights.p_drop / p_attn / drop_seed); every
// nn.Dropout of the reference's Transformer is a SITE with its own Philox key; the inference entry points pass nullptr
struct TrainDrop {
  float p, p_attn;
  unsigned long long seed;
  DropCfg attn(int site) const { return make_drop(p_attn, seed, site); }
};
// y = LayerNorm(dropout(z [+ pos]) + residual) in place, fp32 or bf16 storage
static int drop_res_ln_dt(int dt, void* z, const float* pos, int valid, const void* res, int res_rep, const aitb_lnorm& ln,
                          const TrainDrop& td, int site, int rows, int round_tf, float* rstd, cudaStream_t st);
static bool train_drop_of(const aitb_head_weights* w, TrainDrop& td) {
  td.p = w->p_drop; td.p_attn = w->p_attn; td.seed = w->drop_seed;
  return td.p > 0.f || td.p_attn > 0.f;
}

#define RUN(expr)        \
  do {                   \
    if ((expr)) return 1; \
  } while (0)

static int drop_res_ln_dt(int dt, void* z, const float* pos, int valid, const void* res, int res_rep, const aitb_lnorm& ln,
                          const TrainDrop& td, int site, int rows, int round_tf, float* rstd, cudaStream_t st) {
  if (dt == AITB_F32)
    return drop_res_ln_run((float*)z, pos, 64, valid, (const float*)res, 64, res_rep, ln.gamma, ln.beta, 1e-6f, td.p, td.seed, site,
                           rows, round_tf, rstd, st);
  AITB_REQUIRE(dt == AITB_BF16, "training dropout: fp32 or bf16 storage only");
  return drop_res_ln_run((bf16_t*)z, pos, 64, valid, (const bf16_t*)res, 64, res_rep, ln.gamma, ln.beta, 1e-6f, td.p, td.seed, site,
                         rows, 0, rstd, st);
}

// ---------------------------------------------------------------------------------------------
// AIT: Transformer.forward (system/Models.py:231-280) on token-major inputs
//   pooled [bp*49, 1024], qtok [B*64, 1024]  ->  hb.AIT [bp*64, 1024]
// ---------------------------------------------------------------------------------------------
// out_rows: rows of every 64-token group that are stored (49 = drop the encoder's dead pad rows);
// kv_rows: rows per pair in the K / V buffers
static int mha_block(const aitb_head_weights* w, const aitb_mha& m, const void* qbuf, int ldq, int q_rep,
                     const void* kbuf, const void* vbuf, int ldkv, int G, int mask_mode, int n_keys, void* ao,
                     const void* res, int res_rep, void* out, cudaStream_t st, float* rstd = nullptr, int out_rows = 64,
                     int kv_rows = 64, int res_f16 = 0, int out_f16 = 0, const TrainDrop* td = nullptr, int site_attn = 0,
                     int site_fc = 0) {
  const int dt = w->dtype;
  if (td) {   // training with dropout (fp32 storage, 64-row layout): P <- dropout(P) inside the attention core, then
              // q = LayerNorm(dropout(fc(q)) + residual) as GEMM + one streaming kernel (dropout.cu)
    const DropCfg da = td->attn(site_attn);
    RUN(attn_core_run(qbuf, ldq, q_rep, kbuf, vbuf, ldkv, m.w_sk, m.b_sk, G, mask_mode, n_keys, dt, ao, st, w->round_tf32,
                      kv_rows, &da));
    aitb_gemm_desc d = gemm_base(dt, G * 64, 512, 64, m.w_fc, 256, out, 512, 0);
    view_plain(d, ao, 64);
    RUN(gemm_run(&d, st));
    return drop_res_ln_dt(dt, out, nullptr, 64, res, res_rep, m.ln, *td, site_fc, G * 64, w->round_tf32, rstd, st);
  }
  RUN(attn_core_run(qbuf, ldq, q_rep, kbuf, vbuf, ldkv, m.w_sk, m.b_sk, G, mask_mode, n_keys, dt, ao, st,
                    w->round_tf32, kv_rows));
  // fc (64 -> 512, no bias) + residual + LayerNorm   (SubLayers.py:97-100)
  // bf16 / split storage: the dedicated streaming kernel (fc_ln.cu); tf32 storage and the training forward (which
  // saves 1/sigma) keep the tcgen05 GEMM with the LayerNorm epilogue.  AITB_FC_GEMM=1 forces the GEMM for A/B runs.
  static const bool force_gemm = getenv("AITB_FC_GEMM") != nullptr;
  if (dt != AITB_F32 && rstd == nullptr && !force_gemm)
    return fc_ln_run(dt, ao, m.w_fc, res, m.ln.gamma, m.ln.beta, 1e-6f, out, G * 64, out_rows != 64 ? 64 : G * 64, out_rows != 64 ? out_rows : G * 64,
                     out_rows != 64 ? 1 : 0, 64, res_rep, st, res_f16, out_f16);
  aitb_gemm_desc d = gemm_base(dt, G * 64, 512, 64, m.w_fc, 512, out, 512, w->round_tf32);
  view_plain(d, ao, 64);
  d.flags = AITB_EPI_RES | AITB_EPI_LN;
  d.res = res;
  d.ldr = 512;
  d.res_div = 64;
  d.res_rep = res_rep;
  d.gamma = m.ln.gamma;
  d.beta = m.ln.beta;
  d.ln_rstd = rstd;
  d.res_f16 = res_f16;
  d.out_f16 = out_f16;
  if (out_rows != 64) {   // compaction: the residual stays indexed by the 64-row GEMM row
    d.rows_in = 64;
    d.rows_out = out_rows;
    d.flags |= AITB_EPI_RES_ROW_M;
  }
  return gemm_run(&d, st);
}

// onepass (precision plan, AITB_F32S): x, the hidden tensor and the output are fp16 planes, both GEMMs run one pass
static int ffn_block(const aitb_head_weights* w, const aitb_ffn& f, const void* x, int M, void* hidden, void* out,
                     cudaStream_t st, float* rstd = nullptr, bool onepass = false, const TrainDrop* td = nullptr,
                     int site = 0) {
  const int dt = w->dtype;
  aitb_gemm_desc d1 = gemm_base(dt, M, 2048, 512, f.w1.w, 256, hidden, 2048, w->round_tf32);
  view_plain(d1, x, 512);
  d1.flags = AITB_EPI_BIAS | AITB_EPI_RELU;
  d1.bias = f.w1.bias;
  if (onepass) {   // the hidden tensor feeds the one-pass w_2 GEMM only: its lo plane is never read
    d1.passes = 1; d1.in_f16 = 1; d1.out_f16 = 1;
    d1.flags |= AITB_EPI_HI_ONLY;
  }
  RUN(gemm_run(&d1, st));
  aitb_gemm_desc d2 = gemm_base(dt, M, 512, 2048, f.w2.w, 512, out, 512, w->round_tf32);
  if (onepass) { d2.passes = 1; d2.in_f16 = 1; d2.out_f16 = 1; d2.res_f16 = 1; }
  view_plain(d2, hidden, 2048);
  if (td) {   // x = LayerNorm(dropout(w_2(.) + b_2) + residual)   (SubLayers.py:181-185)
    d2.round_tf32 = 0;
    d2.block_n = 256;             // 512 is the LayerNorm variant
    d2.flags = AITB_EPI_BIAS;
    d2.bias = f.w2.bias;
    RUN(gemm_run(&d2, st));
    return drop_res_ln_dt(dt, out, nullptr, 64, x, 1, f.ln, *td, site, M, w->round_tf32, rstd, st);
  }
  d2.flags = AITB_EPI_BIAS | AITB_EPI_RES | AITB_EPI_LN;
  d2.bias = f.w2.bias;
  d2.res = x;
  d2.ldr = 512;
  d2.gamma = f.ln.gamma;
  d2.beta = f.ln.beta;
  d2.ln_rstd = rstd;
  return gemm_run(&d2, st);
}

// Decoder part that does not depend on the proposals (system/Models.py:250-253 repeats the
// query per proposal): dec_emb + pos + LN, causal self-attention block, cross-attention query projection.
// Computed once per unit on the side stream while the encoder runs.
static int ait_query_side(const aitb_head_weights* w, HeadBufs& hb, int B, cudaStream_t st, const TrainDrop* td = nullptr) {
  const int dt = w->dtype, cb = colsize(dt), rt = w->round_tf32;
  const int RQ = B * 64;
  aitb_gemm_desc d = gemm_base(dt, RQ, 512, 1024, w->dec_emb.w, 512, hb.T0, 512, rt);
  view_plain(d, hb.qtok, 1024);
  d.bias = w->dec_emb.bias;
  if (td) {   // dec_output = LayerNorm(dropout(emb + pos))   (Models.py:152-153)
    d.round_tf32 = 0;
    d.block_n = 256;
    d.flags = AITB_EPI_BIAS;
    RUN(gemm_run(&d, st));
    RUN(drop_res_ln_dt(dt, hb.T0, w->dec_pos, 64, nullptr, 1, w->dec_ln, *td, AITB_DROP_DEC_EMB, RQ, rt, hb.r_T0, st));
  } else {
    d.flags = AITB_EPI_BIAS | AITB_EPI_POS | AITB_EPI_LN;
    d.pos = w->dec_pos;
    d.pos_rows = 64;
    d.gamma = w->dec_ln.gamma;
    d.beta = w->dec_ln.beta;
    d.ln_rstd = hb.r_T0;
    RUN(gemm_run(&d, st));
  }
  aitb_gemm_desc dq = gemm_base(dt, RQ, 1536, 512, w->dec_slf.w_qkv, 256, hb.QKVd, 1536, rt);
  view_plain(dq, hb.T0, 512);
  RUN(gemm_run(&dq, st));
  const uint8_t* qkv = (const uint8_t*)hb.QKVd;
  RUN(mha_block(w, w->dec_slf, qkv, 1536, 1, qkv + 512 * cb, qkv + 1024 * cb, 1536, B, 1, 64, hb.AOd, hb.T0, 1,
                hb.T1, st, hb.r_T1, 64, 64, 0, 0, td, AITB_DROP_DEC_SLF_ATTN, AITB_DROP_DEC_SLF_FC));
  aitb_gemm_desc dc = gemm_base(dt, RQ, 512, 512, w->dec_enc.w_qkv, 256, hb.Qc, 512, rt);
  view_plain(dc, hb.T1, 512);
  return gemm_run(&dc, st);
}

// `query_ready` (optional): event recorded on the side stream after ait_query_side; waited on before the
// cross attention.  NULL = the query side already ran on `st`.
// compact: after the encoder self-attention the 15 pad rows of every pair are dead (they are masked as keys
// everywhere and never read as queries again: SURVEY fact 7, bit-identical output), so the encoder FFN and the
// cross-attention K / V projection run on 49 rows per pair.  The training path keeps the 64-row layout.
static int ait_core(const aitb_head_weights* w, HeadBufs& hb, int B, int P, void* enc_tap, cudaStream_t st,
                    cudaEvent_t query_ready, bool compact = true, const TrainDrop* td = nullptr) {
  const int dt = w->dtype, eb = esize(dt), cb = colsize(dt), rt = w->round_tf32;
  const int bp = B * P, R = bp * 64;
  const int er = compact ? 49 : 64, RE = bp * er;   // encoder rows per pair / in total after the self-attention block
  // precision plan: the encoder-side tensors pooled / X1 / X2 / Hh / ENC are fp16 planes and the GEMMs that read
  // them run one tensor-core pass on the hi planes (11-bit operands); everything downstream keeps three bf16 passes
  const bool plan = dt == AITB_F32S && (w->plan & AITB_PLAN_ENC_ONEPASS) != 0;
  // ---- encoder input: enc_emb (1x1 conv 1024->512 + bias) on the 49 real rows, + pos, LayerNorm
  {
    aitb_gemm_desc d = gemm_base(dt, bp * 49, 512, 1024, w->enc_emb.w, 512, hb.X1, 512, rt);
    view_plain(d, hb.pooled, 1024);
    d.bias = w->enc_emb.bias;
    d.rows_in = 49;
    d.rows_out = 64;
    if (td) {   // enc_output = LayerNorm(dropout(emb + pos)) (Models.py:98-99); the 15 zero-padded rows of a pair are dropped too
                // (LayerNorm(dropout(pos_t))): as queries of the self-attention they enter its head gate (mean over all 64 rows)
      d.round_tf32 = 0;
      d.block_n = 256;
      d.flags = AITB_EPI_BIAS;
      RUN(gemm_run(&d, st));
      RUN(drop_res_ln_dt(dt, hb.X1, w->enc_pos, 49, nullptr, 1, w->enc_ln, *td, AITB_DROP_ENC_EMB, R, rt, hb.r_X1, st));
    } else {
      d.flags = AITB_EPI_BIAS | AITB_EPI_POS | AITB_EPI_LN;
      d.pos = w->enc_pos;
      d.pos_rows = 64;
      d.gamma = w->enc_ln.gamma;
      d.beta = w->enc_ln.beta;
      d.ln_rstd = hb.r_X1;
      if (plan) { d.passes = 1; d.in_f16 = 1; d.out_f16 = 1; }
      RUN(gemm_run(&d, st));
      RUN(ln_pad_rows(hb.X1, dt, bp, 49, w->enc_pos, w->enc_ln, rt, st, plan ? 1 : 0));
    }
  }
  // ---- encoder self-attention
  {
    aitb_gemm_desc d = gemm_base(dt, R, 1536, 512, w->enc_slf.w_qkv, 256, hb.QKV, 1536, rt);
    view_plain(d, hb.X1, 512);
    if (plan) { d.passes = 1; d.in_f16 = 1; }   // Q / K / V themselves stay bf16 planes (the attention core's format)
    RUN(gemm_run(&d, st));
    const uint8_t* qkv = (const uint8_t*)hb.QKV;
    RUN(mha_block(w, w->enc_slf, qkv, 1536, 1, qkv + 512 * cb, qkv + 1024 * cb, 1536, bp, 0, 49, hb.AO, hb.X1, 1,
                  hb.X2, st, hb.r_X2, er, 64, plan ? 1 : 0, plan ? 1 : 0, td, AITB_DROP_ENC_SLF_ATTN, AITB_DROP_ENC_SLF_FC));
  }
  RUN(ffn_block(w, w->enc_ffn, hb.X2, RE, hb.Hh, hb.ENC, st, hb.r_ENC, plan, td, AITB_DROP_ENC_FFN));
  if (enc_tap) {   // [bp, 64, 512] tap: the rows that exist (pad rows of the tap are left untouched when compact)
    cudaError_t e = cudaMemcpy2DAsync(enc_tap, (size_t)64 * 512 * eb, hb.ENC, (size_t)er * 512 * eb, (size_t)er * 512 * eb, bp,
                                      cudaMemcpyDeviceToDevice, st);
    AITB_REQUIRE(e == cudaSuccess, "enc tap copy failed: %s", cudaGetErrorString(e));
  }
  if (query_ready) {
    cudaError_t e = cudaStreamWaitEvent(st, query_ready, 0);
    AITB_REQUIRE(e == cudaSuccess, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
  }
  // ---- cross attention: K, V from the encoder output of each pair, Q shared by the unit's P pairs
  {
    const uint8_t* wkv = (const uint8_t*)w->dec_enc.w_qkv + (size_t)512 * 512 * eb;
    aitb_gemm_desc d = gemm_base(dt, RE, 1024, 512, wkv, 256, hb.KVc, 1024, rt);
    view_plain(d, hb.ENC, 512);
    if (plan) { d.passes = 1; d.in_f16 = 1; }   // K / V stay bf16 planes
    RUN(gemm_run(&d, st));
    if (compact) {   // the last pair's 64-row K / V tile reads 15 rows past the compact buffer: keep them finite
      cudaError_t e = cudaMemsetAsync((uint8_t*)hb.KVc + (size_t)RE * 1024 * eb, 0, (size_t)(64 - er) * 1024 * eb, st);
      AITB_REQUIRE(e == cudaSuccess, "KV pad memset failed: %s", cudaGetErrorString(e));
    }
    const uint8_t* kv = (const uint8_t*)hb.KVc;
    RUN(mha_block(w, w->dec_enc, hb.Qc, 512, P, kv, kv + 512 * cb, 1024, bp, 0, 49, hb.AOc, hb.T1, P, hb.D1, st,
                  hb.r_D1, 64, er, 0, 0, td, AITB_DROP_DEC_ENC_ATTN, AITB_DROP_DEC_ENC_FC));
  }
  RUN(ffn_block(w, w->dec_ffn, hb.D1, R, hb.Hh2, hb.DEC, st, hb.r_DEC, false, td, AITB_DROP_DEC_FFN));
  // ---- dec_trans (1x1 conv 512->1024 + bias); token-major output == NHWC of [bp,1024,8,8]
  {
    aitb_gemm_desc d = gemm_base(dt, R, 1024, 512, w->dec_trans.w, 256, hb.AIT, 1024, rt);
    view_plain(d, hb.DEC, 512);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->dec_trans.bias;
    RUN(gemm_run(&d, st));
  }
  return 0;
}

// SKBlock (bug-compatible: relu(conv1x1_g8(x))^2 + relu(conv3x3_g8(x))^2, blocks_coatt_transformer_sk.py:973-984)
static int sk_block(const aitb_head_weights* w, const aitb_skblock& sk, const void* x, int G, void* out,
                    cudaStream_t st) {
  // One dual-accumulator grouped GEMM: the nine 3x3 taps accumulate into acc0, the 1x1 conv (a tenth,
  // centre tap reading the same A tile) into acc1; the epilogue emits relu(acc1+b1)^2 + relu(acc0+b3)^2.
  const int dt = w->dtype;
  AITB_REQUIRE(sk.w_fused != nullptr, "SKNet fused weights missing");
  aitb_gemm_desc d = gemm_base(dt, G * 64, 1024, 128, sk.w_fused, 128, out, 1024, w->round_tf32);
  d.a_group_c = 128;
  view_map(d, x, 1024, 8, 8, 1, G);
  taps3x3(d);
  d.dual = 1;
  d.flags = AITB_EPI_BIAS | AITB_EPI_RELU | AITB_EPI_SQUARE | AITB_EPI_DUAL;
  d.bias = sk.conv3x3.bias;
  d.bias2 = sk.conv1x1.bias;
  return gemm_run(&d, st);
}

// ResNet-50 layer4 (3 bottlenecks, stride 2 on the first 1x1, frozen BN folded) on [G, 8, 8, 1024] -> [G, 16, 2048]
static int layer4(const aitb_head_weights* w, const void* x, int G, void* c1, void* c2, void* ds, void* y0, void* y1,
                  void** result, cudaStream_t st) {
  const int dt = w->dtype, rt = w->round_tf32;
  const int M = G * 16;
  const void* in = x;
  void* outs[2] = {y0, y1};
  for (int blk = 0; blk < 3; ++blk) {
    const aitb_bottleneck& bn = w->top[blk];
    const int cin = blk == 0 ? 1024 : 2048;
    void* out = outs[blk & 1];
    aitb_gemm_desc d1 = gemm_base(dt, M, 512, cin, bn.conv1.w, 256, c1, 512, rt);
    if (blk == 0) view_map(d1, in, 1024, 8, 4, 2, G); else view_plain(d1, in, 2048);
    d1.flags = AITB_EPI_BIAS | AITB_EPI_RELU;
    d1.bias = bn.conv1.bias;
    RUN(gemm_run(&d1, st));
    aitb_gemm_desc d2 = gemm_base(dt, M, 512, 512, bn.conv2.w, 256, c2, 512, rt);
    view_map(d2, c1, 512, 4, 4, 1, G);
    taps3x3(d2);
    d2.flags = AITB_EPI_BIAS | AITB_EPI_RELU;
    d2.bias = bn.conv2.bias;
    RUN(gemm_run(&d2, st));
    const void* res = in;
    if (blk == 0) {
      AITB_REQUIRE(bn.down.w != nullptr, "layer4: first bottleneck needs the downsample branch");
      aitb_gemm_desc dd = gemm_base(dt, M, 2048, 1024, bn.down.w, 256, ds, 2048, rt);
      view_map(dd, in, 1024, 8, 4, 2, G);
      dd.flags = AITB_EPI_BIAS;
      dd.bias = bn.down.bias;
      RUN(gemm_run(&dd, st));
      res = ds;
    }
    aitb_gemm_desc d3 = gemm_base(dt, M, 2048, 512, bn.conv3.w, 256, out, 2048, rt);
    view_plain(d3, c2, 512);
    d3.flags = AITB_EPI_BIAS | AITB_EPI_RES | AITB_EPI_RES_RELU;
    d3.bias = bn.conv3.bias;
    d3.res = res;
    d3.ldr = 2048;
    RUN(gemm_run(&d3, st));
    in = out;
  }
  *result = const_cast<void*>(in);
  return 0;
}

static int check_weights(const aitb_head_weights* w, bool with_top) {
  AITB_REQUIRE(w != nullptr, "null weights");
  AITB_REQUIRE(w->dtype == AITB_F32 || w->dtype == AITB_BF16 || w->dtype == AITB_F32S, "bad dtype %d", w->dtype);
  AITB_REQUIRE(w->enc_emb.w && w->enc_emb.bias && w->dec_emb.w && w->dec_emb.bias && w->dec_trans.w &&
                   w->dec_trans.bias && w->enc_pos && w->dec_pos && w->enc_ln.gamma && w->dec_ln.gamma,
               "AIT embedding weights missing");
  const aitb_mha* ms[3] = {&w->enc_slf, &w->dec_slf, &w->dec_enc};
  for (int i = 0; i < 3; ++i)
    AITB_REQUIRE(ms[i]->w_qkv && ms[i]->w_sk && ms[i]->b_sk && ms[i]->w_fc && ms[i]->ln.gamma && ms[i]->ln.beta,
                 "attention block %d weights missing", i);
  const aitb_ffn* fs[2] = {&w->enc_ffn, &w->dec_ffn};
  for (int i = 0; i < 2; ++i)
    AITB_REQUIRE(fs[i]->w1.w && fs[i]->w1.bias && fs[i]->w2.w && fs[i]->w2.bias && fs[i]->ln.gamma, "ffn %d missing", i);
  if (with_top) {
    AITB_REQUIRE(w->sk_props.conv1x1.w && w->sk_props.conv3x3.w && w->sk_query.conv1x1.w && w->sk_query.conv3x3.w,
                 "SKNet weights missing");
    for (int i = 0; i < 3; ++i)
      AITB_REQUIRE(w->top[i].conv1.w && w->top[i].conv2.w && w->top[i].conv3.w, "layer4 block %d weights missing", i);
    AITB_REQUIRE(w->w_bbox && w->b_bbox && w->w_cls1 && w->b_cls1 && w->w_cls2 && w->b_cls2, "head weights missing");
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Training step of the AIT module (config 4): forward keeping activations, then the backward.
// Two configurations: fp32 storage + tf32 tensor-core math (AITB_F32), bf16 storage + bf16 tensor-core math (AITB_BF16);
// fp32 accumulation, LayerNorm statistics and parameter gradients in both.  The reference gets this from torch autograd over
// system/Models.py:231-280; the op order below is the exact reverse of ait_query_side / ait_core.
// ---------------------------------------------------------------------------------------------
static void carve_train(Bump& b, HeadBufs& hb, int B, int P, int es = 4) {
  const size_t bp = (size_t)B * P, R = bp * 64, RQ = (size_t)B * 64;
  memset(&hb, 0, sizeof(hb));
  auto f = [&](size_t n) { return b.take(n * es); };
  auto f4 = [&](size_t n) { return (float*)b.take(n * 4); };
  hb.pooled = f(bp * 49 * 1024);
  hb.qtok = f(RQ * 1024);
  hb.X1 = f(R * 512);
  hb.QKV = f(R * 1536);
  hb.AO = f(R * 64);
  hb.X2 = f(R * 512);
  hb.Hh = f(R * 2048);
  hb.ENC = f(R * 512);
  hb.T0 = f(RQ * 512);
  hb.QKVd = f(RQ * 1536);
  hb.AOd = f(RQ * 64);
  hb.T1 = f(RQ * 512);
  hb.Qc = f(RQ * 512);
  hb.KVc = f(R * 1024);
  hb.AOc = f(R * 64);
  hb.D1 = f(R * 512);
  hb.Hh2 = f(R * 2048);
  hb.DEC = f(R * 512);
  hb.AIT = f(R * 1024);
  hb.r_X1 = f4(R);
  hb.r_X2 = f4(R);
  hb.r_ENC = f4(R);
  hb.r_T0 = f4(RQ);
  hb.r_T1 = f4(RQ);
  hb.r_D1 = f4(R);
  hb.r_DEC = f4(R);
}

// dX[M, Np] = epilogue( dY[M, Kp] * W[Kp, Np] )  with W^T ([Np, Kp], K-major) in `wt`: the forward GEMM kernel
template <typename T>
static int dgrad(const T* dy, int M, int Kp, int Np, const T* wt, T* out, int flags, const T* res,
                 int ldr, cudaStream_t st) {
  const int bn = Np % 256 == 0 ? 256 : (Np % 128 == 0 ? 128 : 64);
  // fp32 storage: gradients are rounded to tf32 (RN) where they are produced -- the next MMA would otherwise truncate them
  aitb_gemm_desc d = gemm_base(DtOf<T>::v, M, Np, Kp, wt, bn, out, Np, sizeof(T) == 4 ? 1 : 0);
  view_plain(d, dy, Kp);
  d.flags = flags;
  d.res = res;
  d.ldr = ldr;
  return gemm_run(&d, st);
}

// W [N, K] -> W^T [K, N]  (the packed weights have the storage type of the activations)
template <typename T>
static int transpose_w(const void* w, int N, int K, T* out, cudaStream_t st) {
  return transpose_run(w, DtOf<T>::v, out, DtOf<T>::v, 1, N, K, 1, st);
}

template <typename T>
struct FfnBwd {
  const aitb_ffn* w;
  const aitb_ffn_g* g;
  const T *x, *hid, *y;              // saved: input [R,512], hidden [R,2048], output (post-LN) [R,512]
  const float* rstd;                 // saved 1/sigma
  const TrainDrop* td;               // dropout between w_2 and the residual add (NULL / p == 0: none)
  int site;
  T* gz;                             // [R,512] scratch: the masked gradient of the w_2 output
};

// backward of y = LN(relu(x W1^T + b1) W2^T + b2 + x); gy -> gx (both [R, 512]); gf / gh are scratch
template <typename T>
static int ffn_backward(const FfnBwd<T>& f, int R, const T* gy, T* gf, T* gh, T* gx, T* w1t, T* w2t,
                        cudaStream_t st) {
  RUN(ln_bwd_run(gy, f.y, f.w->ln.gamma, f.w->ln.beta, f.rstd, R, 64, 64, gf, f.g->ln.gamma, f.g->ln.beta, st));
  // gf = gradient of (dropout(z) + x): the residual path (last line) takes it as is, the w_2 path through the mask
  const T* gz = gf;
  if (f.td && f.td->p > 0.f) {
    RUN(drop_bwd_run(gf, f.gz, f.td->p, f.td->seed, f.site, R, 64, 64, st));
    gz = f.gz;
  }
  RUN(colsum_run(gz, 512, R, 512, f.g->w2.bias, st));
  RUN(wgrad_run(gz, 512, f.hid, 2048, R, 512, 2048, f.g->w2.w, 2048, st));
  RUN(transpose_w(f.w->w2.w, 512, 2048, w2t, st));                       // [2048, 512]
  RUN(dgrad<T>(gz, R, 512, 2048, w2t, gh, AITB_EPI_RELU_MASK, f.hid, 2048, st));
  RUN(colsum_run(gh, 2048, R, 2048, f.g->w1.bias, st));
  RUN(wgrad_run(gh, 2048, f.x, 512, R, 2048, 512, f.g->w1.w, 512, st));
  RUN(transpose_w(f.w->w1.w, 2048, 512, w1t, st));                       // [512, 2048]
  return dgrad<T>(gh, R, 2048, 512, w1t, gx, AITB_EPI_RES, gf, 512, st);
}

}  // namespace aitb

using namespace aitb;

extern "C" {

size_t aitb_ait_saved_bytes(int B, int P) {
  Bump b{nullptr, 0};
  HeadBufs hb;
  carve_train(b, hb, B, P);
  return b.off + 1024;
}

size_t aitb_ait_backward_workspace_bytes(int B, int P) {
  const size_t bp = (size_t)B * P, R = bp * 64, RQ = (size_t)B * 64;
  // per-row widths of the gradient buffers taken in aitb_ait_backward (+ transposed weights, attention dz / s)
  size_t n = R * (1024 + 512 + 512 + 2048 + 512 + 512 + 64 + 512 + 1024 + 512 + 512 + 512 + 64 + 1536 + 512 + 512 /* dropout */) +
             bp * 49 * (512 + 1024) + RQ * (512 * 6 + 64 + 1536 + 1024) + 2 * (bp + B) * (512 + 64) +
             (size_t)4 * 1024 * 1024 /* largest W^T pair */;
  return n * 4 + 64 * 1024 /* alignment slack of ~40 buffers */ + 1024;
}

int aitb_ait_forward_train(const aitb_head_weights* w, const float* x_props, const float* x_query, int B, int P,
                           float* out_nchw, void* saved, size_t saved_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RUN(check_weights(w, false));
  AITB_REQUIRE(w->dtype == AITB_F32 || w->dtype == AITB_BF16,
               "aitb_ait_forward_train: the training path runs in the fp32-storage / tf32 or the bf16 configuration");
  AITB_REQUIRE(B > 0 && P > 0 && x_props && x_query && saved, "aitb_ait_forward_train: bad arguments");
  AITB_REQUIRE(w->dtype == AITB_F32 || out_nchw, "aitb_ait_forward_train: the token-major hand-over (out_nchw == NULL) exists in the fp32-storage configuration only");
  const int dt = w->dtype, es = esize(dt);
  AITB_REQUIRE(((uintptr_t)saved & 1023) == 0 && saved_bytes >= aitb_ait_saved_bytes(B, P),
               "aitb_ait_forward_train: `saved` must be 1024-byte aligned and aitb_ait_saved_bytes large");
  const int bp = B * P;
  Bump b{(uint8_t*)saved, 0};
  HeadBufs hb;
  carve_train(b, hb, B, P, es);
  for (int g0 = 0; g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run(x_props + (size_t)g0 * 1024 * 49, AITB_F32, (uint8_t*)hb.pooled + (size_t)g0 * 49 * 1024 * es, dt, gn,
                      1024, 49, 1, st, w->round_tf32));
  }
  RUN(transpose_run(x_query, AITB_F32, hb.qtok, dt, B, 1024, 64, 1, st, w->round_tf32));
  TrainDrop td;
  const bool drop = train_drop_of(w, td);
  AITB_REQUIRE(td.p >= 0.f && td.p < 1.f && td.p_attn >= 0.f && td.p_attn < 1.f, "aitb_ait_forward_train: dropout probabilities must be in [0, 1)");
  RUN(ait_query_side(w, hb, B, st, drop ? &td : nullptr));
  RUN(ait_core(w, hb, B, P, nullptr, st, nullptr, false, drop ? &td : nullptr));
  // out_nchw == NULL: the caller consumes the token-major result in place (aitb_ait_saved_offset(B, P, 1)) -- no NCHW copy
  for (int g0 = 0; out_nchw && g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run((const uint8_t*)hb.AIT + (size_t)g0 * 64 * 1024 * es, dt, out_nchw + (size_t)g0 * 1024 * 64, AITB_F32,
                      gn, 1024, 64, 0, st));
  }
  return 0;
}

size_t aitb_ait_saved_offset(int B, int P, int which) {
  if (B <= 0 || P <= 0) return 0;
  uint8_t* const base = reinterpret_cast<uint8_t*>((uintptr_t)1 << 20);   // Bump hands out pointers only from a non-null base
  Bump b{base, 0};
  HeadBufs hb;
  carve_train(b, hb, B, P);
  const void* p = which == 1 ? hb.AIT : which == 2 ? hb.Hh : which == 3 ? hb.Hh2 : hb.pooled;
  return (size_t)((const uint8_t*)p - base);
}

}  // extern "C"

template <typename T>
static int ait_backward_impl(const aitb_head_weights* w, const float* grad_out_nchw, int grad_token_major, int B, int P,
                             const void* saved, size_t saved_bytes, const aitb_ait_grads* g, float* grad_props,
                             float* grad_query, void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RUN(check_weights(w, false));
  constexpr int dt = DtOf<T>::v;
  AITB_REQUIRE(w->dtype == dt, "aitb_ait_backward: internal dtype dispatch error");
  AITB_REQUIRE(dt == AITB_F32 || !grad_token_major, "aitb_ait_backward_tm: the token-major hand-over exists in the fp32-storage configuration only");
  AITB_REQUIRE(B > 0 && P > 0 && grad_out_nchw && saved && g && grad_props && grad_query && workspace,
               "aitb_ait_backward: bad arguments");
  AITB_REQUIRE(((uintptr_t)saved & 1023) == 0 && saved_bytes >= aitb_ait_saved_bytes(B, P), "aitb_ait_backward: bad `saved`");
  AITB_REQUIRE(((uintptr_t)workspace & 1023) == 0 && workspace_bytes >= aitb_ait_backward_workspace_bytes(B, P),
               "aitb_ait_backward: workspace too small or misaligned");
  const int bp = B * P, R = bp * 64, RQ = B * 64, R49 = bp * 49;
  Bump sb{(uint8_t*)const_cast<void*>(saved), 0};
  HeadBufs hb;
  carve_train(sb, hb, B, P, (int)sizeof(T));
  Bump b{(uint8_t*)workspace, 0};
  auto f = [&](size_t n) { return (T*)b.take(n * sizeof(T)); };
  const T* S_pooled = (const T*)hb.pooled; const T* S_qtok = (const T*)hb.qtok;
  const T* S_X1 = (const T*)hb.X1; const T* S_QKV = (const T*)hb.QKV; const T* S_AO = (const T*)hb.AO;
  const T* S_X2 = (const T*)hb.X2; const T* S_H1 = (const T*)hb.Hh; const T* S_ENC = (const T*)hb.ENC;
  const T* S_T0 = (const T*)hb.T0; const T* S_QKVd = (const T*)hb.QKVd; const T* S_AOd = (const T*)hb.AOd;
  const T* S_T1 = (const T*)hb.T1; const T* S_Qc = (const T*)hb.Qc; const T* S_KVc = (const T*)hb.KVc;
  const T* S_AOc = (const T*)hb.AOc; const T* S_D1 = (const T*)hb.D1; const T* S_H2 = (const T*)hb.Hh2;
  const T* S_DEC = (const T*)hb.DEC;

  TrainDrop td;
  const bool drop = train_drop_of(w, td);
  const TrainDrop* tdp = drop ? &td : nullptr;
  const bool dp = drop && td.p > 0.f;       // the row-wise sites (the attention-probability sites follow p_attn)
  const DropCfg da_enc = td.attn(AITB_DROP_ENC_SLF_ATTN), da_dec = td.attn(AITB_DROP_DEC_SLF_ATTN),
                da_x = td.attn(AITB_DROP_DEC_ENC_ATTN);
  T* wta = f((size_t)2048 * 1024);   // transposed-weight scratch (two slots)
  T* wtb = f((size_t)2048 * 1024);
  T* gZ = dp ? f((size_t)R * 512) : nullptr;   // masked copy of a LayerNorm-input gradient (one site at a time)
  // ---- dec_trans: AIT = DEC Wt^T + b
  T* gAIT_buf = f((size_t)R * 1024);
  const T* gAIT = grad_token_major ? (const T*)(const void*)grad_out_nchw : gAIT_buf;   // token-major (fp32 storage only): already [bp*64, 1024], tf32-rounded
  for (int g0 = 0; !grad_token_major && g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run(grad_out_nchw + (size_t)g0 * 1024 * 64, AITB_F32, gAIT_buf + (size_t)g0 * 64 * 1024, dt, gn, 1024, 64,
                      1, st, 1));
  }
  RUN(colsum_run(gAIT, 1024, R, 1024, g->dec_trans.bias, st));
  RUN(wgrad_run(gAIT, 1024, S_DEC, 512, R, 1024, 512, g->dec_trans.w, 512, st));
  T* gDEC = f((size_t)R * 512);
  RUN(transpose_w(w->dec_trans.w, 1024, 512, wta, st));                  // [512, 1024]
  RUN(dgrad<T>(gAIT, R, 1024, 512, wta, gDEC, 0, (const T*)nullptr, 0, st));
  // ---- decoder FFN
  T* gF = f((size_t)R * 512);
  T* gH = f((size_t)R * 2048);
  T* gD1 = f((size_t)R * 512);
  {
    FfnBwd<T> fb{&w->dec_ffn, &g->dec_ffn, S_D1, S_H2, S_DEC, hb.r_DEC, tdp, AITB_DROP_DEC_FFN, gZ};
    RUN(ffn_backward(fb, R, gDEC, gF, gH, gD1, wta, wtb, st));
  }
  // ---- cross attention: D1 = LN(AOc Wfc^T + T1[unit])
  T* gC1 = f((size_t)R * 512);
  RUN(ln_bwd_run(gD1, S_D1, w->dec_enc.ln.gamma, w->dec_enc.ln.beta, hb.r_D1, R, 64, 64, gC1, g->dec_enc.ln.gamma,
                 g->dec_enc.ln.beta, st));
  const T* gC1m = gC1;                                               // gradient of fc's output: through the dropout mask
  if (dp) {
    RUN(drop_bwd_run(gC1, gZ, td.p, td.seed, AITB_DROP_DEC_ENC_FC, R, 64, 64, st));
    gC1m = gZ;
  }
  RUN(wgrad_run(gC1m, 512, S_AOc, 64, R, 512, 64, g->dec_enc.w_fc, 64, st));
  T* gAO = f((size_t)R * 64);
  RUN(transpose_w(w->dec_enc.w_fc, 512, 64, wta, st));                   // [64, 512]
  RUN(dgrad<T>(gC1m, R, 512, 64, wta, gAO, 0, (const T*)nullptr, 0, st));
  T* gT1res = f((size_t)RQ * 512);
  RUN(bsum_run(gC1, B, P, 64 * 512, gT1res, st));                        // residual T1 is shared by the unit's P pairs
  T* dQpp = f((size_t)R * 512);
  T* gKVc = f((size_t)R * 1024);
  T* dz = f((size_t)(bp + B) * 512);
  T* sv = f((size_t)(bp + B) * 64);
  RUN(attn_bwd_run(S_Qc, 512, P, S_KVc, S_KVc + 512, 1024, w->dec_enc.w_sk, w->dec_enc.b_sk, gAO, bp, 0, 49, dQpp, 512,
                   gKVc, gKVc + 512, 1024, dz, sv, st, &da_x));
  RUN(wgrad_run(dz, 512, sv, 64, bp, 512, 64, g->dec_enc.w_sk, 64, st));
  RUN(colsum_run(dz, 512, bp, 512, g->dec_enc.b_sk, st));
  T* gQc = f((size_t)RQ * 512);
  RUN(bsum_run(dQpp, B, P, 64 * 512, gQc, st));
  const T* Wq = (const T*)w->dec_enc.w_qkv;
  const T* Wkv = Wq + (size_t)512 * 512;
  RUN(wgrad_run(gKVc, 1024, S_ENC, 512, R, 1024, 512, g->dec_enc.w_qkv + (size_t)512 * 512, 512, st));
  T* gENC = f((size_t)R * 512);
  RUN(transpose_w(Wkv, 1024, 512, wta, st));                             // [512, 1024]
  RUN(dgrad<T>(gKVc, R, 1024, 512, wta, gENC, 0, (const T*)nullptr, 0, st));
  RUN(wgrad_run(gQc, 512, S_T1, 512, RQ, 512, 512, g->dec_enc.w_qkv, 512, st));
  T* gT1 = f((size_t)RQ * 512);
  RUN(transpose_w(Wq, 512, 512, wta, st));
  RUN(dgrad<T>(gQc, RQ, 512, 512, wta, gT1, AITB_EPI_RES, gT1res, 512, st));
  // ---- encoder FFN
  T* gX2 = f((size_t)R * 512);
  {
    FfnBwd<T> fb{&w->enc_ffn, &g->enc_ffn, S_X2, S_H1, S_ENC, hb.r_ENC, tdp, AITB_DROP_ENC_FFN, gZ};
    RUN(ffn_backward(fb, R, gENC, gF, gH, gX2, wta, wtb, st));
  }
  // ---- encoder self attention: X2 = LN(AO Wfc^T + X1)
  T* gA1 = f((size_t)R * 512);
  RUN(ln_bwd_run(gX2, S_X2, w->enc_slf.ln.gamma, w->enc_slf.ln.beta, hb.r_X2, R, 64, 64, gA1, g->enc_slf.ln.gamma,
                 g->enc_slf.ln.beta, st));
  const T* gA1m = gA1;
  if (dp) {
    RUN(drop_bwd_run(gA1, gZ, td.p, td.seed, AITB_DROP_ENC_SLF_FC, R, 64, 64, st));
    gA1m = gZ;
  }
  RUN(wgrad_run(gA1m, 512, S_AO, 64, R, 512, 64, g->enc_slf.w_fc, 64, st));
  RUN(transpose_w(w->enc_slf.w_fc, 512, 64, wta, st));
  RUN(dgrad<T>(gA1m, R, 512, 64, wta, gAO, 0, (const T*)nullptr, 0, st));
  T* gQKV = f((size_t)R * 1536);
  RUN(attn_bwd_run(S_QKV, 1536, 1, S_QKV + 512, S_QKV + 1024, 1536, w->enc_slf.w_sk, w->enc_slf.b_sk, gAO, bp, 0, 49, gQKV,
                   1536, gQKV + 512, gQKV + 1024, 1536, dz, sv, st, &da_enc));
  RUN(wgrad_run(dz, 512, sv, 64, bp, 512, 64, g->enc_slf.w_sk, 64, st));
  RUN(colsum_run(dz, 512, bp, 512, g->enc_slf.b_sk, st));
  RUN(wgrad_run(gQKV, 1536, S_X1, 512, R, 1536, 512, g->enc_slf.w_qkv, 512, st));
  T* gX1 = f((size_t)R * 512);
  RUN(transpose_w(w->enc_slf.w_qkv, 1536, 512, wta, st));                // [512, 1536]
  RUN(dgrad<T>(gQKV, R, 1536, 512, wta, gX1, AITB_EPI_RES, gA1, 512, st));
  // ---- encoder input: X1 = LN(enc_emb(pooled) + b + pos) on the 49 real rows (pad rows: LN(pos) -> dgamma / dbeta only)
  T* gE0 = f((size_t)R49 * 512);
  RUN(ln_bwd_run(gX1, S_X1, w->enc_ln.gamma, w->enc_ln.beta, hb.r_X1, R, 64, 49, gE0, g->enc_ln.gamma, g->enc_ln.beta, st));
  if (dp) RUN(drop_bwd_run(gE0, gE0, td.p, td.seed, AITB_DROP_ENC_EMB, R, 64, 49, st));   // in place, compact 49-row layout
  RUN(colsum_run(gE0, 512, R49, 512, g->enc_emb.bias, st));
  RUN(wgrad_run(gE0, 512, S_pooled, 1024, R49, 512, 1024, g->enc_emb.w, 1024, st));
  T* gPooled = f((size_t)R49 * 1024);
  RUN(transpose_w(w->enc_emb.w, 512, 1024, wta, st));                    // [1024, 512]
  RUN(dgrad<T>(gE0, R49, 512, 1024, wta, gPooled, 0, (const T*)nullptr, 0, st));
  for (int g0 = 0; g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run(gPooled + (size_t)g0 * 49 * 1024, dt, grad_props + (size_t)g0 * 1024 * 49, AITB_F32, gn, 1024, 49,
                      0, st));
  }
  // ---- query side: T1 = LN(AOd Wfc^T + T0), T0 = LN(dec_emb(q) + b + pos), Qc = T1 Wq^T
  T* gS1 = f((size_t)RQ * 512);
  RUN(ln_bwd_run(gT1, S_T1, w->dec_slf.ln.gamma, w->dec_slf.ln.beta, hb.r_T1, RQ, 64, 64, gS1, g->dec_slf.ln.gamma,
                 g->dec_slf.ln.beta, st));
  const T* gS1m = gS1;
  if (dp) {
    RUN(drop_bwd_run(gS1, gZ, td.p, td.seed, AITB_DROP_DEC_SLF_FC, RQ, 64, 64, st));
    gS1m = gZ;
  }
  RUN(wgrad_run(gS1m, 512, S_AOd, 64, RQ, 512, 64, g->dec_slf.w_fc, 64, st));
  T* gAOd = f((size_t)RQ * 64);
  RUN(transpose_w(w->dec_slf.w_fc, 512, 64, wta, st));
  RUN(dgrad<T>(gS1m, RQ, 512, 64, wta, gAOd, 0, (const T*)nullptr, 0, st));
  T* gQKVd = f((size_t)RQ * 1536);
  T* dzd = dz + (size_t)bp * 512;
  T* svd = sv + (size_t)bp * 64;
  RUN(attn_bwd_run(S_QKVd, 1536, 1, S_QKVd + 512, S_QKVd + 1024, 1536, w->dec_slf.w_sk, w->dec_slf.b_sk, gAOd, B, 1, 64, gQKVd,
                   1536, gQKVd + 512, gQKVd + 1024, 1536, dzd, svd, st, &da_dec));
  RUN(wgrad_run(dzd, 512, svd, 64, B, 512, 64, g->dec_slf.w_sk, 64, st));
  RUN(colsum_run(dzd, 512, B, 512, g->dec_slf.b_sk, st));
  RUN(wgrad_run(gQKVd, 1536, S_T0, 512, RQ, 1536, 512, g->dec_slf.w_qkv, 512, st));
  T* gT0 = f((size_t)RQ * 512);
  RUN(transpose_w(w->dec_slf.w_qkv, 1536, 512, wta, st));
  RUN(dgrad<T>(gQKVd, RQ, 1536, 512, wta, gT0, AITB_EPI_RES, gS1, 512, st));
  T* gQ0 = f((size_t)RQ * 512);
  RUN(ln_bwd_run(gT0, S_T0, w->dec_ln.gamma, w->dec_ln.beta, hb.r_T0, RQ, 64, 64, gQ0, g->dec_ln.gamma, g->dec_ln.beta, st));
  if (dp) RUN(drop_bwd_run(gQ0, gQ0, td.p, td.seed, AITB_DROP_DEC_EMB, RQ, 64, 64, st));
  RUN(colsum_run(gQ0, 512, RQ, 512, g->dec_emb.bias, st));
  RUN(wgrad_run(gQ0, 512, S_qtok, 1024, RQ, 512, 1024, g->dec_emb.w, 1024, st));
  T* gQtok = f((size_t)RQ * 1024);
  RUN(transpose_w(w->dec_emb.w, 512, 1024, wta, st));
  RUN(dgrad<T>(gQ0, RQ, 512, 1024, wta, gQtok, 0, (const T*)nullptr, 0, st));
  RUN(transpose_run(gQtok, dt, grad_query, AITB_F32, B, 1024, 64, 0, st));
  AITB_REQUIRE(b.off <= workspace_bytes, "aitb_ait_backward: internal workspace accounting error (%zu > %zu)", b.off,
               workspace_bytes);
  return 0;
}

extern "C" {

int aitb_ait_backward(const aitb_head_weights* w, const float* grad_out_nchw, int B, int P, const void* saved,
                      size_t saved_bytes, const aitb_ait_grads* g, float* grad_props, float* grad_query,
                      void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  if (w && w->dtype == AITB_BF16)
    return ait_backward_impl<bf16_t>(w, grad_out_nchw, 0, B, P, saved, saved_bytes, g, grad_props, grad_query, workspace,
                                     workspace_bytes, stream);
  return ait_backward_impl<float>(w, grad_out_nchw, 0, B, P, saved, saved_bytes, g, grad_props, grad_query, workspace, workspace_bytes,
                                  stream);
}

int aitb_ait_backward_tm(const aitb_head_weights* w, const float* grad_out_tm, int B, int P, const void* saved,
                         size_t saved_bytes, const aitb_ait_grads* g, float* grad_props, float* grad_query,
                         void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  AITB_REQUIRE(((uintptr_t)grad_out_tm & 15) == 0, "aitb_ait_backward_tm: grad_out must be 16-byte aligned");
  AITB_REQUIRE(w && w->dtype == AITB_F32, "aitb_ait_backward_tm: the token-major hand-over exists in the fp32-storage configuration only");
  return ait_backward_impl<float>(w, grad_out_tm, 1, B, P, saved, saved_bytes, g, grad_props, grad_query, workspace, workspace_bytes,
                                  stream);
}

int aitb_ln_bwd(const float* g, const float* y, const float* gamma, const float* beta, const float* rstd, int rows, int grp,
                int valid, float* dx, float* dgamma, float* dbeta, aitb_stream_t stream) {
  return ln_bwd_run(g, y, gamma, beta, rstd, rows, grp, valid, dx, dgamma, dbeta, (cudaStream_t)stream);
}
int aitb_colsum(const float* x, int ld, int rows, int cols, float* out, aitb_stream_t stream) {
  return colsum_run(x, ld, rows, cols, out, (cudaStream_t)stream);
}
int aitb_bsum(const float* x, int B, int P, int L, float* out, aitb_stream_t stream) {
  return bsum_run(x, B, P, L, out, (cudaStream_t)stream);
}
int aitb_attn_bwd(const float* q, int ldq, int q_rep, const float* k, const float* v, int ldkv, const float* w_sk,
                  const float* b_sk, const float* dout, int G, int mask_mode, int n_keys, float* dq, int lddq, float* dk,
                  float* dv, int lddkv, float* dz, float* s, aitb_stream_t stream) {
  return attn_bwd_run(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, dout, G, mask_mode, n_keys, dq, lddq, dk, dv, lddkv, dz, s,
                      (cudaStream_t)stream);
}

}  // extern "C"

namespace aitb {
}  // namespace aitb

using namespace aitb;

extern "C" {

const char* aitb_last_error(void) { return g_err; }
int aitb_version(void) { return 100; }

int aitb_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return 1;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("libaitb200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 1;
  }
  return 0;
}

long long aitb_launch_count(int reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

size_t aitb_nms_workspace_bytes(int B, int n_total, int n) { return nms_workspace_bytes(B, n_total, n); }

int aitb_nms_batched(const float* boxes, const int64_t* order, int B, int n_total, int n, float thr, int max_out,
                     int mode, int64_t* keep_out, int32_t* n_keep, float* rois_out, void* workspace,
                     size_t workspace_bytes, aitb_stream_t stream) {
  return nms_run(boxes, order, B, n_total, n, thr, max_out, mode, keep_out, n_keep, rois_out, workspace,
                 workspace_bytes, (cudaStream_t)stream);
}

size_t aitb_topk_workspace_bytes(int B, int n_total, int n) { return topk_workspace_bytes(B, n_total, n); }

int aitb_topk_desc(const float* scores, int B, int n_total, int n, int64_t* order, void* workspace,
                   size_t workspace_bytes, aitb_stream_t stream) {
  return topk_run(scores, B, n_total, n, order, workspace, workspace_bytes, (cudaStream_t)stream);
}

int aitb_roi_align_forward(const void* feat_nhwc, const float* rois, int B, int C, int H, int W, int K,
                           float spatial_scale, int pooled_h, int pooled_w, int sampling_ratio, int dtype,
                           int out_layout, void* out, aitb_stream_t stream) {
  return roi_align_fwd_run(feat_nhwc, rois, B, C, H, W, K, spatial_scale, pooled_h, pooled_w, sampling_ratio, dtype,
                           out_layout, out, (cudaStream_t)stream);
}

int aitb_roi_align_backward(const float* grad, const float* rois, int B, int C, int H, int W, int K,
                            float spatial_scale, int pooled_h, int pooled_w, int sampling_ratio,
                            float* grad_feat_nhwc, aitb_stream_t stream) {
  return roi_align_bwd_run(grad, rois, B, C, H, W, K, spatial_scale, pooled_h, pooled_w, sampling_ratio,
                           grad_feat_nhwc, (cudaStream_t)stream);
}

int aitb_transpose_cs(const void* src, int src_dtype, void* dst, int dst_dtype, int G, int C, int S,
                      int to_channels_last, aitb_stream_t stream) {
  return transpose_run(src, src_dtype, dst, dst_dtype, G, C, S, to_channels_last, (cudaStream_t)stream);
}

int aitb_transpose_cs_round(const void* src, int src_dtype, void* dst, int dst_dtype, int G, int C, int S,
                            int to_channels_last, int round_tf32, aitb_stream_t stream) {
  return transpose_run(src, src_dtype, dst, dst_dtype, G, C, S, to_channels_last, (cudaStream_t)stream, round_tf32);
}

int aitb_gemm(const aitb_gemm_desc* d, aitb_stream_t stream) { return gemm_run(d, (cudaStream_t)stream); }

int aitb_attn_core(const void* q, int ldq, int q_rep, const void* k, const void* v, int ldkv, const float* w_sk,
                   const float* b_sk, int G, int mask_mode, int n_keys, int dtype, void* out, aitb_stream_t stream) {
  return attn_core_run(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, G, mask_mode, n_keys, dtype, out, (cudaStream_t)stream);
}

int aitb_pool_heads(const void* top, int dtype, int G, int P, const float* qfeat, const float* w_bbox,
                    const float* b_bbox, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* feat_out, float* bbox_out, float* cls_prob_out, aitb_stream_t stream) {
  return pool_heads_run(top, dtype, G, P, qfeat, w_bbox, b_bbox, w1, b1, w2, b2, feat_out, bbox_out, cls_prob_out,
                        (cudaStream_t)stream);
}

size_t aitb_coattention_workspace_bytes(int B, int H, int W) {
  const size_t rows = (size_t)B * H * W, rq = (size_t)B * 64;
  const size_t floats = rows * (1024 * 5 + 64 + 512) + rq * (1024 * 4 + 512 * 4) + (size_t)B * 64 * 2 * 2;
  return floats * 4 + 32 * 1024;
}

int aitb_coattention_forward(const aitb_coatt_weights* w, const float* x_img, const float* x_qry, int B, int H, int W,
                             float* non_img, float* non_qry, void* workspace, size_t workspace_bytes,
                             aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(w && w->dtype == AITB_F32, "aitb_coattention_forward: runs in the fp32-storage / tf32 configuration");
  AITB_REQUIRE(w->w_emb_phi && w->b_emb_phi && w->emb.w && w->emb.bias && w->rho.w && w->rho.bias && w->theta.w &&
                   w->theta.bias && w->omega.w && w->omega.bias && w->theta_gn.gamma && w->theta_gn.beta &&
                   w->omega_gn.gamma && w->omega_gn.beta,
               "aitb_coattention_forward: weights missing");
  AITB_REQUIRE(B > 0 && H > 0 && W > 0 && x_img && x_qry && non_img && non_qry && workspace, "aitb_coattention_forward: bad arguments");
  AITB_REQUIRE(((uintptr_t)workspace & 1023) == 0 && workspace_bytes >= aitb_coattention_workspace_bytes(B, H, W),
               "aitb_coattention_forward: workspace too small or misaligned");
  const int Ni = H * W, rows = B * Ni, rq = B * 64, rt = 1;
  Bump b{(uint8_t*)workspace, 0};
  auto f = [&](size_t n) { return (float*)b.take(n * 4); };
  float* Xi = f((size_t)rows * 1024);    // exact tokens (residual)
  float* Xir = f((size_t)rows * 1024);   // tf32-rounded tokens (MMA operand)
  float* EP = f((size_t)rows * 1024);    // [emb | phi] of the image
  float* coT = f((size_t)rows * 64);     // co_attention^T per image: [Ni, 64]
  float* NI = f((size_t)rows * 512);
  float* TH = f((size_t)rows * 1024);
  float* Xq = f((size_t)rq * 1024);
  float* Xqr = f((size_t)rq * 1024);
  float* Eq = f((size_t)rq * 512);
  float* Rq = f((size_t)rq * 512);
  float* EqT = f((size_t)rq * 512);
  float* NQT = f((size_t)rq * 512);
  float* NQ = Eq;                        // Eq is dead once EqT exists
  float* OM = f((size_t)rq * 1024);
  float* OUTq = f((size_t)rq * 1024);
  double* sums = (double*)b.take((size_t)B * 32 * 2 * sizeof(double) * 2);
  float* OUTi = EP;                      // EP is dead after the per-image products

  RUN(transpose_run(x_img, AITB_F32, Xi, AITB_F32, B, 1024, Ni, 1, st, 0));
  RUN(transpose_run(x_img, AITB_F32, Xir, AITB_F32, B, 1024, Ni, 1, st, rt));
  RUN(transpose_run(x_qry, AITB_F32, Xq, AITB_F32, B, 1024, 64, 1, st, 0));
  RUN(transpose_run(x_qry, AITB_F32, Xqr, AITB_F32, B, 1024, 64, 1, st, rt));
  {  // emb | phi of the image, one GEMM  (blocks_coatt...:70-79)
    aitb_gemm_desc d = gemm_base(AITB_F32, rows, 1024, 1024, w->w_emb_phi, 256, EP, 1024, rt);
    view_plain(d, Xir, 1024);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->b_emb_phi;
    RUN(gemm_run(&d, st));
  }
  {  // emb and rho of the query (:73-77)
    aitb_gemm_desc d = gemm_base(AITB_F32, rq, 512, 1024, w->emb.w, 256, Eq, 512, rt);
    view_plain(d, Xqr, 1024);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->emb.bias;
    RUN(gemm_run(&d, st));
    aitb_gemm_desc d2 = gemm_base(AITB_F32, rq, 512, 1024, w->rho.w, 256, Rq, 512, rt);
    view_plain(d2, Xqr, 1024);
    d2.flags = AITB_EPI_BIAS;
    d2.bias = w->rho.bias;
    RUN(gemm_run(&d2, st));
  }
  RUN(transpose_run(Eq, AITB_F32, EqT, AITB_F32, B, 64, 512, 1, st, 0));   // [B, 64, 512] -> [B, 512, 64]
  cudaError_t e = cudaMemsetAsync(NQT, 0, (size_t)rq * 512 * 4, st);
  AITB_REQUIRE(e == cudaSuccess, "aitb_coattention_forward: memset failed: %s", cudaGetErrorString(e));
  for (int i = 0; i < B; ++i) {
    float* EPb = EP + (size_t)i * Ni * 1024;
    float* coTb = coT + (size_t)i * Ni * 64;
    {  // co_attention^T = phi_img^T-major product: [Ni, 64] = phi_img[Ni, 512] . rho_qry[64, 512]^T  (:81)
      aitb_gemm_desc d = gemm_base(AITB_F32, Ni, 64, 512, Rq + (size_t)i * 64 * 512, 64, coTb, 64, rt);
      view_plain(d, EPb + 512, 1024);
      RUN(gemm_run(&d, st));
    }
    {  // non_img = (co^T / N_q) . emb_qry: [Ni, 512]  (:92-93,99)
      aitb_gemm_desc d = gemm_base(AITB_F32, Ni, 512, 64, EqT + (size_t)i * 512 * 64, 256, NI + (size_t)i * Ni * 512, 512, rt);
      view_plain(d, coTb, 64);
      d.out_scale = 1.f / 64.f;
      RUN(gemm_run(&d, st));
    }
    // non_qry^T = emb_img^T . co^T: [512, 64], contraction over the Ni image positions (:104) -- the MN-major
    // row-contraction kernel of the training path; the 1 / N_i of :91 is applied by the omega GEMM
    RUN(wgrad_run(EPb, 1024, coTb, 64, Ni, 512, 64, NQT + (size_t)i * 512 * 64, 64, st));
  }
  {  // theta: 1x1 conv 512 -> 1024, GroupNorm(32), + identity_img  (:100-102)
    aitb_gemm_desc d = gemm_base(AITB_F32, rows, 1024, 512, w->theta.w, 256, TH, 1024, 0);
    view_plain(d, NI, 512);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->theta.bias;
    RUN(gemm_run(&d, st));
    RUN(group_norm_residual_run(TH, Xi, w->theta_gn.gamma, w->theta_gn.beta, B, Ni, 32, 1e-5f, sums, OUTi, st));
    RUN(transpose_run(OUTi, AITB_F32, non_img, AITB_F32, B, 1024, Ni, 0, st, 0));
  }
  {  // omega on the query side (:104-110)
    RUN(transpose_run(NQT, AITB_F32, NQ, AITB_F32, B, 512, 64, 1, st, rt));   // [B, 512, 64] -> [B, 64, 512]
    aitb_gemm_desc d = gemm_base(AITB_F32, rq, 1024, 512, w->omega.w, 256, OM, 1024, 0);
    view_plain(d, NQ, 512);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->omega.bias;
    d.out_scale = 1.f / (float)Ni;
    RUN(gemm_run(&d, st));
    RUN(group_norm_residual_run(OM, Xq, w->omega_gn.gamma, w->omega_gn.beta, B, 64, 32, 1e-5f, sums + (size_t)B * 64, OUTq, st));
    RUN(transpose_run(OUTq, AITB_F32, non_qry, AITB_F32, B, 1024, 64, 0, st, 0));
  }
  AITB_REQUIRE(b.off <= workspace_bytes, "aitb_coattention_forward: internal workspace accounting error");
  return 0;
}

size_t aitb_rpn_workspace_bytes(int B, int H, int W, int dtype) {
  const size_t eb = esize(dtype), rows = (size_t)B * H * W;
  return ((rows * 1024 * eb + 1023) & ~(size_t)1023) + ((rows * 512 * eb + 1023) & ~(size_t)1023) +
         ((rows * 128 * eb + 1023) & ~(size_t)1023) + 1024;
}

int aitb_rpn_forward(const aitb_rpn_weights* w, const float* feat_nchw, int B, int H, int W, const float* base_anchors,
                     const float* im_info, float feat_stride, float* proposals, float* fg_scores, float* rpn_cls_prob,
                     float* rpn_bbox_pred, void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AITB_REQUIRE(w && w->conv.w && w->conv.bias && w->heads.w && w->heads.bias, "aitb_rpn_forward: weights missing");
  AITB_REQUIRE(w->dtype == AITB_F32 || w->dtype == AITB_BF16 || w->dtype == AITB_F32S, "aitb_rpn_forward: bad dtype");
  AITB_REQUIRE(w->A > 0 && (w->n_pad == 64 || w->n_pad == 128) && 6 * w->A <= w->n_pad, "aitb_rpn_forward: bad head width");
  AITB_REQUIRE(B > 0 && H > 0 && W > 0 && W <= 128, "aitb_rpn_forward: map %dx%d unsupported (width <= 128)", H, W);
  AITB_REQUIRE(feat_nchw && workspace && ((uintptr_t)workspace & 1023) == 0 &&
                   workspace_bytes >= aitb_rpn_workspace_bytes(B, H, W, w->dtype),
               "aitb_rpn_forward: bad input / workspace");
  const int dt = w->dtype, rows = B * H * W;
  const size_t eb = esize(dt);
  Bump b{(uint8_t*)workspace, 0};
  void* featT = b.take((size_t)rows * 1024 * eb);
  void* conv = b.take((size_t)rows * 512 * eb);
  void* heads = b.take((size_t)rows * 128 * eb);
  RUN(transpose_run(feat_nchw, AITB_F32, featT, dt, B, 1024, H * W, 1, st, w->round_tf32));
  {  // RPN_Conv 3x3 + bias + ReLU: the map is tiled by boxes of bx x by = 128 positions of one image
    const int bx = W <= 64 ? 64 : 128, by = 128 / bx;
    const int tpg = (H + by - 1) / by;
    aitb_gemm_desc d = gemm_base(dt, B * tpg * 128, 512, 1024, w->conv.w, 256, conv, 512, w->round_tf32);
    d.a.ptr = featT;
    d.a.dims[0] = 1024;
    if (dt == AITB_F32S) {
      d.a.dims[0] = 2048;
      d.a_lo_off = 1024;
    }
    d.a.dims[1] = (uint64_t)W;
    d.a.dims[2] = (uint64_t)H;
    d.a.dims[3] = (uint64_t)B;
    d.a.strides[0] = (uint64_t)1024 * eb;
    d.a.strides[1] = (uint64_t)W * 1024 * eb;
    d.a.strides[2] = (uint64_t)H * W * 1024 * eb;
    d.a.box[0] = 128 / colsize(dt);
    d.a.box[1] = (uint32_t)bx;
    d.a.box[2] = (uint32_t)by;
    d.a.box[3] = 1;
    d.a_m_dim = 2;
    d.a_m_step = 1;
    d.map_w = W;
    d.map_h = H;
    taps3x3(d);
    d.flags = AITB_EPI_BIAS | AITB_EPI_RELU;
    d.bias = w->conv.bias;
    RUN(gemm_run(&d, st));
  }
  {  // RPN_cls_score | RPN_bbox_pred as one 1x1 GEMM
    aitb_gemm_desc d = gemm_base(dt, rows, w->n_pad, 512, w->heads.w, w->n_pad, heads, w->n_pad, 0);
    view_plain(d, conv, 512);
    d.flags = AITB_EPI_BIAS;
    d.bias = w->heads.bias;
    RUN(gemm_run(&d, st));
  }
  return rpn_head_decode_run(heads, dt, w->n_pad, base_anchors, im_info, B, w->A, H, W, feat_stride, proposals, fg_scores,
                             rpn_cls_prob, rpn_bbox_pred, st);
}

int aitb_rpn_decode(const float* scores_nchw, const float* deltas_nchw, const float* base_anchors, const float* im_info,
                    int B, int A, int H, int W, float feat_stride, float* proposals, float* fg_scores,
                    aitb_stream_t stream) {
  return rpn_decode_run(scores_nchw, deltas_nchw, base_anchors, im_info, B, A, H, W, feat_stride, proposals, fg_scores,
                        (cudaStream_t)stream);
}
int aitb_box_decode(const float* boxes, int boxes_stride, int boxes_off, const float* deltas, const float* cls,
                    const float* im_info, int B, int N, const float* h_stds, const float* h_means, float thresh,
                    int divide_by_scale, float* pred, float* key, int32_t* n_valid, aitb_stream_t stream) {
  return box_decode_run(boxes, boxes_stride, boxes_off, deltas, cls, im_info, B, N, h_stds, h_means, thresh,
                        divide_by_scale, pred, key, n_valid, (cudaStream_t)stream);
}
int aitb_det_assemble(const float* pred, const float* cls, const int64_t* order, const int64_t* keep_pos,
                      const int32_t* n_keep, const int32_t* n_valid, int B, int N, int max_per_image, float* dets,
                      int32_t* n_det, aitb_stream_t stream) {
  return det_assemble_run(pred, cls, order, keep_pos, n_keep, n_valid, B, N, max_per_image, dets, n_det,
                          (cudaStream_t)stream);
}

int aitb_wgrad(const float* dy, int ldy, const float* x, int ldx, int M, int N, int K, float* dw, int ldw,
               aitb_stream_t stream) {
  return wgrad_run(dy, ldy, x, ldx, M, N, K, dw, ldw, (cudaStream_t)stream);
}
int aitb_wgrad_bf16(const void* dy, int ldy, const void* x, int ldx, int M, int N, int K, float* dw, int ldw,
                    aitb_stream_t stream) {
  return wgrad_run((const bf16_t*)dy, ldy, (const bf16_t*)x, ldx, M, N, K, dw, ldw, (cudaStream_t)stream);
}

int aitb_wgrad_conv(const float* dy, int ldy, const float* x, int G, int S, int C, int N, int groups, int taps, float* dw,
                    int ldw, aitb_stream_t stream) {
  return wgrad_conv_run(dy, ldy, x, G, S, C, N, groups, taps, dw, ldw, (cudaStream_t)stream);
}

size_t aitb_head_workspace_bytes(int B, int P, int dtype) {
  Bump b{nullptr, 0};
  HeadBufs hb;
  carve(b, hb, B, P, 64 * 64, dtype, true, true);  // map budget: up to 64x64 cells (1024x1024 px input)
  return b.off + 1024;
}

size_t aitb_ait_workspace_bytes(int B, int P, int dtype) {
  Bump b{nullptr, 0};
  HeadBufs hb;
  carve(b, hb, B, P, 0, dtype, false, false);
  return b.off + 1024;
}

int aitb_head_forward(const aitb_head_weights* w, const float* feat_nchw, int H, int W, const float* query_nchw,
                      const float* rois, int B, int P, float* cls_prob, float* bbox_pred, const aitb_head_taps* taps,
                      void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RUN(check_weights(w, true));
  AITB_REQUIRE(B > 0 && P > 0 && H > 0 && W > 0, "aitb_head_forward: bad sizes B=%d P=%d H=%d W=%d", B, P, H, W);
  AITB_REQUIRE(H * W <= 64 * 64, "aitb_head_forward: map %dx%d exceeds the 4096-cell workspace budget", H, W);
  AITB_REQUIRE(feat_nchw && query_nchw && rois && cls_prob && bbox_pred && workspace, "aitb_head_forward: null pointer");
  AITB_REQUIRE(((uintptr_t)workspace & 1023) == 0, "aitb_head_forward: workspace must be 1024-byte aligned");
  AITB_REQUIRE(workspace_bytes >= aitb_head_workspace_bytes(B, P, w->dtype), "aitb_head_forward: workspace too small");
  const int dt = w->dtype, eb = esize(dt);
  const int bp = B * P;
  Bump b{(uint8_t*)workspace, 0};
  HeadBufs hb;
  carve(b, hb, B, P, 64 * 64, dt, true, true);

  // query branch on the side stream: decoder self-attention block, SKNet(query) (RCNN_top(query) rides in the proposal launches).
  // Forked BEFORE the map transpose / ROIAlign: those kernels use no shared memory, so the query side's small GEMM CTAs find room
  // next to them instead of squeezing between the persistent full-shared-memory GEMMs of the encoder
  RUN(transpose_run(query_nchw, AITB_F32, hb.qtok, dt, B, 1024, 64, 1, st, w->round_tf32));
  SideStream* ss = nullptr;
  RUN(side_stream(&ss));
  float* qfeat = (taps && taps->qfeat) ? taps->qfeat : hb.qfeat;
  auto fork_query = [&]() -> int {
    cudaError_t e = cudaEventRecord(ss->fork, st);
    AITB_REQUIRE(e == cudaSuccess, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    e = cudaStreamWaitEvent(ss->stream, ss->fork, 0);
    AITB_REQUIRE(e == cudaSuccess, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
    RUN(ait_query_side(w, hb, B, ss->stream));
    cudaEventRecord(ss->q1, ss->stream);
    RUN(sk_block(w, w->sk_query, hb.qtok, B, hb.SKq, ss->stream));
    cudaEventRecord(ss->q2, ss->stream);
    return 0;
  };
  static const bool fork_late = getenv("AITB_QUERY_FORK_LATE") != nullptr;   // A/B switch: fork after ROIAlign (round-1 order)
  if (!fork_late) RUN(fork_query());
  // a3: ROIAlign from a channels-last copy of the map, token-major output feeding enc_emb
  // exact copy (fp32 in the F32 and F32S configurations): ROIAlign parity
  RUN(transpose_run(feat_nchw, AITB_F32, hb.featT, dt == AITB_BF16 ? AITB_BF16 : AITB_F32, B, 1024, H * W, 1, st));
  for (int k0 = 0; k0 < bp; k0 += 32768) {
    const int kn = bp - k0 < 32768 ? bp - k0 : 32768;
    RUN(roi_align_fwd_run(hb.featT, rois + (size_t)k0 * 5, B, 1024, H, W, kn, 1.f / 16.f, 7, 7, 0, dt, 1,
                          (uint8_t*)hb.pooled + (size_t)k0 * 49 * 1024 * eb, st, taps && taps->pooled ? 0 : w->round_tf32,
                          (dt == AITB_F32S && (w->plan & AITB_PLAN_ENC_ONEPASS)) ? 1 : 0));
  }
  if (taps && taps->pooled) {
    cudaError_t e = cudaMemcpyAsync(taps->pooled, hb.pooled, (size_t)bp * 49 * 1024 * eb, cudaMemcpyDeviceToDevice, st);
    AITB_REQUIRE(e == cudaSuccess, "pooled tap copy failed: %s", cudaGetErrorString(e));
  }
  if (fork_late) RUN(fork_query());
  // a5-a9: AIT
  RUN(ait_core(w, hb, B, P, taps ? taps->enc_out : nullptr, st, ss->q1));
  if (taps && taps->ait_out) {
    cudaError_t e = cudaMemcpyAsync(taps->ait_out, hb.AIT, (size_t)bp * 64 * 1024 * eb, cudaMemcpyDeviceToDevice, st);
    AITB_REQUIRE(e == cudaSuccess, "ait tap copy failed: %s", cudaGetErrorString(e));
  }
  // a10: SKNet, proposal branch
  RUN(sk_block(w, w->sk_props, hb.AIT, bp, hb.SK, st));
  if (taps && taps->sk_out) {
    cudaError_t e = cudaMemcpyAsync(taps->sk_out, hb.SK, (size_t)bp * 64 * 1024 * eb, cudaMemcpyDeviceToDevice, st);
    AITB_REQUIRE(e == cudaSuccess, "sk tap copy failed: %s", cudaGetErrorString(e));
  }
  // a11: RCNN_top over the bp proposal maps and the B query maps in one set of launches (join the query branch first)
  {
    cudaError_t e = cudaStreamWaitEvent(st, ss->q2, 0);
    AITB_REQUIRE(e == cudaSuccess, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
  }
  void* ytop = nullptr;
  RUN(layer4(w, hb.SK, bp + B, hb.c1, hb.c2, hb.ds, hb.y0, hb.y1, &ytop, st));
  // a12: spatial mean of the query maps, then spatial mean + bbox / similarity heads of the pairs
  RUN(pool_heads_run((const uint8_t*)ytop + (size_t)bp * 16 * 2048 * eb, dt, B, 1, nullptr, nullptr, nullptr, nullptr, nullptr,
                     nullptr, nullptr, qfeat, nullptr, nullptr, st));
  RUN(pool_heads_run(ytop, dt, bp, P, qfeat, w->w_bbox, w->b_bbox, w->w_cls1, w->b_cls1, w->w_cls2, w->b_cls2,
                     taps ? taps->feat : nullptr, bbox_pred, cls_prob, st));
  return 0;
}

int aitb_ait_forward(const aitb_head_weights* w, const float* x_props, const float* x_query, int B, int P,
                     float* out_nchw, void* workspace, size_t workspace_bytes, aitb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RUN(check_weights(w, false));
  AITB_REQUIRE(B > 0 && P > 0, "aitb_ait_forward: bad sizes B=%d P=%d", B, P);
  AITB_REQUIRE(x_props && x_query && out_nchw && workspace, "aitb_ait_forward: null pointer");
  AITB_REQUIRE(((uintptr_t)workspace & 1023) == 0, "aitb_ait_forward: workspace must be 1024-byte aligned");
  AITB_REQUIRE(workspace_bytes >= aitb_ait_workspace_bytes(B, P, w->dtype), "aitb_ait_forward: workspace too small");
  const int dt = w->dtype;
  const int bp = B * P;
  Bump b{(uint8_t*)workspace, 0};
  HeadBufs hb;
  carve(b, hb, B, P, 0, dt, false, false);
  for (int g0 = 0; g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run(x_props + (size_t)g0 * 1024 * 49, AITB_F32, (uint8_t*)hb.pooled + (size_t)g0 * 49 * 1024 * esize(dt),
                      dt, gn, 1024, 49, 1, st, w->round_tf32, (dt == AITB_F32S && (w->plan & AITB_PLAN_ENC_ONEPASS)) ? 1 : 0));
  }
  RUN(transpose_run(x_query, AITB_F32, hb.qtok, dt, B, 1024, 64, 1, st, w->round_tf32));
  RUN(ait_query_side(w, hb, B, st));
  RUN(ait_core(w, hb, B, P, nullptr, st, nullptr));
  for (int g0 = 0; g0 < bp; g0 += 32768) {
    const int gn = bp - g0 < 32768 ? bp - g0 : 32768;
    RUN(transpose_run((const uint8_t*)hb.AIT + (size_t)g0 * 64 * 1024 * esize(dt), dt, out_nchw + (size_t)g0 * 1024 * 64,
                      AITB_F32, gn, 1024, 64, 0, st));
  }
  return 0;
}

}  // extern "C"
