/*
 * libaitb200 -- C ABI of the B200-native (sm_100a) AIT detection-head hot path.
 *
 * Drop-in boundary for CAIVIAC/AIT's native extension `model._C` (pybind11 module defined in
 * lib/model/csrc/vision.cpp:7-13) and for the per-(image, query) head that the detector runs on
 * top of it (lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:273-337).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named `h_*`
 *   - the library never allocates or frees device memory and never synchronises the device:
 *     the caller passes outputs, a workspace and the stream
 *   - every entry point returns 0 on success; on failure it returns non-zero and
 *     aitb_last_error() (thread-local) describes why.  There is no CPU fallback.
 *   - re-entrant; no global mutable state except the thread-local error string and the
 *     lazily resolved driver entry point used to encode TMA descriptors
 */
#ifndef AITB200_H_
#define AITB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef void* aitb_stream_t; /* cudaStream_t */

/* Compute configurations ("dtype" arguments):
 *   AITB_F32   fp32 storage, tf32 tensor-core math (10-bit operand mantissa): the fast fp32-storage mode
 *   AITB_BF16  bf16 storage, bf16 math
 *   AITB_F32S  "fp32-class" split mode: every activation / weight x is stored as TWO bf16 planes
 *              hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together) and every product runs as three
 *              bf16 tensor-core passes hi*hi + hi*lo + lo*hi with fp32 accumulation (error ~2^-16 per
 *              operand instead of tf32's 2^-11).  This is the configuration that meets the reference's
 *              fp32 results to 1e-3 on cls_prob.  Storage: a logical [rows, L] matrix is a bf16
 *              [rows, 2L] matrix, hi plane in columns [0, L), lo plane in [L, 2L)  (4 bytes / element). */
enum { AITB_F32 = 0, AITB_BF16 = 1, AITB_F32S = 2 };

const char* aitb_last_error(void);
int aitb_version(void);
/* returns 0 when the current device is sm_100 (compute capability 10.x) */
int aitb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * a1/a2  NMS  -- replaces `_C.nms` (csrc/nms.h:10-28 -> nms_cuda, csrc/cuda/nms.cu:70-131) and
 * the per-image sort/top-N/NMS/pad loop of _ProposalLayer.forward (rpn/proposal_layer.py:129-166).
 *
 *   boxes   [B, n_total, 4] f32 (x1,y1,x2,y2), legacy +1 area convention (nms.cu:13-21)
 *   order   [B, n] int64: per image, indices into n_total of the n candidates in descending
 *           score order (NULL = boxes are already sorted and n == n_total)
 *   thr     suppress when IoU >  thr  (CUDA semantics, nms.cu:60)
 *   mode 0  "proposal": keep_out[b, 0..k) = positions (in score order) of the first
 *           min(kept, max_out) survivors; the greedy scan stops early once max_out are kept
 *   mode 1  "reference nms()": keep_out[b, 0..k) = ORIGINAL indices of all survivors in
 *           ascending index order (nms.cu:127-130); max_out must be >= n
 *   keep_out [B, max_out] int64 (entries beyond n_keep[b] are set to -1), n_keep [B] int32
 *   rois_out (optional, mode 0) [B, max_out, 5] f32: rows (b, x1,y1,x2,y2) of the survivors,
 *           zero-padded, column 0 always = b (proposal_layer.py:160-164)
 * ---------------------------------------------------------------------------------------- */
size_t aitb_nms_workspace_bytes(int B, int n_total, int n);
int aitb_nms_batched(const float* boxes, const int64_t* order, int B, int n_total, int n,
                     float thr, int max_out, int mode, int64_t* keep_out, int32_t* n_keep,
                     float* rois_out, void* workspace, size_t workspace_bytes,
                     aitb_stream_t stream);

/* f1 (next row): K*A-wide descending selection of the top-n scores per image (stable: ties by
 * lower index first), replacing torch.sort + [:pre_nms_topN] (proposal_layer.py:129,144-145).
 *   scores [B, n_total] f32 -> order [B, n] int64 */
size_t aitb_topk_workspace_bytes(int B, int n_total, int n);
int aitb_topk_desc(const float* scores, int B, int n_total, int n, int64_t* order,
                   void* workspace, size_t workspace_bytes, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a3  ROIAlign forward -- replaces `_C.roi_align_forward`
 * (csrc/ROIAlign.h:11-26 -> ROIAlign_forward_cuda, csrc/cuda/ROIAlign_cuda.cu:257-299).
 *   feat_nhwc [B, H, W, C]  channels-last copy of the map (see aitb_nchw_to_nhwc), dtype `dtype`
 *   rois      [K, 5] f32 (batch_idx, x1, y1, x2, y2)
 *   out_layout 0: NCHW [K, C, ph, pw] (the reference layout), 1: token-major [K, ph*pw, C]
 *   out dtype = `dtype`.  sampling_ratio <= 0 selects the adaptive grid ceil(roi/pooled).
 * ---------------------------------------------------------------------------------------- */
int aitb_roi_align_forward(const void* feat_nhwc, const float* rois, int B, int C, int H, int W,
                           int K, float spatial_scale, int pooled_h, int pooled_w,
                           int sampling_ratio, int dtype, int out_layout, void* out,
                           aitb_stream_t stream);

/* a4  ROIAlign backward -- replaces `_C.roi_align_backward` (ROIAlign_cuda.cu:302-346).
 *   grad [K, C, ph, pw] f32 NCHW -> grad_feat_nhwc [B, H, W, C] f32 (must be zeroed by caller),
 *   accumulated with coalesced atomics; convert with aitb_nhwc_to_nchw afterwards. */
int aitb_roi_align_backward(const float* grad, const float* rois, int B, int C, int H, int W,
                            int K, float spatial_scale, int pooled_h, int pooled_w,
                            int sampling_ratio, float* grad_feat_nhwc, aitb_stream_t stream);

/* layout helpers: [G, C, S] <-> [G, S, C] with optional dtype conversion (src/dst dtype enums) */
int aitb_transpose_cs(const void* src, int src_dtype, void* dst, int dst_dtype, int G, int C,
                      int S, int to_channels_last, aitb_stream_t stream);
/* same, fp32 destination rounded to tf32 (nearest) when round_tf32 != 0: the consumer is a tf32 tensor-core GEMM */
int aitb_transpose_cs_round(const void* src, int src_dtype, void* dst, int dst_dtype, int G, int C, int S,
                            int to_channels_last, int round_tf32, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * tcgen05 GEMM building block (used by the head engine below; exported for unit tests)
 *   out[M, N] = epilogue( A[M, K] * W[N, K]^T ),  A through a (<=4-D) strided view so that
 *   3x3 / strided convolutions are expressed as shifted TMA boxes with zero fill.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  uint64_t dims[4];    /* elements, dims[0] innermost (K / channels) */
  uint64_t strides[3]; /* bytes, for dims[1..3] */
  uint32_t box[4];     /* TMA box; box[0]*elem = 128 B, box[1]*box[2]*box[3] = 128 rows */
} aitb_view4;

enum {
  AITB_EPI_BIAS = 1,
  AITB_EPI_RELU = 2,
  AITB_EPI_SQUARE = 4,  /* v = v*v (after relu)                                   */
  AITB_EPI_RES = 8,     /* v += residual[res_row(out_row), n]                     */
  AITB_EPI_POS = 16,    /* v += pos[out_row % pos_rows, n]          (fp32 table)  */
  AITB_EPI_LN = 32,     /* LayerNorm over the full row (requires block_n == N == 512) */
  AITB_EPI_ACCUM = 64,  /* v += out (read-modify-write)                            */
  AITB_EPI_RES_RELU = 128, /* relu applied AFTER the residual add (bottleneck tail) */
  AITB_EPI_DUAL = 256,     /* two accumulators (see `dual`): v = f(acc0 + bias) + f(acc1 + bias2),
                              f = the RELU / SQUARE flags (SKBlock: relu(conv1x1)^2 + relu(conv3x3)^2) */
  AITB_EPI_RELU_MASK = 512, /* backward of ReLU: v = residual[res_row, n] > 0 ? v : 0  (`res` = the saved activation;
                              excludes RES / LN / split) */
  AITB_EPI_RES_ROW_M = 1024, /* the residual row is derived from the GEMM row m instead of the (remapped) output row */
  AITB_EPI_HI_ONLY = 2048    /* AITB_F32S: only the hi plane of the output is consumed (by a one-pass GEMM); the 2-CTA one-pass
                                kernel then skips the lo plane (other kernels write both) */
};

typedef struct {
  int dtype; /* AITB_F32 (tf32 MMA) | AITB_BF16 | AITB_F32S (split bf16 x3; the view, ldo and ldr are given in
                LOGICAL elements of 4 bytes, i.e. strides in bytes of the physical two-plane rows) */
  int M, N, k_per_tap, taps;
  aitb_view4 a;
  int a_m_dim;   /* which A coordinate advances with the m-tile: 1 (plain rows), 3 (small maps: whole maps per tile) or
                    2 (large map, see map_w / map_h) */
  int a_m_step;  /* coordinate step per 128-row m-tile                                      */
  int a_group_c; /* grouped conv: input-channel offset per n-tile (0 otherwise)             */
  int8_t tap_dx[9], tap_dy[9];
  const void* w; /* [N, taps*k_per_tap] K-major, dtype `dtype` (F32S: [N, hi plane | lo plane]) */
  int block_n;   /* 128 | 256 | 512 */
  /* epilogue */
  int flags;
  void* out;
  int ldo;
  int rows_in, rows_out; /* out_row = (m / rows_in) * rows_out + m % rows_in; with rows_out < rows_in only the first
                            rows_out rows of every group of rows_in are stored (compaction) */
  const float* bias;
  const void* res;
  int ldr;
  int res_div, res_rep; /* res_row = ((out_row / res_div) / res_rep) * res_div + out_row % res_div */
  const float* pos;
  int pos_rows;
  const float* gamma;
  const float* beta;
  float eps;
  int round_tf32; /* round stored fp32 activations to tf32 (RN) */
  /* dual != 0 (block_n 128 only): after the `taps` taps, one extra centre tap (dx = dy = 0) whose K block
   * follows them in `w` accumulates into a SECOND accumulator; combined by AITB_EPI_DUAL */
  int dual;
  const float* bias2;
  /* AITB_F32S only: distance (bf16 elements along dims[0]) from the hi plane to the lo plane of A, i.e. the
   * logical row width of the buffer A lives in.  out / res planes are ldo / ldr apart; W's are taps*k_per_tap apart. */
  int a_lo_off;
  /* optional (LN epilogue): 1 / sigma of every normalised row, written at the row's OUTPUT index (training) */
  float* ln_rstd;
  /* a_m_dim == 2: A is a channels-last map [G, map_h, map_w, C] (view dims {C, map_w, map_h, G}) tiled by boxes of
   * box[1] x box[2] = 128 positions of one image (box[1] >= map_w); M = G * ceil(map_h / box[2]) * 128 tile rows;
   * output row = g * map_h * map_w + y * map_w + x (positions outside the map are computed on zero fill and dropped) */
  int map_w, map_h;
  /* the accumulator is multiplied by out_scale before bias / activation (0 = unset = 1) */
  float out_scale;
  /* AITB_F32S only -- the per-stage precision plan of the fp32-class configuration:
   *   passes : 0 / 3 = three MMA passes hi*hi + hi*lo + lo*hi (both operands 16+ bits);  1 = one pass on the hi planes
   *            (honoured by the 2-CTA and the cluster-LayerNorm kernels; other launches run three passes)
   *   in_f16 : the planes of A and W hold IEEE fp16 (11-bit significand; hi alone = tf32-class operands) instead of bf16
   *   out_f16 / res_f16 : element format of the output planes this launch writes / the residual planes it reads
   * fp16 planes saturate at +-65504; they are used only for LayerNorm-bounded AIT tensors (precision plan). */
  int passes, in_f16, out_f16, res_f16;
} aitb_gemm_desc;

int aitb_gemm(const aitb_gemm_desc* d, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a8 core: scaled-dot-product attention over T = 64 tokens, 8 heads of 64, with the
 * selective-head gate (system/SubLayers.py:9-39,89-97; system/Modules.py:16-29):
 *   per group g (a proposal-query pair):  O_h = softmax(mask(Q_h K_h^T / 8)) V_h,
 *   s = mean_T(sum_h O_h),  gate = softmax_h(view(W_sk s + b_sk, [8, 64])),
 *   out[g] = sum_h O_h * gate_h           -> [G, 64, 64]
 *   q rows of group g start at (g / q_rep) * 64 (q_rep > 1 shares one Q block across q_rep groups)
 *   mask_mode 0: keys j >= n_keys masked;  1: causal (j > i masked)
 * ---------------------------------------------------------------------------------------- */
int aitb_attn_core(const void* q, int ldq, int q_rep, const void* k, const void* v, int ldkv,
                   const float* w_sk, const float* b_sk, int G, int mask_mode, int n_keys,
                   int dtype, void* out, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a12: spatial mean over the 4x4 layer-4 map + bbox / similarity-score heads
 * (resnet_coatt_transformer_sk.py:476-485, 419-427; faster_rcnn_coatt_transformer_sk.py:318-337)
 *   top [G, 16, 2048] (dtype) -> feat [G, 2048] f32 (optional), bbox [G, 4] f32,
 *   cls_prob [G] f32 = softmax(W2 (W1 [feat, qfeat[g / P]] + b1) + b2)[1]
 *   qfeat [G / P, 2048] f32 = pooled query feature (same kernel, run first with w_* = NULL)
 * ---------------------------------------------------------------------------------------- */
int aitb_pool_heads(const void* top, int dtype, int G, int P, const float* qfeat,
                    const float* w_bbox, const float* b_bbox, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* feat_out, float* bbox_out,
                    float* cls_prob_out, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The whole head for a batch of B (image, query) units x P proposals:
 * ROIAlign -> AIT (a5-a9) -> SKNet (a10) -> RCNN_top (a11) -> heads (a12).
 * Weights are passed pre-packed (see ait_b200/packing.py): K-major [N, K] matrices in `dtype`,
 * frozen BN folded into the conv weights / biases, fp32 biases and LayerNorm parameters.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* w;     /* [N, K] K-major (dtype) */
  const float* bias; /* [N] or NULL */
} aitb_linear;

typedef struct {
  const float* gamma;
  const float* beta;
} aitb_lnorm;

typedef struct {
  const void* w_qkv; /* [1536, 512] = rows(w_qs; w_ks; w_vs)   (SubLayers.py:51-53) */
  const float* w_sk; /* [512, 64] f32, b_sk [512] f32           (SubLayers.py:22)   */
  const float* b_sk;
  const void* w_fc; /* [512, 64]                                (SubLayers.py:58)   */
  aitb_lnorm ln;
} aitb_mha;

typedef struct {
  aitb_linear w1; /* [2048, 512] + bias */
  aitb_linear w2; /* [512, 2048] + bias */
  aitb_lnorm ln;
} aitb_ffn;

typedef struct {
  aitb_linear conv1; /* [512, Cin]        BN folded */
  aitb_linear conv2; /* [512, 9*512]      tap-major K, BN folded */
  aitb_linear conv3; /* [2048, 512]       BN folded */
  aitb_linear down;  /* [2048, 1024] or w == NULL */
} aitb_bottleneck;

typedef struct {
  aitb_linear conv1x1; /* [1024, 128]   grouped (8 groups of 128 -> 128) */
  aitb_linear conv3x3; /* [1024, 9*128] grouped, tap-major K            */
  const void* w_fused; /* [1024, 10*128] = conv3x3 taps followed by the conv1x1 block (dual-accumulator GEMM) */
} aitb_skblock;

/* aitb_head_weights.plan bits (AITB_F32S configuration only; see DESIGN.md "precision plan") */
#define AITB_PLAN_ENC_ONEPASS 1 /* encoder-side GEMMs (enc_emb, encoder QKV, encoder FFN w_1 / w_2, cross-attention K/V) run ONE
                                   tensor-core pass on fp16 hi planes (11-bit operands) instead of three on bf16 planes; their
                                   weights (enc_emb.w, enc_slf.w_qkv, enc_ffn.w1/w2.w, rows 512.. of dec_enc.w_qkv) must then be
                                   packed as fp16 planes */

typedef struct {
  int dtype;
  int round_tf32;
  int plan; /* AITB_PLAN_* bits, 0 = three passes everywhere */
  /* AIT (system/Models.py:177-220) */
  aitb_linear enc_emb, dec_emb, dec_trans;
  const float* enc_pos; /* [64, 512] f32 (encoder.position_enc.pos_table) */
  const float* dec_pos;
  aitb_lnorm enc_ln, dec_ln;
  aitb_mha enc_slf, dec_slf, dec_enc;
  aitb_ffn enc_ffn, dec_ffn;
  /* SKNet (blocks_coatt_transformer_sk.py:915-998), separate weights for proposals / query */
  aitb_skblock sk_props, sk_query;
  /* RCNN_top = ResNet-50 layer4 (resnet_coatt_transformer_sk.py:416) */
  aitb_bottleneck top[3];
  /* heads (resnet_coatt_transformer_sk.py:419-427), fp32 */
  const float* w_bbox; /* [4, 2048] */
  const float* b_bbox;
  const float* w_cls1; /* [8, 4096] */
  const float* b_cls1;
  const float* w_cls2; /* [2, 8] */
  const float* b_cls2;
  /* training-mode dropout, read by aitb_ait_forward_train / aitb_ait_backward only (every inference entry point ignores
   * them = the reference in .eval()).  p_drop: the nn.Dropout(dropout) sites (Models.py:98,152; SubLayers.py:97,182);
   * p_attn: the attention-probability dropout (Modules.py:9,24 -- always 0.1 in the reference's .train()); drop_seed: the
   * step's seed -- forward and backward of one step must see the same three values.  All zero = no dropout. */
  float p_drop;
  float p_attn;
  unsigned long long drop_seed;
} aitb_head_weights;

/* dropout sites (one Philox key per site; aitb_dropout_mask / aitb_attn_dropout_mask materialise a site's multipliers) */
#define AITB_DROP_ENC_EMB 1      /* rows [bp*64] (incl. the 15 zero-padded rows of a pair: they feed the self-attention's head gate) */
#define AITB_DROP_DEC_EMB 2      /* rows [B*64]: ONE mask per unit, shared by the unit's P pairs (the query side runs once) */
#define AITB_DROP_ENC_SLF_FC 3   /* rows [bp*64] */
#define AITB_DROP_DEC_SLF_FC 4   /* rows [B*64] */
#define AITB_DROP_DEC_ENC_FC 5   /* rows [bp*64] */
#define AITB_DROP_ENC_FFN 6      /* rows [bp*64] */
#define AITB_DROP_DEC_FFN 7      /* rows [bp*64] */
#define AITB_DROP_ENC_SLF_ATTN 8 /* [bp, 8, 64, 64] */
#define AITB_DROP_DEC_SLF_ATTN 9 /* [B, 8, 64, 64] */
#define AITB_DROP_DEC_ENC_ATTN 10 /* [bp, 8, 64, 64] */

typedef struct {
  /* optional taps of intermediates for parity tests (NULL = skip); all in `dtype` unless noted */
  void* pooled;   /* [B*P, 49, 1024] token-major ROIAlign output */
  void* enc_out;  /* [B*P, 64, 512]  */
  void* ait_out;  /* [B*P, 64, 1024] token-major (== [bp,1024,8,8] NCHW transposed) */
  void* sk_out;   /* [B*P, 64, 1024] */
  float* feat;    /* [B*P, 2048] f32  pooled layer-4 feature of each pair */
  float* qfeat;   /* [B, 2048]   f32 */
} aitb_head_taps;

size_t aitb_head_workspace_bytes(int B, int P, int dtype);

/* feat_nchw [B,1024,H,W] f32, query_nchw [B,1024,8,8] f32, rois [B*P,5] f32 (batch idx = unit)
 * -> cls_prob [B*P] f32, bbox_pred [B*P,4] f32 */
int aitb_head_forward(const aitb_head_weights* w, const float* feat_nchw, int H, int W,
                      const float* query_nchw, const float* rois, int B, int P, float* cls_prob,
                      float* bbox_pred, const aitb_head_taps* taps, void* workspace,
                      size_t workspace_bytes, aitb_stream_t stream);

/* The AIT module alone (a5): x_props [bp,1024,7,7] f32 NCHW, x_query [B,1024,8,8] f32 NCHW
 * -> out [bp,1024,8,8] f32 NCHW   (system/Models.py:231-280) */
size_t aitb_ait_workspace_bytes(int B, int P, int dtype);
int aitb_ait_forward(const aitb_head_weights* w, const float* x_props, const float* x_query,
                     int B, int P, float* out_nchw, void* workspace, size_t workspace_bytes,
                     aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f1  front of the proposal layer (rpn/proposal_layer.py:66-118, rpn/bbox_transform.py:77-133):
 * anchors (base anchors [A,4] + cell shifts * feat_stride), bbox_transform_inv, clip_boxes and the
 * `permute(0,2,3,1)` re-ordering of the RPN outputs, in one kernel.
 *   scores_nchw [B, 2A, H, W] (foreground = channels A..2A), deltas_nchw [B, 4A, H, W], im_info [B, 3] (h, w, scale)
 *   -> proposals [B, H*W*A, 4], fg_scores [B, H*W*A]   (then aitb_topk_desc + aitb_nms_batched)
 * ---------------------------------------------------------------------------------------- */
int aitb_rpn_decode(const float* scores_nchw, const float* deltas_nchw, const float* base_anchors,
                    const float* im_info, int B, int A, int H, int W, float feat_stride, float* proposals,
                    float* fg_scores, aitb_stream_t stream);

/* f3 (first half): the RPN head in front of the proposal layer (lib/model/rpn/rpn.py:34-85): 3x3 conv 1024 -> 512 +
 * ReLU on the whole C4 map (nine shifted TMA boxes per K chunk, zero fill = padding), RPN_cls_score | RPN_bbox_pred as
 * ONE 1x1 GEMM, the bg/fg pair softmax, anchors, bbox_transform_inv and clip_boxes.
 *   conv.w  [512, 9*1024] tap-major (dtype), conv.bias [512] f32
 *   heads.w [n_pad, 512] (dtype): rows 0..2A-1 = RPN_cls_score, 2A..6A-1 = RPN_bbox_pred, zero rows up to
 *           n_pad = 64 (6A <= 64) or 128; heads.bias [n_pad] f32
 *   feat_nchw [B, 1024, H, W] f32 (W <= 128) -> proposals [B, H*W*A, 4], fg_scores [B, H*W*A];
 *   optional reference-layout outputs rpn_cls_prob [B, 2A, H, W], rpn_bbox_pred [B, 4A, H, W] (NULL = skip) */
typedef struct {
  int dtype, round_tf32, A, n_pad;
  aitb_linear conv, heads;
} aitb_rpn_weights;
size_t aitb_rpn_workspace_bytes(int B, int H, int W, int dtype);
int aitb_rpn_forward(const aitb_rpn_weights* w, const float* feat_nchw, int B, int H, int W, const float* base_anchors,
                     const float* im_info, float feat_stride, float* proposals, float* fg_scores, float* rpn_cls_prob,
                     float* rpn_bbox_pred, void* workspace, size_t workspace_bytes, aitb_stream_t stream);

/* f3 (second half): the co-attention block in front of the RPN (`CoAttention`, in_ch 1024, c_hidden 512, residual,
 * 'division' normalisation: lib/model/modules/blocks_coatt_transformer_sk.py:17-122, as built by CoAttentionModule,
 * faster_rcnn_coatt_transformer_sk.py:104-124).  fp32 storage + tf32 tensor-core math (dtype must be AITB_F32).
 *   x_img [B,1024,H,W] f32, x_qry [B,1024,8,8] f32 -> non_img [B,1024,H,W], non_qry [B,1024,8,8] */
typedef struct {
  int dtype, round_tf32;
  const void* w_emb_phi;  /* [1024, 1024] = rows(emb.weight; phi.weight), tf32-rounded */
  const float* b_emb_phi; /* [1024] */
  aitb_linear emb, rho;   /* [512, 1024] + bias */
  aitb_linear theta, omega; /* [1024, 512] + bias (theta.0 / omega.0) */
  aitb_lnorm theta_gn, omega_gn; /* GroupNorm(32, 1024) weight / bias (theta.1 / omega.1) */
} aitb_coatt_weights;
size_t aitb_coattention_workspace_bytes(int B, int H, int W);
int aitb_coattention_forward(const aitb_coatt_weights* w, const float* x_img, const float* x_qry, int B, int H, int W,
                             float* non_img, float* non_qry, void* workspace, size_t workspace_bytes,
                             aitb_stream_t stream);

/* Training path of the co-attention block (ait_b200/coatt_train.py composes forward and backward from aitb_gemm / aitb_wgrad /
 * aitb_transpose_cs and these two): GroupNorm(groups, 1024) (+ identity) on token-major [B, N, 1024]
 * fp32, forward keeping the per-(image, group) (sum, sum of squares) in `sums` [B, groups, 2] doubles, and its backward:
 * dx [B, N, 1024] (rounded to tf32 when round_tf32), dgamma / dbeta [1024] ACCUMULATED; bsums: [B, groups, 2] doubles scratch.
 * Reference: nn.GroupNorm inside `theta` / `omega` (blocks_coatt_transformer_sk.py:31-42) + torch autograd. */
int aitb_group_norm_forward(const float* x, const float* identity, const float* gamma, const float* beta, int B, int N,
                            int groups, float eps, double* sums, float* out, aitb_stream_t stream);
int aitb_group_norm_backward(const float* dy, const float* x, const double* sums, const float* gamma, int B, int N,
                             int groups, float eps, int round_tf32, double* bsums, float* dx, float* dgamma, float* dbeta,
                             aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f2  detection post-processing (test_net_voc.py:380-446)
 *   aitb_box_decode: pred = clip(bbox_transform_inv(box, delta * h_stds + h_means)) [/ im_scale];
 *     boxes rows have `boxes_stride` floats with the box at column `boxes_off` (rois [B,N,5]: stride 5, offset 1);
 *     key [B,N] = score > thresh ? score : -inf (sort key), n_valid [B] = number of scores above thresh
 *   aitb_det_assemble: after aitb_topk_desc(key) -> order and aitb_nms_batched(pred, order, nms_thresh, mode 0)
 *     -> keep_pos / n_keep: dets [B, N, 5] = (x1,y1,x2,y2,score) in descending score order, zero-padded,
 *     limited to max_per_image (> 0) with the reference's `>=` rule at the cut; n_det [B]
 * ---------------------------------------------------------------------------------------- */
int aitb_box_decode(const float* boxes, int boxes_stride, int boxes_off, const float* deltas, const float* cls,
                    const float* im_info, int B, int N, const float* h_stds, const float* h_means, float thresh,
                    int divide_by_scale, float* pred, float* key, int32_t* n_valid, aitb_stream_t stream);
int aitb_det_assemble(const float* pred, const float* cls, const int64_t* order, const int64_t* keep_pos,
                      const int32_t* n_keep, const int32_t* n_valid, int B, int N, int max_per_image, float* dets,
                      int32_t* n_det, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training (BASELINE config 4): backward building blocks.  fp32 storage, tf32 tensor-core math.
 * The reference obtains these from torch autograd over the modules of a5-a9.
 * ---------------------------------------------------------------------------------------- */

/* weight gradient of out = x * w^T:  dw[N, ldw] += dy[M, 0..N)^T * x[M, 0..K)   (row-major activations with
 * leading dimensions ldy / ldx; dw must be initialised by the caller, e.g. zeroed).  N % 128 == 0, K % 64 == 0. */
int aitb_wgrad(const float* dy, int ldy, const float* x, int ldx, int M, int N, int K, float* dw, int ldw,
               aitb_stream_t stream);
/* same for bf16 row-major activations (the bf16 training configuration): both operands MN-major bf16 tiles in the ordinary
 * 128-byte swizzle, kind::f16 MMAs of K = 16, dW fp32 accumulated.  ldy / ldx multiples of 8. */
int aitb_wgrad_bf16(const void* dy, int ldy, const void* x, int ldx, int M, int N, int K, float* dw, int ldw,
                    aitb_stream_t stream);
/* Weight gradient of a (grouped) 1x1 / 3x3 stride-1 "same" convolution on a channels-last S x S map (S = 4 | 8) WITHOUT an
 * im2col buffer: x [G,S,S,C] is read through a 4-D TMA view, one shifted box per tap, out-of-map rows zero-filled (= the
 * padding); dW [N, taps*Cg] (tap-major, Cg = C / groups) += dY^T * shifted X.  groups > 1: N / groups must be 128 (one
 * 128-row tile of dW per group).  taps = 1 | 9.  The reference gets this from torch autograd over nn.Conv2d
 * (blocks_coatt_transformer_sk.py:929, resnet_coatt_transformer_sk.py:80). */
int aitb_wgrad_conv(const float* dy, int ldy, const float* x, int G, int S, int C, int N, int groups, int taps, float* dw,
                    int ldw, aitb_stream_t stream);

/* The other backward pieces (exported for unit tests; see ait_b200/csrc/bwd.cu) */
int aitb_ln_bwd(const float* g, const float* y, const float* gamma, const float* beta, const float* rstd,
                int rows, int grp, int valid, float* dx, float* dgamma, float* dbeta, aitb_stream_t stream);
int aitb_colsum(const float* x, int ld, int rows, int cols, float* out, aitb_stream_t stream);
int aitb_bsum(const float* x, int B, int P, int L, float* out, aitb_stream_t stream);
/* selective-head attention backward (one launch per attention block): q / k / v as in aitb_attn_core (fp32),
 * dout [G, 64, 64] -> dq [G*64, lddq] (PER PAIR even when q_rep > 1: sum over the unit's pairs with aitb_bsum),
 * dk / dv [G*64, lddkv] (head h in columns h*64..), dz [G, 512] = d(W_sk s + b_sk), s [G, 64] */
int aitb_attn_bwd(const float* q, int ldq, int q_rep, const float* k, const float* v, int ldkv,
                  const float* w_sk, const float* b_sk, const float* dout, int G, int mask_mode, int n_keys,
                  float* dq, int lddq, float* dk, float* dv, int lddkv, float* dz, float* s,
                  aitb_stream_t stream);

/* AIT training step (config 4): Transformer.forward with the activations the backward needs kept in `saved`
 * (caller-owned, aitb_ait_saved_bytes), then the backward producing the gradients of both inputs and of all
 * parameters the forward uses.  dtype must be AITB_F32; dropout follows aitb_head_weights.p_drop / p_attn / drop_seed. */
typedef struct { float* w; float* bias; } aitb_linear_g;
typedef struct { float* gamma; float* beta; } aitb_lnorm_g;
typedef struct { float* w_qkv; float* w_sk; float* b_sk; float* w_fc; aitb_lnorm_g ln; } aitb_mha_g;
typedef struct { aitb_linear_g w1, w2; aitb_lnorm_g ln; } aitb_ffn_g;
typedef struct {
  aitb_linear_g enc_emb, dec_emb, dec_trans; /* [512,1024]+[512], [512,1024]+[512], [1024,512]+[1024] */
  aitb_lnorm_g enc_ln, dec_ln;
  aitb_mha_g enc_slf, dec_slf, dec_enc;      /* w_qkv [1536,512] = rows(w_qs; w_ks; w_vs) */
  aitb_ffn_g enc_ffn, dec_ffn;
} aitb_ait_grads; /* every buffer fp32, zero-initialised by the caller; gradients are ACCUMULATED into them */

size_t aitb_ait_saved_bytes(int B, int P);
size_t aitb_ait_backward_workspace_bytes(int B, int P);
int aitb_ait_forward_train(const aitb_head_weights* w, const float* x_props, const float* x_query, int B, int P,
                           float* out_nchw, void* saved, size_t saved_bytes, aitb_stream_t stream);
/* grad_out [bp,1024,8,8] -> grad_props [bp,1024,7,7], grad_query [B,1024,8,8] (both overwritten), grads accumulated */
int aitb_ait_backward(const aitb_head_weights* w, const float* grad_out_nchw, int B, int P, const void* saved,
                      size_t saved_bytes, const aitb_ait_grads* grads, float* grad_props, float* grad_query,
                      void* workspace, size_t workspace_bytes, aitb_stream_t stream);
/* Layout hand-over inside a training step (no NCHW round trip between AIT and the next stage): aitb_ait_forward_train
 * accepts out_nchw == NULL and leaves the token-major result [bp*64, 1024] (row = pair*64 + y*8 + x, tf32-rounded) in
 * `saved` at byte offset aitb_ait_saved_offset(B, P, 1) (which = 0: the token-major pooled input [bp*49, 1024]; 2 / 3: the
 * post-ReLU hidden tensors [bp*64, 2048] of the encoder / decoder FFN -- tests read the device's ReLU decisions there;
 * offsets of the fp32-storage configuration);
 * aitb_ait_backward_tm takes the incoming gradient in that same token-major layout, already rounded to tf32. */
size_t aitb_ait_saved_offset(int B, int P, int which);
int aitb_ait_backward_tm(const aitb_head_weights* w, const float* grad_out_tm, int B, int P, const void* saved,
                         size_t saved_bytes, const aitb_ait_grads* grads, float* grad_props, float* grad_query,
                         void* workspace, size_t workspace_bytes, aitb_stream_t stream);

/* Dropout masks of one training step, materialised for parity tests: the multipliers (0 or 1 / (1 - p)) the training
 * forward / backward regenerate on the fly from (seed, site, element index) with Philox4x32-10 -- nothing is stored by the
 * step itself.  The reference's own masks come from torch's generator state and cannot be reproduced; parity is checked by
 * injecting THESE masks into the oracle (tests/test_gpu_train.py).
 *   aitb_dropout_mask: out [rows, 512] fp32 of a row-wise site;  aitb_attn_dropout_mask: out [G, 8, 64, 64] fp32 */
int aitb_dropout_mask(float p, unsigned long long seed, int site, int rows, float* out, aitb_stream_t stream);
int aitb_attn_dropout_mask(float p, unsigned long long seed, int site, int G, float* out, aitb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f4  training-only samplers and losses (ait_b200/csrc/targets.cu)
 *
 * Anchor target layer (lib/model/rpn/anchor_target_layer.py:49-199), two phases around the random sub-sampling:
 *   aitb_anchor_target_assign: anchors = base_anchors [A,4] + cell shifts (cell-major, index cell*A + a); anchors inside
 *     the image of im_info[0] get max / argmax IoU over gt_boxes [B,K,5] and the pre-sampling label
 *     (-1 don't care, 0 bg: max < neg_thr, 1 fg: max >= pos_thr or equal to a gt box's best IoU);
 *     labels [B, H*W*A] int8 (outside anchors -1), argmax [B, H*W*A] (outside -1), counts [B,2] = (#fg, #bg).
 *   aitb_anchor_target_finish: drop [B, 2, ld_drop] bytes (NULL = keep all): drop[b][0][r] != 0 disables the r-th
 *     foreground anchor of image b in ascending anchor order (= `fg_inds[r]`, :132-140), drop[b][1][r] the r-th
 *     background anchor (:146-152); then the reference's four outputs: labels_out [B,1,A*H,W], bbox_targets /
 *     inside_w / outside_w [B,4A,H,W]; n_examples [B] = anchors with label >= 0 (outside weight = 1 / n_examples[B-1],
 *     the reference's `labels[i]` after its loop, :163).
 * Proposal target layer (lib/model/rpn/proposal_target_layer_cascade.py:33-220):
 *   aitb_proposal_target_assign: candidates = rois [B,R,5] ++ gt boxes; max_overlaps / assignment [B,R+K],
 *     cls [B,R+K] int8 (1 fg: max >= fg_thr, 0 bg: bg_lo <= max < bg_hi, -1 neither), counts [B,2].
 *   aitb_proposal_target_sample: picks [B,S] ranks within the image's ascending fg list (first n_fg_pick[b] slots) and
 *     bg list (the rest) -> rois_out [B,S,5], labels_out [B,S], bbox_targets / inside_w / outside_w [B,S,4] (targets
 *     normalised by h_means / h_stds); lists: scratch [B,2,R+K] int32; *bad_flag set to 1 on an out-of-range pick.
 * Losses with their gradients (rpn.py:99-126; faster_rcnn_coatt_transformer_sk.py:334-361; net_utils.py:75-89):
 *   aitb_rpn_loss: losses[0] = cross entropy over anchors with label != -1, losses[1] = smooth L1 (sigma), and
 *     d(gscale[0]*losses[0] + gscale[1]*losses[1]) / d(rpn_cls_score [B,2A,H,W]), / d(rpn_bbox_pred [B,4A,H,W])
 *     (gradient outputs and gscale may be NULL; gscale NULL = ones); acc: 3 doubles of scratch.
 *   aitb_rcnn_loss: score [bs*P,2], bbox_pred [bs*P,4], labels [bs*P] -> losses[0] = cross entropy, losses[1] =
 *     margin_scale * MarginRankingLoss(margin)(|p_i-p_j|, |l_i-l_j|, target), losses[2] = smooth L1 (sigma 1);
 *     cls_prob [bs*P] = softmax(score)[:,1]; gradients as above with gscale[3]; acc: 3 doubles of scratch.
 * Device-side sampling (no host round trip; Philox4x32-10 keyed by seed / image / candidate, same distributions as the
 * reference's numpy draws but an independent stream):
 *   aitb_anchor_target_subsample_device: between _assign and _finish(drop = NULL): keeps the num_fg foreground and
 *     rpn_batch - #fg background anchors with the smallest keys, disables the rest in `labels`.
 *   aitb_proposal_target_picks_device: fills picks / n_fg_pick for aitb_proposal_target_sample from counts; an image
 *     without any candidate sets *bad_flag and -1 picks (the reference raises).
 * ---------------------------------------------------------------------------------------- */
int aitb_anchor_target_subsample_device(int8_t* labels, const int32_t* counts, int B, int total, int num_fg, int rpn_batch,
                                        uint64_t seed, aitb_stream_t stream);
int aitb_proposal_target_picks_device(const int32_t* counts, int B, int S, int fg_per_image, uint64_t seed, int32_t* picks,
                                      int32_t* n_fg_pick, int32_t* bad_flag, aitb_stream_t stream);
size_t aitb_anchor_target_workspace_bytes(int B, int A, int H, int W, int K);
int aitb_anchor_target_assign(const float* base_anchors, const float* gt_boxes, const float* im_info, int B, int A, int H,
                              int W, int K, float feat_stride, float neg_thr, float pos_thr, int clobber_positives,
                              int8_t* labels, int32_t* argmax, int32_t* counts, void* workspace, size_t workspace_bytes,
                              aitb_stream_t stream);
int aitb_anchor_target_finish(const float* base_anchors, const float* gt_boxes, int B, int A, int H, int W, int K,
                              float feat_stride, int8_t* labels, const int32_t* argmax, const uint8_t* drop, int ld_drop,
                              float inside_weight, int32_t* n_examples, float* labels_out, float* bbox_targets,
                              float* inside_w, float* outside_w, aitb_stream_t stream);
int aitb_proposal_target_assign(const float* rois, const float* gt_boxes, int B, int R, int K, float fg_thr, float bg_hi,
                                float bg_lo, float* max_overlaps, int32_t* assignment, int8_t* cls, int32_t* counts,
                                aitb_stream_t stream);
int aitb_proposal_target_sample(const float* rois, const float* gt_boxes, int B, int R, int K, const int8_t* cls,
                                const int32_t* assignment, const int32_t* picks, const int32_t* n_fg_pick, int S,
                                const float* h_means, const float* h_stds, const float* h_inside_w, int32_t* lists,
                                float* rois_out, float* labels_out, float* bbox_targets, float* inside_w, float* outside_w,
                                int32_t* bad_flag, aitb_stream_t stream);
int aitb_rpn_loss(const float* rpn_cls_score, const float* rpn_bbox_pred, const float* labels, const float* bbox_targets,
                  const float* inside_w, const float* outside_w, int B, int A, int H, int W, float sigma,
                  const float* gscale, float* losses, float* d_cls_score, float* d_bbox_pred, double* acc,
                  aitb_stream_t stream);
int aitb_rcnn_loss(const float* score, const float* bbox_pred, const float* labels, const float* bbox_targets,
                   const float* inside_w, const float* outside_w, int bs, int P, float margin, float margin_scale,
                   const float* gscale, float* losses, float* cls_prob, float* d_score, float* d_bbox_pred, double* acc,
                   aitb_stream_t stream);

/* Attention output projection + residual + LayerNorm (lib/model/system/SubLayers.py:97-100) as one streaming kernel
 * (ait_b200/csrc/fc_ln.cu): out[o(m), :] = LN(a[m, 0..64) * w_fc^T + res[r(m), :]) * gamma + beta, N = 512, K = 64.
 * dtype AITB_BF16 or AITB_F32S (two bf16 planes per row: a [M, 64 | 64], w [512, 64 | 64], res / out [*, 512 | 512]).
 * Row maps as in aitb_gemm: o(m) = (m / rows_in) * rows_out + m % rows_in for m % rows_in < rows_out (other rows are
 * dropped); r(m) = ((b / res_div) / res_rep) * res_div + b % res_div with b = m (res_row_m) or o(m). */
int aitb_fc_ln(int dtype, const void* a, const void* w_fc, const void* res, const float* gamma, const float* beta, float eps,
               void* out, int M, int rows_in, int rows_out, int res_row_m, int res_div, int res_rep, aitb_stream_t stream);

/* Training of the head's last layers on the pooled layer-4 features (ait_b200/csrc/heads_train.cu; rows a12 + f4):
 *   forward: feat [G,2048], qfeat [G/P,2048] -> bbox [G,4], hidden [G,8] (kept for the backward), score [G,2] (the LOGITS
 *            the losses of faster_rcnn_coatt_transformer_sk.py:340-361 are taken on)
 *   backward: d_score [G,2], d_bbox [G,4] -> d_feat [G,2048], d_qfeat [G/P,2048] (overwritten) and the parameter gradients
 *            dw_bbox [4,2048], db_bbox [4], dw1 [8,4096], db1 [8], dw2 [2,8], db2 [2], ACCUMULATED (caller zero-initialises)
 *   aitb_mean_pool_backward: d_top [G,16,2048] = d_feat / 16 (adjoint of `_head_to_tail`'s 4x4 mean) */
int aitb_heads_forward_train(const float* feat, const float* qfeat, int G, int P, const float* w_bbox, const float* b_bbox,
                             const float* w1, const float* b1, const float* w2, const float* b2, float* bbox, float* hidden,
                             float* score, aitb_stream_t stream);
size_t aitb_heads_backward_workspace_bytes(int G, int P);
int aitb_heads_backward(const float* feat, const float* qfeat, const float* hidden, const float* d_score, const float* d_bbox,
                        int G, int P, const float* w_bbox, const float* w1, const float* w2, float* d_feat, float* d_qfeat,
                        float* dw_bbox, float* db_bbox, float* dw1, float* db1, float* dw2, float* db2, void* workspace,
                        size_t workspace_bytes, aitb_stream_t stream);
int aitb_mean_pool_backward(const float* d_feat, int G, float* d_top, aitb_stream_t stream);

/* Data-movement helpers of the layer-4 (`RCNN_top`) training path (ait_b200/csrc/train_aux.cu, composed with aitb_gemm /
 * aitb_wgrad by ait_b200/top_train.py), fp32, C % 4 == 0:
 *   aitb_relu_bwd      out[i] = y[i] > 0 ? tf32(dy[i]) : 0   (rounded to nearest: the result only feeds tf32 GEMMs)
 *   aitb_im2col3x3     x [G,s,s,C] -> out [G*s*s, 9*C], tap-major (ky, kx, c), zero padding
 *   aitb_map_subsample x [G,S,S,C] -> out [G,s,s,C] at (stride*y, stride*x);  aitb_map_upsample: its adjoint (zeros elsewhere) */
int aitb_relu_bwd(const float* dy, const float* y, float* out, size_t n, aitb_stream_t stream);
int aitb_im2col3x3(const float* x, int G, int s, int C, float* out, aitb_stream_t stream);
int aitb_map_subsample(const float* x, int G, int S, int s, int stride, int C, float* out, aitb_stream_t stream);
int aitb_map_upsample(const float* x, int G, int S, int s, int stride, int C, float* out, aitb_stream_t stream);
/* channels-last map [B, H, W, C] -> zero-bordered copy [B, H+2, W+2, C] (C % 4 == 0): the 3x3 weight gradient on an arbitrary
 * H x W map (the RPN head's training step, ait_b200/rpn_train.py; rpn.py:18-110 + torch autograd in the reference) is then nine
 * aitb_wgrad launches on row-shifted views of the padded gradient / input maps. */
int aitb_map_pad(const float* x, int B, int H, int W, int C, float* out, aitb_stream_t stream);

/* SKNet training path (lib/model/modules/blocks_coatt_transformer_sk.py:960-998; composed with aitb_gemm / aitb_wgrad by
 * ait_b200/sk_train.py; the reference gets the backward from torch autograd):
 *   aitb_im2col3x3_grouped  like aitb_im2col3x3 with columns ordered (group, ky, kx, channel in group): the weight gradient
 *                           of a grouped 3x3 convolution is one aitb_wgrad per group on a [rows, 9*group_c] slice
 *   aitb_sk_combine         out = r1^2 + r3^2   (r1, r3: the post-ReLU 1x1 / 3x3 branch maps; round_tf32: the consumer is a tf32 GEMM)
 *   aitb_sk_combine_bwd     d1 = 2 dv r1, d3 = 2 dv r3, rounded to tf32 (nearest) */
int aitb_im2col3x3_grouped(const float* x, int G, int s, int C, int group_c, float* out, aitb_stream_t stream);
int aitb_sk_combine(const float* r1, const float* r3, float* out, size_t n, int round_tf32, aitb_stream_t stream);
int aitb_sk_combine_bwd(const float* dv, const float* r1, const float* r3, float* d1, float* d3, size_t n, aitb_stream_t stream);

/* number of kernels launched by this thread through the library since the last reset */
long long aitb_launch_count(int reset);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AITB200_H_ */
