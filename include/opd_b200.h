/*
 * opd_b200.h — C ABI of libopd_b200.so, the B200 (sm_100a) implementation of the
 * Phase 2 -> 3 hot path of Kizuna42/office-person-detection-vit.
 *
 * The reference is 100 % Python and has no FFI of its own; these
 * entry points are what a ctypes binding inside the reference's three engine
 * classes would call.  Each declaration cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a negative opd_status on error;
 *     opd_last_error() returns a thread-local, human readable message;
 *   - no C++ exception crosses the ABI, no torch type appears in a signature;
 *   - all pointers named *_dev are device pointers owned by the caller, all
 *     other pointers are host pointers that are only read during the call;
 *   - work is enqueued on the caller's stream (`stream` is a cudaStream_t passed
 *     as void*); no hidden synchronisation except where stated;
 *   - handles are bound to one device, are not thread-safe, and own only their
 *     own packed tables / weights.
 */
#ifndef OPD_B200_H
#define OPD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPD_ABI_VERSION 1

typedef enum opd_status {
  OPD_OK = 0,
  OPD_ERR_INVALID = -1,     /* bad argument (shape, NULL, range) */
  OPD_ERR_CUDA = -2,        /* CUDA runtime / driver error, see opd_last_error() */
  OPD_ERR_UNSUPPORTED = -3, /* valid request outside the implemented envelope */
  OPD_ERR_NOMEM = -4
} opd_status;

int opd_version(void);
const char* opd_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
int64_t opd_launch_count(void);
/* Kernel-selection knobs for A/B measurements (plans built afterwards see the new value):
 *   "attention_tc" 1 (default): fused attention on tcgen05 / TMEM; 0: the mma.sync flash kernel
 *   "attention_kv" 96 (default), 64 or 128: keys per tile of the tcgen05 attention kernel (4, 4 or 2 CTAs per SM); 65: 64 keys with
 *                  the score tile held in registers
 *   "gemm_res_wide" 1 (default): bias + residual + ReLU GEMMs with K >= 256 use 256-column cta_group::2 tiles; 0: 128-column tiles
 *   "pdl"         1 (default): GEMM / attention launches allow programmatic dependent launch; 0: plain stream order
 *   "mlp_fused"   1 (default): the feed-forward blocks run as one kernel (opd_mlp_ln_bf16); 0: two GEMM launches
 *   "bneck_halo"  1 (default): 64-channel stride-1 bottleneck tails load one halo patch per tile; 0: im2col TMA.
 *   "bneck_pair"  1 (default): 128-channel bottleneck tails run as cta_group::2 pairs; 0: one CTA per tile; 3: tests
 *   "bneck_release" 3 (default): the fused tails hand a residual slot back early in the next epilogue step (bit 0 im2col, bit 1 halo kernel)
 *   "gemm_pair"   1 (default): BLOCK_N = 256 GEMM / convolution layers run as cta_group::2 pairs; 0: off; 3: tests
 *   "gemm_reverse" 1 (default): a GEMM layer walks its row blocks opposite to the launch before it (L2 reuse); 0: always ascending
 *   "dec0_const"  1 (default): decoder layer 0's frame-independent self-attention block runs once per plan; 0: every step
 *   (further measurement variants are listed in csrc/opd_core.cu). */
int opd_set_option(const char* name, int32_t value);

/* ------------------------------------------------------------------------------------------
 * Floor projection + zone classification + counting  (K9 + K10 + K11)
 * replaces: src/transform/homography.py:150-197 (transform_batch), :105-133 (transform_pixel)
 *           src/zone/zone_classifier.py:114-149 (classify), :162-197 (_point_in_polygon)
 *           src/aggregation/aggregator.py:52-75 (get_zone_counts)
 * ------------------------------------------------------------------------------------------ */

typedef struct opd_zone_table opd_zone_table;

#define OPD_MAX_ZONES 64

/* Build the device-side zone table (polygons + uniform-grid accelerator).
 *   verts_xy      [poly_offsets[Z], 2] f64 polygon vertices, zones in declaration order
 *   poly_offsets  [Z+1] first vertex of every polygon (each polygon >= 3 vertices)
 *   priority      [Z] f64, +inf encodes the reference's `priority is None`
 *                 (zone_classifier.py:139-145: key = (priority or +inf, declaration order))
 *   allow_overlap 0: single label = argmin (priority, order); 1: all containing zones
 * Z may be 0 (every point is unclassified).  Z <= OPD_MAX_ZONES.  Synchronous. */
int opd_zone_table_create(const double* verts_xy, const int32_t* poly_offsets, const double* priority,
                          int32_t Z, int32_t allow_overlap, int32_t device, opd_zone_table** out);
void opd_zone_table_destroy(opd_zone_table* zt);
/* Introspection for tests / DESIGN.md: grid size and how many cells need the exact test. */
int opd_zone_table_info(const opd_zone_table* zt, int32_t* grid_w, int32_t* grid_h, int32_t* n_boundary_cells,
                        int32_t* n_classes);

typedef struct opd_floor_params {
  double H[9];          /* row-major 3x3 homography, camera px -> floormap px (homography.py:47-91) */
  double scale_x_mm;    /* FloorMapConfig.scale_x_mm_per_px (floormap_config.py:13-31) */
  double scale_y_mm;
  double map_w_px;      /* FloorMapConfig.width_px / height_px — only for the in-bounds flag */
  double map_h_px;
  int32_t input_is_bbox;/* 0: input rows are points (x, y); 1: rows are boxes (x, y, w, h) and the
                           foot point (x + w/2, y + h) is projected (homography.py:166-169) */
  int32_t skip_projection; /* 1: input rows already are floor px (ZoneClassifier.classify on its own) */
} opd_floor_params;

/* Fused projection + classification + histogram over N rows.  Any output pointer may be NULL.
 *   in_dev        [N,2] or [N,4], f32 (…_f32) or f64 (…_f64)
 *   slot_dev      [N] int32 histogram row (timestamp slot) of each point, or NULL = row 0
 *   floor_px_dev  [N,2], floor_mm_dev [N,2]   same dtype as the input
 *   in_bounds_dev [N] u8   0 <= px < map_w && 0 <= py < map_h (homography.py:181)
 *   zone_idx_dev  [N] int32 zone index in declaration order, -1 = no zone (allow_overlap = 0 tables;
 *                 with allow_overlap = 1 it receives the highest-priority containing zone)
 *   zone_mask_dev [N] u64  bit z set <=> point inside zone z (declaration order)
 *   hist_dev      [T, Z+1] int32, ACCUMULATED (caller zeroes); column Z = "unclassified"
 *                 (aggregator.py:64-75: one count per containing zone, or one "unclassified")
 * The f64 entry follows the reference's float64 arithmetic operation by operation; the f32 entry reads
 * and writes float32 but still projects and classifies in float64, so zone results are identical. */
int opd_floor_project_classify_count_f32(const opd_floor_params* p, const opd_zone_table* zt,
                                         const float* in_dev, const int32_t* slot_dev, int64_t N, int32_t T,
                                         float* floor_px_dev, float* floor_mm_dev, uint8_t* in_bounds_dev,
                                         int32_t* zone_idx_dev, uint64_t* zone_mask_dev, int32_t* hist_dev,
                                         void* stream);
int opd_floor_project_classify_count_f64(const opd_floor_params* p, const opd_zone_table* zt,
                                         const double* in_dev, const int32_t* slot_dev, int64_t N, int32_t T,
                                         double* floor_px_dev, double* floor_mm_dev, uint8_t* in_bounds_dev,
                                         int32_t* zone_idx_dev, uint64_t* zone_mask_dev, int32_t* hist_dev,
                                         void* stream);

/* More than OPD_MAX_ZONES zones (the reference sets no limit, zone_classifier.py:44-112): the caller splits the zones into groups of
 * <= OPD_MAX_ZONES in declaration order, builds one table per group and classifies against each; this keeps, per point, the best
 * group winner by the GLOBAL rank (rank_dev [Z]: position in the order sorted by (priority or +inf, declaration order)):
 * idx_groups_dev [G, N] group-local winner or -1 -> out_idx_dev [N] global zone index (g * OPD_MAX_ZONES + z) or -1;
 * out_count_dev [N] (optional): number of groups that contain the point. */
int opd_zone_combine_groups(const int32_t* idx_groups_dev, const int32_t* rank_dev, int32_t G, int64_t N,
                            int32_t* out_idx_dev, int32_t* out_count_dev, void* stream);

/* Histogram only: Aggregator.get_zone_counts on already classified points (aggregator.py:52-75).
 * Exactly one of zone_idx_dev / zone_mask_dev is non-NULL. */
int opd_zone_histogram(const int32_t* zone_idx_dev, const uint64_t* zone_mask_dev, const int32_t* slot_dev,
                       int64_t N, int32_t Z, int32_t T, int32_t* hist_dev, void* stream);


/* ------------------------------------------------------------------------------------------
 * Tensor-core building blocks of the detector (K4 / K5 / K6).  They are the unit-test surface of
 * the tcgen05 kernels; opd_detr_forward (below) chains them with pre-built plans.
 * replaces (third-party `transformers`, the arithmetic the removed src/detection/vit_detector.py drove):
 *   nn.Linear / 1x1 conv / 3x3 conv + DetrFrozenBatchNorm2d + ReLU / residual  (modeling_detr.py:185-222,
 *   models/resnet/modeling_resnet.py:40-200), residual + LayerNorm (modeling_detr.py:592-631),
 *   eager_attention_forward (modeling_detr.py:386-411)
 * ------------------------------------------------------------------------------------------ */

/* epilogue: 0 bias, 1 bias+ReLU, 2 bias+residual+ReLU, 3 LayerNorm(bias+residual)*gamma+beta (N == 256) */
/* D[M,N] = epilogue(A[M,K] W[N,K]^T + bias); bf16 in/out, fp32 accumulate; N, K multiples of 64.
 * d2_dev (optional) receives bf16(D + pos[row % pos_rows]) with pos [pos_rows, N] f32. */
int opd_gemm_bf16(const void* a_dev, int64_t lda, const void* w_dev, void* d_dev, int64_t ldd, int32_t M,
                  int32_t N, int32_t K, int32_t epilogue, const float* bias_dev, const void* residual_dev,
                  int64_t ldr, const float* gamma_dev, const float* beta_dev, void* d2_dev,
                  const float* pos_dev, int32_t pos_rows, void* stream);
/* y[B,P,Q,N] = epilogue(conv(x[B,H,W,C] NHWC bf16, w[N,KH,KW,C] bf16, stride, pad) + bias); C, N multiples of 64 */
int opd_conv2d_nhwc_bf16(const void* x_dev, int32_t B, int32_t H, int32_t W, int32_t C, const void* w_dev,
                         int32_t N, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t epilogue,
                         const float* bias_dev, const void* residual_dev, void* y_dev, void* stream);
/* Fused bottleneck tail: y = relu(conv1x1(relu(conv3x3(x, stride, pad 1) + bias2)) + bias3 + residual) in one kernel
 * (models/resnet/modeling_resnet.py:134-200, layer.1 + layer.2 + shortcut add, frozen BN folded).
 * x [B,H,W,mid] NHWC bf16, w2 [mid,3,3,mid], w3 [width,mid], residual / y [B,P,Q,width]; mid 64 or 128, width % 128 == 0. */
int opd_bottleneck_tail_bf16(const void* x_dev, int32_t B, int32_t H, int32_t W, int32_t mid, const void* w2_dev,
                             const float* bias2_dev, int32_t stride, const void* w3_dev, const float* bias3_dev,
                             int32_t width, const void* residual_dev, void* y_dev, void* stream);
/* Fused feed-forward block of a DETR encoder / decoder layer (modeling_detr.py:560-640, 643-760: mlp.fc1 -> ReLU -> mlp.fc2 ->
 * residual -> final_layer_norm) in one kernel; the [M, 2048] hidden activations never reach memory:
 *   d  = LayerNorm(relu(x w1^T + b1) w2^T + b2 + x) * gamma + beta,   d2 (optional) = bf16(d + pos[row % pos_rows])
 * x, d, d2 [M, 256] bf16 contiguous; w1 [2048, 256], w2 [256, 2048] bf16; b1 [2048], b2 / gamma / beta [256], pos [pos_rows, 256] f32.
 * Bit-identical to opd_gemm_bf16(epilogue 1) followed by opd_gemm_bf16(epilogue 3). */
int opd_mlp_ln_bf16(const void* x_dev, const void* w1_dev, const float* b1_dev, const void* w2_dev, const float* b2_dev,
                    const float* gamma_dev, const float* beta_dev, void* d_dev, void* d2_dev, const float* pos_dev,
                    int32_t pos_rows, int32_t M, void* stream);
/* o = softmax(q k^T / sqrt(32)) v per (batch, head); head h = columns [32h, 32h+32); row strides in elements */
int opd_attention_bf16(const void* q_dev, int64_t ldq, const void* k_dev, int64_t ldk, const void* v_dev,
                       int64_t ldv, void* o_dev, int64_t ldo, int32_t B, int32_t heads, int32_t Lq, int32_t Lk,
                       void* stream);


/* ------------------------------------------------------------------------------------------
 * DETR-ResNet-50 person detector  (K1-K8)
 * replaces: the removed src/detection/vit_detector.py (ViTDetector.detect_batch, _preprocess_batch,
 *           _postprocess_batch; method table in coverage.json) and the third-party
 *           arithmetic it drove: transformers DetrImageProcessor + DetrForObjectDetection
 *           (models/detr/image_processing_detr.py:687-855, models/detr/modeling_detr.py:185-1402,
 *           models/resnet/modeling_resnet.py:40-240); person filter / xywh / foot point as in
 *           src/detection/yolov8_detector.py:210-225, 229-241.
 * ------------------------------------------------------------------------------------------ */

typedef struct opd_detr opd_detr;

/* One named float32 host tensor of the model's state dict (transformers key names, e.g.
 * "model.encoder.layers.0.self_attn.q_proj.weight"). */
typedef struct opd_tensor_f32 {
  const char* name;
  const float* data;
  int64_t numel;
} opd_tensor_f32;

/* Builds the device copy of the weights: frozen BN folded into the convolutions in float32, tensor-core
 * operands rounded to bf16 and laid out [C_out, kh, kw, C_in], heads kept float32.  Synchronous. */
int opd_detr_create(const opd_tensor_f32* tensors, int32_t n_tensors, int32_t device, opd_detr** out);
void opd_detr_destroy(opd_detr* m);
/* debug != 0: every activation gets its own workspace region (no buffer reuse) so that opd_detr_tap can
 * read any of them after a forward; costs memory, changes no arithmetic. */
int opd_detr_set_debug(opd_detr* m, int32_t debug);
/* do_resize = 0 feeds frames at their own size, like DetrImageProcessor(do_resize=False); default 1. */
int opd_detr_set_resize(opd_detr* m, int32_t do_resize);
/* 0 runs the 3x3 convolution and the 1x1 expansion of ResNet stages 1-2 as separate kernels (default 1: fused). */
int opd_detr_set_fusion(opd_detr* m, int32_t fuse_bottleneck_tail);

/* Model input size for a frame size (DetrImageProcessor shortest-edge 800 / longest-edge 1333 rule,
 * transformers/image_transforms.py:206-242) and the stage-4 feature map size. */
int opd_detr_input_shape(int32_t H0, int32_t W0, int32_t* H_in, int32_t* W_in, int32_t* h_feat, int32_t* w_feat);
int opd_detr_workspace_bytes(const opd_detr* m, int32_t B, int32_t H0, int32_t W0, size_t* bytes);
/* frames_dev [B,H0,W0,3] uint8 (BGR like cv2 frames, or RGB) -> logits_dev [B,100,92] f32, boxes_dev [B,100,4] f32
 * (cxcywh in [0,1]).  All frames of a call have the same size (no padding, pixel_mask = 1).  The launch
 * plan (tensor maps) is cached per (B, H0, W0, workspace_dev).  Enqueues only; CUDA-graph capturable after the
 * first call with the same arguments. */
int opd_detr_forward(opd_detr* m, const uint8_t* frames_dev, int32_t B, int32_t H0, int32_t W0,
                     int32_t frames_are_bgr, void* workspace_dev, size_t workspace_bytes, float* logits_dev,
                     float* boxes_dev, void* stream);
/* Batches that mix frame sizes (the removed ViTDetector._preprocess_batch, coverage.json lines 562-578: DetrImageProcessor pads to
 * the batch maximum and returns pixel_mask; models/detr/image_processing_detr.py:638-667 pad, modeling_detr.py:281-283 mask
 * downsampling, :322-349 mask-aware sine embedding, :386-411 key-padding mask): the batch is given group after group, n frames of one
 * size per group; every frame is resized on its own (800 / 1333 rule), placed in the top-left corner of a canvas of the largest model
 * input size, padded with zeros AFTER normalisation, and the transformer masks the padded feature cells.  Outputs in group order. */
typedef struct opd_frame_group {
  const uint8_t* frames_dev; /* [n, H0, W0, 3] uint8 */
  int32_t n, H0, W0;
  int32_t frames_are_bgr;
} opd_frame_group;
int opd_detr_workspace_bytes_mixed(const opd_detr* m, const opd_frame_group* groups, int32_t n_groups, size_t* bytes);
int opd_detr_forward_mixed(opd_detr* m, const opd_frame_group* groups, int32_t n_groups, void* workspace_dev,
                           size_t workspace_bytes, float* logits_dev, float* boxes_dev, void* stream);
/* Per-kernel timing of one more forward with the arguments of the last opd_detr_forward: a CUDA event is recorded
 * on `stream` between consecutive launches (the kernels still run back to back on that stream).  Synchronises.
 * Call with max_steps = 0 to get *n_steps.  kinds[i] is an opd_step_kind, flops[i] / bytes[i] are the ALGORITHMIC
 * figures of launch i (2*M*N*K; operands + result once), names [n, name_stride] chars. */
typedef enum opd_step_kind {
  OPD_STEP_ELEMENTWISE = 0, OPD_STEP_GEMM = 1, OPD_STEP_CONV = 2, OPD_STEP_ATTENTION = 3, OPD_STEP_HEADS = 4
} opd_step_kind;
int opd_detr_profile(opd_detr* m, void* stream, int32_t max_steps, int32_t* n_steps, int32_t* kinds, double* flops,
                     double* bytes, float* ms, char* names, int32_t name_stride);
/* Named internal activation of the last forward (tests): "pixel_values", "stem", "pool", "stage{s}.{l}",
 * "enc_in", "pos", "enc{i}", "dec{i}", "dec_out", "resized_u8".  rows x cols; *is_f32: 0 bf16, 1 f32, 2 uint8. */
int opd_detr_tap(const opd_detr* m, const char* name, const void** ptr_dev, int64_t* rows, int64_t* cols,
                 int32_t* is_f32);
/* Copies that activation into dst_dev (device, `bytes` = rows * cols * element size) on `stream`. */
int opd_detr_tap_copy(const opd_detr* m, const char* name, void* dst_dev, size_t bytes, void* stream);
/* post_process_object_detection + person filter (image_processing_detr.py:826-843; yolov8_detector.py:210-241):
 * per query: scores_dev [B,Q], labels_dev [B,Q], xyxy_dev [B,Q,4] (pixels of the ORIGINAL H0 x W0 frame);
 * per frame, compacted in query order: det_xywh_dev [B,Q,4] f64 (x1, y1, x2 - x1, y2 - y1 like the reference's Python
 * floats), det_score_dev [B,Q], det_foot_dev [B,Q,2] f64,
 * det_query_dev [B,Q], n_keep_dev [B]  for  score > threshold && label == person_label;
 * det_slot_dev [B,Q] (optional) = slot_base + frame for the compacted rows, -1 for the unused ones: the
 * `slot_dev` argument of opd_floor_project_classify_count_* when the whole [B*Q] block is projected. */
int opd_detr_postprocess(const float* logits_dev, const float* boxes_dev, int32_t B, int32_t Q, int32_t C,
                         int32_t H0, int32_t W0, float threshold, int32_t person_label, float* scores_dev,
                         int32_t* labels_dev, float* xyxy_dev, double* det_xywh_dev, float* det_score_dev,
                         double* det_foot_dev, int32_t* det_query_dev, int32_t* n_keep_dev, int32_t* det_slot_dev,
                         int32_t slot_base, void* stream);

/* Synthetic frames generated on the device ("100 000 synthetic timelapse frames generated on device per
 * rank from seed = 1000 + global_frame_idx // 64", no host I/O in the timed region): frames_dev [B,H,W,3] uint8, frame b is global
 * frame global_frame0 + b, drawn from a counter-based hash of (seed_base + g / 64, g % 64, y, x, c) - large flat-colour blocks,
 * finer blocks and pixel noise like detection/synthetic.py synthetic_frames (its device_frames_reference restates the kernel). */
int opd_synthetic_frames_u8(uint64_t seed_base, int64_t global_frame0, int32_t B, int32_t H, int32_t W, uint8_t* frames_dev,
                            void* stream);

/* ROI features of the compacted detections (the removed ViTDetector.extract_features / _extract_features_from_outputs,
 * coverage.json method table; arithmetic of FeatureExtractor.extract_roi_features + normalize_features,
 * src/tracking/feature_extractor.py:39-88, :21-37): feat_dev [B, fh, fw, D] bf16 = the encoder output of the last
 * forward (opd_detr_tap "enc5"), boxes det_xywh_dev [B,Q,4] f64 in pixels of the img_h x img_w frame, rows
 * r < n_keep_dev[b] valid -> out_dev [B,Q,D] f32: mean over the box's feature cells, divided by (L2 norm + 1e-8);
 * unused rows are zero. */
int opd_roi_features_bf16(const void* feat_dev, int32_t B, int32_t fh, int32_t fw, int32_t D,
                          const double* det_xywh_dev, const int32_t* n_keep_dev, int32_t Q, int32_t img_h,
                          int32_t img_w, float* out_dev, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Piecewise-affine transform (the reference's shipped default, config.yaml:91)
 * Replaces: src/transform/piecewise_affine.py:155-236 (transform_pixel / transform_detection / transform_batch).
 * The table is built from what the reference builds on the host: scipy.spatial.Delaunay(src).transform [T,3,2]
 * (barycentric transforms), the first two rows of every lstsq affine matrix [T,2,3] (piecewise_affine.py:102-125) and
 * the triangle centroids [T,2] (:146-150); eps = 100 * DBL_EPSILON is scipy's find_simplex tolerance.
 * ---------------------------------------------------------------------------------------------- */
typedef struct opd_pwa_table opd_pwa_table;
int opd_pwa_table_create(const double* bary, const double* affine, const double* centroids, int32_t T, double eps,
                         int32_t device, opd_pwa_table** out);
void opd_pwa_table_destroy(opd_pwa_table* t);
/* in_dev [N,2] points or [N,4] (x, y, w, h) boxes (foot point x + w / 2, y + h), float64 -> floor_px_dev [N,2],
 * optional floor_mm_dev [N,2], in_bounds_dev [N], tri_idx_dev [N] (triangle used), extrapolated_dev [N] (1: outside the
 * triangulation, nearest-centroid triangle).  Enqueues one kernel on `stream`. */
int opd_pwa_transform_f64(const opd_pwa_table* t, const double* in_dev, int32_t input_is_bbox, int64_t N,
                          double scale_x_mm, double scale_y_mm, double map_w_px, double map_h_px,
                          double* floor_px_dev, double* floor_mm_dev, uint8_t* in_bounds_dev, int32_t* tri_idx_dev,
                          uint8_t* extrapolated_dev, void* stream);

/* Thin-plate-spline transform (src/transform/piecewise_affine.py:398-545): control points src_points [n,2], the solved
 * weights_x / weights_y [n] and affine_x / affine_y [3] (order: constant, x, y) as the reference computes them (:445-485).
 * Same point / box inputs and optional outputs as opd_pwa_transform_f64 (no triangle outputs). */
typedef struct opd_tps_table opd_tps_table;
int opd_tps_table_create(const double* src_points, const double* weights_x, const double* weights_y, const double* affine_x,
                         const double* affine_y, int32_t n, int32_t device, opd_tps_table** out);
void opd_tps_table_destroy(opd_tps_table* t);
int opd_tps_transform_f64(const opd_tps_table* t, const double* in_dev, int32_t input_is_bbox, int64_t N, double scale_x_mm,
                          double scale_y_mm, double map_w_px, double map_h_px, double* floor_px_dev, double* floor_mm_dev,
                          uint8_t* in_bounds_dev, void* stream);

/* Lens-distortion correction of points (src/calibration/lens_distortion.py:156-203: cv2.undistortPoints(pts, K, dist, P = K)):
 * in_dev [N,2] points (or [N,4] boxes: foot point) float64 -> out_dev [N,2] corrected pixel coordinates. */
int opd_undistort_points_f64(double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, double k3,
                             const double* in_dev, int32_t input_is_bbox, int64_t N, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPD_B200_H */
