/* retinapost.h — C ABI of libretinapost.so: RetinaNet detection post-processing on B200 (sm_100a).
 *
 * This is the drop-in boundary for the decode / top-k / NMS path of srihari-humbarwadi/retinanet-tensorflow2.x.
 * The reference has no native layer (it is Python on TensorFlow ops); every entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  The Python host layers in
 * retinanet-tensorflow2.x_b200/retinanet/ bind these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every d_* pointer is DEVICE memory owned by the caller, h_* is HOST memory;
 *   - all device work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *     no hidden synchronisation, no allocation after rpp_create (rpp_detect_host owns staging set up lazily);
 *   - the handle is immutable after rpp_create: concurrent calls are safe with distinct workspaces (exception:
 *     rpp_detect_host* keep staging buffers in the handle and are NOT re-entrant per handle); a handle belongs to the
 *     CUDA device that was current at rpp_create, and calls made with another device current return RPP_EINVAL;
 *   - return 0 on success, a negative RPP_E* code otherwise; rpp_last_error() gives the message (thread-local);
 *     no exception crosses the ABI.  The Python shim maps RPP_EMODE -> AssertionError, others -> ValueError /
 *     RuntimeError (the reference raises AssertionError for a bad mode, postprocessing_ops.py:194-197).
 *   - inputs are expected to be finite or +-inf; NaN scores are outside the contract (as in TensorFlow, whose
 *     ordering of NaN scores is unspecified): a NaN never becomes a candidate;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns RPP_ECUDA.
 */
#ifndef RETINAPOST_H_
#define RETINAPOST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPP_OK 0
#define RPP_EINVAL (-1)    /* bad argument / shape */
#define RPP_EMODE (-2)     /* unsupported NMS mode (reference: AssertionError) */
#define RPP_ECOMBO (-3)    /* Global* mode fed per-class (4-D) boxes: rank error in the reference */
#define RPP_EWORKSPACE (-4)/* workspace too small */
#define RPP_ECUDA (-5)     /* CUDA runtime error / no device */

/* GenerateDetections._SUPPORTED_NMS_MODES (model/layers/postprocessing_ops.py:177-183), same order. */
enum rpp_mode {
  RPP_COMBINED_NMS = 0,
  RPP_GLOBAL_SOFT_NMS = 1,
  RPP_GLOBAL_HARD_NMS = 2,
  RPP_PER_CLASS_SOFT_NMS = 3,
  RPP_PER_CLASS_HARD_NMS = 4
};

/* Everything the four reference layers read from `params` (configs/...json: input, architecture.feature_fusion,
 * architecture.head.num_classes, anchor_params, encoder_params, inference) — model/builder.py:153-190. */
typedef struct rpp_config {
  int H, W;                       /* input.input_shape */
  int min_level, max_level;       /* architecture.feature_fusion.{min_level,max_level} */
  int num_classes;                /* architecture.head.num_classes */
  int n_areas;  const double* areas;          /* anchor_params.areas (one per level) */
  int n_ratios; const double* aspect_ratios;  /* anchor_params.aspect_ratios */
  int n_scales; const double* scales;         /* anchor_params.scales */
  float box_variance[4];          /* encoder_params.box_variance */
  int scale_box_targets;          /* encoder_params.scale_box_targets */
  int mode;                       /* inference.mode as enum rpp_mode */
  float iou_threshold;            /* inference.iou_threshold */
  float score_threshold;          /* inference.score_threshold */
  float soft_nms_sigma;           /* inference.soft_nms_sigma (config value; the kernel halves it, :255,:450) */
  int pre_nms_top_k;              /* inference.pre_nms_top_k (<= 0: FilterTopKDetections skipped, builder.py:167) */
  int filter_per_class;           /* inference.filter_per_class */
  int max_detections;             /* inference.max_detections */
  int soft_ignores_iou_threshold; /* 1 = TF >= 2.3 NonMaxSuppressionV5 (default); 0 = older kernel form */
  int tpu_semantics;              /* 1 = GlobalHardNMS / PerClassHardNMS run the reference's TPUStrategy branches
                                     (_tpu_global_hard_nms :381-432, _tpu_per_class_hard_nms :288-379: a real global
                                     hard NMS, tf.image.non_max_suppression_padded arithmetic, int32 classes, -1
                                     padding in every field); the reference selects them by detecting a TPUStrategy
                                     (:199-208), which has no counterpart here.  0 (default) = the non-TPU branches.
                                     With any other mode rpp_create returns RPP_EMODE, as the reference's constructor
                                     raises AssertionError under a TPUStrategy (:202-206). */
  int reserved[6];                /* must be zero */
} rpp_config;

/* Lifetime.  rpp_create validates the config (reference: GenerateDetections.__init__ :185-217 and
 * TransformBoxesAndScores.__init__ :61-85), builds the anchor table on the device (AnchorBoxGenerator,
 * dataloader/anchor_generator.py:24-104) and the exact score-threshold pre-image. */
int rpp_create(const rpp_config* cfg, void** handle);
int rpp_destroy(void* handle);
const char* rpp_last_error(void);

/* AnchorBoxGenerator.boxes / .anchor_boundaries (anchor_generator.py:106-112). */
long rpp_num_anchors(void* handle);
int rpp_num_levels(void* handle);
int rpp_anchor_boundaries(void* handle, long* h_out /* num_levels + 1 */);
int rpp_anchors(void* handle, float* d_out_N4 /* [N,4] cx,cy,w,h */, void* stream);

/* Bytes of device scratch the calls below need for a batch of B images (n: rows per image for rpp_topk /
 * rpp_nms; pass 0 for rpp_detect = number of anchors). */
size_t rpp_workspace_bytes(void* handle, int B, long n);

/* TransformBoxesAndScores.call (postprocessing_ops.py:107-117): scores = sigmoid(logits) [B,N,C],
 * boxes = decode(deltas, anchors) / [H,W,H,W] [B,N,4].  Either output may be NULL. */
int rpp_decode(void* handle, const float* d_logits_BNC, const float* d_deltas_BN4, int B,
               float* d_scores_BNC, float* d_boxes_BN4, void* stream);

/* FilterTopKDetections.call (postprocessing_ops.py:163-173) with the handle's pre_nms_top_k / filter_per_class:
 *   per class : scores_out [B,k',C], boxes_out [B,k',C,4], k' = min(k, n)       (:128-147)
 *   global    : scores_out [B,k',C], boxes_out [B,k',4],   k' = min(k, n*C)     (:149-161)
 * Rows are in tf.nn.top_k sorted=True order (value desc, index asc).  d_index_out (optional): the selected
 * anchor index [B,C,k'] (per class) or flat index [B,k'] (global). */
int rpp_topk(void* handle, const float* d_scores_BnC, const float* d_boxes_Bn4, int B, long n,
             float* d_scores_out, float* d_boxes_out, int* d_index_out,
             void* d_workspace, size_t workspace_bytes, void* stream);

/* FilterTopKDetections applied PER PYRAMID LEVEL (optional extension; BASELINE.json north_star "per-level top-k
 * pre-selection", configs[1] "top-1000/level").  The reference has no such mode — its filter runs over the fused
 * anchor axis (postprocessing_ops.py:128-161) — so the contract is defined by composition: the
 * reference's filter on every segment [anchor_boundaries[l], anchor_boundaries[l+1]) of the fused axis
 * (dataloader/anchor_generator.py:42-49), results concatenated along the row axis in level order.
 *   d_scores_levels[l] [B,n_l,C], d_boxes_levels[l] [B,n_l,4]  (contiguous per level, e.g. the head outputs)
 *   per class : scores_out [B,K,C], boxes_out [B,K,C,4], K = sum_l min(k, n_l);  index_out [B,C,K]
 *   global    : scores_out [B,K,C], boxes_out [B,K,4],   K = sum_l min(k, n_l*C); index_out [B,K] (flat)
 * Indices are positions on the FUSED axis (level offsets added).  Workspace: the largest
 * rpp_workspace_bytes(B, n_l) over the levels (short unsampled levels can need more than long sampled ones). */
int rpp_topk_levels(void* handle, int n_levels, const float* const* d_scores_levels,
                    const float* const* d_boxes_levels, const long* n_rows, int B,
                    float* d_scores_out, float* d_boxes_out, int* d_index_out,
                    void* d_workspace, size_t workspace_bytes, void* stream);

/* GenerateDetections.call (postprocessing_ops.py:537-561), non-TPU branches, with the handle's mode/thresholds.
 *   scores [B,n,C]; boxes [B,n,q,4] with q = 1 (3-D boxes) or q = C (after the per-class filter).
 *   outputs: boxes [B,M,4] f32, scores [B,M] f32, classes [B,M] (f32 Combined / int64 Global* / int32 PerClass*),
 *   valid [B] i32 — dtypes and padding per mode as in the reference. */
int rpp_nms(void* handle, const float* d_scores_BnC, const float* d_boxes_Bnq4, int B, long n, int q,
            float* d_boxes_BM4, float* d_scores_BM, void* d_classes_BM, int* d_valid_B,
            void* d_workspace, size_t workspace_bytes, void* stream);

/* The fused path = ModelBuilder.add_post_processing_stage (model/builder.py:153-190) after FuseDetections:
 * TransformBoxesAndScores -> FilterTopKDetections -> GenerateDetections, without materialising the
 * intermediates.  Argument order mirrors the reference's own fused-op precedent, the EfficientNMS_TRT node of
 * onnx_utils.py:13-85: (raw_boxes [B,N,4], class_logits [B,N,C]; anchors live in the handle). */
int rpp_detect(void* handle, const float* d_deltas_BN4, const float* d_logits_BNC, int B,
               float* d_boxes_BM4, float* d_scores_BM, void* d_classes_BM, int* d_valid_B,
               void* d_workspace, size_t workspace_bytes, void* stream);

/* The fused path reading the model's PER-LEVEL head outputs in place: FuseDetections (postprocessing_ops.py:15-56)
 * only reshapes each level's NHWC tensor [B,H_l,W_l,A*C] / [B,H_l,W_l,A*4] to [B,n_l,C] / [B,n_l,4] and concatenates
 * them; here the levels (min_level..max_level, in order) are consumed where they lie, which saves that copy.
 * Covers CombinedNMS / PerClass*NMS with the per-class filter or no filter, num_classes % 4 == 0, 16-byte aligned
 * level tensors; returns RPP_EINVAL otherwise (fuse and call rpp_detect). */
int rpp_detect_levels(void* handle, const float* const* d_deltas_levels, const float* const* d_logits_levels, int B,
                      float* d_boxes_BM4, float* d_scores_BM, void* d_classes_BM, int* d_valid_B,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Generalisation of the two entries above over the element type of the head outputs.  The reference casts whatever
 * the heads emit to fp32 before anything else (tf.cast, postprocessing_ops.py:111-112); RPP_F16 / RPP_BF16 inputs
 * are converted (exactly) as they are loaded, so half-precision heads stream half the bytes.  n_pieces = 1: fused
 * [B,N,C] / [B,N,4] tensors (d_logits[0], d_deltas[0]); n_pieces = number of levels: per-level head outputs.
 * Same mode coverage as rpp_detect_levels; 16-bit inputs need num_classes % 8 == 0. */
#define RPP_F32 0
#define RPP_F16 1
#define RPP_BF16 2
int rpp_detect_typed(void* handle, int n_pieces, const void* const* d_deltas, const void* const* d_logits, int dtype,
                     int B, float* d_boxes_BM4, float* d_scores_BM, void* d_classes_BM, int* d_valid_B,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* EfficientNMS_TRT-compatible entry: the node the reference appends to the exported graph in mode onnx_tensorrt
 * (onnx_utils.py:13-85) — inputs in the node's order (raw-boxes [B,N,4], class-logits [B,N,C], anchor-boxes [1,N,4] =
 * AnchorBoxGenerator.boxes, [cx,cy,w,h] pixels; NULL = the handle's own table), outputs in the node's order
 * (valid_detections [B,1] i32, detection_boxes [B,M,4], detection_scores [B,M], detection_classes [B,M] i32) — with
 * the attributes the reference sets: max_output_boxes = max_detections, score_threshold, iou_threshold,
 * score_activation = True (sigmoid), box_coding = 1 (centre-size boxes decoded against the anchors, no variance
 * scaling, no normalisation; outputs in the same coding, pixels), background_class = -1 (none), class-aware
 * suppression.  The handle's NMS mode and pre_nms_top_k are not used.
 * Semantics as published for TensorRT's efficientNMSPlugin: candidates = sigmoid(logit) >= score_threshold; per image
 * the 4096 best (anchor, class) pairs; greedy in score order, a box is dropped when a kept box of the same class has
 * IoU > iou_threshold; the first max_output_boxes kept; outputs zero-filled beyond the count.  Ties are ordered by
 * flat index (the plugin's own tie order is unspecified).  PARITY UNPINNED: the plugin is not part of the reference
 * tree nor installed here; the oracle restates the same algorithm (rpp_ref_efficient_nms). */
int rpp_efficient_nms(void* handle, const float* d_raw_boxes_BN4, const float* d_class_logits_BNC,
                      const float* d_anchor_boxes_N4, int B, int* d_valid_detections_B1, float* d_detection_boxes_BM4,
                      float* d_detection_scores_BM, int* d_detection_classes_BM,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Same, from HOST buffers (the serving call of export.py:233-253 / evaluate_saved_model.py: tensors arrive from
 * the host and detections are consumed on the host).  Copies H2D in image chunks overlapped with the kernels,
 * then D2H of the four outputs; synchronises before returning.  Pinned host memory is recommended.
 * device: CUDA ordinal.  Staging buffers are allocated on first use and kept in the handle. */
int rpp_detect_host(void* handle, int device, const float* h_deltas_BN4, const float* h_logits_BNC, int B,
                    float* h_boxes_BM4, float* h_scores_BM, void* h_classes_BM, int* h_valid_B);

/* rpp_detect_host over f32 / f16 / bf16 host buffers (same restrictions as rpp_detect_typed for the 16-bit types):
 * half-precision heads halve the PCIe traffic that bounds this call. */
int rpp_detect_host_typed(void* handle, int device, const void* h_deltas_BN4, const void* h_logits_BNC, int dtype,
                          int B, float* h_boxes_BM4, float* h_scores_BM, void* h_classes_BM, int* h_valid_B);

/* COCO post-formatting epilogue = the per-image loop of COCOEvaluator.accumulate_results
 * (eval/coco_evaluator.py:111-134) on the device: rows [0, valid) of every image, boxes divided by
 * (resize_scale / input_shape) tiled to 4 (skipped when d_resize_scale_B2 is NULL: rescale_detections=False),
 * truncated to int32 and turned from x1,y1,x2,y2 into x,y,w,h; class ids optionally remapped through
 * d_class_map[num_classes] (remap_class_ids).  Outputs are COMPACTED in image order: bbox [T,4] i32, category [T] i32,
 * score [T] f32, image index [T] i32 (position in the batch), and T itself; size the outputs for B * max_detections. */
int rpp_coco_format(void* handle, const float* d_boxes_BM4, const float* d_scores_BM, const void* d_classes_BM,
                    const int* d_valid_B, int B, const float* d_resize_scale_B2, const int* d_class_map,
                    int* d_bbox_out_T4, int* d_category_out_T, float* d_score_out_T, int* d_image_out_T,
                    int* d_total_out, void* stream);

/* Number of kernels the last rpp_detect / rpp_nms / rpp_topk / rpp_decode call on this thread launched. */
int rpp_last_launch_count(void);

/* Size in bytes of one element of the classes output for the handle's mode (4, 8 or 4). */
int rpp_classes_itemsize(void* handle);

/* Debug/testing knobs (not part of the reference surface): force the exact slow path of the candidate
 * selection (1), or restore the default sampled pre-threshold (0). */
int rpp_debug_force_exact_scan(void* handle, int on);

/* How often the sampled pre-threshold missed: the number of problems ((image, class) columns, or flat columns of the
 * global filter) that left their candidate list for an exact scan / re-collection of the whole column in the calls on
 * this handle since the last reset — results never depend on it, speed does (DESIGN.md "perf cliff").  Synchronises
 * the device.  No reference counterpart. */
int rpp_debug_exact_scans(void* handle, unsigned long long* h_count, int reset);

/* Test hook (no device needed): the sampling plan the library would use for columns of n rows and C classes —
 * NMS problems (emit = 0, k_lim ignored) or top-k emission of k_lim rows (emit = 1).  h_out[8] = {sampled?, stride,
 * groups G, sampled rows per group, rank of the group maximum used as threshold, list capacity, targeted list
 * length, fine (4x groups) plan?}.  tests/test_host_cpu.py checks the plan's margins by Monte Carlo. */
int rpp_debug_sample_plan(long n, int C, long k_lim, int emit, int* h_out);

/* Per-stage device timing for bench.py's roofline: when on, every rpp_detect / rpp_nms call records CUDA events on
 * the caller's stream at the stage boundaries.  rpp_debug_stage_ms synchronises on the last event and returns the
 * mean milliseconds of the 4 stages {pre-threshold sample, collect stream, NMS problems, merge} over the calls
 * recorded since the last read. */
int rpp_debug_stage_timing(void* handle, int on);
int rpp_debug_stage_ms(void* handle, float* h_ms4, int* n_calls);
/* The same recording by named segment: writes "label=ms;label=ms;..." (mean milliseconds per call, labels in launch
 * order, e.g. "sample", "collect", "nms", "merge", "emit:collect", "rows") into buf and resets the recording. */
int rpp_debug_stage_report(void* handle, char* buf, int cap, int* n_calls);

#ifdef __cplusplus
}
#endif
#endif /* RETINAPOST_H_ */
