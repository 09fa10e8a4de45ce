/*
 * rodet_b200 — C ABI of the B200-native box-level hot path
 * (anchor generation, ARM/ODM matching + encoding, decode, select, top-k, NMS)
 * of the RefineDet-style detector in YoungYoung619/road-object-detection-for-bdd100k.
 *
 * The reference has NO native interface for this path: it is ordinary Python that
 * builds TensorFlow-1 graph ops.  The boundary a maintainer would bind is therefore
 * the set of Python functions listed below; each entry point cites the reference
 * function (file:line under /root/reference) it replaces.  INTEGRATION.md shows the
 * ctypes stub.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Every data pointer is a DEVICE
 *    pointer unless the parameter comment says "host".
 *  - float = IEEE binary32, row-major, innermost dimension contiguous.
 *  - Boxes are y-first: centre [cy,cx,h,w], corner [ymin,xmin,ymax,xmax], normalised.
 *  - Anchors are flattened layer-major, then (fy, fx, a) row-major
 *    (utils/net_tools.py:200,220-223,679-682,731-735).
 *  - Every function is stream-ordered on `stream` (a cudaStream_t passed as void*),
 *    never allocates, never synchronises, never reads results back to the host.
 *  - Return value: 0 on success, a negative ROD_E_* code otherwise; the message is
 *    available from rod_last_error() (thread-local).
 *  - There is no CPU fallback anywhere in this library.
 */
#ifndef RODET_B200_H_
#define RODET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROD_ABI_VERSION 1
#define ROD_MAX_LAYERS 8
#define ROD_MAX_TOPK 1024          /* top_k / NMS candidates per (image, class) */
#define ROD_MAX_CLASSES 64

enum {
  ROD_OK = 0,
  ROD_E_INVALID = -1,      /* bad argument (null pointer, size, unsupported value) */
  ROD_E_CUDA = -2,         /* a CUDA runtime call / launch failed */
  ROD_E_UNSUPPORTED = -3,  /* valid in the reference but not implemented here */
  ROD_E_DLPACK = -4        /* DLPack tensor on wrong device / dtype / layout */
};

/* config.refine_method (config.py:74-77) */
enum { ROD_NEAREST_NEIGHBOR = 0, ROD_JACCARD_BIGGER = 1, ROD_JACCARD_TOPK = 2 };

/* Anchor layout: n_layers feature maps, offset[l]..offset[l+1] is layer l's slice of
 * the flat anchor axis; offset[n_layers] == n_total. */
typedef struct rod_layout {
  int32_t n_layers;
  int32_t n_total;
  int32_t offset[ROD_MAX_LAYERS + 1];
} rod_layout_t;

/* A "list of per-layer tensors" (the reference passes Python lists of 6 tensors of
 * shape [B, fh, fw, A, inner], nets/catch_net.py:306-308,339-341).  base[l] points at
 * element (b=0, anchor 0 of layer l); consecutive images are batch_stride[l] ELEMENTS
 * apart; within an image the layer slice is contiguous. */
typedef struct rod_layered {
  const void* base[ROD_MAX_LAYERS];
  int64_t batch_stride[ROD_MAX_LAYERS];
} rod_layered_t;

const char* rod_last_error(void);
int rod_version(void);
/* host out-params; any may be NULL */
int rod_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);

/* ---- a1-a4  anchors ------------------------------------------------------------
 * Replaces init_anchor / anchors_one_layer / anchors_all_layer
 * (utils/net_tools.py:21-82, 98-122, 125-142) plus the corner and re-derived centre
 * forms every consumer recomputes (utils/net_tools.py:156-171, 203-217, 385-394).
 * feat_h/feat_w/n_anchor: host int32[n_layers]; sizes_px: host double[sum(n_anchor)][2]
 * = pixel (h, w) per anchor as produced by init_anchor.  Cell centres and normalised
 * sizes are evaluated in float64 and rounded to float32 exactly like the NumPy code.
 * Outputs: corner[N,4] (ymin,xmin,ymax,xmax), center[N,4] (acy,acx,ah,aw),
 * yxhw[N,4] (y,x,h,w as anchors_one_layer returns them; may be NULL). */
int rod_anchor_table(int n_layers, const int32_t* feat_h, const int32_t* feat_w,
                     const int32_t* n_anchor, const double* sizes_px, int img_h, int img_w,
                     float* corner, float* center, float* yxhw, void* stream);

/* Same table from an existing anchors_all_layer() result: y,x = device float[sum fh*fw]
 * (layer-major, row-major cells), h,w = device float[sum n_anchor]. */
int rod_anchor_table_from_grid(int n_layers, const int32_t* feat_h, const int32_t* feat_w,
                               const int32_t* n_anchor, const float* y, const float* x,
                               const float* h, const float* w, float* corner, float* center,
                               void* stream);

/* ---- a9  ARM matching + encode ---------------------------------------------------
 * Replaces refine_groundtruth (utils/net_tools.py:270-428), batched over images (the
 * reference runs it per image and tf.train.batch stacks, train.py:109-124).
 * thresholds: host float[n_layers] = config.refine_pos_jac_val_all_layers.
 * center_bboxes[B,gmax,4], labels[B,gmax] (int64 if labels_i64 else int32),
 * gt_counts: device int32[B] (NULL => every image has gmax boxes); each count >= 1.
 * Outputs, flat over the anchor axis: gt[B,N,4], cbboxes[B,N,4], out_labels[B,N],
 * pos_mask[B,N], match_idx[B,N] (argmax over GT, lowest index on ties; may be NULL).
 * method ROD_JACCARD_TOPK returns ROD_E_UNSUPPORTED (reference raises
 * ValueError('Not support now'), utils/net_tools.py:423-424). */
int rod_arm_match_encode(const rod_layout_t* layout, const float* anchors_corner,
                         const float* anchors_center, const float* thresholds,
                         const float* center_bboxes, const void* labels, int labels_i64,
                         const int32_t* gt_counts, int batch, int gmax, int method,
                         float* gt, float* cbboxes, int32_t* out_labels, int32_t* pos_mask,
                         int32_t* match_idx, void* stream);

/* Opt-in extra WITHOUT a reference counterpart, the "forced match" named by BASELINE.json's
 * north star: a post-pass over rod_arm_match_encode's outputs in which every GT box also claims the anchor
 * it overlaps best (argmax over the anchor axis, lowest anchor index on ties, IoU > 0) whatever the layer
 * threshold; several GT boxes on one anchor: highest IoU wins, ties lowest GT index.  The reference's
 * JACCARD_BIGGER has no such step (utils/net_tools.py:405-408); never the default.
 * workspace: rod_arm_forced_match_workspace_bytes(batch, gmax) bytes, 8-byte aligned. */
size_t rod_arm_forced_match_workspace_bytes(int batch, int gmax);
int rod_arm_forced_match(const rod_layout_t* layout, const float* anchors_corner,
                         const float* anchors_center, const float* center_bboxes, const void* labels,
                         int labels_i64, const int32_t* gt_counts, int batch, int gmax, float* gt,
                         float* cbboxes, int32_t* out_labels, int32_t* pos_mask, int32_t* match_idx,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- a10  ODM target generation --------------------------------------------------
 * Replaces det_groundtruth (utils/net_tools.py:431-475).
 * thresholds: host float[n_layers] = config.det_pos_jac_val_all_layers.
 * Inputs are per-layer lists: refine_out/offset_gt/cbboxes inner 4 (float),
 * refine_labels/refine_pos_mask inner 1 (int32).
 * Outputs flat: det_gt[B,N,4], mask[B,N], det_labels[B,N], iou[B,N]. */
int rod_odm_target(const rod_layout_t* layout, const float* anchors_center,
                   const float* thresholds, const rod_layered_t* refine_out,
                   const rod_layered_t* offset_gt, const rod_layered_t* cbboxes,
                   const rod_layered_t* refine_labels, const rod_layered_t* refine_pos_mask,
                   int batch, float* det_gt, int32_t* mask, int32_t* det_labels, float* iou,
                   void* stream);

/* ---- a9 + a10 fused: training-target generation in one launch ----------------------
 * Replaces the call sequence train.py:109-113 -> :147-149, i.e. refine_groundtruth
 * (utils/net_tools.py:382-421, JACCARD_BIGGER) followed by det_groundtruth (:431-475) on the
 * ARM head output.  Same inputs as rod_arm_match_encode plus refine_out (per-layer list, inner 4)
 * and the ODM thresholds; the same eight outputs as the two calls, bit for bit.  cbboxes,
 * out_labels and match_idx may be NULL (a training step only consumes gt / pos_mask and the four
 * ODM outputs).  Per anchor: 16 B read + 68 B written instead of 124 B for the two-call path.
 * workspace: rod_target_fused_workspace_bytes() (= 8) bytes of device memory, 8-byte aligned, holding
 * the work-item counter of the kernel's dynamic tile scheduler: it must be ZERO before the first call
 * and every call leaves it zero again, so one cudaMemset at allocation time is enough.  Calls that may
 * run concurrently (different streams) need different workspaces. */
size_t rod_target_fused_workspace_bytes(void);
int rod_target_fused(const rod_layout_t* layout, const float* anchors_corner,
                     const float* anchors_center, const float* arm_thresholds,
                     const float* odm_thresholds, const float* center_bboxes, const void* labels,
                     int labels_i64, const int32_t* gt_counts, int batch, int gmax,
                     const rod_layered_t* refine_out, float* gt, float* cbboxes,
                     int32_t* out_labels, int32_t* pos_mask, int32_t* match_idx, float* det_gt,
                     int32_t* det_mask, int32_t* det_labels, float* iou, void* workspace,
                     void* stream);

/* ---- a7 / a17  decode ------------------------------------------------------------
 * Replaces decode_locations_one_layer (utils/net_tools.py:182-234) and the inference
 * call site c2c(decode(anchors, refine_out + det_out)) (evaluate.py:139-143,
 * predict.py:130-134).  det_out may be NULL (single decode).  to_corner != 0 applies
 * centerBboxes_2_cornerBboxes.  out flat [B,N,4]. */
int rod_decode(const rod_layout_t* layout, const float* anchors_center,
               const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
               int to_corner, float* out, void* stream);
/* Opt-in extra WITHOUT a reference counterpart (the reference decodes refine_out + det_out once,
 * SURVEY.md 0.5): the RefineDet cascade named by BASELINE.json's north star — det_out decoded against
 * the refined anchors decode(anchors, refine_out), i.e. decode_locations_one_layer applied twice with its
 * corner -> re-derived-centre anchor step (utils/net_tools.py:156-171) in between.  Never the default. */
int rod_decode_cascade(const rod_layout_t* layout, const float* anchors_center,
                       const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                       int to_corner, float* out, void* stream);

/* ---- a6  encode one box against every anchor -------------------------------------
 * Replaces encode_locations_one_layer (utils/net_tools.py:147-179): center_bbox is a
 * device float[4]; out[n,4] for anchors [first, first+n). */
int rod_encode_one_box(const float* anchors_center, int first, int n, const float* center_bbox,
                       float* out, void* stream);

/* ---- a5, a8, a14, a16  element-wise box helpers -----------------------------------
 * n = number of boxes; every array is [n,4] unless noted.
 * rod_center_to_corner / rod_corner_to_center: utils/common_tools.py:16-35 / 38-57.
 * rod_jaccard: net_tools.jaccard (utils/net_tools.py:237-267); b is [n,4] or, when
 *   b_broadcast != 0, a single box [4]; out[n]; plain divide.
 * rod_bboxes_jaccard / rod_bboxes_intersection: utils/tf_extended/bboxes.py:452-479 /
 *   482-508 (safe_divide, utils/tf_extended/math.py:25-38); ref is [4] or [n,4].
 * rod_bboxes_clip / rod_bboxes_resize: utils/tf_extended/bboxes.py:103-136 / 139-163. */
int rod_center_to_corner(const float* in, float* out, int64_t n, void* stream);
int rod_corner_to_center(const float* in, float* out, int64_t n, void* stream);
int rod_jaccard(const float* a, const float* b, int b_broadcast, float* out, int64_t n,
                void* stream);
int rod_bboxes_jaccard(const float* ref, int ref_broadcast, const float* boxes, float* out,
                       int64_t n, void* stream);
int rod_bboxes_intersection(const float* ref, int ref_broadcast, const float* boxes,
                            float* out, int64_t n, void* stream);
int rod_bboxes_clip(const float* ref, int ref_broadcast, const float* boxes, float* out,
                    int64_t n, void* stream);
int rod_bboxes_resize(const float* ref, const float* boxes, float* out, int64_t n,
                      void* stream);

/* ---- a11  select -----------------------------------------------------------------
 * Replaces bboxes_select_one_layer / bboxes_select_all_layers
 * (utils/net_tools.py:658-736): for class c != ignore_class,
 * scores_c = p_c * (p_c >= thr), bboxes_c = loc * (p_c >= thr).
 * predictions: per-layer list, inner n_classes; localizations: per-layer list, inner 4.
 * Outputs class-major: out_scores[n_classes][B][N], out_bboxes[n_classes][B][N][4]
 * (slot ignore_class is left untouched). */
int rod_bboxes_select(const rod_layout_t* layout, const rod_layered_t* predictions,
                      const rod_layered_t* localizations, int batch, int n_classes,
                      int ignore_class, float select_threshold, float* out_scores,
                      float* out_bboxes, void* stream);

/* ---- a12  sort / top-k -----------------------------------------------------------
 * Replaces tfe.bboxes_sort (utils/tf_extended/bboxes.py:60-100): per row of
 * scores[rows,n] take the top_k largest (descending, equal scores -> lower index
 * first, TF TopKV2), gather boxes[rows,n,4].  Outputs out_scores[rows,k],
 * out_bboxes[rows,k,4], out_idx[rows,k] (may be NULL).  1 <= k <= min(n, ROD_MAX_TOPK).
 * workspace: none. */
int rod_bboxes_sort(const float* scores, const float* bboxes, int64_t rows, int n, int top_k,
                    float* out_scores, float* out_bboxes, int32_t* out_idx, void* stream);

/* ---- a13  NMS ---------------------------------------------------------------------
 * Replaces tfe.bboxes_nms / bboxes_nms_batch (utils/tf_extended/bboxes.py:166-232)
 * incl. the zero padding of pad_axis (utils/tf_extended/tensors.py:59-86):
 * greedy tf.image.non_max_suppression over all n candidates of each row (descending
 * score, ties -> lower index; suppress iff IoU > nms_threshold), outputs in selection
 * order padded with zeros to keep_top_k.  n <= ROD_MAX_TOPK.
 * out_scores[rows,keep], out_bboxes[rows,keep,4], out_idx[rows,keep] (-1 padded; may be
 * NULL). */
int rod_bboxes_nms_batch(const float* scores, const float* bboxes, int64_t rows, int n,
                         float nms_threshold, int keep_top_k, float* out_scores,
                         float* out_bboxes, int32_t* out_idx, void* stream);

/* ---- a15  fused post-process ------------------------------------------------------
 * Replaces detected_bboxes (utils/net_tools.py:739-758) = select -> top_k -> NMS ->
 * pad (-> clip), optionally with the decode call site fused in front
 * (evaluate.py:139-151):
 *   localizations != NULL : corner boxes are given (drop-in detected_bboxes);
 *   localizations == NULL : boxes are c2c(decode(anchors_center, refine_out+det_out))
 *                            evaluated only for the top_k candidates.
 * clip_box: device float[4] or NULL.  Outputs class-major over classes
 * c = 0..n_classes-1 (slot ignore_class untouched):
 *   out_scores[n_classes][B][keep], out_bboxes[n_classes][B][keep][4],
 *   out_counts[n_classes][B] int32 = number of detections with non-zero score
 *   (may be NULL; this is what the per-rank NCCL count all-gather ships).
 * workspace: rod_detect_workspace_bytes(batch, n_classes, top_k) bytes, 256-aligned.  For full speed
 * the rod_detect_workspace_clean_bytes(...) bytes starting at rod_detect_flags_offset(...) should be
 * zero before the FIRST call with a workspace (n_classes == 11, select_threshold > 0): they hold the
 * sampled score histogram that steers the candidate cuts, and the library leaves them zero after
 * every call instead of clearing them in front of every call.  Results never depend on it: stale
 * contents only send segments to the exact general kernels on that one call. */
size_t rod_detect_workspace_bytes(const rod_layout_t* layout, int batch, int n_classes,
                                  int top_k);
size_t rod_detect_workspace_clean_bytes(const rod_layout_t* layout, int batch, int n_classes,
                                        int top_k);
/* Diagnostics: byte offset inside the workspace of uint32 flags[n_classes][B] that the last
 * rod_detect / rod_detect_logits call with select_threshold > 0 left behind: non-zero = the segment
 * (class, image) was handed to the exact general kernels (candidate list overflow, sampled score cut
 * too high, massive ties).  The result never depends on it; bench.py reports the rate. */
size_t rod_detect_flags_offset(const rod_layout_t* layout, int batch, int n_classes, int top_k);
int rod_detect(const rod_layout_t* layout, const float* anchors_center,
               const rod_layered_t* predictions, const rod_layered_t* localizations,
               const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
               int n_classes, int ignore_class, float select_threshold, float nms_threshold,
               int top_k, int keep_top_k, const float* clip_box, float* out_scores,
               float* out_bboxes, int32_t* out_counts, void* workspace, size_t workspace_bytes,
               void* stream);

/* ---- f-1  softmax in front of the select stage -----------------------------------------
 * rod_softmax replaces slim.softmax(clf_out[i]) (evaluate.py:136-137, predict.py:127-128) for one
 * [rows, n_classes] tensor (in place allowed).  rod_detect_logits is rod_detect with the class
 * LOGITS as `logits`: the softmax is fused into the select pass (no probability tensor is written);
 * its result is bit-identical to rod_softmax followed by rod_detect.  Probabilities follow the
 * float32 softmax to < 1e-6 relative (hardware ex2), i.e. tolerance parity with the reference.
 * Same workspace as rod_detect. */
int rod_softmax(const float* logits, int64_t rows, int n_classes, float* out, void* stream);
int rod_detect_logits(const rod_layout_t* layout, const float* anchors_center,
                      const rod_layered_t* logits, const rod_layered_t* localizations,
                      const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                      int n_classes, int ignore_class, float select_threshold, float nms_threshold,
                      int top_k, int keep_top_k, const float* clip_box, float* out_scores,
                      float* out_bboxes, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- f-2  evaluation TP / FP matching ----------------------------------------------
 * Replaces tfe.bboxes_matching / bboxes_matching_batch for one class
 * (utils/tf_extended/bboxes.py:246-380): scores[rows,n] (unused by the matching itself, kept for
 * signature parity), bboxes[rows,n,4] detections in score order; glabels[rows,g], gbboxes[rows,g,4],
 * gdifficults[rows,g] (int64 if labels_i64 else int32; zero-padded GT has label 0).
 * Outputs: out_n_gbboxes[rows] int64 (non-difficult GT of this class), out_tp / out_fp [rows,n] bool. */
int rod_bboxes_matching_batch(int64_t label, const float* scores, const float* bboxes, const void* glabels,
                              const float* gbboxes, const void* gdifficults, int labels_i64, int rows,
                              int n, int g_n, float matching_threshold, int64_t* out_n_gbboxes,
                              uint8_t* out_tp, uint8_t* out_fp, void* stream);

/* ---- f-2  evaluation metrics (evaluate.py:162-197) ----------------------------------
 * rod_tpfp_append replaces the update of tfe.streaming_tp_fp_arrays
 * (utils/tf_extended/metrics.py:133-204) for `rows` classes at once: scores / tp / fp [rows, n]
 * (n = batch * keep_top_k detections in image order); keeps, in order, the entries with tp | fp and
 * score > rm_threshold (all entries when remove_zero_scores == 0, as the reference does) and appends
 * them to v_scores / v_tp / v_fp [rows, capacity] at v_count[rows] (device counters, updated);
 * v_ids (optional) receives id_base + position, a global detection id that makes merges across
 * ranks order-independent; v_nobjects[rows] += sum(num_gbboxes[rows, n_gb]).  The caller guarantees
 * capacity >= v_count + n (entries beyond capacity are dropped, the counter still advances).
 * rod_precision_recall replaces tfe.precision_recall (:100-130) AFTER the descending score sort:
 * float64 cumulative sums, recall = tp_c / num_gbboxes, precision = tp_c / (tp_c + fp_c), 0 where
 * the denominator is <= 0.  num_gbboxes: device int64 scalar.
 * rod_average_precision replaces tfe.average_precision_voc07 and _voc12 (:210-258):
 * out[0] = VOC07 11-point AP (thresholds07 = the 11 recall levels np.arange(0., 1.1, 0.1), host),
 * out[1] = VOC12 area AP, both float64 on the device. */
int rod_tpfp_append(const float* scores, const uint8_t* tp, const uint8_t* fp, int rows, int64_t n,
                    const int64_t* num_gbboxes, int n_gb, int remove_zero_scores, float rm_threshold,
                    int64_t id_base, float* v_scores, uint8_t* v_tp, uint8_t* v_fp, int64_t* v_ids,
                    int64_t capacity, int64_t* v_count, int64_t* v_nobjects, void* stream);
/* rod_sort_scores_desc replaces the sort in front of it: tf.nn.top_k(scores, k, sorted=True) + gather
 * (utils/tf_extended/metrics.py:117-123): the k best of n scores in descending order, equal scores in index
 * order; emits any of tp_sorted / fp_sorted (uint8), scores_sorted, idx_sorted (NULL to skip). */
size_t rod_sort_scores_workspace_bytes(int64_t n);
int rod_sort_scores_desc(const float* scores, int64_t n, int64_t k, const uint8_t* tp, const uint8_t* fp,
                         uint8_t* tp_sorted, uint8_t* fp_sorted, float* scores_sorted,
                         int32_t* idx_sorted, void* workspace, size_t workspace_bytes, void* stream);
size_t rod_precision_recall_workspace_bytes(int64_t n);
int rod_precision_recall(const uint8_t* tp_sorted, const uint8_t* fp_sorted, int64_t n,
                         const int64_t* num_gbboxes, double* precision, double* recall,
                         void* workspace, size_t workspace_bytes, void* stream);
int rod_average_precision(const double* precision, const double* recall, int64_t n,
                          const double* thresholds07, double* out_voc07_voc12, void* stream);

/* ---- f-3  losses on the ARM / ODM targets (utils/net_tools.py:478-623) ---------------
 * rod_smooth_l1_loss: out_loss[0] = sum_{b,n,k} smooth_l1((y - x) * mask) / batch over all layers
 * (smooth_l1 :478-489; refine_loss :492-516; det_loss :538-551).  y, x: per-layer [B,..,4] f32,
 * mask: per-layer [B,..,1] i32.  grad_x (optional, flat [B,N,4]) receives d(loss * batch * grad_scale)/dx.
 * rod_clf_loss: classification loss of det_clf_loss (:553-615).  logits per-layer [B,..,C] f32, labels and
 * mask per-layer [B,..,1] i32 (det_labels, det_pos_mask), iou per-layer [B,..] f32 (iou_all_layers).
 * out6 = {clf_loss, pos_loss, neg_loss, max_hard_pred, n_pos, n_neg} (device floats).  The workspace keeps
 * what rod_clf_loss_grad needs: grad_logits (flat [B,N,C]) = upstream * d clf_loss / d logits, with masks,
 * labels and the IoU factor treated as constants.  Tolerance parity (1e-5 relative). */
size_t rod_smooth_l1_workspace_bytes(const rod_layout_t* layout, int batch);
int rod_smooth_l1_loss(const rod_layout_t* layout, const rod_layered_t* y, const rod_layered_t* x,
                       const rod_layered_t* mask, int batch, float* out_loss, float* grad_x,
                       float grad_scale, void* workspace, size_t workspace_bytes, void* stream);
size_t rod_clf_loss_workspace_bytes(const rod_layout_t* layout, int batch);
int rod_clf_loss(const rod_layout_t* layout, const rod_layered_t* logits, const rod_layered_t* labels,
                 const rod_layered_t* mask, const rod_layered_t* iou, int batch, int n_classes,
                 float negative_ratio, float* out6, void* workspace, size_t workspace_bytes, void* stream);
int rod_clf_loss_grad(const rod_layout_t* layout, const rod_layered_t* logits, const rod_layered_t* labels,
                      const rod_layered_t* mask, const rod_layered_t* iou, int batch, int n_classes,
                      float upstream, const void* workspace, float* grad_logits, void* stream);

/* ---- f-4  ground-truth boxes of the training input pipeline -------------------------
 * The box half of process_raw_data_train (utils/data_pileline_tools.py:88-108), the step right
 * before refine_groundtruth, fused and batched: per image b
 *   bboxes = tfe.bboxes_resize(distort_bbox[b], bboxes)           (bboxes.py:139-163; skipped if NULL)
 *   labels, bboxes = tfe.bboxes_filter_overlap(labels, bboxes, threshold, assign_negative)
 *                                                                 (bboxes.py:408-428; if filter_overlap)
 *   bboxes = flip_bboxes(bboxes) where mirror[b] != 0             (tf_image.py:284-289; skipped if NULL)
 *   bboxes = min(max(bboxes, 0), 1)                               (data_pileline_tools.py:107-108; if clamp01)
 * bboxes [batch,gmax,4] corner form, labels [batch,gmax] (int64 if labels_i64 else int32),
 * counts[batch] valid boxes per image (NULL = gmax).  Kept boxes stay in order, the rest of each
 * row is zeroed, out_counts[batch] receives the new counts. */
int rod_gt_boxes_update(const float* bboxes, const void* labels, int labels_i64, const int32_t* counts,
                        int batch, int gmax, const float* distort_bbox, const uint8_t* mirror,
                        int filter_overlap, float threshold, int assign_negative, int clamp01,
                        float* out_bboxes, void* out_labels, int32_t* out_counts, void* stream);

/* ---- f-4  ground truth from the on-disk format --------------------------------------
 * The reference stores one tf.train.Example per image in TFRecord files (writer
 * dataset/pascalvoc_to_tfrecords.py:128-170, schema dataset/pascalvoc_common.py:75-98) and reads
 * them through slim's DatasetDataProvider (utils/data_pileline_tools.py:33-71).  These calls
 * replace the ground-truth half of that reader: a HOST parser of the file bytes (image decoding is
 * out of scope) and a device gather that assembles padded batches.
 *   rod_tfrecord_index   : counts the records and their objects (`data` = n_bytes of one or more
 *                          concatenated TFRecord files in host memory; verify_crc checks the masked
 *                          CRC-32C of every record).
 *   rod_tfrecord_read_gt : fills host arrays — ymin/xmin/ymax/xmax[n_objects] float32 (the
 *                          'image/object/bbox/{ymin,...}' features), label/difficult/truncated[n_objects]
 *                          int64 (difficult / truncated may be NULL), offsets[n_records + 1] (objects
 *                          of record r = [offsets[r], offsets[r+1])), shape[n_records][3] (may be NULL).
 *   rod_gt_gather        : DEVICE arrays (the ones above, uploaded once) + indices[batch] (record per
 *                          image; NULL = 0..batch-1) -> bboxes[batch][gmax][4] corner form
 *                          (ymin,xmin,ymax,xmax), labels[batch][gmax], difficults[batch][gmax] (may be
 *                          NULL), counts[batch] = min(#objects, gmax); rows zero padded.  An index
 *                          outside [0, n_records) gives a zero row and counts = -1.  This is the
 *                          input form of rod_gt_boxes_update / rod_corner_to_center / rod_arm_match_encode. */
int rod_tfrecord_index(const void* data, size_t n_bytes, int verify_crc, int64_t* n_records,
                       int64_t* n_objects);
int rod_tfrecord_read_gt(const void* data, size_t n_bytes, int verify_crc, int64_t n_records,
                         int64_t n_objects, float* ymin, float* xmin, float* ymax, float* xmax,
                         int64_t* label, int64_t* difficult, int64_t* truncated, int64_t* offsets,
                         int64_t* shape);
int rod_gt_gather(const float* ymin, const float* xmin, const float* ymax, const float* xmax,
                  const int64_t* label, const int64_t* difficult, const int64_t* offsets,
                  const int64_t* indices, int64_t n_records, int batch, int gmax, float* bboxes,
                  int64_t* labels, int64_t* difficults, int32_t* counts, void* stream);

/* ---- measurement helpers (bench.py) ------------------------------------------------
 * rod_peak_fp32_nofma: runs a dependent-chain FADD/FMUL (no FMA) kernel and returns in
 * *ops the number of FP32 instructions-lanes issued; time it with events on `stream`. */
int rod_peak_fp32_nofma(int iters, float* sink, double* ops, void* stream);
/* rod_l2_flush: writes `bytes` bytes of `buf` (use > L2 size) */
int rod_l2_flush(void* buf, size_t bytes, void* stream);

/* ---- DLPack front door --------------------------------------------------------------
 * Same operations, tensors passed as borrowed `DLTensor*` (dlpack.h v0.8 layout; the
 * Python host hands `DLManagedTensor*` from torch.utils.dlpack.to_dlpack, whose first
 * member is the DLTensor).  The library never calls the deleter.  Device type, dtype,
 * rank/shape and contiguity are validated here (ROD_E_DLPACK on mismatch).
 * Per-layer lists are host arrays of n_layers `DLTensor*`. */
struct DLTensor;
/* Fills a rod_layered_t from a host array of n_layers borrowed DLTensors [B, ..., inner]
 * (validated: CUDA device, dtype (code,bits), per-image element count, contiguity after the batch
 * dim, 16-byte alignment for inner == 4).  *batch: in = expected batch or -1, out = batch found. */
int rod_dl_layered(const rod_layout_t* layout, const struct DLTensor* const* tensors, int inner,
                   int dtype_code, int dtype_bits, rod_layered_t* out, int* batch);
int rod_dl_arm_match_encode(const rod_layout_t* layout, const struct DLTensor* anchors_corner,
                            const struct DLTensor* anchors_center, const float* thresholds,
                            const struct DLTensor* center_bboxes, const struct DLTensor* labels,
                            const struct DLTensor* gt_counts, int method,
                            const struct DLTensor* gt, const struct DLTensor* cbboxes,
                            const struct DLTensor* out_labels, const struct DLTensor* pos_mask,
                            const struct DLTensor* match_idx, void* stream);
int rod_dl_odm_target(const rod_layout_t* layout, const struct DLTensor* anchors_center,
                      const float* thresholds, const struct DLTensor* const* refine_out,
                      const struct DLTensor* const* offset_gt, const struct DLTensor* const* cbboxes,
                      const struct DLTensor* const* refine_labels,
                      const struct DLTensor* const* refine_pos_mask, const struct DLTensor* det_gt,
                      const struct DLTensor* mask, const struct DLTensor* det_labels,
                      const struct DLTensor* iou, void* stream);
int rod_dl_target_fused(const rod_layout_t* layout, const struct DLTensor* anchors_corner,
                        const struct DLTensor* anchors_center, const float* arm_thresholds,
                        const float* odm_thresholds, const struct DLTensor* center_bboxes,
                        const struct DLTensor* labels, const struct DLTensor* gt_counts,
                        const struct DLTensor* const* refine_out, const struct DLTensor* gt,
                        const struct DLTensor* cbboxes, const struct DLTensor* out_labels,
                        const struct DLTensor* pos_mask, const struct DLTensor* match_idx,
                        const struct DLTensor* det_gt, const struct DLTensor* det_mask,
                        const struct DLTensor* det_labels, const struct DLTensor* iou,
                        const struct DLTensor* workspace, void* stream);
int rod_dl_decode(const rod_layout_t* layout, const struct DLTensor* anchors_center,
                  const struct DLTensor* const* refine_out, const struct DLTensor* const* det_out,
                  int to_corner, const struct DLTensor* out, void* stream);
int rod_dl_detect(const rod_layout_t* layout, const struct DLTensor* anchors_center,
                  const struct DLTensor* const* predictions,
                  const struct DLTensor* const* localizations,
                  const struct DLTensor* const* refine_out, const struct DLTensor* const* det_out,
                  int ignore_class, float select_threshold, float nms_threshold, int top_k,
                  int keep_top_k, const struct DLTensor* clip_box, const struct DLTensor* out_scores,
                  const struct DLTensor* out_bboxes, const struct DLTensor* out_counts,
                  const struct DLTensor* workspace, void* stream);

/* rod_dl_detect with class logits (see rod_detect_logits); rod_dl_softmax: out may alias logits. */
int rod_dl_detect_logits(const rod_layout_t* layout, const struct DLTensor* anchors_center,
                         const struct DLTensor* const* logits,
                         const struct DLTensor* const* localizations,
                         const struct DLTensor* const* refine_out, const struct DLTensor* const* det_out,
                         int ignore_class, float select_threshold, float nms_threshold, int top_k,
                         int keep_top_k, const struct DLTensor* clip_box, const struct DLTensor* out_scores,
                         const struct DLTensor* out_bboxes, const struct DLTensor* out_counts,
                         const struct DLTensor* workspace, void* stream);
int rod_dl_softmax(const struct DLTensor* logits, const struct DLTensor* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RODET_B200_H_ */
