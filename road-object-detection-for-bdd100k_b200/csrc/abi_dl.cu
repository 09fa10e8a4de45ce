// DLPack front door of the C ABI: validates device / dtype / shape / contiguity of borrowed
// DLTensors and forwards to the raw-pointer entry points.  Never calls a deleter.
#include "common.cuh"
#include "../../include/rodet_dlpack.h"

namespace rod {

#define DL_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      rod::set_error(__VA_ARGS__); \
      return ROD_E_DLPACK;         \
    }                              \
  } while (0)

static inline void* dl_ptr(const DLTensor* t) { return static_cast<char*>(t->data) + t->byte_offset; }

static int64_t dl_numel(const DLTensor* t, int from = 0) {
  int64_t n = 1;
  for (int i = from; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}

// dims [from, ndim) form a compact row-major block
static bool dl_compact(const DLTensor* t, int from) {
  if (!t->strides) return true;
  int64_t expect = 1;
  for (int i = t->ndim - 1; i >= from; --i) {
    if (t->shape[i] != 1 && t->strides[i] != expect) return false;
    expect *= t->shape[i];
  }
  return true;
}

static int dl_check(const DLTensor* t, const char* name, int code, int bits) {
  DL_REQUIRE(t != nullptr, "%s: DLTensor is NULL", name);
  DL_REQUIRE(t->device.device_type == ROD_kDLCUDA || t->device.device_type == ROD_kDLCUDAManaged,
             "%s: tensor is not on a CUDA device (device_type=%d); there is no CPU path", name, t->device.device_type);
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess)
    DL_REQUIRE(t->device.device_id == dev, "%s: tensor lives on cuda:%d but the current device is cuda:%d", name,
               t->device.device_id, dev);
  DL_REQUIRE(t->dtype.code == code && t->dtype.bits == bits && t->dtype.lanes == 1,
             "%s: dtype (code=%d,bits=%d,lanes=%d) != expected (code=%d,bits=%d)", name, t->dtype.code, t->dtype.bits,
             t->dtype.lanes, code, bits);
  DL_REQUIRE(t->data != nullptr || dl_numel(t) == 0, "%s: data pointer is NULL", name);
  return ROD_OK;
}

static int dl_flat(const DLTensor* t, const char* name, int code, int bits, int64_t numel) {
  int rc = dl_check(t, name, code, bits);
  if (rc) return rc;
  DL_REQUIRE(dl_compact(t, 0), "%s: tensor must be contiguous", name);
  DL_REQUIRE(dl_numel(t) == numel, "%s: has %lld elements, expected %lld", name, (long long)dl_numel(t), (long long)numel);
  return ROD_OK;
}

// list of per-layer tensors [B, ..., inner] -> rod_layered_t
static int dl_layered(const DLTensor* const* ts, const rod_layout_t* L, int inner, int code, int bits, int* batch,
                      const char* name, rod_layered_t* out) {
  DL_REQUIRE(ts != nullptr, "%s: list is NULL", name);
  for (int l = 0; l < L->n_layers; ++l) {
    const DLTensor* t = ts[l];
    char nm[96];
    snprintf(nm, sizeof(nm), "%s[%d]", name, l);
    int rc = dl_check(t, nm, code, bits);
    if (rc) return rc;
    DL_REQUIRE(t->ndim >= 2, "%s: needs at least 2 dims [B, ...]", nm);
    const int64_t per = (int64_t)(L->offset[l + 1] - L->offset[l]) * inner;
    DL_REQUIRE(dl_numel(t, 1) == per, "%s: %lld elements per image, expected %lld (= anchors of the layer x %d)", nm,
               (long long)dl_numel(t, 1), (long long)per, inner);
    DL_REQUIRE(dl_compact(t, 1), "%s: dims after the batch dim must be contiguous", nm);
    if (*batch < 0) *batch = (int)t->shape[0];
    DL_REQUIRE(t->shape[0] == *batch, "%s: batch %lld != %d", nm, (long long)t->shape[0], *batch);
    out->base[l] = dl_ptr(t);
    out->batch_stride[l] = (t->strides && t->shape[0] > 1) ? t->strides[0] : per;
    if (inner == 4)
      DL_REQUIRE(((uintptr_t)out->base[l] & 15u) == 0 && (out->batch_stride[l] % 4) == 0,
                 "%s: box tensors must be 16-byte aligned", nm);
  }
  return ROD_OK;
}

}  // namespace rod

using namespace rod;

extern "C" int rod_dl_layered(const rod_layout_t* layout, const DLTensor* const* tensors, int inner, int dtype_code,
                              int dtype_bits, rod_layered_t* out, int* batch) {
  int rc = check_layout(layout);
  if (rc) return rc;
  ROD_REQUIRE(out != nullptr && batch != nullptr && inner >= 1, "rod_dl_layered: bad arguments");
  return dl_layered(tensors, layout, inner, dtype_code, dtype_bits, batch, "tensors", out);
}

extern "C" int rod_dl_arm_match_encode(const rod_layout_t* layout, const DLTensor* anchors_corner,
                                       const DLTensor* anchors_center, const float* thresholds,
                                       const DLTensor* center_bboxes, const DLTensor* labels,
                                       const DLTensor* gt_counts, int method, const DLTensor* gt,
                                       const DLTensor* cbboxes, const DLTensor* out_labels,
                                       const DLTensor* pos_mask, const DLTensor* match_idx, void* stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int64_t N = layout->n_total;
  if ((rc = dl_flat(anchors_corner, "anchors_corner", ROD_kDLFloat, 32, N * 4))) return rc;
  if ((rc = dl_flat(anchors_center, "anchors_center", ROD_kDLFloat, 32, N * 4))) return rc;
  if ((rc = dl_check(center_bboxes, "center_bboxes", ROD_kDLFloat, 32))) return rc;
  DL_REQUIRE(center_bboxes->ndim == 3 && center_bboxes->shape[2] == 4 && dl_compact(center_bboxes, 0),
             "center_bboxes: expected contiguous [B, G, 4]");
  const int B = (int)center_bboxes->shape[0], G = (int)center_bboxes->shape[1];
  DL_REQUIRE(labels != nullptr, "labels: DLTensor is NULL");
  const int i64 = labels->dtype.bits == 64;
  if ((rc = dl_flat(labels, "labels", ROD_kDLInt, i64 ? 64 : 32, (int64_t)B * G))) return rc;
  if (gt_counts && (rc = dl_flat(gt_counts, "gt_counts", ROD_kDLInt, 32, B))) return rc;
  if ((rc = dl_flat(gt, "gt", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if ((rc = dl_flat(cbboxes, "cbboxes", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if ((rc = dl_flat(out_labels, "out_labels", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(pos_mask, "pos_mask", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if (match_idx && (rc = dl_flat(match_idx, "match_idx", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  return rod_arm_match_encode(layout, (const float*)dl_ptr(anchors_corner), (const float*)dl_ptr(anchors_center),
                              thresholds, (const float*)dl_ptr(center_bboxes), dl_ptr(labels), i64,
                              gt_counts ? (const int32_t*)dl_ptr(gt_counts) : nullptr, B, G, method,
                              (float*)dl_ptr(gt), (float*)dl_ptr(cbboxes), (int32_t*)dl_ptr(out_labels),
                              (int32_t*)dl_ptr(pos_mask), match_idx ? (int32_t*)dl_ptr(match_idx) : nullptr, stream);
}

extern "C" int rod_dl_odm_target(const rod_layout_t* layout, const DLTensor* anchors_center,
                                 const float* thresholds, const DLTensor* const* refine_out,
                                 const DLTensor* const* offset_gt, const DLTensor* const* cbboxes,
                                 const DLTensor* const* refine_labels, const DLTensor* const* refine_pos_mask,
                                 const DLTensor* det_gt, const DLTensor* mask, const DLTensor* det_labels,
                                 const DLTensor* iou, void* stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int64_t N = layout->n_total;
  if ((rc = dl_flat(anchors_center, "anchors_center", ROD_kDLFloat, 32, N * 4))) return rc;
  int B = -1;
  rod_layered_t ro, og, cb, lb, pm;
  if ((rc = dl_layered(refine_out, layout, 4, ROD_kDLFloat, 32, &B, "refine_out", &ro))) return rc;
  if ((rc = dl_layered(offset_gt, layout, 4, ROD_kDLFloat, 32, &B, "offset_gt", &og))) return rc;
  if ((rc = dl_layered(cbboxes, layout, 4, ROD_kDLFloat, 32, &B, "cbboxes", &cb))) return rc;
  if ((rc = dl_layered(refine_labels, layout, 1, ROD_kDLInt, 32, &B, "refine_labels", &lb))) return rc;
  if ((rc = dl_layered(refine_pos_mask, layout, 1, ROD_kDLInt, 32, &B, "refine_pos_mask", &pm))) return rc;
  if ((rc = dl_flat(det_gt, "det_gt", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if ((rc = dl_flat(mask, "mask", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(det_labels, "det_labels", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(iou, "iou", ROD_kDLFloat, 32, (int64_t)B * N))) return rc;
  return rod_odm_target(layout, (const float*)dl_ptr(anchors_center), thresholds, &ro, &og, &cb, &lb, &pm, B,
                        (float*)dl_ptr(det_gt), (int32_t*)dl_ptr(mask), (int32_t*)dl_ptr(det_labels),
                        (float*)dl_ptr(iou), stream);
}

extern "C" int rod_dl_target_fused(const rod_layout_t* layout, const DLTensor* anchors_corner,
                                   const DLTensor* anchors_center, const float* arm_thresholds,
                                   const float* odm_thresholds, const DLTensor* center_bboxes, const DLTensor* labels,
                                   const DLTensor* gt_counts, const DLTensor* const* refine_out, const DLTensor* gt,
                                   const DLTensor* cbboxes, const DLTensor* out_labels, const DLTensor* pos_mask,
                                   const DLTensor* match_idx, const DLTensor* det_gt, const DLTensor* det_mask,
                                   const DLTensor* det_labels, const DLTensor* iou, const DLTensor* workspace,
                                   void* stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int64_t N = layout->n_total;
  if ((rc = dl_flat(anchors_corner, "anchors_corner", ROD_kDLFloat, 32, N * 4))) return rc;
  if ((rc = dl_flat(anchors_center, "anchors_center", ROD_kDLFloat, 32, N * 4))) return rc;
  if ((rc = dl_check(center_bboxes, "center_bboxes", ROD_kDLFloat, 32))) return rc;
  DL_REQUIRE(center_bboxes->ndim == 3 && center_bboxes->shape[2] == 4 && dl_compact(center_bboxes, 0),
             "center_bboxes: expected contiguous [B, G, 4]");
  int B = (int)center_bboxes->shape[0];
  const int G = (int)center_bboxes->shape[1];
  DL_REQUIRE(labels != nullptr, "labels: DLTensor is NULL");
  const int i64 = labels->dtype.bits == 64;
  if ((rc = dl_flat(labels, "labels", ROD_kDLInt, i64 ? 64 : 32, (int64_t)B * G))) return rc;
  if (gt_counts && (rc = dl_flat(gt_counts, "gt_counts", ROD_kDLInt, 32, B))) return rc;
  rod_layered_t ro;
  if ((rc = dl_layered(refine_out, layout, 4, ROD_kDLFloat, 32, &B, "refine_out", &ro))) return rc;
  if ((rc = dl_flat(gt, "gt", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if (cbboxes && (rc = dl_flat(cbboxes, "cbboxes", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if (out_labels && (rc = dl_flat(out_labels, "out_labels", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(pos_mask, "pos_mask", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if (match_idx && (rc = dl_flat(match_idx, "match_idx", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(det_gt, "det_gt", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  if ((rc = dl_flat(det_mask, "det_mask", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(det_labels, "det_labels", ROD_kDLInt, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(iou, "iou", ROD_kDLFloat, 32, (int64_t)B * N))) return rc;
  if ((rc = dl_flat(workspace, "workspace", ROD_kDLInt, 32, 2))) return rc;
  return rod_target_fused(layout, (const float*)dl_ptr(anchors_corner), (const float*)dl_ptr(anchors_center),
                          arm_thresholds, odm_thresholds, (const float*)dl_ptr(center_bboxes), dl_ptr(labels), i64,
                          gt_counts ? (const int32_t*)dl_ptr(gt_counts) : nullptr, B, G, &ro, (float*)dl_ptr(gt),
                          cbboxes ? (float*)dl_ptr(cbboxes) : nullptr, out_labels ? (int32_t*)dl_ptr(out_labels) : nullptr,
                          (int32_t*)dl_ptr(pos_mask), match_idx ? (int32_t*)dl_ptr(match_idx) : nullptr,
                          (float*)dl_ptr(det_gt), (int32_t*)dl_ptr(det_mask), (int32_t*)dl_ptr(det_labels),
                          (float*)dl_ptr(iou), dl_ptr(workspace), stream);
}

extern "C" int rod_dl_decode(const rod_layout_t* layout, const DLTensor* anchors_center,
                             const DLTensor* const* refine_out, const DLTensor* const* det_out, int to_corner,
                             const DLTensor* out, void* stream) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int64_t N = layout->n_total;
  if ((rc = dl_flat(anchors_center, "anchors_center", ROD_kDLFloat, 32, N * 4))) return rc;
  int B = -1;
  rod_layered_t ro, dt;
  if ((rc = dl_layered(refine_out, layout, 4, ROD_kDLFloat, 32, &B, "refine_out", &ro))) return rc;
  if (det_out && (rc = dl_layered(det_out, layout, 4, ROD_kDLFloat, 32, &B, "det_out", &dt))) return rc;
  if ((rc = dl_flat(out, "out", ROD_kDLFloat, 32, (int64_t)B * N * 4))) return rc;
  return rod_decode(layout, (const float*)dl_ptr(anchors_center), &ro, det_out ? &dt : nullptr, B, to_corner,
                    (float*)dl_ptr(out), stream);
}

static int dl_detect_impl(const rod_layout_t* layout, const DLTensor* anchors_center,
                             const DLTensor* const* predictions, const DLTensor* const* localizations,
                             const DLTensor* const* refine_out, const DLTensor* const* det_out, int ignore_class,
                             float select_threshold, float nms_threshold, int top_k, int keep_top_k,
                             const DLTensor* clip_box, const DLTensor* out_scores, const DLTensor* out_bboxes,
                             const DLTensor* out_counts, const DLTensor* workspace, void* stream, int logits) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int64_t N = layout->n_total;
  DL_REQUIRE(predictions && predictions[0], "predictions: list is NULL");
  DL_REQUIRE(predictions[0]->ndim >= 2, "predictions[0]: needs at least 2 dims");
  const int C = (int)predictions[0]->shape[predictions[0]->ndim - 1];
  DL_REQUIRE(C >= 1 && C <= ROD_MAX_CLASSES, "predictions: last dim (classes) = %d not in [1,%d]", C, ROD_MAX_CLASSES);
  int B = -1;
  rod_layered_t pr, lo, ro, dt;
  if ((rc = dl_layered(predictions, layout, C, ROD_kDLFloat, 32, &B, "predictions", &pr))) return rc;
  if (localizations) {
    if ((rc = dl_layered(localizations, layout, 4, ROD_kDLFloat, 32, &B, "localizations", &lo))) return rc;
  } else {
    DL_REQUIRE(refine_out && det_out && anchors_center, "need localizations, or refine_out + det_out + anchors_center");
    if ((rc = dl_flat(anchors_center, "anchors_center", ROD_kDLFloat, 32, N * 4))) return rc;
    if ((rc = dl_layered(refine_out, layout, 4, ROD_kDLFloat, 32, &B, "refine_out", &ro))) return rc;
    if ((rc = dl_layered(det_out, layout, 4, ROD_kDLFloat, 32, &B, "det_out", &dt))) return rc;
  }
  if (clip_box && (rc = dl_flat(clip_box, "clip_box", ROD_kDLFloat, 32, 4))) return rc;
  if ((rc = dl_flat(out_scores, "out_scores", ROD_kDLFloat, 32, (int64_t)C * B * keep_top_k))) return rc;
  if ((rc = dl_flat(out_bboxes, "out_bboxes", ROD_kDLFloat, 32, (int64_t)C * B * keep_top_k * 4))) return rc;
  if (out_counts && (rc = dl_flat(out_counts, "out_counts", ROD_kDLInt, 32, (int64_t)C * B))) return rc;
  if ((rc = dl_check(workspace, "workspace", ROD_kDLUInt, 8))) return rc;
  DL_REQUIRE(dl_compact(workspace, 0), "workspace: must be contiguous");
  return (logits ? rod_detect_logits : rod_detect)(layout, anchors_center ? (const float*)dl_ptr(anchors_center) : nullptr, &pr,
                    localizations ? &lo : nullptr, localizations ? nullptr : &ro, localizations ? nullptr : &dt, B, C,
                    ignore_class, select_threshold, nms_threshold, top_k, keep_top_k,
                    clip_box ? (const float*)dl_ptr(clip_box) : nullptr, (float*)dl_ptr(out_scores),
                    (float*)dl_ptr(out_bboxes), out_counts ? (int32_t*)dl_ptr(out_counts) : nullptr, dl_ptr(workspace),
                    (size_t)dl_numel(workspace), stream);
}

extern "C" int rod_dl_detect(const rod_layout_t* layout, const DLTensor* anchors_center,
                             const DLTensor* const* predictions, const DLTensor* const* localizations,
                             const DLTensor* const* refine_out, const DLTensor* const* det_out, int ignore_class,
                             float select_threshold, float nms_threshold, int top_k, int keep_top_k,
                             const DLTensor* clip_box, const DLTensor* out_scores, const DLTensor* out_bboxes,
                             const DLTensor* out_counts, const DLTensor* workspace, void* stream) {
  return dl_detect_impl(layout, anchors_center, predictions, localizations, refine_out, det_out, ignore_class,
                        select_threshold, nms_threshold, top_k, keep_top_k, clip_box, out_scores, out_bboxes, out_counts,
                        workspace, stream, 0);
}

extern "C" int rod_dl_detect_logits(const rod_layout_t* layout, const DLTensor* anchors_center,
                                    const DLTensor* const* logits, const DLTensor* const* localizations,
                                    const DLTensor* const* refine_out, const DLTensor* const* det_out, int ignore_class,
                                    float select_threshold, float nms_threshold, int top_k, int keep_top_k,
                                    const DLTensor* clip_box, const DLTensor* out_scores, const DLTensor* out_bboxes,
                                    const DLTensor* out_counts, const DLTensor* workspace, void* stream) {
  return dl_detect_impl(layout, anchors_center, logits, localizations, refine_out, det_out, ignore_class,
                        select_threshold, nms_threshold, top_k, keep_top_k, clip_box, out_scores, out_bboxes, out_counts,
                        workspace, stream, 1);
}

extern "C" int rod_dl_softmax(const DLTensor* logits, const DLTensor* out, void* stream) {
  int rc;
  if ((rc = dl_check(logits, "logits", ROD_kDLFloat, 32))) return rc;
  if ((rc = dl_check(out, "out", ROD_kDLFloat, 32))) return rc;
  DL_REQUIRE(logits->ndim >= 1, "logits: needs at least 1 dim");
  DL_REQUIRE(dl_compact(logits, 0) && dl_compact(out, 0), "logits / out: must be contiguous");
  const int64_t C = logits->shape[logits->ndim - 1], total = dl_numel(logits);
  DL_REQUIRE(dl_numel(out) == total, "out: %lld elements, logits has %lld", (long long)dl_numel(out), (long long)total);
  DL_REQUIRE(C >= 1 && C <= ROD_MAX_CLASSES, "logits: last dim (classes) = %lld not in [1,%d]", (long long)C, ROD_MAX_CLASSES);
  return rod_softmax((const float*)dl_ptr(logits), C ? total / C : 0, (int)C, (float*)dl_ptr(out), stream);
}
