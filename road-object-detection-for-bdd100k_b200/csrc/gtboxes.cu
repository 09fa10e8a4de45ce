// f-4: the ground-truth-box half of the training input pipeline, the step right before ARM matching
// (utils/data_pileline_tools.py:88-108):
//   tfe.bboxes_resize(distort_bbox, bboxes)                   utils/tf_extended/bboxes.py:139-163
//   tfe.bboxes_filter_overlap(labels, bboxes, 0.3, False)     :408-428 (uses bboxes_intersection :480-504)
//   flip_bboxes when the image is mirrored                    utils/augmentation/tf_image.py:284-289
//   tf.maximum(bboxes, 0.), tf.minimum(bboxes, 1.)            utils/data_pileline_tools.py:107-108
// fused and batched over images: one warp per image, order-preserving compaction with ballots.
#include "common.cuh"

namespace rod {

constexpr int kGtWarps = 4;

template <typename LabelT>
__global__ void __launch_bounds__(32 * kGtWarps)
gt_boxes_update_kernel(const float* __restrict__ bboxes, const LabelT* __restrict__ labels, const int32_t* __restrict__ counts,
                       int batch, int gmax, const float* __restrict__ crop, const uint8_t* __restrict__ mirror,
                       int filter, float threshold, int assign_negative, int clamp01, float* __restrict__ out_bboxes,
                       LabelT* __restrict__ out_labels, int32_t* __restrict__ out_counts) {
  const int b = blockIdx.x * kGtWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= batch) return;
  const int n = counts ? min(max(counts[b], 0), gmax) : gmax;
  float4 c = make_float4(0.f, 0.f, 1.f, 1.f);
  if (crop) c = ldg4(crop + 4ll * b);
  const float sh = __fsub_rn(c.z, c.x), sw = __fsub_rn(c.w, c.y);
  const bool flip = mirror != nullptr && mirror[b] != 0;
  const float* src = bboxes + 4ll * b * gmax;
  float* dst = out_bboxes + 4ll * b * gmax;
  const LabelT* lsrc = labels + (long long)b * gmax;
  LabelT* ldst = out_labels + (long long)b * gmax;
  int base = 0;
  for (int g0 = 0; g0 < n; g0 += 32) {
    const int g = g0 + lane;
    bool keep = false;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    LabelT lab = 0;
    if (g < n) {
      v = ldg4(src + 4ll * g);
      lab = lsrc[g];
      if (crop)                                        // translate, then scale (two roundings, as the reference)
        v = make_float4(__fdiv_rn(__fsub_rn(v.x, c.x), sh), __fdiv_rn(__fsub_rn(v.y, c.y), sw),
                        __fdiv_rn(__fsub_rn(v.z, c.x), sh), __fdiv_rn(__fsub_rn(v.w, c.y), sw));
      keep = true;
      if (filter) {
        // bboxes_intersection([0,0,1,1], v): clipped area over box area, 0 where the box area <= 0
        const float h = fmaxf(__fsub_rn(fminf(v.z, 1.f), fmaxf(v.x, 0.f)), 0.f);
        const float w = fmaxf(__fsub_rn(fminf(v.w, 1.f), fmaxf(v.y, 0.f)), 0.f);
        const float vol = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
        const float score = vol > 0.f ? __fdiv_rn(__fmul_rn(h, w), vol) : 0.f;
        const bool in = score > threshold;
        if (assign_negative) lab = in ? lab : (LabelT)(-lab);
        else keep = in;
      }
      if (flip) v = make_float4(v.x, __fsub_rn(1.f, v.w), v.z, __fsub_rn(1.f, v.y));
      if (clamp01) v = make_float4(fminf(fmaxf(v.x, 0.f), 1.f), fminf(fmaxf(v.y, 0.f), 1.f),
                                   fminf(fmaxf(v.z, 0.f), 1.f), fminf(fmaxf(v.w, 0.f), 1.f));
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      st4(dst + 4ll * pos, v);
      ldst[pos] = lab;
    }
    base += __popc(m);
  }
  for (int g = base + lane; g < gmax; g += 32) {          // zero padding behind the kept boxes
    st4(dst + 4ll * g, make_float4(0.f, 0.f, 0.f, 0.f));
    ldst[g] = 0;
  }
  if (lane == 0 && out_counts) out_counts[b] = base;
}

}  // namespace rod

extern "C" int rod_gt_boxes_update(const float* bboxes, const void* labels, int labels_i64, const int32_t* counts, int batch,
                                   int gmax, const float* distort_bbox, const uint8_t* mirror, int filter_overlap,
                                   float threshold, int assign_negative, int clamp01, float* out_bboxes, void* out_labels,
                                   int32_t* out_counts, void* stream) {
  using namespace rod;
  ROD_REQUIRE(batch >= 0 && gmax >= 0, "rod_gt_boxes_update: batch=%d gmax=%d invalid", batch, gmax);
  if (batch == 0) return ROD_OK;
  ROD_REQUIRE(gmax == 0 || (bboxes && labels && out_bboxes && out_labels), "rod_gt_boxes_update: NULL pointer argument");
  ROD_REQUIRE(gmax == 0 || bboxes != out_bboxes || !filter_overlap || assign_negative, "rod_gt_boxes_update: in-place only without compaction");
  const unsigned grid = (unsigned)((batch + kGtWarps - 1) / kGtWarps);
  cudaStream_t st = (cudaStream_t)stream;
  if (labels_i64)
    gt_boxes_update_kernel<long long><<<grid, 32 * kGtWarps, 0, st>>>(
        bboxes, (const long long*)labels, counts, batch, gmax, distort_bbox, mirror, filter_overlap, threshold, assign_negative,
        clamp01, out_bboxes, (long long*)out_labels, out_counts);
  else
    gt_boxes_update_kernel<int32_t><<<grid, 32 * kGtWarps, 0, st>>>(
        bboxes, (const int32_t*)labels, counts, batch, gmax, distort_bbox, mirror, filter_overlap, threshold, assign_negative,
        clamp01, out_bboxes, (int32_t*)out_labels, out_counts);
  ROD_LAUNCH_CHECK("gt_boxes_update_kernel");
  return ROD_OK;
}
