// Opt-in extra WITHOUT a reference counterpart: the "forced match" of BASELINE.json's north star.
// The reference's JACCARD_BIGGER only thresholds max_G IoU per anchor (utils/net_tools.py:405-408): a ground-truth box
// no anchor reaches the layer threshold for gets no positive at all.  This post-pass of rod_arm_match_encode adds the
// SSD / RefineDet bipartite step: every GT box also claims the anchor it overlaps best (argmax over the ANCHOR axis,
// lowest anchor index on ties, IoU > 0), whatever the threshold; if several GT boxes claim one anchor, the highest IoU
// wins (ties: lowest GT index).  Same IoU arithmetic as the matching kernel (net_tools.jaccard op order).  Never the default.
#include "common.cuh"

namespace rod {

constexpr int kFmBlock = 256;

// one CTA per (GT box, image): max over all anchors of (IoU, lowest anchor index)
__global__ void __launch_bounds__(kFmBlock)
gt_best_anchor_kernel(int n_anchors, const float* __restrict__ corner, const float* __restrict__ gtb,
                      const int32_t* __restrict__ counts, int gmax, unsigned long long* __restrict__ best) {
  __shared__ unsigned long long s_w[kFmBlock / 32];
  const int g = blockIdx.x, b = blockIdx.y;
  int count = counts ? counts[b] : gmax;
  count = min(max(count, 0), gmax);
  unsigned long long key = 0ull;                              // (IoU bits, ~anchor): larger = better; 0 = no overlap at all
  if (g < count) {
    const float4 gc = center_to_corner(ldg4(gtb + 4ll * ((long long)b * gmax + g)));
    const float area_g = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
    if ((gc.z > gc.x) && (gc.w > gc.y)) {
      for (int n = threadIdx.x; n < n_anchors; n += kFmBlock) {
        const float4 a = ldg4(corner + 4ll * n);
        const float h = __fsub_rn(fminf(a.z, gc.z), fmaxf(a.x, gc.x)), w = __fsub_rn(fminf(a.w, gc.w), fmaxf(a.y, gc.y));
        if (h > 0.f && w > 0.f) {
          const float inter = __fmul_rn(h, w);
          const float jac = __fdiv_rn(inter, __fadd_rn(__fsub_rn(box_vol(a), inter), area_g));
          if (jac > 0.f) {
            const unsigned long long k = ((unsigned long long)__float_as_uint(jac) << 32) | (unsigned)(~(unsigned)n);
            key = k > key ? k : key;
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, key, o);
    key = x > key ? x : key;
  }
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kFmBlock / 32; ++w) key = s_w[w] > key ? s_w[w] : key;
    best[(long long)b * gmax + g] = key;
  }
}

// one CTA per image: conflicts between GT boxes that claim the same anchor, then the winners rewrite that anchor's row
template <typename LabelT>
__global__ void __launch_bounds__(128)
forced_apply_kernel(int n_anchors, const float* __restrict__ center, const float* __restrict__ gtb,
                    const LabelT* __restrict__ labels, const int32_t* __restrict__ counts, int gmax,
                    const unsigned long long* __restrict__ best, float* __restrict__ out_gt, float* __restrict__ out_cb,
                    int32_t* __restrict__ out_lab, int32_t* __restrict__ out_pos, int32_t* __restrict__ out_idx) {
  extern __shared__ unsigned long long s_key[];
  const int b = blockIdx.x;
  int count = counts ? counts[b] : gmax;
  count = min(max(count, 0), gmax);
  for (int g = threadIdx.x; g < count; g += blockDim.x) s_key[g] = best[(long long)b * gmax + g];
  __syncthreads();
  for (int g = threadIdx.x; g < count; g += blockDim.x) {
    const unsigned long long k = s_key[g];
    if (k == 0ull) continue;
    const unsigned anchor = ~(unsigned)(k & 0xffffffffull), iou = (unsigned)(k >> 32);
    bool wins = true;
    for (int o = 0; o < count && wins; ++o) {
      if (o == g || s_key[o] == 0ull) continue;
      const unsigned a2 = ~(unsigned)(s_key[o] & 0xffffffffull), i2 = (unsigned)(s_key[o] >> 32);
      if (a2 == anchor && (i2 > iou || (i2 == iou && o < g))) wins = false;
    }
    if (!wins) continue;
    const float4 gcen = ldg4(gtb + 4ll * ((long long)b * gmax + g));
    const float4 e = encode_center(ldg4(center + 4ll * anchor), gcen);
    const long long o = (long long)b * n_anchors + anchor;
    st4(out_gt + 4 * o, make_float4(__fadd_rn(e.x, 0.f), __fadd_rn(e.y, 0.f), __fadd_rn(e.z, 0.f), __fadd_rn(e.w, 0.f)));
    st4(out_cb + 4 * o, make_float4(__fadd_rn(gcen.x, 0.f), __fadd_rn(gcen.y, 0.f), __fadd_rn(gcen.z, 0.f), __fadd_rn(gcen.w, 0.f)));
    out_lab[o] = (int)labels[(long long)b * gmax + g];
    out_pos[o] = 1;
    if (out_idx) out_idx[o] = g;
  }
}

}  // namespace rod

extern "C" size_t rod_arm_forced_match_workspace_bytes(int batch, int gmax) {
  return batch <= 0 || gmax <= 0 ? 256 : (size_t)batch * gmax * 8 + 256;
}

extern "C" int rod_arm_forced_match(const rod_layout_t* layout, const float* anchors_corner, const float* anchors_center,
                                    const float* center_bboxes, const void* labels, int labels_i64, const int32_t* gt_counts,
                                    int batch, int gmax, float* gt, float* cbboxes, int32_t* out_labels, int32_t* pos_mask,
                                    int32_t* match_idx, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  ROD_REQUIRE(anchors_corner && anchors_center && center_bboxes && labels, "rod_arm_forced_match: NULL input pointer");
  ROD_REQUIRE(gt && cbboxes && out_labels && pos_mask, "rod_arm_forced_match: NULL output pointer");
  ROD_REQUIRE(batch >= 0 && batch <= 65535 && gmax >= 1 && gmax <= 65535, "rod_arm_forced_match: batch=%d gmax=%d invalid", batch, gmax);
  ROD_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0 &&
                  workspace_bytes >= rod_arm_forced_match_workspace_bytes(batch, gmax),
              "rod_arm_forced_match: workspace NULL, misaligned or too small");
  const size_t smem = (size_t)gmax * 8;
  ROD_REQUIRE(smem <= 48 * 1024, "rod_arm_forced_match: gmax=%d too large (at most 6144 boxes per image)", gmax);
  if (batch == 0) return ROD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* best = static_cast<unsigned long long*>(workspace);
  gt_best_anchor_kernel<<<dim3(gmax, batch), kFmBlock, 0, st>>>(layout->n_total, anchors_corner, center_bboxes, gt_counts, gmax, best);
  ROD_LAUNCH_CHECK("gt_best_anchor_kernel");
  if (labels_i64)
    forced_apply_kernel<long long><<<batch, 128, smem, st>>>(layout->n_total, anchors_center, center_bboxes, (const long long*)labels,
                                                             gt_counts, gmax, best, gt, cbboxes, out_labels, pos_mask, match_idx);
  else
    forced_apply_kernel<int><<<batch, 128, smem, st>>>(layout->n_total, anchors_center, center_bboxes, (const int*)labels, gt_counts,
                                                       gmax, best, gt, cbboxes, out_labels, pos_mask, match_idx);
  ROD_LAUNCH_CHECK("forced_apply_kernel");
  return ROD_OK;
}
