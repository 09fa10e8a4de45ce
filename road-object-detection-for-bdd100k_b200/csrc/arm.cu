// a9: ARM anchor<->ground-truth matching + box encoding, batched over images.
// Replaces refine_groundtruth (utils/net_tools.py:270-428).
//
// JACCARD_BIGGER (:382-421): per anchor, max / first-argmax over the image's GT boxes of
// net_tools.jaccard, thresholded per layer; the matched GT's centre box, its encoding and
// its label are emitted for positive anchors, zeros elsewhere.
//
// Kernel shape: grid = (anchor tiles, images); a CTA owns BLOCK consecutive anchors of one
// image.  Only ~6 % of (anchor, GT) pairs intersect on street-scene data, and a pair that
// does not intersect contributes IoU == 0 exactly, which can never win the strict-'>' argmax
// that starts at (0, index 0).  So each CTA first culls the image's GT list against the
// bounding box of its anchors (ordered compaction => ascending GT index => lowest-index
// tie-break is preserved) into shared memory, and the per-anchor loop walks survivors only.
// This turns the FP32-bound N x G loop into an HBM-write-bound kernel (40 B / anchor).
#include "common.cuh"

namespace rod {

constexpr int kArmBlock = 256;

template <typename LabelT>
__global__ void __launch_bounds__(kArmBlock)
arm_jaccard_bigger_kernel(const __grid_constant__ Layout L, const __grid_constant__ Thresholds T, const float* __restrict__ corner,
                          const float* __restrict__ center, const float* __restrict__ gtb,
                          const LabelT* __restrict__ labels, const int32_t* __restrict__ counts,
                          int gmax, float* __restrict__ out_gt, float* __restrict__ out_cb,
                          int32_t* __restrict__ out_lab, int32_t* __restrict__ out_pos,
                          int32_t* __restrict__ out_idx) {
  extern __shared__ float4 s_box[];                       // [gmax] surviving GT corners
  float* s_area = reinterpret_cast<float*>(s_box + gmax);  // [gmax]
  int* s_id = reinterpret_cast<int*>(s_area + gmax);       // [gmax] original GT index
  __shared__ float s_red[4][kArmBlock / 32];
  __shared__ int s_wcount[kArmBlock / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int N = L.n_total;
  const int n = blockIdx.x * kArmBlock + tid;
  const bool valid = n < N;

  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) a = ldg4(corner + 4ll * n);

  // ---- bounding box of this CTA's anchors
  float r0 = valid ? a.x : INFINITY, r1 = valid ? a.y : INFINITY;
  float r2 = valid ? a.z : -INFINITY, r3 = valid ? a.w : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r0 = fminf(r0, __shfl_xor_sync(0xffffffffu, r0, o));
    r1 = fminf(r1, __shfl_xor_sync(0xffffffffu, r1, o));
    r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    r3 = fmaxf(r3, __shfl_xor_sync(0xffffffffu, r3, o));
  }
  if (lane == 0) { s_red[0][warp] = r0; s_red[1][warp] = r1; s_red[2][warp] = r2; s_red[3][warp] = r3; }
  __syncthreads();
  float t_ymin = s_red[0][0], t_xmin = s_red[1][0], t_ymax = s_red[2][0], t_xmax = s_red[3][0];
#pragma unroll
  for (int w = 1; w < kArmBlock / 32; ++w) {
    t_ymin = fminf(t_ymin, s_red[0][w]); t_xmin = fminf(t_xmin, s_red[1][w]);
    t_ymax = fmaxf(t_ymax, s_red[2][w]); t_xmax = fmaxf(t_xmax, s_red[3][w]);
  }

  // ---- cull + ordered compaction of the GT list
  int count = counts ? counts[b] : gmax;
  count = min(max(count, 0), gmax);
  const float* gt_img = gtb + 4ll * b * gmax;
  int total = 0;
  for (int base = 0; base < count; base += kArmBlock) {
    const int g = base + tid;
    bool hit = false;
    float4 gc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < count) {
      gc = center_to_corner(ldg4(gt_img + 4ll * g));   // net_tools.py:323 c2c(center_bboxes[i])
      hit = (gc.z > t_ymin) && (gc.x < t_ymax) && (gc.w > t_xmin) && (gc.y < t_xmax);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    __syncthreads();                                   // s_wcount reuse across iterations
    if (lane == 0) s_wcount[warp] = __popc(m);
    __syncthreads();
    int pre = total, all = 0;
#pragma unroll
    for (int w = 0; w < kArmBlock / 32; ++w) {
      const int c = s_wcount[w];
      pre += (w < warp) ? c : 0;
      all += c;
    }
    if (hit) {
      const int p = pre + __popc(m & ((1u << lane) - 1u));
      s_box[p] = gc;
      // (g_ymax-g_ymin)*(g_xmax-g_xmin), net_tools.py:265
      s_area[p] = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
      s_id[p] = g;
    }
    total += all;
  }
  __syncthreads();

  if (!valid) return;

  // ---- per-anchor max / first-argmax over surviving GT (ascending original index)
  const float vol_a = box_vol(a);
  float best = 0.f;
  int bi = 0;
  for (int j = 0; j < total; ++j) {
    const float4 g = s_box[j];
    const float h = __fsub_rn(fminf(a.z, g.z), fmaxf(a.x, g.x));
    const float w = __fsub_rn(fminf(a.w, g.w), fmaxf(a.y, g.y));
    if (h > 0.f && w > 0.f) {
      const float inter = __fmul_rn(h, w);
      const float uni = __fadd_rn(__fsub_rn(vol_a, inter), s_area[j]);
      const float jac = __fdiv_rn(inter, uni);
      if (jac > best) { best = jac; bi = s_id[j]; }
    }
  }

  // ---- threshold, gather the matched GT, encode (net_tools.py:405-416, 334-343)
  const int l = layer_of(L, n);
  const bool pos = best >= T.v[l];
  float4 ogt = make_float4(0.f, 0.f, 0.f, 0.f), ocb = ogt;
  int lab = 0;
  if (pos && count > 0) {
    const float4 gcen = ldg4(gt_img + 4ll * bi);
    const float4 e = encode_center(ldg4(center + 4ll * n), gcen);
    // the reference accumulates 0 + 1*v, which turns -0.0 into +0.0
    ogt = make_float4(__fadd_rn(e.x, 0.f), __fadd_rn(e.y, 0.f), __fadd_rn(e.z, 0.f), __fadd_rn(e.w, 0.f));
    ocb = make_float4(__fadd_rn(gcen.x, 0.f), __fadd_rn(gcen.y, 0.f), __fadd_rn(gcen.z, 0.f), __fadd_rn(gcen.w, 0.f));
    lab = (int)labels[(long long)b * gmax + bi];       // :342 cast to int32
  }
  const long long o = (long long)b * N + n;
  st4_cs(out_gt + 4 * o, ogt);
  st4_cs(out_cb + 4 * o, ocb);
  __stcs(out_lab + o, lab);
  __stcs(out_pos + o, (pos && count > 0) ? 1 : 0);
  if (out_idx) __stcs(out_idx + o, bi);
}

// NEAREST_NEIGHBOR (:354-380, 283-312): argmin over GT of |encode(anchor, gt_i)|^2, every anchor
// positive.  Not used by any reference script; kept simple (no culling possible).
template <typename LabelT>
__global__ void __launch_bounds__(kArmBlock)
arm_nearest_kernel(Layout L, const float* __restrict__ center, const float* __restrict__ gtb,
                   const LabelT* __restrict__ labels, const int32_t* __restrict__ counts, int gmax,
                   float* __restrict__ out_gt, float* __restrict__ out_cb, int32_t* __restrict__ out_lab,
                   int32_t* __restrict__ out_pos, int32_t* __restrict__ out_idx) {
  const int b = blockIdx.y;
  const int N = L.n_total;
  const int n = blockIdx.x * kArmBlock + threadIdx.x;
  if (n >= N) return;
  int count = counts ? counts[b] : gmax;
  count = min(max(count, 0), gmax);
  const float* gt_img = gtb + 4ll * b * gmax;
  const float4 ac = ldg4(center + 4ll * n);
  float best = INFINITY;
  int bi = 0;
  float4 be = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g = 0; g < count; ++g) {
    const float4 e = encode_center(ac, ldg4(gt_img + 4ll * g));
    // tf.reduce_sum over the last axis of 4 (:364): ((e0^2 + e1^2) + e2^2) + e3^2
    const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)),
                              __fmul_rn(e.w, e.w));
    if (g == 0 || d < best) { best = d; bi = g; be = e; }
  }
  const long long o = (long long)b * N + n;
  float4 ogt = make_float4(0.f, 0.f, 0.f, 0.f), ocb = ogt;
  int lab = 0;
  if (count > 0) {
    const float4 gcen = ldg4(gt_img + 4ll * bi);
    ogt = make_float4(__fadd_rn(be.x, 0.f), __fadd_rn(be.y, 0.f), __fadd_rn(be.z, 0.f), __fadd_rn(be.w, 0.f));
    ocb = make_float4(__fadd_rn(gcen.x, 0.f), __fadd_rn(gcen.y, 0.f), __fadd_rn(gcen.z, 0.f), __fadd_rn(gcen.w, 0.f));
    lab = (int)labels[(long long)b * gmax + bi];
  }
  st4_cs(out_gt + 4 * o, ogt);
  st4_cs(out_cb + 4 * o, ocb);
  __stcs(out_lab + o, lab);
  __stcs(out_pos + o, 1);                               // :376 pos_mask = ones
  if (out_idx) __stcs(out_idx + o, bi);
}

}  // namespace rod

extern "C" int rod_arm_match_encode(const rod_layout_t* layout, const float* anchors_corner,
                                    const float* anchors_center, const float* thresholds,
                                    const float* center_bboxes, const void* labels, int labels_i64,
                                    const int32_t* gt_counts, int batch, int gmax, int method,
                                    float* gt, float* cbboxes, int32_t* out_labels, int32_t* pos_mask,
                                    int32_t* match_idx, void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  if (method == ROD_JACCARD_TOPK) {
    set_error("Not support now");                      // utils/net_tools.py:424
    return ROD_E_UNSUPPORTED;
  }
  ROD_REQUIRE(method == ROD_JACCARD_BIGGER || method == ROD_NEAREST_NEIGHBOR,
              "Function parameter \"method\" wrong");   // utils/net_tools.py:426
  ROD_REQUIRE(anchors_corner && anchors_center && center_bboxes && labels, "NULL input pointer");
  ROD_REQUIRE(gt && cbboxes && out_labels && pos_mask, "NULL output pointer");
  ROD_REQUIRE(batch >= 0 && gmax >= 1, "batch=%d gmax=%d invalid", batch, gmax);
  ROD_REQUIRE(batch <= 65535, "batch=%d exceeds 65535", batch);
  if (batch == 0) return ROD_OK;
  const Layout L = to_layout(layout);
  Thresholds T;
  for (int i = 0; i < ROD_MAX_LAYERS; ++i) T.v[i] = 0.f;
  if (method == ROD_JACCARD_BIGGER) {
    ROD_REQUIRE(thresholds != nullptr, "thresholds is NULL");
    for (int i = 0; i < L.n_layers; ++i) T.v[i] = thresholds[i];
  }
  const dim3 grid((L.n_total + kArmBlock - 1) / kArmBlock, batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (method == ROD_JACCARD_BIGGER) {
    const size_t smem = (size_t)gmax * (sizeof(float4) + sizeof(float) + sizeof(int));
    ROD_REQUIRE(smem <= 200 * 1024, "gmax=%d too large for the shared-memory GT list", gmax);
    if (labels_i64) {
      auto k = arm_jaccard_bigger_kernel<long long>;
      if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, kArmBlock, smem, st>>>(L, T, anchors_corner, anchors_center, center_bboxes,
                                       (const long long*)labels, gt_counts, gmax, gt, cbboxes, out_labels,
                                       pos_mask, match_idx);
    } else {
      auto k = arm_jaccard_bigger_kernel<int>;
      if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, kArmBlock, smem, st>>>(L, T, anchors_corner, anchors_center, center_bboxes,
                                       (const int*)labels, gt_counts, gmax, gt, cbboxes, out_labels,
                                       pos_mask, match_idx);
    }
    ROD_LAUNCH_CHECK("arm_jaccard_bigger_kernel");
  } else {
    if (labels_i64)
      arm_nearest_kernel<long long><<<grid, kArmBlock, 0, st>>>(L, anchors_center, center_bboxes,
                                                                 (const long long*)labels, gt_counts, gmax, gt,
                                                                 cbboxes, out_labels, pos_mask, match_idx);
    else
      arm_nearest_kernel<int><<<grid, kArmBlock, 0, st>>>(L, anchors_center, center_bboxes, (const int*)labels,
                                                           gt_counts, gmax, gt, cbboxes, out_labels, pos_mask,
                                                           match_idx);
    ROD_LAUNCH_CHECK("arm_nearest_kernel");
  }
  return ROD_OK;
}
