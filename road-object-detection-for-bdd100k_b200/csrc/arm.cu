// a9: ARM anchor<->ground-truth matching + box encoding, batched over images.
// Replaces refine_groundtruth (utils/net_tools.py:270-428).
//
// JACCARD_BIGGER (:382-421): per anchor, max / first-argmax over the image's GT boxes of
// net_tools.jaccard, thresholded per layer; the matched GT's centre box, its encoding and
// its label are emitted for positive anchors, zeros elsewhere.
//
// Kernel shape: grid = (anchor tiles, images); a CTA owns 256 consecutive anchors of one image, a
// WARP owns 32 of them.  Only ~6 % of (anchor, GT) pairs intersect on street-scene data, and a pair
// that does not intersect contributes IoU == 0 exactly, which can never win the strict-'>' argmax
// that starts at (0, index 0).  The CTA stages the image's GT boxes (corner form + area) in shared
// memory once; each warp then culls them against the bounding box of ITS 32 anchors with one
// ballot per 32 GT boxes and walks only the surviving bits (ascending GT index => lowest-index
// tie-break is preserved), broadcasting each survivor from shared memory.  That turns the
// FP32-bound N x G loop into an HBM-write-bound kernel (40 B / anchor).
#include "common.cuh"

namespace rod {

constexpr int kArmBlock = 128;

constexpr int kArmPer = 1;                               // anchors per lane (2 measured slower on B200: 41 us vs 35 us at B=32)
constexpr int kArmTile = kArmBlock * kArmPer;

template <typename LabelT>
__global__ void __launch_bounds__(kArmBlock)
arm_jaccard_bigger_kernel(const __grid_constant__ Layout L, const __grid_constant__ Thresholds T,
                          const float* __restrict__ corner, const float* __restrict__ center,
                          const float* __restrict__ gtb, const LabelT* __restrict__ labels,
                          const int32_t* __restrict__ counts, int gmax, float* __restrict__ out_gt,
                          float* __restrict__ out_cb, int32_t* __restrict__ out_lab, int32_t* __restrict__ out_pos,
                          int32_t* __restrict__ out_idx) {
  extern __shared__ float4 s_box[];                       // [gmax] GT corners (net_tools.py:323)
  float* s_area = reinterpret_cast<float*>(s_box + gmax);  // [gmax] (g_ymax-g_ymin)*(g_xmax-g_xmin), :265
  // 32-bit shared-window addresses in opaque registers: through generic pointers ptxas re-derives the window base
  // (S2UR SR_CgaCtaId + 3 uniform ops) in every iteration of the GT walk (see target_fused.cu)
  unsigned a_box, a_area;
  asm volatile("mov.u32 %0, %1;" : "=r"(a_box) : "r"((unsigned)__cvta_generic_to_shared(s_box)));
  asm volatile("mov.u32 %0, %1;" : "=r"(a_area) : "r"((unsigned)__cvta_generic_to_shared(s_area)));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int N = L.n_total;
  const float4 none = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);   // intersects nothing

  int count = counts ? counts[b] : gmax;
  count = min(max(count, 0), gmax);
  const float* gt_img = gtb + 4ll * b * gmax;
  for (int g = tid; g < count; g += kArmBlock) {
    const float4 gc = center_to_corner(ldg4(gt_img + 4ll * g));
    const bool ok = (gc.z > gc.x) && (gc.w > gc.y);    // zero-extent GT: intersection 0 with everything
    s_box[g] = ok ? gc : none;
    s_area[g] = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
  }
  // lane owns anchors n0 + lane and n0 + 32 + lane of the warp's 64 (coalesced 512 B rows)
  // the last tiles hold the big-anchor layers, whose warps keep most GT boxes: start them first
  const int n0 = (gridDim.x - 1 - blockIdx.x) * kArmTile + warp * (32 * kArmPer);
  int n[kArmPer];
  float4 a[kArmPer];
  float vol_a[kArmPer], best[kArmPer];
  int bi[kArmPer];
#pragma unroll
  for (int u = 0; u < kArmPer; ++u) {
    n[u] = n0 + u * 32 + lane;
    a[u] = n[u] < N ? ldg4(corner + 4ll * n[u]) : none;
    vol_a[u] = box_vol(a[u]);
    best[u] = 0.f;
    bi[u] = 0;
  }
  // bounding box of the warp's anchors
  float t_ymin = a[0].x, t_xmin = a[0].y, t_ymax = a[0].z, t_xmax = a[0].w;
#pragma unroll
  for (int u = 1; u < kArmPer; ++u) {
    t_ymin = fminf(t_ymin, a[u].x); t_xmin = fminf(t_xmin, a[u].y);
    t_ymax = fmaxf(t_ymax, a[u].z); t_xmax = fmaxf(t_xmax, a[u].w);
  }
  t_ymin = warp_min(t_ymin); t_xmin = warp_min(t_xmin);
  t_ymax = warp_max(t_ymax); t_xmax = warp_max(t_xmax);
  __syncthreads();

  // ---- per-anchor max / first-argmax over the GT boxes that can intersect this warp's anchors
  for (int base = 0; base < count; base += 32) {
    const int g = base + lane;
    bool hit = false;
    if (g < count) {
      const float4 gc = lds_f4(a_box + 16u * (unsigned)g);
      hit = (gc.z > t_ymin) && (gc.x < t_ymax) && (gc.w > t_xmin) && (gc.y < t_xmax);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {                                        // ascending GT index
      const int j = base + __ffs(m) - 1;
      m &= m - 1;
      const float4 gc = lds_f4(a_box + 16u * (unsigned)j);   // broadcast
#pragma unroll
      for (int u = 0; u < kArmPer; ++u) {
        // positive intersection <=> min(a.max, g.max) > max(a.min, g.min) on both axes (both boxes
        // have positive extents, or are the `none` box)
        if ((a[u].z > gc.x) && (gc.z > a[u].x) && (a[u].w > gc.y) && (gc.w > a[u].y)) {
          const float h = __fsub_rn(fminf(a[u].z, gc.z), fmaxf(a[u].x, gc.x));
          const float w = __fsub_rn(fminf(a[u].w, gc.w), fmaxf(a[u].y, gc.y));
          const float inter = __fmul_rn(h, w);
          const float uni = __fadd_rn(__fsub_rn(vol_a[u], inter), lds_f1(a_area + 4u * (unsigned)j));
          const float jac = __fdiv_rn(inter, uni);
          if (jac > best[u]) { best[u] = jac; bi[u] = j; }
        }
      }
    }
  }

  // ---- threshold, gather the matched GT, encode (net_tools.py:405-416, 334-343)
#pragma unroll
  for (int u = 0; u < kArmPer; ++u) {
    if (n[u] >= N) continue;
    const int l = layer_of(L, n[u]);
    const bool pos = (best[u] >= T.v[l]) && count > 0;
    float4 ogt = make_float4(0.f, 0.f, 0.f, 0.f), ocb = ogt;
    int lab = 0;
    if (pos) {
      const float4 gcen = ldg4(gt_img + 4ll * bi[u]);
      const float4 e = encode_center(ldg4(center + 4ll * n[u]), gcen);
      // the reference accumulates 0 + 1*v, which turns -0.0 into +0.0
      ogt = make_float4(__fadd_rn(e.x, 0.f), __fadd_rn(e.y, 0.f), __fadd_rn(e.z, 0.f), __fadd_rn(e.w, 0.f));
      ocb = make_float4(__fadd_rn(gcen.x, 0.f), __fadd_rn(gcen.y, 0.f), __fadd_rn(gcen.z, 0.f), __fadd_rn(gcen.w, 0.f));
      lab = (int)labels[(long long)b * gmax + bi[u]];  // :342 cast to int32
    }
    const long long o = (long long)b * N + n[u];
    st4_cs(out_gt + 4 * o, ogt);
    st4_cs(out_cb + 4 * o, ocb);
    __stcs(out_lab + o, lab);
    __stcs(out_pos + o, pos ? 1 : 0);
    if (out_idx) __stcs(out_idx + o, bi[u]);
  }
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
  const dim3 grid((L.n_total + kArmBlock - 1) / kArmBlock, batch);       // nearest-neighbour kernel
  const dim3 grid_jb((L.n_total + kArmTile - 1) / kArmTile, batch);      // two anchors per thread
  cudaStream_t st = (cudaStream_t)stream;
  if (method == ROD_JACCARD_BIGGER) {
    const size_t smem = (size_t)gmax * (sizeof(float4) + sizeof(float));
    ROD_REQUIRE(smem <= 200 * 1024, "gmax=%d too large for the shared-memory GT list", gmax);
    if (labels_i64) {
      auto k = arm_jaccard_bigger_kernel<long long>;
      if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid_jb, kArmBlock, smem, st>>>(L, T, anchors_corner, anchors_center, center_bboxes,
                                       (const long long*)labels, gt_counts, gmax, gt, cbboxes, out_labels,
                                       pos_mask, match_idx);
    } else {
      auto k = arm_jaccard_bigger_kernel<int>;
      if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid_jb, kArmBlock, smem, st>>>(L, T, anchors_corner, anchors_center, center_bboxes,
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
