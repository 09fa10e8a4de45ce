// Fused training-target generation: a9 ARM matching + encoding (refine_groundtruth,
// utils/net_tools.py:270-428, JACCARD_BIGGER branch :382-421) and a10 ODM target generation
// (det_groundtruth, :431-475) in ONE kernel — the call sequence train.py:109-113 -> :147-149.
//
// The two-call path writes 40 B / anchor in ARM and reads them back in ODM.  Here the matched GT box,
// its encoding, label and mask stay in registers: per anchor the kernel reads refine_out (16 B) and
// writes the eight outputs (40 + 28 B), 84 B instead of 124 B of HBM traffic, and ARM's issue-bound
// GT walk overlaps ODM's loads / stores inside one launch.  Results are bit-identical to
// rod_arm_match_encode followed by rod_odm_target (same device functions, same op order).
//
// Shape: one resident wave of persistent CTAs that pull work items (one 128-anchor tile of one image)
// from a global counter, most expensive first (the tiles of the big-anchor layers meet every GT box, the
// 64x64 layer's tiles only ~6 % of them; an image with 100 GT boxes costs ~10x one with 5): a static
// partition leaves 40 % of the issue slots idle behind the slowest CTA.  A CTA stages the item's GT
// boxes (corner form + area) in shared memory; inside the tile a WARP owns 32 consecutive anchors: it
// culls the GT list against the bounding box of its anchors with one ballot per 32 GT boxes and walks
// the surviving bits in ascending GT index (= lowest-index tie-break of tf.argmax), exactly like
// arm_jaccard_bigger_kernel.  The counter pair lives in an 8-byte caller workspace that must be zero
// before the first call and is left zero by every call (the last CTA to leave resets it).
#include "common.cuh"

namespace rod {

constexpr int kTfBlock = 128;

// (the GT walk reads shared memory through lds_f4 / lds_f1 on a 32-bit window address kept in an opaque register:
// through a generic pointer ptxas re-derives the window base — S2UR SR_CgaCtaId + 3 uniform ops — in every iteration)

struct FusedParams {
  Layout L;
  Thresholds Ta, To;                    // ARM / ODM per-layer IoU thresholds (config.py:79-80)
  const float* corner;                  // [N,4] anchor corners
  const float* center;                  // [N,4] anchor centre form re-derived from the corners
  const float* gtb;                     // [B,gmax,4] GT centre boxes
  const void* labels;                   // [B,gmax] int32 / int64
  const int32_t* counts;                // [B] or NULL
  LayeredF refine_out;                  // ARM head output, per layer [B,...,4]
  float* out_gt; float* out_cb; int32_t* out_lab; int32_t* out_pos; int32_t* out_idx;     // ARM outputs (cb / lab / idx optional)
  float* det_gt; int32_t* det_mask; int32_t* det_lab; float* iou;                          // ODM outputs
  int gmax, tiles, batch;
  unsigned* sched;                      // [0] next work item, [1] CTAs that have left
};

// (10 resident CTAs per SM asked of the compiler = 47 registers: 36.9 us at B = 32; 8 / 12 / 14 CTAs measured 38.5 /
// 37.3 / 40.3 us — the kernel is bound by its instruction count, not by occupancy)
template <typename LabelT>
__global__ void __launch_bounds__(kTfBlock, 10)
target_fused_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ float4 s_box[];                         // [gmax] GT corners (net_tools.py:323)
  float* s_area = reinterpret_cast<float*>(s_box + P.gmax);  // [gmax] GT areas (:265)
  unsigned a_box, a_area;                                   // (opaque moves: ptxas otherwise rematerialises the window base per use)
  asm volatile("mov.u32 %0, %1;" : "=r"(a_box) : "r"((unsigned)__cvta_generic_to_shared(s_box)));
  asm volatile("mov.u32 %0, %1;" : "=r"(a_area) : "r"((unsigned)__cvta_generic_to_shared(s_area)));

  __shared__ int s_tile, s_img;                             // the current work item, decoded by thread 0 (tile < 0: none left)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.L.n_total;
  const float4 none = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);   // intersects nothing
  const unsigned total = (unsigned)P.tiles * (unsigned)P.batch;

  // item -> (tile, image): all images' last tile first, then the one before, ... (descending cost); one division per
  // item per CTA (thread 0), not per thread
  auto fetch = [&]() {
    const unsigned item = atomicAdd(P.sched, 1u);
    const bool any = item < total;
    s_tile = any ? P.tiles - 1 - (int)(item / (unsigned)P.batch) : -1;
    s_img = any ? (int)(item % (unsigned)P.batch) : 0;
  };
  if (tid == 0) fetch();
  __syncthreads();
  int t = s_tile, b = s_img;
  while (t >= 0) {
    __syncthreads();                                        // every warp has read the item and is done with the GT list
    if (tid == 0) fetch();                                  // next item: the fetch overlaps the GT staging
    int count = P.counts ? P.counts[b] : P.gmax;
    count = min(max(count, 0), P.gmax);
    const float* gt_img = P.gtb + 4ll * b * P.gmax;
    for (int g = tid; g < count; g += kTfBlock) {
      const float4 gc = center_to_corner(ldg4(gt_img + 4ll * g));
      const bool ok = (gc.z > gc.x) && (gc.w > gc.y);       // zero-extent GT: intersection 0 with everything
      s_box[g] = ok ? gc : none;
      s_area[g] = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
    }
    __syncthreads();
    const int t_next = s_tile, b_next = s_img;
    {
    const int n = t * kTfBlock + warp * 32 + lane;
    const bool valid = n < N;
    const int l = layer_of(P.L, valid ? n : N - 1);
    const long long li = 4ll * ((valid ? n : N - 1) - P.L.offset[l]);
    // ODM's only input: issued now, consumed after the GT walk
    const float4 ro = ldg4(P.refine_out.base[l] + (long long)b * P.refine_out.stride[l] + li);
    const float4 a = valid ? ldg4(P.corner + 4ll * n) : none;
    const float vol_a = box_vol(a);
    // bounding box of the warp's anchors
    const float t_ymin = warp_min(a.x), t_xmin = warp_min(a.y), t_ymax = warp_max(a.z), t_xmax = warp_max(a.w);

    // ---- ARM: per-anchor max / first-argmax over the GT boxes that can intersect this warp's anchors
    float best = 0.f;
    int bi = 0;
    for (int base = 0; base < count; base += 32) {
      const int g = base + lane;
      bool hit = false;
      if (g < count) {
        const float4 gc = lds_f4(a_box + 16u * (unsigned)g);
        hit = (gc.z > t_ymin) && (gc.x < t_ymax) && (gc.w > t_xmin) && (gc.y < t_xmax);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {                                           // ascending GT index
        const int j = base + __ffs(m) - 1;
        m &= m - 1;
        const float4 gc = lds_f4(a_box + 16u * (unsigned)j);  // broadcast
        // positive intersection <=> both extents > 0 (x - y > 0 <=> x > y in IEEE arithmetic without FTZ)
        const float h = __fsub_rn(fminf(a.z, gc.z), fmaxf(a.x, gc.x));
        const float w = __fsub_rn(fminf(a.w, gc.w), fmaxf(a.y, gc.y));
        if (h > 0.f && w > 0.f) {
          const float inter = __fmul_rn(h, w);
          const float uni = __fadd_rn(__fsub_rn(vol_a, inter), lds_f1(a_area + 4u * (unsigned)j));
          const float jac = __fdiv_rn(inter, uni);
          if (jac > best) { best = jac; bi = j; }
        }
      }
    }
    if (valid) {

    // ---- ARM epilogue: threshold, gather the matched GT, encode (net_tools.py:405-416, 334-343)
    const bool pos = (best >= P.Ta.v[l]) && count > 0;
    const float4 ac = ldg4(P.center + 4ll * n);
    float4 ogt = make_float4(0.f, 0.f, 0.f, 0.f), ocb = ogt;
    int lab = 0;
    if (pos) {
      const float4 gcen = ldg4(gt_img + 4ll * bi);
      const float4 e = encode_center(ac, gcen);
      // the reference accumulates 0 + 1*v, which turns -0.0 into +0.0
      ogt = make_float4(__fadd_rn(e.x, 0.f), __fadd_rn(e.y, 0.f), __fadd_rn(e.z, 0.f), __fadd_rn(e.w, 0.f));
      ocb = make_float4(__fadd_rn(gcen.x, 0.f), __fadd_rn(gcen.y, 0.f), __fadd_rn(gcen.z, 0.f), __fadd_rn(gcen.w, 0.f));
      lab = (int)static_cast<const LabelT*>(P.labels)[(long long)b * P.gmax + bi];    // :342 cast to int32
    }
    const long long o = (long long)b * N + n;
    st4_cs(P.out_gt + 4 * o, ogt);
    if (P.out_cb) st4_cs(P.out_cb + 4 * o, ocb);
    if (P.out_lab) __stcs(P.out_lab + o, lab);
    __stcs(P.out_pos + o, pos ? 1 : 0);
    if (P.out_idx) __stcs(P.out_idx + o, bi);

    // ---- ODM (net_tools.py:454-473) on the values ARM just produced.  An unmatched anchor carries
    // cbboxes == 0, hence GT corner box (0,0,0,0), inter == 0 and iou == 0 / vol == 0 exactly as long as the
    // refined box has a positive finite volume — guaranteed by the bounds below without evaluating exp()
    // (same argument as odm_target_kernel).
    const float cy = __fadd_rn(__fmul_rn(ro.x, ac.z), ac.x), cx = __fadd_rn(__fmul_rn(ro.y, ac.w), ac.y);
    const bool trivial = !pos && ro.z >= -2.f && ro.z <= 40.f && ro.w >= -2.f && ro.w <= 40.f && fabsf(cy) <= 16.f &&
                         fabsf(cx) <= 16.f && ac.z >= 0.0009765625f && ac.w >= 0.0009765625f;
    float j = 0.f;
    if (!__all_sync(__activemask(), trivial)) {
      const float4 ra = center_to_corner(decode_center(ac, ro));     // :459-460
      const float4 gc = center_to_corner(ocb);                       // :463
      const float area_g = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
      j = jaccard_ref(ra, box_vol(ra), gc, area_g);                  // :465
    }
    const int mk = ((j >= P.To.v[l]) ? 1 : 0) * (pos ? 1 : 0);       // :468-469
    const float mf = (float)mk;
    const float4 d = make_float4(__fmul_rn(__fsub_rn(ogt.x, ro.x), mf), __fmul_rn(__fsub_rn(ogt.y, ro.y), mf),
                                 __fmul_rn(__fsub_rn(ogt.z, ro.z), mf), __fmul_rn(__fsub_rn(ogt.w, ro.w), mf));   // :471
    st4_cs(P.det_gt + 4 * o, d);
    __stcs(P.det_mask + o, mk);
    __stcs(P.det_lab + o, lab * mk);                                 // :472
    __stcs(P.iou + o, j);
    }   // valid
    }   // tile
    t = t_next; b = b_next;
  }
  // leave: the last CTA resets the counters for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(P.sched + 1, 1u) == gridDim.x - 1u) {
      P.sched[0] = 0u;
      P.sched[1] = 0u;
    }
  }
}

static int g_tf_ctas_per_sm = 0;

}  // namespace rod

extern "C" size_t rod_target_fused_workspace_bytes(void) { return 8; }

extern "C" int rod_target_fused(const rod_layout_t* layout, const float* anchors_corner, const float* anchors_center,
                                const float* arm_thresholds, const float* odm_thresholds, const float* center_bboxes,
                                const void* labels, int labels_i64, const int32_t* gt_counts, int batch, int gmax,
                                const rod_layered_t* refine_out, float* gt, float* cbboxes, int32_t* out_labels,
                                int32_t* pos_mask, int32_t* match_idx, float* det_gt, int32_t* det_mask,
                                int32_t* det_labels, float* iou, void* workspace, void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(refine_out, nl, "refine_out"))) return rc;
  ROD_REQUIRE(anchors_corner && anchors_center && arm_thresholds && odm_thresholds && center_bboxes && labels,
              "rod_target_fused: NULL input pointer");
  ROD_REQUIRE(gt && pos_mask && det_gt && det_mask && det_labels && iou, "rod_target_fused: NULL output pointer");
  ROD_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "rod_target_fused: workspace is NULL or not 8-byte aligned");
  ROD_REQUIRE(batch >= 0 && gmax >= 1, "rod_target_fused: batch=%d gmax=%d invalid", batch, gmax);
  ROD_REQUIRE(batch <= 65535, "rod_target_fused: batch=%d exceeds 65535", batch);
  if (batch == 0) return ROD_OK;
  FusedParams P;
  P.L = to_layout(layout);
  for (int i = 0; i < ROD_MAX_LAYERS; ++i) {
    P.Ta.v[i] = i < nl ? arm_thresholds[i] : 0.f;
    P.To.v[i] = i < nl ? odm_thresholds[i] : 0.f;
  }
  P.corner = anchors_corner; P.center = anchors_center; P.gtb = center_bboxes; P.labels = labels; P.counts = gt_counts;
  P.refine_out = to_layered_f(refine_out, nl);
  P.out_gt = gt; P.out_cb = cbboxes; P.out_lab = out_labels; P.out_pos = pos_mask; P.out_idx = match_idx;
  P.det_gt = det_gt; P.det_mask = det_mask; P.det_lab = det_labels; P.iou = iou;
  P.gmax = gmax;
  P.tiles = (P.L.n_total + kTfBlock - 1) / kTfBlock;
  const size_t smem = (size_t)gmax * (sizeof(float4) + sizeof(float));
  ROD_REQUIRE(smem <= 200 * 1024, "rod_target_fused: gmax=%d too large for the shared-memory GT list", gmax);
  auto k = labels_i64 ? target_fused_kernel<long long> : target_fused_kernel<int>;
  if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (g_tf_ctas_per_sm == 0) {
    int per = 0;
    ROD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, target_fused_kernel<long long>, kTfBlock, 2048));
    g_tf_ctas_per_sm = per > 0 ? per : 1;
  }
  // one resident wave of persistent CTAs (fewer when there is less work)
  P.batch = batch;
  P.sched = static_cast<unsigned*>(workspace);
  const long long resident = (long long)g_tf_ctas_per_sm * sm_count(), items = (long long)P.tiles * batch;
  const unsigned grid = (unsigned)(items < resident ? items : resident);
  k<<<grid, kTfBlock, smem, (cudaStream_t)stream>>>(P);
  ROD_LAUNCH_CHECK("target_fused_kernel");
  return ROD_OK;
}
