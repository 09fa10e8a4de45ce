// a10 ODM target generation (det_groundtruth, utils/net_tools.py:431-475),
// a7/a17 decode (decode_locations_one_layer :182-234; evaluate.py:139-143),
// a6 encode of one box against every anchor (:147-179).
// All element-wise over (image, anchor): one thread per anchor, float4 traffic,
// grid = (anchor tiles, images).  Inputs arrive as per-layer lists (rod_layered_t).
#include "common.cuh"

namespace rod {

constexpr int kEwBlock = 256;

__device__ __forceinline__ const float* lp4(const LayeredF& t, const Layout& L, int l, int b, int n) {
  return t.base[l] + (long long)b * t.stride[l] + 4ll * (n - L.offset[l]);
}
__device__ __forceinline__ const int32_t* lp1(const LayeredI& t, const Layout& L, int l, int b, int n) {
  return t.base[l] + (long long)b * t.stride[l] + (n - L.offset[l]);
}

__global__ void __launch_bounds__(kEwBlock)
odm_target_kernel(const __grid_constant__ Layout L, const __grid_constant__ Thresholds T,
                  const float* __restrict__ center, const __grid_constant__ LayeredF refine_out,
                  const __grid_constant__ LayeredF offset_gt, const __grid_constant__ LayeredF cbboxes,
                  const __grid_constant__ LayeredI labels, const __grid_constant__ LayeredI pos_mask,
                  float* __restrict__ det_gt, int32_t* __restrict__ mask, int32_t* __restrict__ det_labels,
                  float* __restrict__ iou) {
  const int n = blockIdx.x * kEwBlock + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= L.n_total) return;
  const int l = layer_of(L, n);
  const float4 ro = ldg4(lp4(refine_out, L, l, b, n));
  const float4 og = ldg4(lp4(offset_gt, L, l, b, n));
  const float4 cb = ldg4(lp4(cbboxes, L, l, b, n));
  const int lab = __ldg(lp1(labels, L, l, b, n));
  const int pm = __ldg(lp1(pos_mask, L, l, b, n));
  const float4 ac = ldg4(center + 4ll * n);
  // Fast path for warps without any matched anchor (most of them).  An unmatched anchor carries
  // cbboxes == (0,0,0,0), so its GT corner box is (0,0,0,0): the y-extent of the intersection is
  // max(min(ymax,0) - max(ymin,0), 0) == 0 whatever the refined box is, hence inter == 0 and
  // iou == 0 / vol_a == 0 exactly as long as vol_a is a positive finite number.  The bounds below
  // guarantee that without evaluating exp(): h = exp(o2)*ah >= e^-2 * 2^-10 dwarfs the rounding of
  // cy +- h/2 for |cy| <= 16, and e^40 * ah stays far from overflow.
  const float cy = __fadd_rn(__fmul_rn(ro.x, ac.z), ac.x), cx = __fadd_rn(__fmul_rn(ro.y, ac.w), ac.y);
  const bool trivial = cb.x == 0.f && cb.y == 0.f && cb.z == 0.f && cb.w == 0.f && ro.z >= -2.f && ro.z <= 40.f &&
                       ro.w >= -2.f && ro.w <= 40.f && fabsf(cy) <= 16.f && fabsf(cx) <= 16.f &&
                       ac.z >= 0.0009765625f && ac.w >= 0.0009765625f;
  float j = 0.f;
  if (!__all_sync(__activemask(), trivial)) {
    // :459-460 refined anchors, corner form ; :463 matched GT, corner form
    const float4 ra = center_to_corner(decode_center(ac, ro));
    const float4 gc = center_to_corner(cb);
    // :465 element-wise jaccard(refined anchor, its assigned GT)
    const float area_g = __fmul_rn(__fsub_rn(gc.z, gc.x), __fsub_rn(gc.w, gc.y));
    j = jaccard_ref(ra, box_vol(ra), gc, area_g);
  }
  // :468-469
  const int m = ((j >= T.v[l]) ? 1 : 0) * pm;
  const float mf = (float)m;
  // :471 (offset_gt - refine_out) * float(mask)
  const float4 d = make_float4(__fmul_rn(__fsub_rn(og.x, ro.x), mf), __fmul_rn(__fsub_rn(og.y, ro.y), mf),
                               __fmul_rn(__fsub_rn(og.z, ro.z), mf), __fmul_rn(__fsub_rn(og.w, ro.w), mf));
  const long long o = (long long)b * L.n_total + n;
  st4_cs(det_gt + 4 * o, d);
  __stcs(mask + o, m);
  __stcs(det_labels + o, lab * m);    // :472
  __stcs(iou + o, j);
}

// has_det: 0 single decode of refine_out; 1 the reference's call site decode(anchors, refine_out + det_out)
// (evaluate.py:141); 2 (opt-in extra, no reference counterpart) the RefineDet cascade: the refined anchors
// decode(anchors, refine_out) are treated as a new anchor layer and det_out is decoded against them, i.e.
// decode_locations_one_layer applied twice with its own corner -> re-derived centre step in between (:156-171).
__global__ void __launch_bounds__(kEwBlock)
decode_kernel(const __grid_constant__ Layout L, const float* __restrict__ center,
              const __grid_constant__ LayeredF refine_out, const __grid_constant__ LayeredF det_out,
              int has_det, int to_corner, float* __restrict__ out) {
  const int n = blockIdx.x * kEwBlock + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= L.n_total) return;
  const int l = layer_of(L, n);
  float4 o = ldg4(lp4(refine_out, L, l, b, n));
  float4 r;
  if (has_det == 2) {
    const float4 d = ldg4(lp4(det_out, L, l, b, n));
    const float4 refined = corner_to_center(center_to_corner(decode_center(ldg4(center + 4ll * n), o)));
    r = decode_center(refined, d);
  } else {
    if (has_det) {                       // evaluate.py:141 (refine_out + det_out)
      const float4 d = ldg4(lp4(det_out, L, l, b, n));
      o = make_float4(__fadd_rn(o.x, d.x), __fadd_rn(o.y, d.y), __fadd_rn(o.z, d.z), __fadd_rn(o.w, d.w));
    }
    r = decode_center(ldg4(center + 4ll * n), o);
  }
  if (to_corner) r = center_to_corner(r);   // evaluate.py:142
  st4_cs(out + 4 * ((long long)b * L.n_total + n), r);
}

__global__ void __launch_bounds__(kEwBlock)
encode_one_box_kernel(const float* __restrict__ center, int first, int n, const float* __restrict__ box,
                      float* __restrict__ out) {
  const int i = blockIdx.x * kEwBlock + threadIdx.x;
  if (i >= n) return;
  const float4 g = ldg4(box);
  st4(out + 4ll * i, encode_center(ldg4(center + 4ll * (first + i)), g));
}

}  // namespace rod

extern "C" int rod_odm_target(const rod_layout_t* layout, const float* anchors_center,
                              const float* thresholds, const rod_layered_t* refine_out,
                              const rod_layered_t* offset_gt, const rod_layered_t* cbboxes,
                              const rod_layered_t* refine_labels, const rod_layered_t* refine_pos_mask,
                              int batch, float* det_gt, int32_t* mask, int32_t* det_labels, float* iou,
                              void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(refine_out, nl, "refine_out"))) return rc;
  if ((rc = check_layered(offset_gt, nl, "offset_gt"))) return rc;
  if ((rc = check_layered(cbboxes, nl, "cbboxes"))) return rc;
  if ((rc = check_layered(refine_labels, nl, "refine_labels"))) return rc;
  if ((rc = check_layered(refine_pos_mask, nl, "refine_pos_mask"))) return rc;
  ROD_REQUIRE(anchors_center && thresholds, "NULL input pointer");
  ROD_REQUIRE(det_gt && mask && det_labels && iou, "NULL output pointer");
  ROD_REQUIRE(batch >= 0 && batch <= 65535, "batch=%d invalid", batch);
  if (batch == 0) return ROD_OK;
  const Layout L = to_layout(layout);
  Thresholds T;
  for (int i = 0; i < ROD_MAX_LAYERS; ++i) T.v[i] = i < nl ? thresholds[i] : 0.f;
  const dim3 grid((L.n_total + kEwBlock - 1) / kEwBlock, batch);
  odm_target_kernel<<<grid, kEwBlock, 0, (cudaStream_t)stream>>>(
      L, T, anchors_center, to_layered_f(refine_out, nl), to_layered_f(offset_gt, nl), to_layered_f(cbboxes, nl),
      to_layered_i(refine_labels, nl), to_layered_i(refine_pos_mask, nl), det_gt, mask, det_labels, iou);
  ROD_LAUNCH_CHECK("odm_target_kernel");
  return ROD_OK;
}

static int decode_impl(const rod_layout_t* layout, const float* anchors_center, const rod_layered_t* refine_out,
                       const rod_layered_t* det_out, int batch, int to_corner, float* out, void* stream, int cascade);

extern "C" int rod_decode(const rod_layout_t* layout, const float* anchors_center,
                          const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                          int to_corner, float* out, void* stream) {
  return decode_impl(layout, anchors_center, refine_out, det_out, batch, to_corner, out, stream, 0);
}

extern "C" int rod_decode_cascade(const rod_layout_t* layout, const float* anchors_center,
                                  const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                                  int to_corner, float* out, void* stream) {
  if (det_out == nullptr) {
    rod::set_error("rod_decode_cascade: det_out is NULL");
    return ROD_E_INVALID;
  }
  return decode_impl(layout, anchors_center, refine_out, det_out, batch, to_corner, out, stream, 1);
}

static int decode_impl(const rod_layout_t* layout, const float* anchors_center, const rod_layered_t* refine_out,
                       const rod_layered_t* det_out, int batch, int to_corner, float* out, void* stream, int cascade) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(refine_out, nl, "refine_out"))) return rc;
  if (det_out && (rc = check_layered(det_out, nl, "det_out"))) return rc;
  ROD_REQUIRE(anchors_center && out, "NULL pointer argument");
  ROD_REQUIRE(batch >= 0 && batch <= 65535, "batch=%d invalid", batch);
  if (batch == 0) return ROD_OK;
  const Layout L = to_layout(layout);
  const dim3 grid((L.n_total + kEwBlock - 1) / kEwBlock, batch);
  const LayeredF ro = to_layered_f(refine_out, nl);
  decode_kernel<<<grid, kEwBlock, 0, (cudaStream_t)stream>>>(L, anchors_center, ro,
                                                              det_out ? to_layered_f(det_out, nl) : ro,
                                                              det_out ? (cascade ? 2 : 1) : 0, to_corner, out);
  ROD_LAUNCH_CHECK("decode_kernel");
  return ROD_OK;
}

extern "C" int rod_encode_one_box(const float* anchors_center, int first, int n, const float* center_bbox,
                                  float* out, void* stream) {
  using namespace rod;
  ROD_REQUIRE(anchors_center && center_bbox && out, "NULL pointer argument");
  ROD_REQUIRE(first >= 0 && n >= 0, "first=%d n=%d invalid", first, n);
  if (n == 0) return ROD_OK;
  encode_one_box_kernel<<<(n + kEwBlock - 1) / kEwBlock, kEwBlock, 0, (cudaStream_t)stream>>>(
      anchors_center, first, n, center_bbox, out);
  ROD_LAUNCH_CHECK("encode_one_box_kernel");
  return ROD_OK;
}
