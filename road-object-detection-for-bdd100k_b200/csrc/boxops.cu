// a5 / a8 / a14 / a16: element-wise box helpers of utils/common_tools.py and
// utils/tf_extended/bboxes.py.  One thread per box, float4 in / out.
#include "common.cuh"

namespace rod {

constexpr int kBoxBlock = 256;

enum BoxOp { kC2Corner, kC2Center, kJaccard, kTfeJaccard, kTfeIntersection, kClip, kResize };

// safe_divide, utils/tf_extended/math.py:25-38
__device__ __forceinline__ float safe_div(float n, float d) { return d > 0.f ? __fdiv_rn(n, d) : 0.f; }

template <int OP>
__global__ void __launch_bounds__(kBoxBlock)
boxop_kernel(const float* __restrict__ a, const float* __restrict__ b, int b_broadcast,
             float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * kBoxBlock + threadIdx.x;
  if (i >= n) return;
  if (OP == kC2Corner) { st4(out + 4 * i, center_to_corner(ldg4(a + 4 * i))); return; }
  if (OP == kC2Center) { st4(out + 4 * i, corner_to_center(ldg4(a + 4 * i))); return; }
  if (OP == kJaccard) {          // net_tools.jaccard(anchors=a, corner_bbox=b), utils/net_tools.py:254-266
    const float4 x = ldg4(a + 4 * i);
    const float4 g = ldg4(b + (b_broadcast ? 0 : 4 * i));
    out[i] = jaccard_ref(x, box_vol(x), g, __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y)));
    return;
  }
  // the tfe ops take (bbox_ref=a, bboxes=b): a is [4] or [n,4]
  const float4 r = ldg4(a + (b_broadcast ? 0 : 4 * i));
  const float4 x = ldg4(b + 4 * i);
  if (OP == kResize) {           // utils/tf_extended/bboxes.py:154-162
    const float sy = __fsub_rn(r.z, r.x), sx = __fsub_rn(r.w, r.y);
    st4(out + 4 * i, make_float4(__fdiv_rn(__fsub_rn(x.x, r.x), sy), __fdiv_rn(__fsub_rn(x.y, r.y), sx),
                                 __fdiv_rn(__fsub_rn(x.z, r.x), sy), __fdiv_rn(__fsub_rn(x.w, r.y), sx)));
    return;
  }
  const float ymin = fmaxf(x.x, r.x), xmin = fmaxf(x.y, r.y), ymax = fminf(x.z, r.z), xmax = fminf(x.w, r.w);
  if (OP == kClip) {             // utils/tf_extended/bboxes.py:128-135
    st4(out + 4 * i, make_float4(fminf(ymin, ymax), fminf(xmin, xmax), ymax, xmax));
    return;
  }
  const float h = fmaxf(__fsub_rn(ymax, ymin), 0.f), w = fmaxf(__fsub_rn(xmax, xmin), 0.f);
  const float inter = __fmul_rn(h, w);
  const float vol_b = __fmul_rn(__fsub_rn(x.z, x.x), __fsub_rn(x.w, x.y));
  if (OP == kTfeIntersection) { out[i] = safe_div(inter, vol_b); return; }   // :506-507
  // :475-478 union = ((-inter) + area_b) + area_ref
  const float vol_r = __fmul_rn(__fsub_rn(r.z, r.x), __fsub_rn(r.w, r.y));
  out[i] = safe_div(inter, __fadd_rn(__fadd_rn(-inter, vol_b), vol_r));
}

template <int OP>
static int launch(const float* a, const float* b, int bc, float* out, long long n, void* stream, const char* name) {
  ROD_REQUIRE(a && out && n >= 0, "%s: NULL pointer or negative size", name);
  if (n == 0) return ROD_OK;
  const long long blocks = (n + kBoxBlock - 1) / kBoxBlock;
  ROD_REQUIRE(blocks < 2147483647ll, "%s: too many boxes", name);
  boxop_kernel<OP><<<(unsigned)blocks, kBoxBlock, 0, (cudaStream_t)stream>>>(a, b, bc, out, n);
  ROD_LAUNCH_CHECK(name);
  return ROD_OK;
}

}  // namespace rod

extern "C" {
int rod_center_to_corner(const float* in, float* out, int64_t n, void* stream) {
  return rod::launch<rod::kC2Corner>(in, nullptr, 0, out, n, stream, "rod_center_to_corner");
}
int rod_corner_to_center(const float* in, float* out, int64_t n, void* stream) {
  return rod::launch<rod::kC2Center>(in, nullptr, 0, out, n, stream, "rod_corner_to_center");
}
int rod_jaccard(const float* a, const float* b, int b_broadcast, float* out, int64_t n, void* stream) {
  ROD_REQUIRE(b != nullptr, "rod_jaccard: b is NULL");
  return rod::launch<rod::kJaccard>(a, b, b_broadcast, out, n, stream, "rod_jaccard");
}
int rod_bboxes_jaccard(const float* ref, int ref_broadcast, const float* boxes, float* out, int64_t n, void* stream) {
  ROD_REQUIRE(boxes != nullptr, "rod_bboxes_jaccard: boxes is NULL");
  return rod::launch<rod::kTfeJaccard>(ref, boxes, ref_broadcast, out, n, stream, "rod_bboxes_jaccard");
}
int rod_bboxes_intersection(const float* ref, int ref_broadcast, const float* boxes, float* out, int64_t n, void* stream) {
  ROD_REQUIRE(boxes != nullptr, "rod_bboxes_intersection: boxes is NULL");
  return rod::launch<rod::kTfeIntersection>(ref, boxes, ref_broadcast, out, n, stream, "rod_bboxes_intersection");
}
int rod_bboxes_clip(const float* ref, int ref_broadcast, const float* boxes, float* out, int64_t n, void* stream) {
  ROD_REQUIRE(boxes != nullptr, "rod_bboxes_clip: boxes is NULL");
  return rod::launch<rod::kClip>(ref, boxes, ref_broadcast, out, n, stream, "rod_bboxes_clip");
}
int rod_bboxes_resize(const float* ref, const float* boxes, float* out, int64_t n, void* stream) {
  ROD_REQUIRE(boxes != nullptr, "rod_bboxes_resize: boxes is NULL");
  return rod::launch<rod::kResize>(ref, boxes, 1, out, n, stream, "rod_bboxes_resize");
}
}
