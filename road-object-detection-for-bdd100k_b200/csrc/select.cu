// a11: per-class score thresholding.  Replaces bboxes_select_one_layer /
// bboxes_select_all_layers (utils/net_tools.py:658-736): for every class c != ignore_class,
//   scores_c = p_c * (p_c >= thr)      bboxes_c = loc * (p_c >= thr)
// with the layers concatenated on the anchor axis.  Kept for drop-in completeness; the fused
// rod_detect never materialises these 200 B / anchor.
#include "common.cuh"

namespace rod {

constexpr int kSelBlock = 256;

__global__ void __launch_bounds__(kSelBlock)
select_kernel(const __grid_constant__ Layout L, const __grid_constant__ LayeredF probs,
              const __grid_constant__ LayeredF loc, int n_classes, int ignore_class, float thr, int batch,
              float* __restrict__ out_scores, float* __restrict__ out_boxes) {
  const int n = blockIdx.x * kSelBlock + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= L.n_total) return;
  const int l = layer_of(L, n);
  const float* p = probs.base[l] + (long long)b * probs.stride[l] + (long long)(n - L.offset[l]) * n_classes;
  const float4 bx = ldg4(loc.base[l] + (long long)b * loc.stride[l] + 4ll * (n - L.offset[l]));
  for (int c = 0; c < n_classes; ++c) {
    if (c == ignore_class) continue;
    const float s = __ldg(p + c);
    const float fm = (s >= thr) ? 1.f : 0.f;                       // :690
    const long long o = ((long long)c * batch + b) * L.n_total + n;
    __stcs(out_scores + o, __fmul_rn(s, fm));                      // :691
    st4_cs(out_boxes + 4 * o, make_float4(__fmul_rn(bx.x, fm), __fmul_rn(bx.y, fm), __fmul_rn(bx.z, fm),
                                           __fmul_rn(bx.w, fm)));   // :692
  }
}

}  // namespace rod

extern "C" int rod_bboxes_select(const rod_layout_t* layout, const rod_layered_t* predictions,
                                 const rod_layered_t* localizations, int batch, int n_classes,
                                 int ignore_class, float select_threshold, float* out_scores,
                                 float* out_bboxes, void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(predictions, nl, "predictions"))) return rc;
  if ((rc = check_layered(localizations, nl, "localizations"))) return rc;
  ROD_REQUIRE(out_scores && out_bboxes, "rod_bboxes_select: NULL output pointer");
  ROD_REQUIRE(batch >= 0 && batch <= 65535 && n_classes >= 1 && n_classes <= ROD_MAX_CLASSES,
              "rod_bboxes_select: batch=%d n_classes=%d invalid", batch, n_classes);
  if (batch == 0) return ROD_OK;
  const Layout L = to_layout(layout);
  const dim3 grid((L.n_total + kSelBlock - 1) / kSelBlock, batch);
  select_kernel<<<grid, kSelBlock, 0, (cudaStream_t)stream>>>(L, to_layered_f(predictions, nl),
                                                               to_layered_f(localizations, nl), n_classes,
                                                               ignore_class, select_threshold, batch, out_scores,
                                                               out_bboxes);
  ROD_LAUNCH_CHECK("select_kernel");
  return ROD_OK;
}
