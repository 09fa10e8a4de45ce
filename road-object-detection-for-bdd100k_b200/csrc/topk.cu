// a12: exact segmented top-k (tf.nn.top_k semantics: descending, equal scores -> lower
// index first), one CTA per (image, class) segment.  Replaces tfe.bboxes_sort
// (utils/tf_extended/bboxes.py:60-100) and, with the SelectedScores source, the
// select + sort front half of detected_bboxes (utils/net_tools.py:745-750).
//
// The algorithm is topk_row (select_topk.cuh).
#include "select_topk.cuh"

namespace rod {

constexpr int kTopkBlock = 1024;

template <typename Src>
__global__ void __launch_bounds__(kTopkBlock)
topk_segment_kernel(const __grid_constant__ Src src, long long rows, int k, float* __restrict__ out_scores,
                    int32_t* __restrict__ out_idx,
                    const float* __restrict__ gather_boxes, float* __restrict__ out_boxes) {
  pdl_wait();                                          // (fused path: launched while the segment kernel still runs)
  if (!src.any_active()) return;
  // persistent over rows: the fused path launches a small grid that usually finds nothing to do
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    if (!src.row_active(r)) continue;
    topk_row<Src, kTopkBlock>(src, r, k, out_scores, out_idx, gather_boxes, out_boxes);
  }
}

int launch_topk_selected(const SelectedScores& src, long long rows, int k, float* out_scores, int32_t* out_idx,
                         cudaStream_t st) {
  const unsigned grid = src.over_cnt ? (unsigned)(rows < 2 * sm_count() ? rows : 2 * sm_count()) : (unsigned)rows;
  ROD_CUDA(launch_pdl(4, topk_segment_kernel<SelectedScores>, dim3(grid), dim3(kTopkBlock), 0, st, src, rows, k, out_scores, out_idx,
                      (const float*)nullptr, (float*)nullptr));
  return ROD_OK;
}

}  // namespace rod

extern "C" int rod_bboxes_sort(const float* scores, const float* bboxes, int64_t rows, int n, int top_k,
                               float* out_scores, float* out_bboxes, int32_t* out_idx, void* stream) {
  using namespace rod;
  ROD_REQUIRE(scores && out_scores, "rod_bboxes_sort: NULL pointer argument");
  ROD_REQUIRE((bboxes == nullptr) == (out_bboxes == nullptr), "rod_bboxes_sort: bboxes and out_bboxes go together");
  ROD_REQUIRE(rows >= 0 && rows < 2147483647ll && n >= 1, "rod_bboxes_sort: rows=%lld n=%d invalid", (long long)rows, n);
  ROD_REQUIRE(top_k >= 1 && top_k <= n, "rod_bboxes_sort: top_k=%d must be in [1, n=%d] (tf.nn.top_k requires k <= n)", top_k, n);
  if (top_k > ROD_MAX_TOPK) {
    set_error("rod_bboxes_sort: top_k=%d > %d is not supported", top_k, ROD_MAX_TOPK);
    return ROD_E_UNSUPPORTED;
  }
  if (rows == 0) return ROD_OK;
  DenseScores src{scores, n};
  topk_segment_kernel<DenseScores><<<(unsigned)rows, kTopkBlock, 0, (cudaStream_t)stream>>>(
      src, rows, top_k, out_scores, out_idx, bboxes, out_bboxes);
  ROD_LAUNCH_CHECK("topk_segment_kernel<DenseScores>");
  return ROD_OK;
}
