// a12: exact segmented top-k (tf.nn.top_k semantics: descending, equal scores -> lower
// index first), one CTA per (image, class) segment.  Replaces tfe.bboxes_sort
// (utils/tf_extended/bboxes.py:60-100) and, with the SelectedScores source, the
// select + sort front half of detected_bboxes (utils/net_tools.py:745-750).
//
// Algorithm: 4 x 8-bit MSB-first radix select on order-preserving uint32 keys (warp-aggregated
// shared-memory histograms), then an index-ordered compaction that takes every key above the
// k-th and the FIRST `need` keys equal to it, then a bitonic sort of the k survivors on the
// 64-bit composite (key, ~index).  Every tie rule is therefore decided by index, never by
// thread scheduling.
#include "select_topk.cuh"

namespace rod {

constexpr int kTopkBlock = 1024;

template <typename Src>
__global__ void __launch_bounds__(kTopkBlock)
topk_segment_kernel(const __grid_constant__ Src src, long long rows, int k, float* __restrict__ out_scores,
                    int32_t* __restrict__ out_idx,
                    const float* __restrict__ gather_boxes, float* __restrict__ out_boxes) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned long long s_sel[ROD_MAX_TOPK];
  __shared__ int s_warp[kTopkBlock / 32];
  __shared__ unsigned s_bin, s_above, s_gt_count;

  pdl_wait();                                          // (fused path: launched while the segment kernel still runs)
  if (!src.any_active()) return;
  const int n = src.size();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // persistent over rows: the fused path launches a small grid that usually finds nothing to do
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
  if (!src.row_active(r)) continue;
  __syncthreads();

  // ---------------- radix select of the k-th largest key
  unsigned prefix = 0, pmask = 0;
  int need = k;                       // how many still to take among keys matching the prefix
  unsigned eq_total = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += kTopkBlock) {
      const int i = i0 + tid;
      bool real;
      const bool in = i < n;
      const unsigned key = in ? float_key(src.fetch(r, i, real)) : 0u;
      const bool match = in && ((key & pmask) == prefix);
      const unsigned active = __ballot_sync(0xffffffffu, match);
      if (match) {
        const unsigned bin = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(active, bin);
        if (lane == (__ffs(peers) - 1)) atomicAdd(&s_hist[bin], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    if (warp == 0) {
      // lane j owns bins [255-8j-7, 255-8j] (descending order of key)
      unsigned loc[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { loc[q] = s_hist[255 - (8 * lane + q)]; sum += loc[q]; }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned above = incl - sum;     // keys in strictly higher bins than this lane's group
      if (above < (unsigned)need && incl >= (unsigned)need) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (above + loc[q] >= (unsigned)need) { s_bin = 255 - (8 * lane + q); s_above = above; break; }
          above += loc[q];
        }
      }
    }
    __syncthreads();
    const unsigned bin = s_bin;
    need -= (int)s_above;
    prefix |= bin << shift;
    pmask |= 255u << shift;
    eq_total = s_hist[bin];
    __syncthreads();
  }
  const unsigned kth = prefix;         // k-th largest key; `need` of the eq_total equal keys are taken
  const int n_gt = k - need;
  (void)eq_total;

  // ---------------- index-ordered compaction
  if (tid == 0) s_gt_count = 0;
  __syncthreads();
  int eq_seen = 0;
  for (int i0 = 0; i0 < n; i0 += kTopkBlock) {
    const int i = i0 + tid;
    bool real;
    const bool in = i < n;
    const unsigned key = in ? float_key(src.fetch(r, i, real)) : 0u;
    const bool gt = in && key > kth;
    const bool eq = in && key == kth;
    if (gt) {
      const unsigned p = atomicAdd(&s_gt_count, 1u);
      s_sel[p] = ((unsigned long long)key << 32) | (unsigned)(~(unsigned)i);
    }
    const unsigned m = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int pre = eq_seen, all = 0;
#pragma unroll 8
    for (int w = 0; w < kTopkBlock / 32; ++w) {
      const int c = s_warp[w];
      pre += (w < warp) ? c : 0;
      all += c;
    }
    if (eq) {
      const int rank = pre + __popc(m & ((1u << lane) - 1u));
      if (rank < need) s_sel[n_gt + rank] = ((unsigned long long)key << 32) | (unsigned)(~(unsigned)i);
    }
    eq_seen += all;
    __syncthreads();
  }

  // ---------------- bitonic sort (descending) of the k survivors, padded with 0 (= lowest)
  int k2 = 1;
  while (k2 < k) k2 <<= 1;
  for (int i = k + tid; i < k2; i += kTopkBlock) s_sel[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= k2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (k2 >> 1); t += kTopkBlock) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s_sel[lo], b = s_sel[hi];
        if ((a < b) == desc) { s_sel[lo] = b; s_sel[hi] = a; }
      }
      __syncthreads();
    }
  }

  // ---------------- emit
  for (int j = tid; j < k; j += kTopkBlock) {
    const int i = (int)(~(unsigned)(s_sel[j] & 0xffffffffull));
    bool real;
    const float s = src.fetch(r, i, real);
    out_scores[r * k + j] = s;
    if (out_idx) out_idx[r * k + j] = real ? i : (i | (int)0x80000000);
    if (out_boxes) st4(out_boxes + 4 * (r * k + j), ldg4(gather_boxes + 4 * (r * n + i)));
  }
  }  // rows
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
