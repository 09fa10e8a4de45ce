// Score sources and the exact segmented top-k used by rod_bboxes_sort and rod_detect.
#pragma once
#include "common.cuh"

namespace rod {

// Row r of a dense [rows, n] score matrix (tfe.bboxes_sort on tensors).
struct DenseScores {
  const float* scores;
  int n;
  __device__ __forceinline__ int size() const { return n; }
  __device__ __forceinline__ bool any_active() const { return true; }
  __device__ __forceinline__ bool row_active(long long) const { return true; }
  // returns the (masked) score; real == false marks entries whose box was zeroed by select
  __device__ __forceinline__ float fetch(long long r, int i, bool& real) const {
    real = true;
    return __ldg(scores + r * n + i);
  }
};

// Fused select (utils/net_tools.py:686-695) over the per-layer prediction list: row r =
// c * batch + b reads class column c of image b; scores = p * (p >= thr).
struct SelectedScores {
  LayeredF probs;
  Layout L;
  int n_classes, ignore_class, batch;
  int logits;                 // predictions are logits: softmax on the fly (same bits as the fused select)
  float thr;
  const unsigned* over_cnt;   // when set: only segments whose streaming list overflowed are active
  const unsigned* over_any;   // when set: non-zero iff any segment is flagged (lets the whole grid leave at once)
  unsigned over_cap;
  __device__ __forceinline__ int size() const { return L.n_total; }
  __device__ __forceinline__ bool any_active() const { return over_any == nullptr || *over_any != 0u; }
  __device__ __forceinline__ bool row_active(long long r) const {
    return (int)(r / batch) != ignore_class && (over_cnt == nullptr || over_cnt[r] > over_cap);
  }
  __device__ __forceinline__ float fetch(long long r, int i, bool& real) const {
    const int c = (int)(r / batch), b = (int)(r % batch);
    const int l = layer_of(L, i);
    const float* row = probs.base[l] + (long long)b * probs.stride[l] + (long long)(i - L.offset[l]) * n_classes;
    const float p = logits ? softmax_pick(row, n_classes, c) : __ldg(row + c);
    real = p >= thr;
    return __fmul_rn(p, real ? 1.f : 0.f);
  }
};

// Exact top-k of row r (tf.nn.top_k order: descending, equal scores -> lower index first), all BLOCK threads of the CTA.
// 4 x 8-bit MSB-first radix select on order-preserving uint32 keys (warp-aggregated shared-memory histograms), then an
// index-ordered compaction that takes every key above the k-th and the FIRST `need` keys equal to it, then a bitonic
// sort of the k survivors on the 64-bit composite (key, ~index).  Every tie rule is decided by index, never by thread
// scheduling.  Starts and ends with a block barrier.
template <typename Src, int BLOCK>
__device__ __forceinline__ void topk_row(const Src& src, long long r, int k, float* __restrict__ out_scores,
                                         int32_t* __restrict__ out_idx, const float* __restrict__ gather_boxes,
                                         float* __restrict__ out_boxes) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned long long s_sel[ROD_MAX_TOPK];
  __shared__ int s_warp[BLOCK / 32];
  __shared__ unsigned s_bin, s_above, s_gt_count;
  const int n = src.size();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();

  // ---------------- radix select of the k-th largest key
  unsigned prefix = 0, pmask = 0;
  int need = k;                       // how many still to take among keys matching the prefix
  unsigned eq_total = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += BLOCK) {
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
  for (int i0 = 0; i0 < n; i0 += BLOCK) {
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
    for (int w = 0; w < BLOCK / 32; ++w) {
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
  for (int i = k + tid; i < k2; i += BLOCK) s_sel[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= k2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (k2 >> 1); t += BLOCK) {
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
  for (int j = tid; j < k; j += BLOCK) {
    const int i = (int)(~(unsigned)(s_sel[j] & 0xffffffffull));
    bool real;
    const float s = src.fetch(r, i, real);
    out_scores[r * k + j] = s;
    if (out_idx) out_idx[r * k + j] = real ? i : (i | (int)0x80000000);
    if (out_boxes) st4(out_boxes + 4 * (r * k + j), ldg4(gather_boxes + 4 * (r * n + i)));
  }
  __syncthreads();
}

}  // namespace rod
