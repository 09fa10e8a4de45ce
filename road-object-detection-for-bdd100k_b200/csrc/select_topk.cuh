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

}  // namespace rod
