// f-2: the descending score sort in front of precision / recall.  tfe.precision_recall sorts the accumulated
// detections with tf.nn.top_k(scores, k = num_detections, sorted=True) and gathers tp / fp in that order
// (utils/tf_extended/metrics.py:117-123): descending score, equal scores in index order.
//
// One 64-bit composite key per detection, (~order_key(score) << 32) | index, makes every key distinct, so an
// ascending bitonic sort needs no stability argument: tiles of 2048 keys are sorted / merged in shared memory,
// the strides above a tile are plain coalesced compare-exchange passes over global memory (L2-resident: the
// accumulated arrays of an evaluation are a few MB).  The gather of tp / fp / scores is fused into the last pass.
#include "common.cuh"

namespace rod {

constexpr int kSortTile = 2048;      // keys per CTA in the shared-memory passes
constexpr int kSortBlock = 512;

__global__ void __launch_bounds__(256)
sort_build_kernel(const float* __restrict__ scores, long long n, long long n_pad, unsigned long long* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_pad) return;
  keys[i] = i < n ? (((unsigned long long)(~float_key(scores[i])) << 32) | (unsigned)i) : ~0ull;   // padding sorts last
}

__device__ __forceinline__ void cmpx(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

// all compare-exchange steps with stride < kSortTile of the stages size_lo .. size_hi (powers of two), in smem.
// size_lo == 2: the initial tile sort; otherwise size_lo == size_hi == the global stage whose small strides remain.
__global__ void __launch_bounds__(kSortBlock)
sort_tile_kernel(unsigned long long* __restrict__ keys, long long size_lo, long long size_hi) {
  __shared__ unsigned long long s[kSortTile];
  const long long base = (long long)blockIdx.x * kSortTile;
  for (int i = threadIdx.x; i < kSortTile; i += kSortBlock) s[i] = keys[base + i];
  __syncthreads();
  for (long long size = size_lo; size <= size_hi; size <<= 1) {
    for (int stride = (int)((size >> 1) < kSortTile ? (size >> 1) : (kSortTile >> 1)); stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kSortTile / 2; t += kSortBlock) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (((base + lo) & size) == 0);       // ascending half of the bitonic stage
        cmpx(s[lo], s[hi], up);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kSortTile; i += kSortBlock) keys[base + i] = s[i];
}

// one compare-exchange step of stage `size` with stride >= kSortTile over global memory
__global__ void __launch_bounds__(256)
sort_global_kernel(unsigned long long* __restrict__ keys, long long n_pad, long long size, long long stride) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n_pad / 2) return;
  const long long lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
  unsigned long long a = keys[lo], b = keys[hi];
  const bool up = ((lo & size) == 0);
  if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
}

__global__ void __launch_bounds__(256)
sort_emit_kernel(const unsigned long long* __restrict__ keys, long long k, const float* __restrict__ scores,
                 const uint8_t* __restrict__ tp, const uint8_t* __restrict__ fp, uint8_t* __restrict__ tp_s,
                 uint8_t* __restrict__ fp_s, float* __restrict__ scores_s, int32_t* __restrict__ idx_s) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= k) return;
  const unsigned j = (unsigned)(keys[i] & 0xffffffffull);
  if (idx_s) idx_s[i] = (int32_t)j;
  if (tp_s) tp_s[i] = tp[j];
  if (fp_s) fp_s[i] = fp[j];
  if (scores_s) scores_s[i] = scores[j];
}

static long long sort_padded(long long n) {
  long long p = kSortTile;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace rod

extern "C" size_t rod_sort_scores_workspace_bytes(int64_t n) {
  return n <= 0 ? 256 : (size_t)rod::sort_padded(n) * 8 + 256;
}

extern "C" int rod_sort_scores_desc(const float* scores, int64_t n, int64_t k, const uint8_t* tp, const uint8_t* fp,
                                    uint8_t* tp_sorted, uint8_t* fp_sorted, float* scores_sorted, int32_t* idx_sorted,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rod;
  ROD_REQUIRE(n >= 0 && k >= 0 && k <= n, "rod_sort_scores_desc: k=%lld must be in [0, n=%lld] (tf.nn.top_k)", (long long)k, (long long)n);
  ROD_REQUIRE(n < 2147483647ll, "rod_sort_scores_desc: n=%lld too large", (long long)n);
  if (k == 0) return ROD_OK;
  ROD_REQUIRE(scores && workspace && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "rod_sort_scores_desc: NULL / misaligned pointer");
  ROD_REQUIRE(workspace_bytes >= rod_sort_scores_workspace_bytes(n), "rod_sort_scores_desc: workspace too small");
  ROD_REQUIRE((tp_sorted == nullptr || tp != nullptr) && (fp_sorted == nullptr || fp != nullptr), "rod_sort_scores_desc: tp / fp missing");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = static_cast<unsigned long long*>(workspace);
  const long long n_pad = sort_padded(n);
  sort_build_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(scores, n, n_pad, keys);
  ROD_LAUNCH_CHECK("sort_build_kernel");
  const unsigned tiles = (unsigned)(n_pad / kSortTile);
  sort_tile_kernel<<<tiles, kSortBlock, 0, st>>>(keys, 2, kSortTile);
  ROD_LAUNCH_CHECK("sort_tile_kernel");
  for (long long size = 2ll * kSortTile; size <= n_pad; size <<= 1) {
    for (long long stride = size >> 1; stride >= kSortTile; stride >>= 1) {
      sort_global_kernel<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, st>>>(keys, n_pad, size, stride);
      ROD_LAUNCH_CHECK("sort_global_kernel");
    }
    sort_tile_kernel<<<tiles, kSortBlock, 0, st>>>(keys, size, size);
    ROD_LAUNCH_CHECK("sort_tile_kernel");
  }
  sort_emit_kernel<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(keys, k, scores, tp, fp, tp_sorted, fp_sorted, scores_sorted, idx_sorted);
  ROD_LAUNCH_CHECK("sort_emit_kernel");
  return ROD_OK;
}
