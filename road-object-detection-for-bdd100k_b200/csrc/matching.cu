// f-2: evaluation TP / FP matching of detections against ground truth.
// Replaces tfe.bboxes_matching / bboxes_matching_batch (utils/tf_extended/bboxes.py:246-380):
// for every detection, in the given (score-sorted) order: IoU (tfe.bboxes_jaccard, safe_divide)
// against the image's GT boxes of the same class, first-argmax, match iff IoU > threshold,
// TP iff matched & GT not matched before & not difficult, FP iff not difficult & (no match or GT
// already matched); matched GT boxes are remembered.
// One warp per (image) row of one class: the detections are inherently sequential, lanes split the
// GT boxes (kept in shared memory), the argmax is a warp reduction with lowest-index tie-break.
#include "common.cuh"

namespace rod {

constexpr int kMatchWarps = 4;

template <typename LabelT>
__global__ void __launch_bounds__(32 * kMatchWarps)
matching_kernel(long long label, const float* __restrict__ scores, const float* __restrict__ bboxes,
                const LabelT* __restrict__ glabels, const float* __restrict__ gbboxes,
                const LabelT* __restrict__ gdifficults, int rows, int n, int g_n, float thr,
                long long* __restrict__ out_n, unsigned char* __restrict__ out_tp,
                unsigned char* __restrict__ out_fp) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kMatchWarps + warp;
  // per warp: GT boxes, areas, flags (bit0 same class, bit1 difficult, bit2 matched)
  float4* s_g = reinterpret_cast<float4*>(s_raw) + (size_t)warp * g_n;
  float* s_ga = reinterpret_cast<float*>(reinterpret_cast<float4*>(s_raw) + (size_t)kMatchWarps * g_n) + (size_t)warp * g_n;
  unsigned char* s_fl = reinterpret_cast<unsigned char*>(reinterpret_cast<float*>(reinterpret_cast<float4*>(s_raw) +
                        (size_t)kMatchWarps * g_n) + (size_t)kMatchWarps * g_n) + (size_t)warp * g_n;
  if (r >= rows) return;
  (void)scores;
  int n_gb = 0;
  for (int g = lane; g < g_n; g += 32) {
    const float4 b = ldg4(gbboxes + 4ll * ((long long)r * g_n + g));
    const bool same = (long long)glabels[(long long)r * g_n + g] == label;      // :274, :293
    const bool diff = gdifficults[(long long)r * g_n + g] != 0;                  // :273 cast to bool
    s_g[g] = b;
    s_ga[g] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    s_fl[g] = (unsigned char)((same ? 1 : 0) | (diff ? 2 : 0));
    n_gb += (same && !diff) ? 1 : 0;
  }
  n_gb = __reduce_add_sync(0xffffffffu, n_gb);
  if (lane == 0) out_n[r] = n_gb;                                               // :274-275
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const float4 d = ldg4(bboxes + 4ll * ((long long)r * n + i));
    const float ad = __fmul_rn(__fsub_rn(d.z, d.x), __fsub_rn(d.w, d.y));
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int g = lane; g < g_n; g += 32) {
      const float4 b = s_g[g];
      // tfe.bboxes_jaccard(bbox_ref = detection, bboxes = GT), utils/tf_extended/bboxes.py:467-478
      const float h = fmaxf(__fsub_rn(fminf(b.z, d.z), fmaxf(b.x, d.x)), 0.f);
      const float w = fmaxf(__fsub_rn(fminf(b.w, d.w), fmaxf(b.y, d.y)), 0.f);
      const float inter = __fmul_rn(h, w);
      const float uni = __fadd_rn(__fadd_rn(-inter, s_ga[g]), ad);
      float j = uni > 0.f ? __fdiv_rn(inter, uni) : 0.f;
      j = __fmul_rn(j, (s_fl[g] & 1) ? 1.f : 0.f);                                // :293
      if (j > best) { best = j; bi = g; }                                        // first max within the lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                                           // first max across lanes
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      bool tp = false, fp = false;
      if (g_n > 0) {
        const unsigned char fl = s_fl[bi];
        const bool match = best > thr;                                           // :298
        const bool existing = fl & 4, not_diff = !(fl & 2);
        tp = not_diff && match && !existing;                                     // :304-305
        fp = not_diff && (existing || !match);                                   // :307-308
        if (not_diff && match) s_fl[bi] = fl | 4;                                // :311-313
      }
      out_tp[(long long)r * n + i] = tp;
      out_fp[(long long)r * n + i] = fp;
    }
    __syncwarp();
  }
}

}  // namespace rod

extern "C" int rod_bboxes_matching_batch(int64_t label, const float* scores, const float* bboxes, const void* glabels,
                                         const float* gbboxes, const void* gdifficults, int labels_i64, int rows,
                                         int n, int g_n, float matching_threshold, int64_t* out_n_gbboxes,
                                         uint8_t* out_tp, uint8_t* out_fp, void* stream) {
  using namespace rod;
  ROD_REQUIRE(bboxes && glabels && gbboxes && gdifficults && out_n_gbboxes && out_tp && out_fp,
              "rod_bboxes_matching_batch: NULL pointer argument");
  ROD_REQUIRE(rows >= 0 && n >= 0 && g_n >= 0, "rod_bboxes_matching_batch: negative size");
  if (rows == 0) return ROD_OK;
  const size_t smem = (size_t)kMatchWarps * g_n * (sizeof(float4) + sizeof(float) + 1) + 16;
  ROD_REQUIRE(smem <= 200 * 1024, "rod_bboxes_matching_batch: %d ground-truth boxes per image do not fit in shared memory", g_n);
  const unsigned grid = (unsigned)((rows + kMatchWarps - 1) / kMatchWarps);
  cudaStream_t st = (cudaStream_t)stream;
  if (labels_i64) {
    auto k = matching_kernel<long long>;
    if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 32 * kMatchWarps, smem, st>>>(label, scores, bboxes, (const long long*)glabels, gbboxes,
                                            (const long long*)gdifficults, rows, n, g_n, matching_threshold,
                                            (long long*)out_n_gbboxes, out_tp, out_fp);
  } else {
    auto k = matching_kernel<int>;
    if (smem > 48 * 1024) ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 32 * kMatchWarps, smem, st>>>(label, scores, bboxes, (const int*)glabels, gbboxes,
                                            (const int*)gdifficults, rows, n, g_n, matching_threshold,
                                            (long long*)out_n_gbboxes, out_tp, out_fp);
  }
  ROD_LAUNCH_CHECK("matching_kernel");
  return ROD_OK;
}
