// f-2 (second half): evaluation metrics behind evaluate.py:162-197 —
//   tfe.streaming_tp_fp_arrays   utils/tf_extended/metrics.py:133-204  -> rod_tpfp_append
//   tfe.precision_recall         :100-130 (after the score sort)       -> rod_precision_recall
//   tfe.average_precision_voc07 / _voc12  :210-258                     -> rod_average_precision
// The score sort itself (tf.nn.top_k over all accumulated detections) is done by the host layer.
#include "common.cuh"

namespace rod {

constexpr int kMetBlock = 1024;

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  __syncthreads();                                   // s_warp may still be read from the previous call
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int wsum = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0, winc = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += x;
  }
  const int wbase = __shfl_sync(0xffffffffu, winc - wsum, warp);
  *total = __shfl_sync(0xffffffffu, winc, 31);
  return wbase + inc - v;
}

// One CTA per row (class).  Appends, in order, the detections with (tp | fp) [and score > rm_threshold]
// to the row's accumulated arrays; adds the batch's ground-truth counts to v_nobjects.
__global__ void __launch_bounds__(kMetBlock)
tpfp_append_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ tp, const uint8_t* __restrict__ fp,
                   long long n, const long long* __restrict__ num_gbboxes, int n_gb, int remove_zero_scores,
                   float rm_threshold, long long id_base, float* __restrict__ v_scores, uint8_t* __restrict__ v_tp,
                   uint8_t* __restrict__ v_fp, long long* __restrict__ v_ids, long long capacity,
                   long long* __restrict__ v_count, long long* __restrict__ v_nobjects) {
  __shared__ int s_warp[32];
  __shared__ long long s_base;
  const long long r = blockIdx.x;
  scores += r * n; tp += r * n; fp += r * n;
  v_scores += r * capacity; v_tp += r * capacity; v_fp += r * capacity;
  if (v_ids) v_ids += r * capacity;
  if (threadIdx.x == 0) s_base = v_count[r];
  if (threadIdx.x < 32 && num_gbboxes != nullptr) {        // v_nobjects += reduce_sum(num_gbboxes)   (:176-177)
    long long acc = 0;
    for (int i = threadIdx.x; i < n_gb; i += 32) acc += num_gbboxes[r * n_gb + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) v_nobjects[r] += acc;
  }
  __syncthreads();
  long long base = s_base;
  for (long long i0 = 0; i0 < n; i0 += kMetBlock) {
    const long long i = i0 + threadIdx.x;
    bool t = false, f = false, keep = false;
    float s = 0.f;
    if (i < n) {
      t = tp[i] != 0; f = fp[i] != 0; s = scores[i];
      // mask = tp | fp, and (scores > 1e-4) — the reference filters only when remove_zero_scores (:164-170)
      keep = remove_zero_scores ? ((t || f) && s > rm_threshold) : true;
    }
    int total;
    const int off = block_exclusive_scan(keep ? 1 : 0, s_warp, &total);
    if (keep) {
      const long long pos = base + off;
      if (pos < capacity) {
        v_scores[pos] = s; v_tp[pos] = t; v_fp[pos] = f;
        if (v_ids) v_ids[pos] = id_base + i;
      }
    }
    base += total;
  }
  if (threadIdx.x == 0) v_count[r] = base;
}

// precision / recall of detections already sorted by descending score (:117-130):
// tp_c = cumsum(tp), fp_c = cumsum(fp) in float64 (exact integers), recall = tp_c / n_gb, precision = tp_c / (tp_c + fp_c),
// both 0 where the denominator is <= 0 (_safe_div).
constexpr int kPrPer = 4, kPrTile = kMetBlock * kPrPer;

__global__ void __launch_bounds__(kMetBlock)
pr_block_sums_kernel(const uint8_t* __restrict__ tp, const uint8_t* __restrict__ fp, long long n, int* __restrict__ sums) {
  __shared__ int s_t[32], s_f[32];
  const long long i0 = (long long)blockIdx.x * kPrTile;
  int t = 0, f = 0;
  for (int q = 0; q < kPrPer; ++q) {
    const long long i = i0 + q * kMetBlock + threadIdx.x;
    if (i < n) { t += tp[i] != 0; f += fp[i] != 0; }
  }
  t = __reduce_add_sync(0xffffffffu, t); f = __reduce_add_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0) { s_t[threadIdx.x >> 5] = t; s_f[threadIdx.x >> 5] = f; }
  __syncthreads();
  if (threadIdx.x < 32) {
    t = __reduce_add_sync(0xffffffffu, s_t[threadIdx.x]); f = __reduce_add_sync(0xffffffffu, s_f[threadIdx.x]);
    if (threadIdx.x == 0) { sums[2 * blockIdx.x] = t; sums[2 * blockIdx.x + 1] = f; }
  }
}

__global__ void __launch_bounds__(kMetBlock)
pr_apply_kernel(const uint8_t* __restrict__ tp, const uint8_t* __restrict__ fp, long long n, const int* __restrict__ sums,
                const long long* __restrict__ num_gbboxes, double* __restrict__ precision, double* __restrict__ recall) {
  __shared__ int s_warp[32];
  __shared__ long long s_pt, s_pf;
  // prefix of the preceding tiles
  long long pt = 0, pf = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kMetBlock) { pt += sums[2 * b]; pf += sums[2 * b + 1]; }
  if (threadIdx.x == 0) { s_pt = 0; s_pf = 0; }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { pt += __shfl_xor_sync(0xffffffffu, pt, o); pf += __shfl_xor_sync(0xffffffffu, pf, o); }
  if ((threadIdx.x & 31) == 0 && (pt | pf)) { atomicAdd((unsigned long long*)&s_pt, (unsigned long long)pt); atomicAdd((unsigned long long*)&s_pf, (unsigned long long)pf); }
  __syncthreads();
  long long base_t = s_pt, base_f = s_pf;
  const double ngb = (double)num_gbboxes[0];
  const long long i0 = (long long)blockIdx.x * kPrTile;
  for (int q = 0; q < kPrPer; ++q) {
    const long long i = i0 + q * kMetBlock + threadIdx.x;
    const int t = (i < n && tp[i] != 0) ? 1 : 0, f = (i < n && fp[i] != 0) ? 1 : 0;
    int tot_t, tot_f;
    const int et = block_exclusive_scan(t, s_warp, &tot_t);
    const int ef = block_exclusive_scan(f, s_warp, &tot_f);
    if (i < n) {
      const double ct = (double)(base_t + et + t), cf = (double)(base_f + ef + f);
      recall[i] = ngb > 0.0 ? ct / ngb : 0.0;
      precision[i] = (ct + cf) > 0.0 ? ct / (ct + cf) : 0.0;
    }
    base_t += tot_t; base_f += tot_f;
  }
}

struct ApThresholds { double t[11]; };

// Single CTA.  M_i = max(p_i .. p_{n-1}, 0) (reverse cummax with the appended 0).
// voc12 (:210-232): sum_i M_i * (r_i - r_{i-1}), r_{-1} = 0  (the final [0 * (1 - r_{n-1})] term is 0).
// voc07 (:235-258): sum_t max{p_i : r_i >= t, or the appended 0} / 11, accumulated in threshold order.
__global__ void __launch_bounds__(kMetBlock)
average_precision_kernel(const double* __restrict__ precision, const double* __restrict__ recall, long long n,
                         const ApThresholds T, double* __restrict__ out) {
  __shared__ double s_wmax[32], s_red[32], s_v07[32][11];
  __shared__ double s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc12 = 0.0, v07[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) v07[k] = 0.0;
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  const long long ntile = (n + kMetBlock - 1) / kMetBlock;
  for (long long tile = ntile - 1; tile >= 0; --tile) {
    const long long i = tile * kMetBlock + threadIdx.x;
    const bool in = i < n;
    const double p = in ? precision[i] : 0.0, r = in ? recall[i] : 0.0;
    const double rprev = (in && i > 0) ? recall[i - 1] : 0.0;
    // inclusive suffix max inside the tile: higher threads first
    double m = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double x = __shfl_down_sync(0xffffffffu, m, o);
      if (lane + o < 32) m = fmax(m, x);
    }
    if (lane == 0) s_wmax[warp] = m;
    __syncthreads();
    double above = s_carry;                            // max over all later tiles (and the appended 0)
    for (int w = warp + 1; w < 32; ++w) above = fmax(above, s_wmax[w]);
    m = fmax(m, above);
    if (in) {
      acc12 += m * (r - rprev);
#pragma unroll
      for (int k = 0; k < 11; ++k)
        if (r >= T.t[k]) v07[k] = fmax(v07[k], p);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = fmax(s_carry, fmax(s_wmax[0], above));   // thread 0's m covers the whole tile
    __syncthreads();
  }
  // deterministic block reductions (fixed tree)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc12 += __shfl_xor_sync(0xffffffffu, acc12, o);
#pragma unroll
  for (int k = 0; k < 11; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v07[k] = fmax(v07[k], __shfl_xor_sync(0xffffffffu, v07[k], o));
  }
  if (lane == 0) {
    s_red[warp] = acc12;
    for (int k = 0; k < 11; ++k) s_v07[warp][k] = v07[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a12 = 0.0;
    for (int w = 0; w < 32; ++w) a12 += s_red[w];
    double a07 = 0.0;
    for (int k = 0; k < 11; ++k) {
      double v = 0.0;
      for (int w = 0; w < 32; ++w) v = fmax(v, s_v07[w][k]);
      a07 = k == 0 ? v / 11.0 : a07 + v / 11.0;      // tf.add_n(l_aps): in order
    }
    out[0] = a07;
    out[1] = a12;
  }
}

}  // namespace rod

extern "C" int rod_tpfp_append(const float* scores, const uint8_t* tp, const uint8_t* fp, int rows, int64_t n,
                               const int64_t* num_gbboxes, int n_gb, int remove_zero_scores, float rm_threshold,
                               int64_t id_base, float* v_scores, uint8_t* v_tp, uint8_t* v_fp, int64_t* v_ids,
                               int64_t capacity, int64_t* v_count, int64_t* v_nobjects, void* stream) {
  using namespace rod;
  ROD_REQUIRE(rows >= 0 && n >= 0 && capacity >= 0 && n_gb >= 0, "rod_tpfp_append: rows=%d n=%lld capacity=%lld invalid", rows,
              (long long)n, (long long)capacity);
  if (rows == 0) return ROD_OK;
  ROD_REQUIRE(v_count && v_nobjects && (n == 0 || (scores && tp && fp && v_scores && v_tp && v_fp)),
              "rod_tpfp_append: NULL pointer argument");
  tpfp_append_kernel<<<rows, kMetBlock, 0, (cudaStream_t)stream>>>(
      scores, tp, fp, n, reinterpret_cast<const long long*>(num_gbboxes), n_gb, remove_zero_scores, rm_threshold, id_base,
      v_scores, v_tp, v_fp, reinterpret_cast<long long*>(v_ids), capacity, reinterpret_cast<long long*>(v_count),
      reinterpret_cast<long long*>(v_nobjects));
  ROD_LAUNCH_CHECK("tpfp_append_kernel");
  return ROD_OK;
}

extern "C" size_t rod_precision_recall_workspace_bytes(int64_t n) {
  return (size_t)((n + rod::kPrTile - 1) / rod::kPrTile + 1) * 2 * sizeof(int);
}

extern "C" int rod_precision_recall(const uint8_t* tp_sorted, const uint8_t* fp_sorted, int64_t n,
                                    const int64_t* num_gbboxes, double* precision, double* recall, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  using namespace rod;
  ROD_REQUIRE(n >= 0 && n < (1ll << 31) * kPrTile, "rod_precision_recall: n=%lld invalid", (long long)n);
  if (n == 0) return ROD_OK;
  ROD_REQUIRE(tp_sorted && fp_sorted && num_gbboxes && precision && recall && workspace, "rod_precision_recall: NULL pointer argument");
  ROD_REQUIRE(workspace_bytes >= rod_precision_recall_workspace_bytes(n), "rod_precision_recall: workspace too small");
  const unsigned nblk = (unsigned)((n + kPrTile - 1) / kPrTile);
  int* sums = reinterpret_cast<int*>(workspace);
  pr_block_sums_kernel<<<nblk, kMetBlock, 0, (cudaStream_t)stream>>>(tp_sorted, fp_sorted, n, sums);
  ROD_LAUNCH_CHECK("pr_block_sums_kernel");
  pr_apply_kernel<<<nblk, kMetBlock, 0, (cudaStream_t)stream>>>(tp_sorted, fp_sorted, n, sums,
                                                                  reinterpret_cast<const long long*>(num_gbboxes), precision, recall);
  ROD_LAUNCH_CHECK("pr_apply_kernel");
  return ROD_OK;
}

extern "C" int rod_average_precision(const double* precision, const double* recall, int64_t n, const double* thresholds07,
                                     double* out_voc07_voc12, void* stream) {
  using namespace rod;
  ROD_REQUIRE(n >= 0 && out_voc07_voc12 && thresholds07 && (n == 0 || (precision && recall)),
              "rod_average_precision: invalid argument");
  ApThresholds T;
  for (int k = 0; k < 11; ++k) T.t[k] = thresholds07[k];
  average_precision_kernel<<<1, kMetBlock, 0, (cudaStream_t)stream>>>(precision, recall, n, T, out_voc07_voc12);
  ROD_LAUNCH_CHECK("average_precision_kernel");
  return ROD_OK;
}
