// f-3: the losses that consume the ARM / ODM targets (utils/net_tools.py:478-623):
//   smooth_l1 :478-489, refine_loss :492-516 and the det_loss half of det_clf_loss :538-551
//                                                                -> rod_smooth_l1_loss
//   clf_loss half of det_clf_loss :553-615 (softmax, hard-negative mining by the k-th smallest
//   background probability, IoU-factor weighting, two cross entropies) -> rod_clf_loss (+ _grad)
// Floating-point parity with the reference is by tolerance (1e-5 relative): TF's float32 reductions
// have their own order and its exp / log are TF kernels; sums are accumulated here in float64 in a
// fixed order (deterministic run to run).  Masks, labels and the IoU factor are constants of the
// backward pass (gradients w.r.t. the head outputs only).
#include "common.cuh"

namespace rod {

constexpr int kLossBlock = 256;

// deterministic block reduction of a double: fixed shuffle tree + fixed order over warps; result in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_w) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
  return t;
}

// sum of n partials by one warp in a fixed order (n is a few hundred: one partial per persistent block)
__device__ __forceinline__ double warp_sum_partials(const double* p, int n) {
  double t = 0.0;
  for (int i = threadIdx.x & 31; i < n; i += 32) t += __ldcg(p + i);          // (written by other blocks of the same kernel in the fused epilogues)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// ------------------------------------------------------------------------------------------
// smooth-L1 over masked offsets:  sum_{b,n,k} smooth_l1((y - x) * mask)        (:478-489, :505-512, :543-550)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLossBlock)
smooth_l1_kernel(const __grid_constant__ Layout L, const __grid_constant__ LayeredF y, const __grid_constant__ LayeredF x,
                 const __grid_constant__ LayeredI mask, int batch, double* __restrict__ partial, float* __restrict__ grad_x,
                 float grad_scale) {
  __shared__ double s_w[kLossBlock / 32];
  const int tiles_per_image = (L.n_total + kLossBlock - 1) / kLossBlock;
  double acc = 0.0;
  for (int t = blockIdx.x; t < tiles_per_image * batch; t += gridDim.x) {      // persistent: one partial per block
    const int b = t / tiles_per_image, n = (t - b * tiles_per_image) * kLossBlock + threadIdx.x;
    if (n >= L.n_total) continue;
    const int l = layer_of(L, n);
    const long long i = n - L.offset[l];
    const float4 yv = ldg4(y.base[l] + (long long)b * y.stride[l] + 4 * i);
    const float4 xv = ldg4(x.base[l] + (long long)b * x.stride[l] + 4 * i);
    const float m = (float)mask.base[l][(long long)b * mask.stride[l] + i];           // tf.cast(mask, dtype)
    const float z[4] = {__fmul_rn(__fsub_rn(yv.x, xv.x), m), __fmul_rn(__fsub_rn(yv.y, xv.y), m),
                        __fmul_rn(__fsub_rn(yv.z, xv.z), m), __fmul_rn(__fsub_rn(yv.w, xv.w), m)};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = fabsf(z[k]), mn = fminf(a, 1.f);
      acc += (double)__fmul_rn(0.5f, __fadd_rn(__fmul_rn(__fsub_rn(a, 1.f), mn), a));   // 0.5 * ((|z| - 1) * min(|z|, 1) + |z|)
      g[k] = -grad_scale * m * fminf(fmaxf(z[k], -1.f), 1.f);                          // d/dx: -m * clamp(z, -1, 1)
    }
    if (grad_x) st4(grad_x + 4 * ((long long)b * L.n_total + n), make_float4(g[0], g[1], g[2], g[3]));
  }
  const double tot = block_sum(acc, s_w);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(32)
smooth_l1_final_kernel(const double* __restrict__ partial, int n, float inv_bs, float* __restrict__ out) {
  const double t = warp_sum_partials(partial, n);
  if (threadIdx.x == 0) out[0] = (float)(t * (double)inv_bs);
}

// ------------------------------------------------------------------------------------------
// classification loss
// ------------------------------------------------------------------------------------------
struct ClfState {            // device scalars shared by the kernels of one rod_clf_loss call
  unsigned prefix, pmask;    // radix select state: bits of the k-th smallest value found so far
  int k_rem;                 // rank still to resolve inside the current prefix
  int n_pos, n_neg_total, n_neg;
  float max_hard_pred;
  unsigned ticket;           // blocks of the running kernel that have delivered their partial results (0 between kernels)
  unsigned hist[256];
};

// "Last block finishes the job": every block calls this after its global writes; exactly one block — the last to
// arrive — gets true, with the writes of all other blocks visible to loads that bypass L1 (__ldcg).  That block runs
// the few-thread epilogue (plan, radix pick, final sum) that used to be a <<<1, 32>>> launch of its own, and leaves the
// ticket at 0 for the next kernel.  No block ever waits for another one.
__device__ __forceinline__ bool last_block_arrives(unsigned* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned n = gridDim.x * gridDim.y;
    const bool last = atomicAdd(ticket, 1u) == n - 1u;
    if (last) *ticket = 0u;
    s_last = last ? 1 : 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// per (layer, image): statistics of the IoU map for the IoU factor (:590-600)
//   z = (iou - mean) / sqrt(var + 1e-8);  z += 0 - min(z);  z /= max(z) + 1e-8;  factor = z^4
__global__ void __launch_bounds__(kLossBlock)
iou_stats_kernel(const __grid_constant__ Layout L, const __grid_constant__ LayeredF iou, float4* __restrict__ stats,
                 ClfState* __restrict__ st) {
  __shared__ double s_w[kLossBlock / 32];
  if (blockIdx.x == 0 && blockIdx.y == 0) {            // first kernel of a rod_clf_loss call: select histogram and ticket start at 0
    st->hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) st->ticket = 0u;
  }
  static_assert(kLossBlock == 256, "one histogram bin per thread");
  __shared__ float s_mean, s_min[kLossBlock / 32], s_max[kLossBlock / 32];
  const int l = blockIdx.x, b = blockIdx.y, n = L.offset[l + 1] - L.offset[l];
  const float* p = iou.base[l] + (long long)b * iou.stride[l];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kLossBlock) acc += (double)p[i];
  const double tot = block_sum(acc, s_w);
  if (threadIdx.x == 0) s_mean = (float)(tot / (double)n);
  __syncthreads();
  const float mean = s_mean;
  acc = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += kLossBlock) {
    const float v = p[i], d = __fsub_rn(v, mean);
    acc += (double)__fmul_rn(d, d);
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
  const double vs = block_sum(acc, s_w);               // (syncs: s_min / s_max are visible afterwards)
  if (threadIdx.x == 0) {
    for (int w = 1; w < kLossBlock / 32; ++w) { mn = fminf(mn, s_min[w]); mx = fmaxf(mx, s_max[w]); }
    const float var = (float)(vs / (double)n);
    const float sd = sqrtf(__fadd_rn(var, 1e-8f));
    const float zmin = __fdiv_rn(__fsub_rn(mn, mean), sd), zmax = __fdiv_rn(__fsub_rn(mx, mean), sd);
    const float shift = __fsub_rn(0.f, zmin);
    const float denom = __fadd_rn(__fadd_rn(zmax, shift), 1e-8f);
    stats[(size_t)b * L.n_layers + l] = make_float4(mean, sd, shift, denom);
  }
}

__device__ __forceinline__ float iou_factor(float v, float4 st) {
  float z = __fdiv_rn(__fsub_rn(v, st.x), st.y);
  z = __fdiv_rn(__fadd_rn(z, st.z), st.w);
  const float z2 = __fmul_rn(z, z);
  return __fmul_rn(z2, z2);                             // tf.pow(z, 4)
}

struct ClfParams {
  Layout L;
  LayeredF logits, iou;
  LayeredI labels, mask;
  int batch, C;
};

// Copies the C logits of the block's anchors [n0, n0 + kLossBlock) of image b into shared memory with
// coalesced loads (one contiguous chunk per layer the block touches); thread t's row is s_tile + t * C
// (stride C words: conflict-free for odd C).  Ends with a block barrier.
__device__ __forceinline__ void stage_rows(const ClfParams& P, int b, int n0, float* s_tile) {
  const int n1 = min(n0 + kLossBlock, P.L.n_total);
  for (int l = 0; l < P.L.n_layers; ++l) {
    const int lo = max(n0, P.L.offset[l]), hi = min(n1, P.L.offset[l + 1]);
    if (lo >= hi) continue;
    const float* src = P.logits.base[l] + (long long)b * P.logits.stride[l] + (long long)(lo - P.L.offset[l]) * P.C;
    float* dst = s_tile + (lo - n0) * P.C;
    const int len = (hi - lo) * P.C;
    for (int i = threadIdx.x; i < len; i += kLossBlock) dst[i] = __ldg(src + i);
  }
  __syncthreads();
}

// softmax pieces of one anchor: max, sum of exp(x - max), log of it
template <typename F>
__device__ __forceinline__ void row_softmax(const float* __restrict__ row, int C, float& m, float& s, F&& each) {
  m = row[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
  s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float e = softmax_exp(row[c], m);
    s = c == 0 ? e : __fadd_rn(s, e);
    each(c, e);
  }
}

// n_positives, number of negatives to keep (:564, :577-581), radix-select start state.  One warp (of the last block of
// clf_pass1_kernel); returns the rank to select (= n_neg) in every lane.
__device__ __forceinline__ int clf_plan(const int* partial_npos, int nparts, long long total, int batch, float negative_ratio,
                                        ClfState* st) {
  int c = 0;
  for (int i = threadIdx.x; i < nparts; i += 32) c += __ldcg(partial_npos + i);
  c = __reduce_add_sync(0xffffffffu, c);
  const int max_neg = (int)(total - c);
  int n_neg = (int)(negative_ratio * (float)c) + batch;                  // tf.cast(3. * n_positives, int32) + bs
  n_neg = n_neg < max_neg ? n_neg : max_neg;
  if (threadIdx.x == 0) {
    st->n_pos = c; st->n_neg_total = max_neg; st->n_neg = n_neg;
    st->max_hard_pred = 0.f;
  }
  return n_neg;
}

// One warp: given the complete histogram of the current byte among the values matching (prefix, pmask), extend the
// prefix by the bin that holds the k-th smallest (1-based; k == 0: nothing to select), store the new select state
// and clear the histogram for the next pass.  lane j owns bins [8j, 8j+8) in ascending order.
__device__ __forceinline__ void radix_pick(int shift, int k_in, unsigned prefix, unsigned pmask, ClfState* st) {
  unsigned loc[8], sum = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { loc[q] = __ldcg(&st->hist[8 * threadIdx.x + q]); sum += loc[q]; }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned x = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)threadIdx.x >= o) inc += x;
  }
  const unsigned before = inc - sum;
  const unsigned k = (unsigned)k_in;
  const bool own = k >= 1 && before < k && inc >= k;     // at most one lane: the k-th smallest lies in my bins
  int k_rem = k_in;
  if (own) {
    unsigned acc = before;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (acc + loc[q] >= k) {
        prefix |= (unsigned)(8 * threadIdx.x + q) << shift;
        pmask |= 255u << shift;
        k_rem = (int)(k - acc);
        break;
      }
      acc += loc[q];
    }
  }
  const unsigned owners = __ballot_sync(0xffffffffu, own);
  if (own || (owners == 0u && threadIdx.x == 0)) {       // (no owner — k == 0: the state stays as it was)
    st->prefix = prefix; st->pmask = pmask; st->k_rem = k_rem;
    if (shift == 0 && own) st->max_hard_pred = __uint_as_float(prefix);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) st->hist[8 * threadIdx.x + q] = 0u;
}

// pass 1: per anchor cross entropies and the hard-negative-mining value (:566-575, :602-612)
__global__ void __launch_bounds__(kLossBlock)
clf_pass1_kernel(const __grid_constant__ ClfParams P, const float4* __restrict__ stats, float* __restrict__ nvalues,
                 float* __restrict__ ce0, double* __restrict__ partial_pos, int* __restrict__ partial_npos, float negative_ratio,
                 ClfState* __restrict__ st) {
  __shared__ double s_w[kLossBlock / 32];
  __shared__ int s_c[kLossBlock / 32];
  __shared__ unsigned s_h[256];                          // first radix pass (top byte of the nvalues) rides along
  s_h[threadIdx.x] = 0u;
  __syncthreads();
  unsigned cur_bin = 0u, cur_cnt = 0u;                   // run-length in registers: nearly all values share one top byte
  auto count = [&](float v) {
    const unsigned bin = __float_as_uint(v) >> 24;
    if (bin != cur_bin) {
      if (cur_cnt) atomicAdd(&s_h[cur_bin], cur_cnt);
      cur_bin = bin; cur_cnt = 0u;
    }
    ++cur_cnt;
  };
  const int tiles_per_image = (P.L.n_total + kLossBlock - 1) / kLossBlock;
  double pos_acc = 0.0;
  int npos = 0;
  for (int t = blockIdx.x; t < tiles_per_image * P.batch; t += gridDim.x) {
    const int b = t / tiles_per_image, n = (t - b * tiles_per_image) * kLossBlock + threadIdx.x;
    if (n >= P.L.n_total) continue;
    const int l = layer_of(P.L, n);
    const long long i = n - P.L.offset[l];
    const float* row = P.logits.base[l] + (long long)b * P.logits.stride[l] + i * P.C;      // L1 serves the 44 B stride
    const int m_i = P.mask.base[l][(long long)b * P.mask.stride[l] + i];
    const int lab = P.labels.base[l][(long long)b * P.labels.stride[l] + i];
    float mx, s, e0 = 0.f;
    row_softmax(row, P.C, mx, s, [&](int c, float e) { if (c == 0) e0 = e; });
    const float lse = (float)log((double)s);
    const long long o = (long long)b * P.L.n_total + n;
    ce0[o] = __fsub_rn(lse, __fsub_rn(row[0], mx));                       // CE against class 0 (:612)
    if (m_i != 0) {
      nvalues[o] = 1.f;                                                   // tf.where(nmask, p0, 1. - fnmask)
      count(1.f);
      // CE against the matched class (:605).  A label outside [0, C) gives NaN, like TF's GPU kernel of
      // sparse_softmax_cross_entropy_with_logits (its CPU kernel raises), instead of an out-of-bounds read.
      const float ce = (lab >= 0 && lab < P.C) ? __fsub_rn(lse, __fsub_rn(row[lab], mx)) : __int_as_float(0x7fc00000);
      const float f = iou_factor(P.iou.base[l][(long long)b * P.iou.stride[l] + i], stats[(size_t)b * P.L.n_layers + l]);
      pos_acc += (double)__fmul_rn(ce, f);
      npos += 1;
    } else {
      const float p0 = __fmul_rn(e0, __frcp_rn(s));                       // background probability
      nvalues[o] = p0;
      count(p0);
    }
  }
  if (cur_cnt) atomicAdd(&s_h[cur_bin], cur_cnt);
  const double tot = block_sum(pos_acc, s_w);
  npos = __reduce_add_sync(0xffffffffu, npos);
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = npos;
  __syncthreads();                                       // (also: every thread's histogram contribution is in s_h)
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < kLossBlock / 32; ++w) c += s_c[w];
    partial_pos[blockIdx.x] = tot;
    partial_npos[blockIdx.x] = c;
  }
  if (s_h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_h[threadIdx.x]);
  if (last_block_arrives(&st->ticket) && threadIdx.x < 32) {
    const int k = clf_plan(partial_npos, (int)gridDim.x, (long long)P.L.n_total * P.batch, P.batch, negative_ratio, st);
    radix_pick(24, k, 0u, 0u, st);
  }
}

// k-th smallest of the nvalues (all >= 0, so the float bits order like unsigned integers): 4 passes of an 8-bit radix
// histogram; the first rides along in clf_pass1_kernel, and the pick after each pass is done by the last block to
// deliver its counts (last_block_arrives) instead of by a launch of its own.
__global__ void __launch_bounds__(kLossBlock)
radix_hist_kernel(const float* __restrict__ v, long long total, int shift, ClfState* __restrict__ st) {
  __shared__ unsigned s_h[256];
  s_h[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned prefix = st->prefix, pmask = st->pmask;
  const int k = st->k_rem;
  for (long long i = (long long)blockIdx.x * kLossBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kLossBlock) {
    const unsigned key = __float_as_uint(v[i]);
    if ((key & pmask) == prefix) atomicAdd(&s_h[(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (s_h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_h[threadIdx.x]);
  if (last_block_arrives(&st->ticket) && threadIdx.x < 32) radix_pick(shift, k, prefix, pmask, st);
}

// pass 2: sum of the class-0 cross entropy over the mined negatives: nmask & (nvalues < max_hard_pred)  (:585-587, :612-613)
__global__ void __launch_bounds__(kLossBlock)
clf_pass2_kernel(const __grid_constant__ ClfParams P, const float* __restrict__ nvalues, const float* __restrict__ ce0,
                 ClfState* __restrict__ st, const double* __restrict__ partial_pos, double* __restrict__ partial_neg, float inv_bs,
                 float* __restrict__ out) {
  __shared__ double s_w[kLossBlock / 32];
  const int tiles_per_image = (P.L.n_total + kLossBlock - 1) / kLossBlock;
  const float thr = st->max_hard_pred;
  double acc = 0.0;
  for (int t = blockIdx.x; t < tiles_per_image * P.batch; t += gridDim.x) {
    const int b = t / tiles_per_image, n = (t - b * tiles_per_image) * kLossBlock + threadIdx.x;
    if (n >= P.L.n_total) continue;
    const int l = layer_of(P.L, n);
    const long long i = n - P.L.offset[l], o = (long long)b * P.L.n_total + n;
    const int m_i = P.mask.base[l][(long long)b * P.mask.stride[l] + i];
    if (m_i == 0 && nvalues[o] < thr) acc += (double)ce0[o];
  }
  const double tot = block_sum(acc, s_w);
  if (threadIdx.x == 0) partial_neg[blockIdx.x] = tot;
  // the last block to deliver its partial sums both losses up in the fixed order of the partials (deterministic)
  if (last_block_arrives(&st->ticket) && threadIdx.x < 32) {
    const double pos = warp_sum_partials(partial_pos, (int)gridDim.x), neg = warp_sum_partials(partial_neg, (int)gridDim.x);
    if (threadIdx.x == 0) {
      const float pos_loss = (float)(pos * (double)inv_bs), neg_loss = (float)(neg * (double)inv_bs);
      out[0] = __fadd_rn(__fdiv_rn(neg_loss, 2.f), pos_loss);             // clf_loss = neg_loss / 2. + pos_loss  (:615)
      out[1] = pos_loss;
      out[2] = neg_loss;
      out[3] = thr;
      out[4] = (float)st->n_pos;
      out[5] = (float)st->n_neg;
    }
  }
}

// backward: d clf_loss / d logits = w * (softmax - onehot(target)), w = factor / bs for positives (target = label),
// 0.5 / bs for the mined negatives (target = 0), 0 elsewhere
__global__ void __launch_bounds__(kLossBlock)
clf_grad_kernel(const __grid_constant__ ClfParams P, const float4* __restrict__ stats, const float* __restrict__ nvalues,
                const ClfState* __restrict__ st, float scale, float* __restrict__ grad) {
  extern __shared__ float s_tile[];
  const int b = blockIdx.y, n0 = blockIdx.x * kLossBlock, n = n0 + threadIdx.x;
  stage_rows(P, b, n0, s_tile);
  if (n < P.L.n_total) {
    const int l = layer_of(P.L, n);
    const long long i = n - P.L.offset[l], o = (long long)b * P.L.n_total + n;
    const int m_i = P.mask.base[l][(long long)b * P.mask.stride[l] + i];
    float w = 0.f;
    int target = 0;
    if (m_i != 0) {
      target = P.labels.base[l][(long long)b * P.labels.stride[l] + i];
      w = scale * iou_factor(P.iou.base[l][(long long)b * P.iou.stride[l] + i], stats[(size_t)b * P.L.n_layers + l]);
    } else if (nvalues[o] < st->max_hard_pred) {
      w = 0.5f * scale;
    }
    float* row = s_tile + threadIdx.x * P.C;             // gradient overwrites the logits in place
    if (target < 0 || target >= P.C) {                   // invalid label: NaN row, as the forward loss (NaN * 0 = NaN)
      for (int c = 0; c < P.C; ++c) row[c] = __int_as_float(0x7fc00000);
    } else if (w == 0.f) {
      for (int c = 0; c < P.C; ++c) row[c] = 0.f;
    } else {
      float mx, s;
      row_softmax(row, P.C, mx, s, [&](int, float) {});
      const float rinv = __frcp_rn(s);
      for (int c = 0; c < P.C; ++c) {
        const float p = __fmul_rn(softmax_exp(row[c], mx), rinv);
        row[c] = w * (p - (c == target ? 1.f : 0.f));
      }
    }
  }
  __syncthreads();
  // coalesced store of the tile: grad is flat [B, N, C]
  const int len = (min(n0 + kLossBlock, P.L.n_total) - n0) * P.C;
  float* dst = grad + ((long long)b * P.L.n_total + n0) * P.C;
  for (int i = threadIdx.x; i < len; i += kLossBlock) __stcs(dst + i, s_tile[i]);
}

static size_t al(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace rod

// persistent grids: one partial per block, a few hundred blocks
static int loss_grid(const rod_layout_t* layout, int batch) {
  const long long tiles = (long long)((layout->n_total + rod::kLossBlock - 1) / rod::kLossBlock) * batch;
  const long long cap = 4ll * rod::sm_count();
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

extern "C" size_t rod_smooth_l1_workspace_bytes(const rod_layout_t* layout, int batch) {
  if (!layout || batch <= 0) return 256;
  return rod::al((size_t)loss_grid(layout, batch) * sizeof(double)) + 256;
}

extern "C" int rod_smooth_l1_loss(const rod_layout_t* layout, const rod_layered_t* y, const rod_layered_t* x,
                                  const rod_layered_t* mask, int batch, float* out_loss, float* grad_x, float grad_scale,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rod;
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(y, nl, "y")) || (rc = check_layered(x, nl, "x")) || (rc = check_layered(mask, nl, "mask"))) return rc;
  ROD_REQUIRE(batch >= 1 && out_loss && workspace && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
              "rod_smooth_l1_loss: invalid argument (workspace must be 8-byte aligned)");
  ROD_REQUIRE(workspace_bytes >= rod_smooth_l1_workspace_bytes(layout, batch), "rod_smooth_l1_loss: workspace too small");
  const Layout L = to_layout(layout);
  const int grid = loss_grid(layout, batch);
  double* partial = reinterpret_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  smooth_l1_kernel<<<grid, kLossBlock, 0, st>>>(L, to_layered_f(y, nl), to_layered_f(x, nl), to_layered_i(mask, nl), batch, partial,
                                                 grad_x, grad_scale);
  ROD_LAUNCH_CHECK("smooth_l1_kernel");
  smooth_l1_final_kernel<<<1, 32, 0, st>>>(partial, grid, 1.f / (float)batch, out_loss);
  ROD_LAUNCH_CHECK("smooth_l1_final_kernel");
  return ROD_OK;
}

namespace rod {
struct ClfWs {
  float4* stats; float* nvalues; float* ce0; double* ppos; double* pneg; int* pn; ClfState* st; size_t bytes; int nblk;
};
static ClfWs clf_ws(const rod_layout_t* layout, int batch, void* base) {
  ClfWs w;
  const size_t N = layout->n_total, tot = N * batch;
  w.nblk = loss_grid(layout, batch);                    // partials: one per persistent block
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* q = p ? p + off : nullptr; off += al(bytes); return q; };
  w.st = (ClfState*)take(sizeof(ClfState));
  w.stats = (float4*)take(sizeof(float4) * layout->n_layers * batch);
  w.nvalues = (float*)take(sizeof(float) * tot);
  w.ce0 = (float*)take(sizeof(float) * tot);
  w.ppos = (double*)take(sizeof(double) * w.nblk);
  w.pneg = (double*)take(sizeof(double) * w.nblk);
  w.pn = (int*)take(sizeof(int) * w.nblk);
  w.bytes = off + 256;
  return w;
}
static int clf_params(const rod_layout_t* layout, const rod_layered_t* logits, const rod_layered_t* labels,
                      const rod_layered_t* mask, const rod_layered_t* iou, int batch, int n_classes, ClfParams* P) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(logits, nl, "logits")) || (rc = check_layered(labels, nl, "labels")) ||
      (rc = check_layered(mask, nl, "mask")) || (rc = check_layered(iou, nl, "iou")))
    return rc;
  ROD_REQUIRE(batch >= 1 && n_classes >= 1 && n_classes <= ROD_MAX_CLASSES, "rod_clf_loss: batch=%d n_classes=%d invalid", batch, n_classes);
  ROD_REQUIRE(batch <= 65535, "rod_clf_loss: batch=%d exceeds 65535 (grid y dimension)", batch);
  P->L = to_layout(layout);
  P->logits = to_layered_f(logits, nl); P->iou = to_layered_f(iou, nl);
  P->labels = to_layered_i(labels, nl); P->mask = to_layered_i(mask, nl);
  P->batch = batch; P->C = n_classes;
  return ROD_OK;
}
}  // namespace rod

extern "C" size_t rod_clf_loss_workspace_bytes(const rod_layout_t* layout, int batch) {
  if (!layout || batch <= 0) return 256;
  return rod::clf_ws(layout, batch, nullptr).bytes;
}

extern "C" int rod_clf_loss(const rod_layout_t* layout, const rod_layered_t* logits, const rod_layered_t* labels,
                            const rod_layered_t* mask, const rod_layered_t* iou, int batch, int n_classes,
                            float negative_ratio, float* out6, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rod;
  ClfParams P;
  int rc = clf_params(layout, logits, labels, mask, iou, batch, n_classes, &P);
  if (rc) return rc;
  ROD_REQUIRE(out6 && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "rod_clf_loss: NULL / misaligned output or workspace");
  ROD_REQUIRE(workspace_bytes >= rod_clf_loss_workspace_bytes(layout, batch), "rod_clf_loss: workspace too small");
  const ClfWs W = clf_ws(layout, batch, workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)P.L.n_total * batch;
  const int grid = W.nblk;
  // six launches: the <<<1, 32>>> plan / pick / final steps run in the last block of the kernel in front of them
  iou_stats_kernel<<<dim3(P.L.n_layers, batch), kLossBlock, 0, st>>>(P.L, P.iou, W.stats, W.st);
  ROD_LAUNCH_CHECK("iou_stats_kernel");
  clf_pass1_kernel<<<grid, kLossBlock, 0, st>>>(P, W.stats, W.nvalues, W.ce0, W.ppos, W.pn, negative_ratio, W.st);
  ROD_LAUNCH_CHECK("clf_pass1_kernel");
  const unsigned hgrid = (unsigned)std::min<long long>((total + kLossBlock * 8 - 1) / (kLossBlock * 8), 4ll * sm_count());
  for (int pass = 1; pass < 4; ++pass) {                 // (the top byte was counted by clf_pass1_kernel)
    radix_hist_kernel<<<hgrid, kLossBlock, 0, st>>>(W.nvalues, total, 24 - 8 * pass, W.st);
    ROD_LAUNCH_CHECK("radix_hist_kernel");
  }
  clf_pass2_kernel<<<grid, kLossBlock, 0, st>>>(P, W.nvalues, W.ce0, W.st, W.ppos, W.pneg, 1.f / (float)batch, out6);
  ROD_LAUNCH_CHECK("clf_pass2_kernel");
  return ROD_OK;
}

extern "C" int rod_clf_loss_grad(const rod_layout_t* layout, const rod_layered_t* logits, const rod_layered_t* labels,
                                 const rod_layered_t* mask, const rod_layered_t* iou, int batch, int n_classes,
                                 float upstream, const void* workspace, float* grad_logits, void* stream) {
  using namespace rod;
  ClfParams P;
  int rc = clf_params(layout, logits, labels, mask, iou, batch, n_classes, &P);
  if (rc) return rc;
  ROD_REQUIRE(workspace && grad_logits, "rod_clf_loss_grad: NULL pointer argument");
  const ClfWs W = clf_ws(layout, batch, const_cast<void*>(workspace));
  const dim3 grid((P.L.n_total + kLossBlock - 1) / kLossBlock, batch);
  const size_t tile = (size_t)kLossBlock * n_classes * sizeof(float);
  ROD_CUDA(cudaFuncSetAttribute(clf_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile));
  clf_grad_kernel<<<grid, kLossBlock, tile, (cudaStream_t)stream>>>(P, W.stats, W.nvalues, W.st, upstream / (float)batch, grad_logits);
  ROD_LAUNCH_CHECK("clf_grad_kernel");
  return ROD_OK;
}
