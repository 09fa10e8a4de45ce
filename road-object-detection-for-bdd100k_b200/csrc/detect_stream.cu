// a15 fast path: streaming select + exact top-k + decode + NMS for select_threshold > 0.
// Replaces detected_bboxes (utils/net_tools.py:739-758) with the decode call site of
// evaluate.py:139-143 fused in (boxes are only decoded for the top_k candidates).
//
// With thr > 0 every entry the select stage zeroes (score*0, box*0) ranks below every real
// candidate and is indistinguishable from pad_axis's zero padding in the output, so only the
// candidates with p >= thr matter.  Per (class, image) segment:
//   S   sample_kernel   coarse histogram of the candidates of every 19th ANCHOR (a spatially stratified
//                       sample: any cluster of candidates — one object, one image quadrant, one layer — is
//                       sampled at the same rate; 19 is coprime with the 6 / 9 anchor shapes per cell, so
//                       consecutive samples come from different cells and cycle through all shapes).
//   A   scan_kernel     ONE pass over the [B,N,C] scores (or logits: <.,true> computes the softmax of each
//                       anchor in registers first, f-1).  256-anchor tiles are staged into shared memory
//                       with TMA bulk copies (cp.async.bulk + mbarrier, double buffered); one thread per
//                       anchor.  Every CTA derives, per class, TWO score cuts from the sampled histogram:
//                       about 1.6 * top_k candidates of the segment lie above cut_hi, about 4 * top_k above
//                       cut_lo.  Candidates >= cut_hi go to the CTA's private slice of the segment's tier-1
//                       list (when a slice is full — clustered candidates — to the segment's shared spill
//                       list, one global atomic per spilled entry); candidates in [cut_lo, cut_hi) go to a
//                       tier-2 slice that is only read when tier 1 turns out to hold fewer than top_k.
//   B   segment_kernel  one CTA per segment: histogram of the listed candidates -> threshold bin (the lowest
//                       bin still inside the top_k), counting sort on the score bins + exact in-bin rank
//                       (score desc, anchor asc = tf.nn.top_k order), keep the first top_k, gather + decode
//                       their boxes on demand, greedy NMS in batches against the kept list, write keep_top_k
//                       rows zero padded.
// The cuts are a performance device only: a segment whose two tiers together hold fewer than top_k entries
// although candidates below cut_lo were dropped (a 4x estimation error), whose spill list overflowed, or whose
// threshold bin holds massive ties, is flagged and redone by the exact general path (nms_kernel<fused>: top-k by
// topk_row, then NMS, in one CTA), which runs only for flagged rows.
// Prediction depths other than 11 use the plain-load two-pass kernels (hist_kernel, thresh_kernel,
// collect_kernel: full histogram, exact threshold bin) in front of the same segment kernel.
#include <cuda_fp16.h>

#include "select_topk.cuh"

namespace rod {

constexpr int kBins = 1024;
constexpr int kStreamBlock = 256;
constexpr int kSegBlock = 256;
constexpr int kSegWarps = kSegBlock / 32;
static_assert(kSegBlock == 256, "segment_kernel zeroes its 256-bin sample histogram row with one store per thread");

__device__ __forceinline__ int score_bin(float s) {
  // monotone non-decreasing in s for s > 0; only resolution (never correctness) depends on it
  return min((int)(s * (float)kBins), kBins - 1);     // float -> int conversion saturates, NaN -> 0
}

// Walks the float range [F0, F1) of image b's concatenated [N*C] score array (layer-major),
// calling fn(score, anchor, class) for every element.  float4 loads on the 16 B-aligned body.
template <typename Fn>
__device__ __forceinline__ void for_each_score(const Layout& L, const LayeredF& probs, int C, int b, int F0, int F1,
                                               Fn fn) {
  for (int l = 0; l < L.n_layers; ++l) {
    const int lo = max(F0, L.offset[l] * C), hi = min(F1, L.offset[l + 1] * C);
    if (lo >= hi) continue;
    const float* slab = probs.base[l] + (long long)b * probs.stride[l] - (long long)L.offset[l] * C;
    const int mis = (int)((reinterpret_cast<uintptr_t>(slab + lo) >> 2) & 3);
    const int head = min((4 - mis) & 3, hi - lo);
    const int nvec = (hi - lo - head) >> 2;
    const int tail0 = lo + head + 4 * nvec;
    // scalar head / tail
    if ((int)threadIdx.x < head) {
      const int f = lo + threadIdx.x;
      fn(__ldg(slab + f), f / C, f % C);
    }
    if ((int)threadIdx.x < hi - tail0) {
      const int f = tail0 + threadIdx.x;
      fn(__ldg(slab + f), f / C, f % C);
    }
    // vector body: thread-private (anchor, class) cursor advanced without divisions
    int f = lo + head + 4 * threadIdx.x;
    int n = f / C, c = f - n * C;
    const int step = 4 * blockDim.x, step_n = step / C, step_c = step - step_n * C;
    for (int v = threadIdx.x; v < nvec; v += 4 * blockDim.x) {
      // 4 independent 16 B loads in flight per thread before any is consumed
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v + u * (int)blockDim.x < nvec) x[u] = ldg4(slab + f + u * step);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v + u * (int)blockDim.x < nvec) {
          int nn = n, cc = c;
          fn(x[u].x, nn, cc); if (++cc == C) { cc = 0; ++nn; }
          fn(x[u].y, nn, cc); if (++cc == C) { cc = 0; ++nn; }
          fn(x[u].z, nn, cc); if (++cc == C) { cc = 0; ++nn; }
          fn(x[u].w, nn, cc);
        }
        f += step; n += step_n; c += step_c;
        if (c >= C) { c -= C; ++n; }
      }
    }
  }
}

struct StreamParams {
  LayeredF probs;
  Layout L;
  int C, ignore_class, batch, chunk;   // chunk = floats per CTA (multiple of 4)
  float thr;
};

__global__ void __launch_bounds__(kStreamBlock)
hist_kernel(const __grid_constant__ StreamParams P, unsigned* __restrict__ g_hist) {
  extern __shared__ unsigned s_hist[];            // [C][kBins]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < P.C * kBins; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const int F0 = blockIdx.x * P.chunk, F1 = min(F0 + P.chunk, P.L.n_total * P.C);
  const float thr = P.thr;
  const int ign = P.ignore_class;
  for_each_score(P.L, P.probs, P.C, b, F0, F1, [&](float s, int, int c) {
    if (s >= thr && c != ign) atomicAdd(&s_hist[c * kBins + score_bin(s)], 1u);
  });
  __syncthreads();
  for (int i = threadIdx.x; i < P.C * kBins; i += blockDim.x) {
    const unsigned v = s_hist[i];
    if (v) {
      const int c = i / kBins, bin = i - c * kBins;
      atomicAdd(&g_hist[((size_t)c * P.batch + b) * kBins + bin], v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// TMA-staged scan (prediction depth C known at compile time)
// ------------------------------------------------------------------------------------------
constexpr int kScanBlock = 256;     // threads = anchors per tile
constexpr int kScanStages = 2;      // TMA tiles in flight per CTA (ring of shared-memory buffers; 4 measured slower again in round 2: 79 vs 72 us per step)
constexpr int kListCap = 4096;      // per segment candidate list entries (8 B each)
constexpr int kMaxChunks = 64;      // CTAs per image; each owns kListCap / chunks list slots per class

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// 1-D bulk copy global -> shared, completion signalled on `bar` (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

constexpr int kSampleBins = 256;    // coarse bins (4 score bins each) of the sampled histogram
constexpr int kSampleStride = 19;   // every 19th anchor is sampled (coprime with the anchor shapes per cell)
constexpr int kSamplePhase = 9;     // anchors 9, 28, 47, ...
constexpr int kSpillCap = 2048;     // per segment: entries that did not fit their CTA's slice

struct ScanParams {
  LayeredF probs;
  Layout L;
  int ignore_class, batch, chunk;      // chunk = anchors per CTA (multiple of 256)
  int chunks, spc;                     // CTAs per image, list slots per (class, CTA)
  float thr;
  unsigned* g_shist;                   // [rows][kSampleBins] histogram of the sampled tiles (sample_kernel); zero between calls
  unsigned* g_zero4;                   // [4][rows] + [4]: over flags, dense counters, spill counters, tier-2 flags, any: cleared by sample_kernel
  int target_hi, target_lo;            // sampled candidates that must lie at or above cut_hi / cut_lo
  int* g_est;                          // [rows] cut_lo as a score bin (0: nothing below it was dropped)
  unsigned* g_cnt1;                    // [rows][kMaxChunks] tier-1 candidates written by each CTA
  unsigned* g_cnt2;                    // [rows][kMaxChunks] tier-2 candidates written by each CTA
  unsigned long long* g_list2;         // [rows][kListCap]  per-CTA tier-2 slices
  unsigned* g_lo_over;                 // [rows] set when a tier-2 slice overflowed (tier 2 incomplete)
  unsigned* g_any;                     // [1] set when any segment is flagged for the exact general kernels
  int b0, nb;                          // images [b0, b0 + nb) of the batch are scanned by this launch
  unsigned* g_over;                    // [rows] set when a CTA ran out of list slots: exact general kernels redo the segment
  unsigned long long* g_list;          // [rows][kListCap]  per-CTA slices
  unsigned* g_spill_cnt;               // [rows] entries appended to the spill list (may exceed kSpillCap: overflow)
  unsigned long long* g_spill;         // [rows][kSpillCap] entries whose slice was full
};

template <int C>
struct ScanShared {
  unsigned cnt[C], cnt2[C];            // tier-1 / tier-2 entries of this CTA
  float cut_hi[C], cut_lo[C];          // per-class tier thresholds: max(thr, estimated cut)
  unsigned spill_full[C];              // set once the segment's spill list is exhausted
  unsigned lo_over[C];                 // set when this CTA's tier-2 slice overflowed
};

// probabilities of one anchor in registers: e[] holds the C scores, or (LOGITS, f-1) the logits, which are
// replaced by their softmax; mx / rinv let a single class be recomputed later with the same instructions
template <int C, bool LOGITS>
__device__ __forceinline__ void row_probs(float (&e)[C], float& mx, float& rinv) {
  if constexpr (LOGITS) {
    mx = e[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, e[c]);
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = softmax_exp(e[c], mx);
    float sum = e[0];
#pragma unroll
    for (int c = 1; c < C; ++c) sum = __fadd_rn(sum, e[c]);
    rinv = __frcp_rn(sum);
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = __fmul_rn(e[c], rinv);
  }
}

// Pre-pass: histogram (coarse bins) of the candidates of every kSampleStride-th anchor.  From it the scan
// pass estimates, per segment, the two score cuts.  The estimates cannot affect the result: a segment whose
// lists then hold fewer than top_k entries although candidates were dropped (or overflow) is redone by the
// exact general kernels.
template <int C, bool LOGITS>
__global__ void __launch_bounds__(256)
sample_kernel(const __grid_constant__ ScanParams P) {
  __shared__ float s_rows[8][32 * C];                  // per warp: its 32 sampled rows
  pdl_launch_dependents();                             // the scan pass may start its prologue now
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the flags / counters the LATER kernels of this call raise (this kernel uses none of them): cleared here instead
  // of by a memset node in front of every call
  if (blockIdx.x == 0 && threadIdx.x < 4 * C) {
    const size_t rows = (size_t)C * P.batch;
    P.g_zero4[(threadIdx.x / C) * rows + (size_t)(threadIdx.x % C) * P.batch + b] = 0u;
    if (b == 0 && threadIdx.x == 0) P.g_zero4[4 * rows] = 0u;
  }
  const int n = ((blockIdx.x * 8 + warp) * 32 + lane) * kSampleStride + kSamplePhase;
  const bool valid = n < P.L.n_total;
  const float* row = nullptr;
  if (valid) {
    const int l = layer_of(P.L, n);
    row = P.probs.base[l] + (long long)b * P.probs.stride[l] + (long long)(n - P.L.offset[l]) * C;
  }
  // The rows of a warp lie kSampleStride * C floats apart: a lane-per-row load would touch 32 cache lines per
  // instruction.  Stage them with lane-per-ELEMENT loads instead (element idx of the warp's 32 * C floats belongs
  // to row idx / C): ~3 lines per instruction, then every lane reads its own row from shared memory.
  float* mine = s_rows[warp];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int idx = j * 32 + lane, r = idx / C, k = idx - r * C;
    const unsigned long long pr = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)row, r);
    if (pr) mine[idx] = __ldg(reinterpret_cast<const float*>((uintptr_t)pr) + k);
  }
  __syncwarp();
  if (!valid) return;
  float e[C], mx = 0.f, rinv = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) e[c] = mine[lane * C + c];          // stride C words: conflict-free for odd C
  row_probs<C, LOGITS>(e, mx, rinv);
  unsigned cand = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) cand |= (unsigned)(e[c] >= P.thr) << c;
  cand &= ~(1u << P.ignore_class);
  unsigned* h = P.g_shist + (size_t)b * kSampleBins;
  while (cand) {
    const int c = __ffs(cand) - 1;
    cand &= cand - 1;
    float s = mine[lane * C + c];
    if constexpr (LOGITS) s = __fmul_rn(softmax_exp(s, mx), rinv);
    atomicAdd(h + (size_t)c * P.batch * kSampleBins + (score_bin(s) >> 2), 1u);
  }
}

// One anchor per lane; `row` points at its C scores (shared-memory tile, or global for the few
// unaligned head / tail anchors).  Candidates at or above the class's cut are appended to this CTA's
// private slice of the segment's list while it has room.
// LOGITS: the row holds logits; probabilities come from softmax_exp / __frcp_rn (common.cuh), and a
// candidate's probability is recomputed with the same instructions when its key is built.
template <int C, bool LOGITS>
__device__ __forceinline__ void scan_anchor(const ScanParams& P, ScanShared<C>& S, const float (&cut)[C], bool valid,
                                            const float* __restrict__ row, int n, int b, int chunk_id) {
  const size_t slice_off = (size_t)b * kListCap + (size_t)chunk_id * P.spc;
  const size_t slice_stride = (size_t)P.batch * kListCap;
  unsigned cand = 0;
  float mx = 0.f, rinv = 0.f;
  if (valid) {
    float e[C];
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = row[c];
    row_probs<C, LOGITS>(e, mx, rinv);
#pragma unroll
    for (int c = 0; c < C; ++c) cand |= (unsigned)(e[c] >= cut[c]) << c;
    cand &= ~(1u << P.ignore_class);
  }
  const unsigned nkey = (unsigned)(~(unsigned)n);
  while (cand) {
    const int c = __ffs(cand) - 1;
    cand &= cand - 1;
    float s = row[c];
    if constexpr (LOGITS) s = __fmul_rn(softmax_exp(s, mx), rinv);
    const unsigned long long entry = ((unsigned long long)__float_as_uint(s) << 32) | nkey;
    if (s < S.cut_hi[c]) {                                // tier 2: only read when tier 1 holds fewer than top_k
      bool placed2 = false;
      if (S.cnt2[c] < (unsigned)P.spc) {
        const unsigned pos = atomicAdd(&S.cnt2[c], 1u);
        if (pos < (unsigned)P.spc) {
          P.g_list2[slice_off + (size_t)c * slice_stride + pos] = entry;
          placed2 = true;
        }
      }
      if (!placed2) S.lo_over[c] = 1u;
      continue;
    }
    bool placed = false;
    if (S.cnt[c] < (unsigned)P.spc) {                     // (no shared atomic once the slice is full)
      const unsigned pos = atomicAdd(&S.cnt[c], 1u);
      if (pos < (unsigned)P.spc) {
        P.g_list[slice_off + (size_t)c * slice_stride + pos] = entry;
        placed = true;
      }
    }
    if (!placed && !S.spill_full[c]) {                    // slice full: the segment's shared spill list
      const size_t r = (size_t)c * P.batch + b;
      const unsigned q = atomicAdd(&P.g_spill_cnt[r], 1u);
      if (q < (unsigned)kSpillCap) P.g_spill[r * kSpillCap + q] = entry;
      else S.spill_full[c] = 1u;                          // (dense inputs: no global atomic once the spill list is full)
    }
  }
}

template <int C, bool LOGITS, int kScanStages>
__global__ void __launch_bounds__(kScanBlock)
scan_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  float* s_tiles = reinterpret_cast<float*>(s_dyn);                      // kScanStages x [256][C]
  ScanShared<C>& S = *reinterpret_cast<ScanShared<C>*>(s_dyn + kScanStages * sizeof(float) * kScanBlock * C);
  __shared__ __align__(8) unsigned long long s_bar[kScanStages];

  const int tid = threadIdx.x;
  pdl_launch_dependents();                             // the next kernel of the stream may be scheduled (it waits for this grid)
  if (tid == 0) {
    for (int q = 0; q < kScanStages; ++q) mbar_init(&s_bar[q], 1);
    fence_mbar_init();
  }
  pdl_wait();                                          // sampled histogram complete
  unsigned phases = 0;                                 // bit q = parity of barrier q
  const int b = P.b0 + blockIdx.y, chunk_id = blockIdx.x;   // grid = (chunks per image, images)
  if (tid < C) { S.cnt[tid] = 0u; S.cnt2[tid] = 0u; S.spill_full[tid] = 0u; S.lo_over[tid] = 0u; }
  // per-class cuts from the sampled histogram: the highest coarse bin t with count(bins >= t) >= target.
  // One HALF warp per class (16 lanes x 16 bins, four 16-byte loads each): all C <= 16 classes in one round of
  // loads instead of two rounds of whole warps — this prologue sits in front of every CTA's first compare.
  static_assert(C <= kScanBlock / 16 && kSampleBins == 256, "one half warp per class, 16 bins per lane");
  {
    const int c = tid >> 4, hl = tid & 15;
    unsigned v[16], sum = 0;
    if (c < C) {
      const uint4* h = reinterpret_cast<const uint4*>(P.g_shist + ((size_t)c * P.batch + b) * kSampleBins + hl * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 x = h[q];
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        sum += x.x + x.y + x.z + x.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = 0u;
    }
    unsigned suf = sum;                                // inclusive suffix over the half warp (lane 15 owns the top bins)
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const unsigned x = __shfl_down_sync(0xffffffffu, suf, o, 16);
      if (hl + o < 16) suf += x;
    }
    const unsigned above = suf - sum;
    int tb[2];
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const unsigned target = (unsigned)(which == 0 ? P.target_hi : P.target_lo);
      int t = 0;
      if (above < target && suf >= target) {           // exactly one lane when the row holds >= target samples
        unsigned acc = above;
#pragma unroll
        for (int q = 15; q >= 0; --q) {
          acc += v[q];
          if (acc >= target) { t = hl * 16 + q; break; }
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) t = max(t, __shfl_xor_sync(0xffffffffu, t, o, 16));
      tb[which] = t;
    }
    if (hl == 0 && c < C) {
      // score_bin(s) >= 4t  <=>  s >= 4t / 1024
      const float hi = fmaxf(P.thr, (float)(4 * tb[0]) * (1.f / (float)kBins));
      const float lo = fmaxf(P.thr, (float)(4 * tb[1]) * (1.f / (float)kBins));
      S.cut_hi[c] = hi;
      S.cut_lo[c] = fminf(lo, hi);
      if (chunk_id == 0) P.g_est[(size_t)c * P.batch + b] = lo > P.thr ? 4 * tb[1] : 0;
    }
  }
  __syncthreads();
  float cut[C];
#pragma unroll
  for (int c = 0; c < C; ++c) cut[c] = S.cut_lo[c];

  const int A0 = chunk_id * P.chunk, A1 = min(A0 + P.chunk, P.L.n_total);
  for (int l = 0; l < P.L.n_layers; ++l) {
    const int lo = max(A0, P.L.offset[l]), hi = min(A1, P.L.offset[l + 1]);
    if (lo >= hi) continue;
    // row(n) = slab + n * C for the global anchor index n
    const float* slab = P.probs.base[l] + (long long)b * P.probs.stride[l] - (long long)P.L.offset[l] * C;
    // bulk copies need 16 B aligned sources: peel up to 3 head anchors (C odd: 4C*h covers every residue)
    int head = 0;
    while (head < 4 && ((reinterpret_cast<uintptr_t>(slab + (long long)(lo + head) * C)) & 15u) != 0) ++head;
    int t0 = lo + head, t1 = hi;
    if (head == 4 || t0 >= hi) { t0 = hi; t1 = hi; }
    t1 = t0 + ((t1 - t0) & ~3);                          // whole multiples of 4 anchors = 16 B multiples
    // ---- TMA tiles of 256 anchors through a ring of kScanStages buffers: start the first copies,
    //      then do the scalar anchors while they are in flight
    const int ntiles = (t1 - t0 + kScanBlock - 1) / kScanBlock;
    auto issue_tile = [&](int t) {
      const int a = t0 + t * kScanBlock;
      const int q = t % kScanStages;
      tma_load_1d(s_tiles + q * kScanBlock * C, slab + (long long)a * C, (unsigned)(min(kScanBlock, t1 - a) * C * 4),
                  &s_bar[q]);
    };
    if (tid == 0)
      for (int t = 0; t < min(ntiles, kScanStages - 1); ++t) issue_tile(t);
    // ---- scalar anchors: [lo, t0) and [t1, hi), at most a handful, straight from global memory
    const int nscalar = (t0 - lo) + (hi - t1);
    for (int i = tid; i < nscalar; i += kScanBlock) {
      const int n = i < t0 - lo ? lo + i : t1 + (i - (t0 - lo));
      scan_anchor<C, LOGITS>(P, S, cut, true, slab + (long long)n * C, n, b, chunk_id);
    }
    for (int t = 0; t < ntiles; ++t) {
      const int q = t % kScanStages;
      const int a0 = t0 + t * kScanBlock;
      const int cnt = min(kScanBlock, t1 - a0);
      // buffer (t-1) % stages was released by the barrier at the end of the previous iteration
      if (t + kScanStages - 1 < ntiles && tid == 0) issue_tile(t + kScanStages - 1);
      mbar_wait(&s_bar[q], (phases >> q) & 1u);
      phases ^= 1u << q;
      // row stride C words: conflict-free across lanes for odd C
      scan_anchor<C, LOGITS>(P, S, cut, tid < cnt, s_tiles + q * kScanBlock * C + tid * C, a0 + tid, b, chunk_id);
      __syncthreads();                                   // tile consumed: its buffer may be refilled
    }
  }
  __syncthreads();
  if (tid < C) {
    const size_t r = (size_t)tid * P.batch + b;
    const unsigned c = S.cnt[tid], c2 = S.cnt2[tid];
    P.g_cnt1[r * kMaxChunks + chunk_id] = c < (unsigned)P.spc ? c : (unsigned)P.spc;
    P.g_cnt2[r * kMaxChunks + chunk_id] = c2 < (unsigned)P.spc ? c2 : (unsigned)P.spc;
    if (S.lo_over[tid]) P.g_lo_over[r] = 1u;
    if (S.spill_full[tid]) { P.g_over[r] = 1u; *P.g_any = 1u; }   // the spill list overflowed: exact general kernels redo the segment
  }
}

// one warp per segment: the smallest bin t with count(bins >= t) >= k  (0 when fewer than k candidates)
__global__ void __launch_bounds__(256)
thresh_kernel(const unsigned* __restrict__ g_hist, int rows, int k, int* __restrict__ tbin) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const uint4* h = reinterpret_cast<const uint4*>(g_hist + (size_t)r * kBins + lane * 32);
  unsigned v[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 x = __ldg(h + q);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
  unsigned sum = 0;
#pragma unroll
  for (int q = 0; q < 32; ++q) sum += v[q];
  // inclusive suffix sum over lanes (lane 31 owns the highest bins)
  unsigned suf = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned x = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += x;
  }
  const unsigned above = suf - sum;                 // candidates in higher lanes' bins
  const unsigned total = __shfl_sync(0xffffffffu, suf, 0);
  if (total < (unsigned)k) {
    if (lane == 0) tbin[r] = 0;
    return;
  }
  if (above < (unsigned)k && suf >= (unsigned)k) {
    unsigned acc = above;
    int t = lane * 32;
#pragma unroll
    for (int q = 31; q >= 0; --q) {
      acc += v[q];
      if (acc >= (unsigned)k) { t = lane * 32 + q; break; }
    }
    tbin[r] = t;
  }
}

__global__ void __launch_bounds__(kStreamBlock)
collect_kernel(const __grid_constant__ StreamParams P, const int* __restrict__ tbin, unsigned* __restrict__ g_cnt,
               unsigned long long* __restrict__ g_list, int cap) {
  __shared__ int s_tbin[ROD_MAX_CLASSES];
  const int b = blockIdx.y;
  if ((int)threadIdx.x < P.C) s_tbin[threadIdx.x] = tbin[threadIdx.x * P.batch + b];
  __syncthreads();
  const int F0 = blockIdx.x * P.chunk, F1 = min(F0 + P.chunk, P.L.n_total * P.C);
  const float thr = P.thr;
  const int ign = P.ignore_class;
  for_each_score(P.L, P.probs, P.C, b, F0, F1, [&](float s, int n, int c) {
    if (s >= thr && c != ign && score_bin(s) >= s_tbin[c]) {
      const size_t r = (size_t)c * P.batch + b;
      const unsigned pos = atomicAdd(&g_cnt[r], 1u);
      if (pos < (unsigned)cap)
        g_list[r * cap + pos] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(~(unsigned)n);
    }
  });
}

// ------------------------------------------------------------------------------------------
// B: per-segment sort + decode + NMS
// ------------------------------------------------------------------------------------------
struct SegParams {
  LayeredF loc, refine, det;
  const float* center;
  Layout L;
  int has_loc, batch, ignore_class, cap, k, keep;
  float nms_thr;
  const float* clip;
  // candidate lists: per-CTA slices of the scan pass ([rows][kListCap], counts [rows][kMaxChunks]), or
  // (force_dense: prediction depths without the TMA scan) one list per segment ([rows][cap], cnt2)
  const unsigned long long* list1;
  const unsigned* cnt1;
  const unsigned long long* list2;   // tier-2 slices ([rows][kListCap], counts cnt1b [rows][kMaxChunks]): candidates in [cut_lo, cut_hi)
  const unsigned* cnt1b;
  const unsigned* lo_over;           // [rows] tier 2 is incomplete (a slice overflowed)
  const unsigned* spill_cnt;         // [rows] spill entries (values above kSpillCap: overflow, flagged by the scan pass)
  const unsigned long long* spill;   // [rows][kSpillCap]
  const unsigned* cnt2;
  const int* est;            // [rows] score bin below which the scan pass dropped candidates (0: none dropped)
  long long* dbg;            // optional per-segment phase timestamps (rod_debug_set_timing), else NULL
  unsigned* over;            // out: 1 when the segment must be redone by the exact general kernels
  unsigned* any;             // out: 1 when any segment is
  unsigned* shist;           // sampled histogram [rows][kSampleBins] of the scan pass (or NULL): every CTA zeroes its row for the next call
  int b0, nb;                // images [b0, b0 + nb) of the batch: one CTA per (class, image of the range)
  int chunks, spc, force_dense;
};

// Region A of the segment kernel's shared memory: the sort keys first, then (aliased) the NMS working set:
//   kept rows in packed half form (keep x 16 B), per-warp column masks (64 x 8 x 4 B),
//   candidates' shrunk boxes in packed half form (k x 8 B), candidates' boxes as packed half rows (k x 16 B)
__host__ __device__ inline size_t seg_qh_offset(int keep) { return (size_t)keep * 16 + 64 * 8 * 4; }
__host__ __device__ inline size_t seg_nh_offset(int k, int keep) { return ((seg_qh_offset(keep) + (size_t)k * 8) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t seg_region_a(int cap, int k, int keep) {
  const size_t a = (size_t)cap * 16 + 1024 * 4, b = seg_nh_offset(k, keep) + (size_t)k * 16;
  return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// Pair predicate in half precision, two candidates per instruction.  "Row r can intersect candidate box q"
// is r.ymax > q.ymin && r.ymin < q.ymax && r.xmax > q.xmin && r.xmin < q.xmax.  Rows are stored as
// (ymax rounded UP, ymin rounded DOWN, xmax UP, xmin DOWN), each duplicated into both halves of a half2;
// candidates as (ymin DOWN, ymax UP, xmin DOWN, xmax UP), the two candidates of a lane side by side.  Rounding
// outwards keeps the test a NECESSARY condition of the float predicate (x > y  =>  up(x) > down(y)), so it only
// filters: every surviving pair still goes through the exact IoU test.  Four HSET2 + two LOP3 test one row
// against two candidates (eight FSETP before).
__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint4 pack_row(const float4& b) {          // b = (ymin, xmin, ymax, xmax)
  const __half zu = __float2half_ru(b.z), xd = __float2half_rd(b.x), wu = __float2half_ru(b.w), yd = __float2half_rd(b.y);
  return make_uint4(h2u(__halves2half2(zu, zu)), h2u(__halves2half2(xd, xd)), h2u(__halves2half2(wu, wu)), h2u(__halves2half2(yd, yd)));
}
__device__ __forceinline__ uint2 pack_cand(const float4& q) {         // (ymin DOWN | ymax UP), (xmin DOWN | xmax UP)
  return make_uint2(h2u(__halves2half2(__float2half_rd(q.x), __float2half_ru(q.z))),
                    h2u(__halves2half2(__float2half_rd(q.y), __float2half_ru(q.w))));
}
// 0xFFFF in the low half: the row can intersect candidate 0, in the high half: candidate 1
__device__ __forceinline__ unsigned pair_mask(const uint4& row, __half2 qymin, __half2 qymax, __half2 qxmin, __half2 qxmax) {
  return __hgt2_mask(u2h(row.x), qymin) & __hlt2_mask(u2h(row.y), qymax) & __hgt2_mask(u2h(row.z), qxmin) & __hlt2_mask(u2h(row.w), qxmax);
}

// exact "fdiv_rn(inter, den) > thr" with a division-free fast path (thr >= 0, den > 0)
__device__ __forceinline__ bool iou_exceeds(float inter, float den, float thr) {
  const float p = __fmul_rn(thr, den);
  if (inter > __fmul_rn(p, 1.00000095367431640625f)) return true;     // 1 + 2^-20
  if (inter < __fmul_rn(p, 0.99999904632568359375f)) return false;    // 1 - 2^-20
  return __fdiv_rn(inter, den) > thr;
}

__global__ void __launch_bounds__(kSegBlock, 5)
segment_kernel(const __grid_constant__ SegParams P, const unsigned long long* __restrict__ g_list, float* __restrict__ out_scores,
               float* __restrict__ out_boxes, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  // region A: sort keys (cap x 8 B); after the gather it is re-used for the NMS working set
  //           (kept boxes, kept areas, overlap words, intra-batch rows)
  // then: boxes (k x 16), normalised boxes (k x 16), area, score (k x 4 each), kept positions (keep x 4)
  const int cap = P.cap, k = P.k, keep = P.keep;
  const size_t regionA = seg_region_a(cap, k, keep);
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_raw);
  uint4* s_kh = reinterpret_cast<uint4*>(s_raw);                                        // [keep] kept rows, packed half form
  unsigned* s_cmw = reinterpret_cast<unsigned*>(s_kh + keep);                          // [64][kSegWarps] partial column masks
  uint2* s_qh = reinterpret_cast<uint2*>(s_raw + seg_qh_offset(keep));                 // [k] shrunk candidate boxes, packed half form
  uint4* s_nh = reinterpret_cast<uint4*>(s_raw + seg_nh_offset(k, keep));              // [k] candidate boxes as packed half rows
  float4* s_box = reinterpret_cast<float4*>(s_raw + regionA);
  float4* s_nbox = s_box + k;
  float* s_area = reinterpret_cast<float*>(s_nbox + k);
  float* s_score = s_area + k;
  int* s_selected = reinterpret_cast<int*>(s_score + k);
  __shared__ unsigned long long s_deadw[kSegWarps], s_sel;      // per-warp "suppressed by the kept list" masks of a batch
  __shared__ int s_nsel;

  const int c = (int)(blockIdx.x / P.nb), b = P.b0 + (int)(blockIdx.x % P.nb);
  const long long r = (long long)c * P.batch + b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();                             // the fallback kernels may be scheduled (they wait for this grid)
  if (out_counts && tid == 0) out_counts[r] = 0;       // (flagged rows: the general kernels add theirs)
  if (c == P.ignore_class) return;
  pdl_wait();                                          // candidate lists of the preceding grid complete
  if (P.shist) P.shist[r * kSampleBins + tid] = 0u;    // (the scan pass has read it; kSampleBins == kSegBlock)
#define SEG_T(i) do { if (P.dbg != nullptr && tid == 0) P.dbg[r * 8 + (i)] = clock64(); } while (0)
  SEG_T(0);
  // ---- 0. histogram of the segment's listed candidates -> threshold bin (the lowest bin still inside
  // the top_k) and the start offset of every bin in descending order.
  __shared__ int s_n, s_tbv;
  __shared__ unsigned s_part[kMaxChunks], s_part2[kMaxChunks];
  __shared__ unsigned s_wsum[kSegWarps];
  __shared__ int s_tot, s_has2;
  const bool dense = P.force_dense != 0;
  const unsigned long long* base;
  const unsigned long long* base2 = nullptr;
  const unsigned long long* spill = nullptr;
  unsigned n_spill = 0;
  int parts, pstride;
  if (tid == 0) s_has2 = 0;
  __syncthreads();
  if (dense) {
    const unsigned n_in = P.cnt2[r];
    if (n_in > (unsigned)cap) {                       // massive ties in the threshold bin: exact kernels redo it
      if (tid == 0) { P.over[r] = 1u; *P.any = 1u; }
      return;
    }
    base = g_list + r * cap; parts = 1; pstride = 0;
    if (tid == 0) s_part[0] = n_in;
  } else {
    if (P.over[r] != 0u) return;                      // a scan CTA ran out of list slots: exact kernels redo it
    base = P.list1 + r * kListCap; parts = P.chunks; pstride = P.spc;
    base2 = P.list2 + r * kListCap;
    if (tid < parts) {
      s_part[tid] = P.cnt1[r * kMaxChunks + tid];
      const unsigned n2 = P.cnt1b[r * kMaxChunks + tid];
      s_part2[tid] = n2;
      if (n2) s_has2 = 1;
    }
    n_spill = min(P.spill_cnt[r], (unsigned)kSpillCap);
    spill = P.spill + r * kSpillCap;
  }
  unsigned long long* s_tmp = s_keys + cap;
  unsigned* s_h = reinterpret_cast<unsigned*>(s_keys + 2 * cap);          // [kBins]
  bool use2 = false;                                   // tier 2 joins when tier 1 holds fewer than top_k
  // Visits every listed entry: a warp takes whole slices (a slice holds ~100 entries once the scan pass cuts
  // at ~2 * top_k per segment), four independent loads per lane; a single long list is split over the block.
  auto for_each_entry = [&](auto&& fn) {
    const int stride = parts == 1 ? kSegBlock : 32, first = parts == 1 ? tid : lane;
    const int nparts = use2 ? 2 * parts : parts;       // parts .. 2 * parts - 1: the tier-2 slices
    for (int pp = parts == 1 ? 0 : warp; pp < nparts; pp += kSegWarps) {
      const bool second = pp >= parts;
      const int part = second ? pp - parts : pp;
      const unsigned n = second ? s_part2[part] : s_part[part];
      const unsigned long long* sl = (second ? base2 : base) + (size_t)part * pstride;
      for (unsigned j0 = 0; j0 < n; j0 += 4 * stride) {
        unsigned long long ev[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned j = j0 + u * stride + first;
          ev[u] = j < n ? sl[j] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if ((unsigned)(ev[u] >> 32) != 0u) fn(ev[u]);   // empty slots are 0; real entries have score >= thr > 0
      }
    }
    for (unsigned j0 = 0; j0 < n_spill; j0 += 4 * kSegBlock) {       // spill list (usually empty): the whole block
      unsigned long long ev[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned j = j0 + u * kSegBlock + tid;
        ev[u] = j < n_spill ? spill[j] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if ((unsigned)(ev[u] >> 32) != 0u) fn(ev[u]);
    }
  };
  for (int attempt = 0; attempt < 2; ++attempt) {
  for (int i = tid; i < kBins; i += kSegBlock) s_h[i] = 0u;
  __syncthreads();
  for_each_entry([&](unsigned long long e) { atomicAdd(&s_h[score_bin(__uint_as_float((unsigned)(e >> 32)))], 1u); });
  __syncthreads();
  {
    // suffix sums over bins (higher bins first): thread t owns bins [4t, 4t+4)
    static_assert(kBins == 4 * kSegBlock, "one thread per four bins");
    const unsigned h0 = s_h[4 * tid], h1 = s_h[4 * tid + 1], h2 = s_h[4 * tid + 2], h3 = s_h[4 * tid + 3];
    const unsigned mine = h0 + h1 + h2 + h3;
    unsigned suf = mine;                               // inclusive suffix over the warp's higher lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned x = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += x;
    }
    if (lane == 0) s_wsum[warp] = suf;
    __syncthreads();
    unsigned above = suf - mine;
    for (int w = warp + 1; w < kSegWarps; ++w) above += s_wsum[w];
    const unsigned c3 = above + h3, c2 = c3 + h2, c1 = c2 + h1, c0 = c1 + h0;     // count(bins >= 4t+3 .. 4t)
    const unsigned kk = (unsigned)k;
    if (above < kk && c0 >= kk) {                      // exactly one thread: the k-th candidate is in my bins
      const int q = c3 >= kk ? 3 : (c2 >= kk ? 2 : (c1 >= kk ? 1 : 0));
      s_tbv = 4 * tid + q;
      s_n = (int)(q == 3 ? c3 : (q == 2 ? c2 : (q == 1 ? c1 : c0)));
    } else if (tid == 0 && c0 < kk) {                  // fewer than k listed candidates: all of them
      s_tbv = 0;
      s_n = (int)c0;
    }
    if (tid == 0) s_tot = (int)c0;
    s_h[4 * tid + 3] = above;                          // start offset of every bin in the sorted order
    s_h[4 * tid + 2] = c3;
    s_h[4 * tid + 1] = c2;
    s_h[4 * tid] = c1;
  }
  __syncthreads();
  // tier 1 holds fewer than top_k candidates (cut_hi was estimated too high): take tier 2 in as well
  if (dense || use2 || s_tot >= k || !s_has2) break;
  use2 = true;
  __syncthreads();
  }
  const int cnt = s_n, tb = s_tbv;
  // fewer than k listed although the scan pass dropped candidates below cut_lo (or lost tier-2 entries), or
  // massive ties in the threshold bin: the exact general kernels redo the segment
  if ((!dense && cnt < k && (P.est[r] > 0 || P.lo_over[r] != 0u)) || cnt > cap) {
    if (tid == 0) { P.over[r] = 1u; *P.any = 1u; }
    return;
  }
  SEG_T(1);
  // ---- 1. sort: descending (score bits, ~anchor) = tf.nn.top_k order.  Counting sort on the score
  // bins (monotone in the score): scatter the entries at or above the threshold bin to their bin's
  // range, then an exact rank inside each bin (bins hold a handful of entries; all-tied inputs stay
  // correct, just slower).
  for_each_entry([&](unsigned long long e) {
    const int bin = score_bin(__uint_as_float((unsigned)(e >> 32)));
    if (bin >= tb) s_tmp[atomicAdd(&s_h[bin], 1u)] = e;
  });
  __syncthreads();                                    // now s_h[bin] = end of the bin's range = start of bin - 1's
  for (int j = tid; j < cnt; j += kSegBlock) {
    const unsigned long long e = s_tmp[j];
    const int bin = score_bin(__uint_as_float((unsigned)(e >> 32)));
    const int s0 = bin < kBins - 1 ? (int)s_h[bin + 1] : 0, s1 = (int)s_h[bin];
    int rank = 0;
    for (int x = s0; x < s1; ++x) rank += (s_tmp[x] > e) ? 1 : 0;
    s_keys[s0 + rank] = e;
  }
  __syncthreads();
  SEG_T(2);
  const int m = min(cnt, k);                          // real candidates entering NMS

  const float thr = P.nms_thr;
  // Pair predicate.  iou > thr needs inter > thr * max(area_r, area_c), hence an overlap of more
  // than thr * h_c in y and thr * w_c in x (inter <= ih * w_c and inter <= h_c * iw, in the rounded
  // arithmetic too: fl() is monotone).  So the candidate's box shrunk by 0.99 * thr * (h, w) on every
  // side must still intersect the row: the same four compares as the plain intersection test, ~3x
  // fewer IoU evaluations.  The 1 % slack covers the rounding of the shrunk corners as long as the
  // coordinates are < 2^16 times the shrink; otherwise the plain box is used.
  const float tq = __fmul_rn(thr, 0.99f);
  auto shrunk = [&](const float4& c, float area) -> float4 {
    const float sh = __fmul_rn(tq, __fsub_rn(c.z, c.x)), sw = __fmul_rn(tq, __fsub_rn(c.w, c.y));
    const bool ok = area > 1e-30f && fmaxf(fabsf(c.x), fabsf(c.z)) < __fmul_rn(65536.f, sh) &&
                    fmaxf(fabsf(c.y), fabsf(c.w)) < __fmul_rn(65536.f, sw);
    return ok ? make_float4(__fadd_rn(c.x, sh), __fadd_rn(c.y, sw), __fsub_rn(c.z, sh), __fsub_rn(c.w, sw)) : c;
  };
  // ---- 2. gather / decode boxes (evaluate.py:141-142), select-stage mask is 1 for all of them.
  // Only the first 256 candidates now; the greedy loop usually stops before it needs more (keep_top_k
  // survivors), and fetches further chunks on demand.  Until then s_area[j] parks the anchor index.
  auto gather = [&](int j, int i) {
    const int l = layer_of(P.L, i);
    const long long off = 4ll * (i - P.L.offset[l]);
    float4 v;
    if (P.has_loc) {
      v = ldg4(P.loc.base[l] + (long long)b * P.loc.stride[l] + off);
    } else {
      float4 o = ldg4(P.refine.base[l] + (long long)b * P.refine.stride[l] + off);
      const float4 d = ldg4(P.det.base[l] + (long long)b * P.det.stride[l] + off);
      o = make_float4(__fadd_rn(o.x, d.x), __fadd_rn(o.y, d.y), __fadd_rn(o.z, d.z), __fadd_rn(o.w, d.w));
      v = center_to_corner(decode_center(ldg4(P.center + 4ll * i), o));
    }
    const float4 nb = make_float4(fminf(v.x, v.z), fminf(v.y, v.w), fmaxf(v.x, v.z), fmaxf(v.y, v.w));
    const float area = __fmul_rn(__fsub_rn(nb.z, nb.x), __fsub_rn(nb.w, nb.y));
    s_box[j] = v;
    // boxes with area <= 0 never overlap anything (TF IOU returns 0): make them unreachable
    const float4 nbv = (area > 0.f) ? nb : make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    s_nbox[j] = nbv;
    s_qh[j] = pack_cand(shrunk(nbv, area));
    s_nh[j] = pack_row(nbv);
    s_area[j] = area;
  };
  int gathered = 0;
  for (int j = tid; j < m; j += kSegBlock) {
    const unsigned long long key = s_keys[j];
    s_score[j] = __uint_as_float((unsigned)(key >> 32));
    s_area[j] = __int_as_float((int)(~(unsigned)(key & 0xffffffffull)));
  }
  __syncthreads();                                    // keys are dead from here: region A becomes the mask

  SEG_T(3);
  // ---- 3. greedy NMS in batches of 64 candidates against the kept list.
  // A candidate is kept iff no earlier KEPT box has IoU > thr with it (tf.image.non_max_suppression).
  // Every lane owns two candidates of the batch (boxes + areas in registers).  Phase 1 (all warps):
  // the lane tests its candidates against the kept boxes of its warp's share (rows broadcast from
  // shared memory; exact 4-compare intersection predicate, IoU only when it holds) and against the
  // batch rows of its warp, accumulating "dead" bits and per-column suppressor masks in registers.
  // Phase 2 (warp 0): resolve the 64 x 64 intra-batch dependencies and append the survivors.
  const float4 none = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  auto suppresses = [&](const float4& r, float ar, const float4& c, float ac) -> bool {
    const float ih = __fsub_rn(fminf(r.z, c.z), fmaxf(r.x, c.x));
    const float iw = __fsub_rn(fminf(r.w, c.w), fmaxf(r.y, c.y));
    const float inter = __fmul_rn(ih, iw);
    const float den = __fsub_rn(__fadd_rn(ar, ac), inter);
    return (thr >= 0.f && den > 1e-30f) ? iou_exceeds(inter, den, thr) : (__fdiv_rn(inter, den) > thr);
  };
  int nk = 0;
  int bsz = 64;
  for (int p0 = 0; p0 < m && nk < keep; p0 += bsz) {
    // close to keep_top_k a half batch (one candidate per lane) is enough: the last full batch would
    // spend a quarter of all pair tests on a handful of survivors
    bsz = (keep - nk <= 24) ? 32 : 64;
    if (p0 + bsz > gathered && gathered < m) {           // block-uniform: fetch the next boxes (256 first, then 128 at a time)
      const int g1 = min(m, gathered + (gathered == 0 ? 256 : 128));
      for (int j = gathered + tid; j < g1; j += kSegBlock) gather(j, __float_as_int(s_area[j]));
      gathered = g1;
      __syncthreads();
    }
    const bool two = bsz == 64;
    const int nb = min(bsz, m - p0);
    const int c0 = p0 + lane, c1 = c0 + 32;
    const bool v1 = two && c1 < m;
    // the lane's two candidates side by side in half2 registers (shrunk boxes, computed once in the gather phase);
    // the IoU tests (rare) re-read box and area
    const uint2 qa = c0 < m ? s_qh[c0] : pack_cand(none);
    const uint2 qb = v1 ? s_qh[c1] : pack_cand(none);
    const __half2 qymin = u2h(__byte_perm(qa.x, qb.x, 0x5410)), qymax = u2h(__byte_perm(qa.x, qb.x, 0x7632));
    const __half2 qxmin = u2h(__byte_perm(qa.y, qb.y, 0x5410)), qxmax = u2h(__byte_perm(qa.y, qb.y, 0x7632));
    // -- vs the kept list: warp w takes rows w, w+8, ...  First a branch-free pass that only records which rows
    // can intersect the lane's candidates (pipelined broadcast loads, 16 rows per pass: bit i = candidate 0,
    // bit 16 + i = candidate 1), then the exact IoU test for the recorded rows only (a few per lane).
    bool d0 = false, d1 = false;
    for (int jb = warp; jb < nk; jb += 16 * kSegWarps) {
      unsigned h = 0u;
      const int ni = min(16, (nk - jb + kSegWarps - 1) / kSegWarps);      // rows of this pass (warp-uniform)
#pragma unroll 4
      for (int i = 0; i < ni; ++i) h |= pair_mask(s_kh[jb + i * kSegWarps], qymin, qymax, qxmin, qxmax) & (0x00010001u << i);
      unsigned h0 = h & 0xffffu, h1 = h >> 16;
      while (h0 && !d0) {
        const int p = s_selected[jb + (__ffs(h0) - 1) * kSegWarps];
        h0 &= h0 - 1;
        d0 = suppresses(s_nbox[p], s_area[p], s_nbox[c0], s_area[c0]);
      }
      while (h1 && !d1) {
        const int p = s_selected[jb + (__ffs(h1) - 1) * kSegWarps];
        h1 &= h1 - 1;
        d1 = suppresses(s_nbox[p], s_area[p], s_nbox[c1], s_area[c1]);
      }
    }
    // -- vs the batch itself: warp w takes rows 8w .. 8w+7; bit i of cm = row 8w+i suppresses my column
    unsigned cm0 = 0u, cm1 = 0u;
    {
      unsigned h = 0u;
      const int r0 = warp * (64 / kSegWarps);
#pragma unroll
      for (int rr = 0; rr < 64 / kSegWarps; ++rr)
        if (r0 + rr < nb) h |= pair_mask(s_nh[p0 + r0 + rr], qymin, qymax, qxmin, qxmax) & (0x00010001u << rr);
      // only EARLIER rows of the batch can suppress a candidate (rows >= 32 only the second candidate)
      unsigned h0 = h & 0xffffu, h1 = h >> 16;
      while (h0) {
        const int rr = __ffs(h0) - 1;
        h0 &= h0 - 1;
        if (r0 + rr < lane) cm0 |= (unsigned)suppresses(s_nbox[p0 + r0 + rr], s_area[p0 + r0 + rr], s_nbox[c0], s_area[c0]) << rr;
      }
      while (h1) {
        const int rr = __ffs(h1) - 1;
        h1 &= h1 - 1;
        if (r0 + rr < lane + 32) cm1 |= (unsigned)suppresses(s_nbox[p0 + r0 + rr], s_area[p0 + r0 + rr], s_nbox[c1], s_area[c1]) << rr;
      }
    }
    s_cmw[lane * kSegWarps + warp] = cm0;
    s_cmw[(lane + 32) * kSegWarps + warp] = cm1;
    {
      const unsigned long long dw = (unsigned long long)__ballot_sync(0xffffffffu, d0) |
                                    ((unsigned long long)__ballot_sync(0xffffffffu, d1) << 32);
      if (lane == 0) s_deadw[warp] = dw;
    }
    __syncthreads();                                   // partial masks of every warp visible
    // -- resolve (warp 0): candidate q is dead if a KEPT earlier candidate suppresses it, kept once no
    // undecided earlier candidate could; every round decides at least the lowest undecided one.
    if (warp == 0) {
      unsigned long long cmA = 0ull, cmB = 0ull;       // suppressor rows of my two candidates
#pragma unroll
      for (int w = 0; w < kSegWarps; ++w) {
        cmA |= (unsigned long long)s_cmw[lane * kSegWarps + w] << (w * (64 / kSegWarps));
        cmB |= (unsigned long long)s_cmw[(lane + 32) * kSegWarps + w] << (w * (64 / kSegWarps));
      }
      unsigned long long dead = 0ull;
#pragma unroll
      for (int w = 0; w < kSegWarps; ++w) dead |= s_deadw[w];
      unsigned long long U = ~dead;
      if (nb < 64) U &= (1ull << nb) - 1ull;
      unsigned long long K = 0ull;
      while (U) {
        const bool u0 = (U >> lane) & 1ull, u1 = (U >> (lane + 32)) & 1ull;
        const bool x0 = u0 && (cmA & K), x1 = u1 && (cmB & K);
        const bool k0 = u0 && !x0 && !(cmA & U), k1 = u1 && !x1 && !(cmB & U);
        const unsigned long long newK = (unsigned long long)__ballot_sync(0xffffffffu, k0) |
                                        ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32);
        const unsigned long long newD = (unsigned long long)__ballot_sync(0xffffffffu, x0) |
                                        ((unsigned long long)__ballot_sync(0xffffffffu, x1) << 32);
        K |= newK;
        U &= ~(newK | newD);
      }
      const int room = keep - nk;                     // greedy stops once keep boxes are selected
      if (__popcll(K) > room) {                       // keep only the first `room` survivors
        const unsigned lo32 = (unsigned)K, hi32 = (unsigned)(K >> 32);
        const int cl = __popc(lo32);
        const int pos = room <= cl ? (int)__fns(lo32, 0u, room) : 32 + (int)__fns(hi32, 0u, room - cl);
        K &= (pos >= 63) ? ~0ull : ((2ull << pos) - 1ull);
      }
      // append the survivors to the kept list, in order (lane owns candidates lane and lane + 32)
      if ((K >> lane) & 1ull) {
        const int pos = nk + __popcll(K & ((1ull << lane) - 1ull));
        s_selected[pos] = c0; s_kh[pos] = s_nh[c0];
      }
      if ((K >> (lane + 32)) & 1ull) {
        const int pos = nk + __popcll(K & ((1ull << (lane + 32)) - 1ull));
        s_selected[pos] = c1; s_kh[pos] = s_nh[c1];
      }
      if (lane == 0) s_sel = K;
    }
    __syncthreads();
    nk += __popcll(s_sel);
  }

  if (tid == 0) s_nsel = nk;
  __syncthreads();

  SEG_T(4);
  // ---- 5. emit keep rows: survivors in order, then pad_axis zeros (clip applies to all rows)
  const int nsel = s_nsel;
  float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (P.clip) cb = ldg4(P.clip);
  int nonzero = 0;
  for (int j = tid; j < keep; j += kSegBlock) {
    float sc = 0.f;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < nsel) {
      const int p = s_selected[j];
      sc = s_score[p];
      bx = s_box[p];
    }
    if (P.clip) {
      const float ymin = fmaxf(bx.x, cb.x), xmin = fmaxf(bx.y, cb.y), ymax = fminf(bx.z, cb.z), xmax = fminf(bx.w, cb.w);
      bx = make_float4(fminf(ymin, ymax), fminf(xmin, xmax), ymax, xmax);
    }
    __stcs(out_scores + r * keep + j, sc);
    st4_cs(out_boxes + 4 * (r * keep + j), bx);
    nonzero += (sc != 0.f) ? 1 : 0;
  }
  if (out_counts) {
    nonzero = __reduce_add_sync(0xffffffffu, nonzero);
    if (lane == 0) s_wsum[warp] = (unsigned)nonzero;
    __syncthreads();
    if (tid == 0) {
      unsigned total = 0;
      for (int w = 0; w < kSegWarps; ++w) total += s_wsum[w];
      out_counts[r] = (int)total;
    }
  }
  SEG_T(5);
  if (P.dbg != nullptr && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    P.dbg[r * 8 + 6] = nk;
    P.dbg[r * 8 + 7] = (long long)cnt | ((long long)smid << 32);        // (debug hook: which SM ran the segment)
  }
#undef SEG_T
}

static long long* g_seg_dbg = nullptr;
static size_t seg_smem_bytes(int cap, int k, int keep) {
  return seg_region_a(cap, k, keep) + (size_t)k * (16 + 16 + 4 + 4) + (size_t)keep * 4 + 16;
}
static int stream_cap(int k) {
  int p = 1;
  while (p < k) p <<= 1;
  return p * 2 < 1024 ? 1024 : p * 2;
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// bytes at the start of the streaming workspace that hold flags / counters / the sampled histogram
size_t stream_clean_bytes(int batch, int n_classes) {
  return (size_t)batch * n_classes * (4 + 4 + 4 + 4 + kSampleBins * 4) + 16;
}

size_t stream_workspace_bytes(int batch, int n_classes, int top_k) {
  const size_t rows = (size_t)batch * n_classes;
  return align256(rows * (4 + 4 + 4 + 4 + kSampleBins * 4 + kBins * 4) + 16) + align256(rows * 4) + 2 * align256(rows * kMaxChunks * 4) +
         2 * align256(rows * (size_t)kListCap * 8) + align256(rows * (size_t)kSpillCap * 8) +
         align256(rows * (size_t)stream_cap(top_k) * 8) + 256;
}

// Enqueues S, A, B.  *over_out (device, [rows]) is non-zero for segments the exact general
// kernels must redo.
int launch_detect_stream(const Layout& L, const float* anchors_center, const LayeredF& probs, const LayeredF* loc,
                         const LayeredF* refine, const LayeredF* det, int batch, int C, int logits, int ignore_class,
                         float select_thr, float nms_thr, int top_k, int keep, const float* clip, float* out_scores,
                         float* out_boxes, int32_t* out_counts, void* ws, const unsigned** over_out, int* cap_out,
                         const unsigned** any_out, cudaStream_t st) {
  ROD_REQUIRE(!logits || C == 11, "launch_detect_stream: fused softmax needs 11 classes (got %d)", C);
  const size_t rows = (size_t)batch * C;
  const int cap = stream_cap(top_k);
  unsigned char* p = reinterpret_cast<unsigned char*>(ws);
  // zero-initialised head of the workspace: over flags, dense counters, spill counters, sampled histogram (TMA
  // path) and the full histogram (plain-load path only)
  unsigned* g_over = reinterpret_cast<unsigned*>(p);
  unsigned* g_cnt2 = g_over + rows;
  unsigned* g_spill_cnt = g_cnt2 + rows;
  unsigned* g_lo_over = g_spill_cnt + rows;
  unsigned* g_any = g_lo_over + rows;                    // [4] (one flag, padded)
  unsigned* g_shist = g_any + 4;
  unsigned* g_hist = g_shist + rows * kSampleBins;
  const size_t zero_tma = rows * (4 + 4 + 4 + 4 + kSampleBins * 4) + 16, zero_all = zero_tma + rows * kBins * 4;
  p += align256(zero_all);
  int* g_tbin = reinterpret_cast<int*>(p);                     // plain-load path: threshold bins; TMA path: estimated cuts
  p += align256(rows * 4);
  unsigned* g_cnt1 = reinterpret_cast<unsigned*>(p);           // [rows][kMaxChunks], fully written by the scan pass
  p += align256(rows * kMaxChunks * 4);
  unsigned* g_cnt1b = reinterpret_cast<unsigned*>(p);          // the same for tier 2
  p += align256(rows * kMaxChunks * 4);
  unsigned long long* g_list = reinterpret_cast<unsigned long long*>(p);
  p += align256(rows * (size_t)kListCap * 8);
  unsigned long long* g_list_lo = reinterpret_cast<unsigned long long*>(p);
  p += align256(rows * (size_t)kListCap * 8);
  unsigned long long* g_spill = reinterpret_cast<unsigned long long*>(p);
  p += align256(rows * (size_t)kSpillCap * 8);
  unsigned long long* g_list2 = reinterpret_cast<unsigned long long*>(p);   // [rows][cap]
  int n_chunks = 1, spc = 512;
  // The zero-initialised head: prediction depth 11 with a sample pass keeps it clean itself (sample_kernel clears the
  // flags of the call, every segment CTA zeroes its histogram row for the next call) — no memset node per call.
  // The sampled histogram only steers the cuts, so a workspace that was never zeroed costs speed on its first
  // call (flagged segments go to the exact kernels), never results.
  bool self_cleaning = false;
  const size_t seg_smem = seg_smem_bytes(cap, top_k, keep);
  ROD_REQUIRE(seg_smem <= 220 * 1024, "rod_detect: top_k=%d keep=%d needs %zu B of shared memory", top_k, keep, seg_smem);
  ROD_CUDA(cudaFuncSetAttribute(segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem));
  auto fill_seg_params = [&](SegParams& G) {
    G.has_loc = loc ? 1 : 0;
    G.loc = loc ? *loc : *refine;
    G.refine = loc ? *loc : *refine;
    G.det = loc ? *loc : *det;
    G.center = anchors_center;
    G.L = L; G.batch = batch; G.ignore_class = ignore_class; G.cap = cap; G.k = top_k; G.keep = keep;
    G.nms_thr = nms_thr; G.clip = clip;
    G.cnt2 = g_cnt2; G.over = g_over; G.any = g_any; G.dbg = g_seg_dbg;
    G.list1 = g_list; G.cnt1 = g_cnt1; G.est = g_tbin; G.spill_cnt = g_spill_cnt; G.spill = g_spill;
    G.list2 = g_list_lo; G.cnt1b = g_cnt1b; G.lo_over = g_lo_over; G.shist = nullptr;
  };

  if (C == 11) {
    // ---- sampled pre-pass + TMA-staged single pass
    ScanParams SP;
    SP.probs = probs; SP.L = L; SP.batch = batch; SP.thr = select_thr;
    SP.ignore_class = (ignore_class >= 0 && ignore_class < 11) ? ignore_class : 31;   // 31: no class bit ever matches
    SP.g_shist = g_shist; SP.g_est = g_tbin; SP.g_cnt1 = g_cnt1; SP.g_over = g_over; SP.g_list = g_list;
    SP.g_spill_cnt = g_spill_cnt; SP.g_spill = g_spill;
    SP.g_cnt2 = g_cnt1b; SP.g_list2 = g_list_lo; SP.g_lo_over = g_lo_over; SP.g_any = g_any;
    // sampled anchors: kSamplePhase, kSamplePhase + kSampleStride, ...  Tier 1 keeps about 1.6 * top_k candidates per
    // segment (a too-high cut_hi only costs reading tier 2 as well), tiers 1 + 2 together about 4 * top_k: falling
    // below top_k there takes a 4x estimation error (> 8 sigma of the binomial sample at top_k = 400)
    const int sampled = L.n_total > kSamplePhase ? (L.n_total - kSamplePhase + kSampleStride - 1) / kSampleStride : 0;
    const int stiles = (sampled + 255) / 256;
    const double frac = L.n_total > 0 ? (double)sampled / (double)L.n_total : 0.0;
    SP.target_hi = (int)(1.6 * top_k * frac + 0.5);
    SP.target_lo = (int)(4.0 * top_k * frac + 0.5);
    const bool use_cut = SP.target_hi >= 24 && sampled >= 512;                    // too few samples: never cut
    if (!use_cut) SP.target_hi = SP.target_lo = 0x7fffffff;
    // (ROD_WS_MEMSET=1 brings the memset node back: with several independent calls in flight on different streams it
    // measured 2 us per step faster — 52.4 vs 54.7 — while one call at a time is 0.9 us slower, 66.5 vs 65.7)
    static const int ws_memset = [] { const char* e = getenv("ROD_WS_MEMSET"); return e ? atoi(e) : 0; }();
    self_cleaning = use_cut && !ws_memset;
    if (!self_cleaning) ROD_CUDA(cudaMemsetAsync(g_over, 0, zero_tma, st));
    SP.g_zero4 = g_over;
    int chunks = (4 * sm_count() + batch - 1) / batch;          // one wave of ~4 CTAs per SM (8 per SM measured slower)
    chunks = chunks < 8 ? 8 : (chunks > 32 ? 32 : chunks);     // >= 8: list slices of at most 512 entries
    int chunk = (L.n_total + chunks - 1) / chunks;
    chunk = ((chunk + kScanBlock - 1) / kScanBlock) * kScanBlock;
    chunks = (L.n_total + chunk - 1) / chunk;
    ROD_REQUIRE(chunks <= kMaxChunks, "rod_detect: %d anchors need more than %d CTAs per image", L.n_total, kMaxChunks);
    SP.chunk = chunk;
    SP.chunks = n_chunks = chunks;
    SP.spc = spc = kListCap / chunks < 512 ? kListCap / chunks : 512;
    SP.b0 = 0; SP.nb = batch;
    if (use_cut) {
      auto ks = logits ? sample_kernel<11, true> : sample_kernel<11, false>;
      ks<<<dim3(stiles, batch), 256, 0, st>>>(SP);
      ROD_LAUNCH_CHECK("sample_kernel");
    }
    // One stream; the segment and general kernels are programmatic launches (scheduled while the predecessor
    // drains, pdl_wait before they read its results); see launch_pdl for what that buys.  (Measured and dropped: persistent scan CTAs walking
    // the images in order with per-image completion counters so that segments start under the scan — the scan
    // needs its ~4 CTAs per SM, 108 vs 77 us; and scanning the batch in 2 / 4 / 8 slices with the segment
    // kernels forked to side streams through events — 103 / 120 / 165 us, the cross-stream edges cost more
    // than the overlap returns.)
    const size_t smem0 = (size_t)kScanStages * sizeof(float) * kScanBlock * 11 + sizeof(ScanShared<11>);
    auto k0 = logits ? scan_kernel<11, true, kScanStages> : scan_kernel<11, false, kScanStages>;
    ROD_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    ROD_CUDA(launch_pdl(1, k0, dim3(chunks, batch), dim3(kScanBlock), smem0, st, SP));
    SegParams G;
    fill_seg_params(G);
    G.chunks = n_chunks; G.spc = spc; G.force_dense = 0;
    G.shist = self_cleaning ? g_shist : nullptr;
    G.b0 = 0; G.nb = batch;
    ROD_CUDA(launch_pdl(2, segment_kernel, dim3((unsigned)rows), dim3(kSegBlock), seg_smem, st, G, (const unsigned long long*)g_list2,
                        out_scores, out_boxes, out_counts));
  } else {
    ROD_CUDA(cudaMemsetAsync(g_over, 0, zero_all, st));
    // ---- generic prediction depth: plain-load two-pass kernels, every segment takes the dense route
    StreamParams SP;
    SP.probs = probs; SP.L = L; SP.C = C; SP.ignore_class = ignore_class; SP.batch = batch; SP.thr = select_thr;
    const int total_f = L.n_total * C;
    int chunks = (4 * sm_count() + batch - 1) / batch;
    chunks = chunks < 4 ? 4 : (chunks > 64 ? 64 : chunks);
    int chunk = (total_f + chunks - 1) / chunks;
    chunk = ((chunk + 1023) / 1024) * 1024;
    chunks = (total_f + chunk - 1) / chunk;
    SP.chunk = chunk;
    const dim3 grid(chunks, batch);
    const size_t hsmem = (size_t)C * kBins * 4;
    ROD_REQUIRE(hsmem <= 200 * 1024, "rod_detect: %d classes need %zu B of shared memory for the histogram", C, hsmem);
    ROD_CUDA(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
    hist_kernel<<<grid, kStreamBlock, hsmem, st>>>(SP, g_hist);
    ROD_LAUNCH_CHECK("hist_kernel");
    thresh_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(g_hist, (int)rows, top_k, g_tbin);
    ROD_LAUNCH_CHECK("thresh_kernel");
    collect_kernel<<<grid, kStreamBlock, 0, st>>>(SP, g_tbin, g_cnt2, g_list2, cap);
    ROD_LAUNCH_CHECK("collect_kernel");
    SegParams G;
    fill_seg_params(G);
    G.chunks = 1; G.spc = 512; G.force_dense = 1;
    G.b0 = 0; G.nb = batch;
    segment_kernel<<<(unsigned)rows, kSegBlock, seg_smem, st>>>(G, g_list2, out_scores, out_boxes, out_counts);
    ROD_LAUNCH_CHECK("segment_kernel");
  }
  *over_out = g_over;
  *any_out = g_any;
  *cap_out = 0;                                                 // fallback kernels run where over[r] > 0
  return ROD_OK;
}

}  // namespace rod

// Debug hook (not part of the public header): per-segment clock64() stamps of the segment kernel's
// phases, 8 x int64 per (class, image) row.  Pass NULL to switch it off.
extern "C" void rod_debug_set_timing(long long* device_buf) { rod::g_seg_dbg = device_buf; }

