// a13 + a15: batched per-class greedy NMS (tf.image.non_max_suppression semantics, see
// oracle/tf_shim) with zero padding to keep_top_k, and the fused post-process driver.
// Replaces tfe.bboxes_nms / bboxes_nms_batch (utils/tf_extended/bboxes.py:166-232),
// pad_axis (utils/tf_extended/tensors.py:59-86), tfe.bboxes_clip (:103-136) and
// detected_bboxes (utils/net_tools.py:739-758).
//
// One CTA per (class, image) problem of n <= 1024 candidates:
//   1. candidates -> shared memory in NMS visiting order (score desc, position asc);
//      the fused path receives them already in that order from the top-k kernel and gathers /
//      decodes their boxes on the fly (only top_k boxes per segment are ever decoded);
//   2. upper-triangular suppression bitmask, 64 candidates per 64-bit word;
//   3. one warp sweeps the candidates in order, OR-ing mask rows of the survivors;
//   4. survivors are written in selection order, zero padded.
// Candidates whose box was zeroed by the select stage ("dummies": score*0, box*0) have zero
// area, so they neither suppress nor get suppressed (TF: area <= 0 -> IoU 0); the mask is only
// built up to the last non-dummy candidate.
#include "select_topk.cuh"

namespace rod {

constexpr int kNmsBlock = 256;

// IoU of tensorflow/core/kernels/non_max_suppression_op.cc on pre-normalised boxes
// (b = ymin,xmin,ymax,xmax with min<=max) and their areas; true iff iou > thr.
__device__ __forceinline__ bool nms_overlaps(float4 a, float area_a, float4 b, float area_b, float thr) {
  if (area_a <= 0.f || area_b <= 0.f) return false;
  const float ih = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float iw = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(ih, iw);
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return iou > thr;
}

struct DenseBoxes {        // standalone: boxes[rows, n, 4], scores[rows, n] in arbitrary order
  const float* scores;
  const float* boxes;
};
struct GatherBoxes {       // fused: candidates come sorted from the top-k kernel
  const float* scores;     // [rows, n]
  const int32_t* idx;      // [rows, n] anchor index, bit 31 = dummy
  LayeredF loc;            // corner boxes (has_loc) ...
  LayeredF refine, det;    // ... or offsets to decode
  const float* center;     // anchors (acy,acx,ah,aw)
  Layout L;
  int has_loc, batch;
  const unsigned* over_cnt;   // when set: only segments whose streaming list overflowed run here
  const unsigned* over_any;   // when set: non-zero iff any segment is flagged
  unsigned over_cap;
  // fuse_topk: the CTA first computes the row's top-k itself (topk_row) into scores_w / idx_w (= scores / idx) — the
  // flagged-segment path of rod_detect, where a separate top-k launch would cost ~1.5 us per call just to find no work
  int fuse_topk;
  SelectedScores topk_src;
  float* scores_w;
  int32_t* idx_w;
};

template <bool FUSED>
__global__ void __launch_bounds__(kNmsBlock)
nms_kernel(const __grid_constant__ DenseBoxes dsrc, const __grid_constant__ GatherBoxes gsrc, long long rows, int n, float thr, int keep, int ignore_class,
           const float* __restrict__ clip, float* __restrict__ out_scores, float* __restrict__ out_boxes,
           int32_t* __restrict__ out_idx, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  // layout: raw boxes | normalised boxes | mask words | sort keys (dense only) | area | score | src | selected
  const int W = (n + 63) >> 6;
  float4* s_box = reinterpret_cast<float4*>(s_raw);                                       // [n] as given
  float4* s_nbox = s_box + n;                                                            // [n] min/max normalised
  unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_nbox + n);        // [n*W]
  unsigned long long* s_sort = s_mask + (size_t)n * W;                                   // [1024] (dense only)
  float* s_area = reinterpret_cast<float*>(s_sort + (FUSED ? 0 : 1024));                 // [n]
  float* s_score = s_area + n;                                                           // [n]
  int* s_src = reinterpret_cast<int*>(s_score + n);                                      // [n] source position
  int* s_selected = s_src + n;                                                           // [keep]
  __shared__ int s_last, s_nsel;

  if (FUSED) {
    pdl_wait();                                        // launched while the top-k kernel still runs
    if (gsrc.over_any != nullptr && *gsrc.over_any == 0u) return;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // persistent over rows (the fused path launches a small grid that usually finds nothing to do)
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
  if (FUSED && (int)(r / gsrc.batch) == ignore_class) continue;
  if (FUSED && gsrc.over_cnt != nullptr && gsrc.over_cnt[r] <= gsrc.over_cap) continue;
  if constexpr (FUSED) {
    if (gsrc.fuse_topk) topk_row<SelectedScores, kNmsBlock>(gsrc.topk_src, r, n, gsrc.scores_w, gsrc.idx_w, nullptr, nullptr);
  }
  __syncthreads();
  if (tid == 0) { s_last = 0; s_nsel = 0; }
  __syncthreads();

  // ---------------- 1. load candidates in visiting order
  if (FUSED) {
    const int b = (int)(r % gsrc.batch);
    for (int j = tid; j < n; j += kNmsBlock) {
      // (written by this CTA a moment ago when fuse_topk: not through the read-only path)
      const float sc = gsrc.fuse_topk ? __ldcg(gsrc.scores + r * n + j) : __ldg(gsrc.scores + r * n + j);
      const int raw = gsrc.fuse_topk ? __ldcg(gsrc.idx + r * n + j) : __ldg(gsrc.idx + r * n + j);
      const bool dummy = raw < 0;
      const int i = raw & 0x7fffffff;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      const int l = layer_of(gsrc.L, i);
      const long long off = 4ll * (i - gsrc.L.offset[l]);
      float4 v;
      if (gsrc.has_loc) {
        v = ldg4(gsrc.loc.base[l] + (long long)b * gsrc.loc.stride[l] + off);
      } else if (!dummy) {
        float4 o = ldg4(gsrc.refine.base[l] + (long long)b * gsrc.refine.stride[l] + off);
        const float4 d = ldg4(gsrc.det.base[l] + (long long)b * gsrc.det.stride[l] + off);
        o = make_float4(__fadd_rn(o.x, d.x), __fadd_rn(o.y, d.y), __fadd_rn(o.z, d.z), __fadd_rn(o.w, d.w));
        v = center_to_corner(decode_center(ldg4(gsrc.center + 4ll * i), o));
      } else {
        v = bx;
      }
      // select stage: bboxes * fmask (utils/net_tools.py:692)
      const float fm = dummy ? 0.f : 1.f;
      bx = (gsrc.has_loc || !dummy) ? make_float4(__fmul_rn(v.x, fm), __fmul_rn(v.y, fm), __fmul_rn(v.z, fm), __fmul_rn(v.w, fm)) : bx;
      s_box[j] = bx;
      s_score[j] = sc;
      s_src[j] = dummy ? -1 : j;
      if (!dummy) atomicMax(&s_last, j + 1);
    }
  } else {
    // sort positions by (score desc, position asc) - bitonic on (key, ~pos)
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int j = tid; j < n2; j += kNmsBlock)
      s_sort[j] = j < n ? (((unsigned long long)float_key(__ldg(dsrc.scores + r * n + j)) << 32) | (unsigned)(~(unsigned)j)) : 0ull;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (n2 >> 1); t += kNmsBlock) {
          const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long a = s_sort[lo], c = s_sort[hi];
          if ((a < c) == desc) { s_sort[lo] = c; s_sort[hi] = a; }
        }
        __syncthreads();
      }
    for (int j = tid; j < n; j += kNmsBlock) {
      const int p = (int)(~(unsigned)(s_sort[j] & 0xffffffffull));
      s_box[j] = ldg4(dsrc.boxes + 4 * (r * n + p));
      s_score[j] = __ldg(dsrc.scores + r * n + p);
      s_src[j] = p;
    }
    if (tid == 0) s_last = n;
  }
  __syncthreads();
  const int last = s_last;              // candidates >= last are dummies (always selected)
  const int Wl = (last + 63) >> 6;
  for (int j = tid; j < last; j += kNmsBlock) {
    const float4 b = s_box[j];
    const float4 nb = make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
    s_nbox[j] = nb;
    s_area[j] = __fmul_rn(__fsub_rn(nb.z, nb.x), __fsub_rn(nb.w, nb.y));
  }
  __syncthreads();

  // ---------------- 2. suppression bitmask (upper triangle)
  for (int t = tid; t < last * Wl; t += kNmsBlock) {
    const int i = t / Wl, w = t - i * Wl;
    unsigned long long bits = 0ull;
    if (w >= (i >> 6)) {
      const float4 bi = s_nbox[i];
      const float ai = s_area[i];
      const int j0 = w << 6;
      const int jbeg = max(j0, i + 1), jend = min(j0 + 64, last);
      for (int j = jbeg; j < jend; ++j)
        if (nms_overlaps(bi, ai, s_nbox[j], s_area[j], thr)) bits |= 1ull << (j - j0);
    }
    s_mask[(size_t)i * Wl + w] = bits;
  }
  __syncthreads();

  // ---------------- 3. greedy sweep by warp 0 (lane w owns dead-word w)
  if (warp == 0) {
    unsigned long long dead = 0ull;
    int nsel = 0;
    for (int p = 0; p < last && nsel < keep; ++p) {
      const unsigned long long dw = __shfl_sync(0xffffffffu, dead, p >> 6);
      if (!((dw >> (p & 63)) & 1ull)) {
        if (lane == 0) s_selected[nsel] = p;
        ++nsel;
        if (lane < Wl) dead |= s_mask[(size_t)p * Wl + lane];
      }
    }
    for (int p = last + lane; p < n; p += 32) {     // trailing dummies fill the remaining slots in order
      const int slot = nsel + (p - last);
      if (slot < keep) s_selected[slot] = p;
    }
    nsel = min(keep, nsel + (n - last));
    if (lane == 0) s_nsel = nsel;
  }
  __syncthreads();

  // ---------------- 4. emit: gather + pad_axis zero padding (+ optional clip)
  const int nsel = s_nsel;
  float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (clip) cb = ldg4(clip);
  int nonzero = 0;
  for (int j = tid; j < keep; j += kNmsBlock) {
    float sc = 0.f;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    int src = -1;
    if (j < nsel) {
      const int p = s_selected[j];
      sc = s_score[p];
      bx = s_box[p];
      src = s_src[p];
    }
    if (clip) {   // tfe.bboxes_clip, utils/tf_extended/bboxes.py:128-135 (applied to padding rows too)
      const float ymin = fmaxf(bx.x, cb.x), xmin = fmaxf(bx.y, cb.y), ymax = fminf(bx.z, cb.z), xmax = fminf(bx.w, cb.w);
      bx = make_float4(fminf(ymin, ymax), fminf(xmin, xmax), ymax, xmax);
    }
    out_scores[r * keep + j] = sc;
    st4(out_boxes + 4 * (r * keep + j), bx);
    if (out_idx) out_idx[r * keep + j] = src;
    nonzero += (sc != 0.f) ? 1 : 0;
  }
  if (out_counts) {
    nonzero = __reduce_add_sync(0xffffffffu, nonzero);
    if (lane == 0 && nonzero) atomicAdd(out_counts + r, nonzero);
  }
  }  // rows
}

static size_t nms_smem_bytes(int n, int keep, bool dense) {
  const int W = (n + 63) >> 6;
  size_t b = (size_t)n * W * 8 + (size_t)n * (16 + 16 + 4 + 4 + 4) + (size_t)keep * 4 + 16;
  if (dense) b += 1024 * 8;
  return b;
}

int launch_topk_selected(const SelectedScores& src, long long rows, int k, float* out_scores, int32_t* out_idx,
                         cudaStream_t st);
size_t stream_workspace_bytes(int batch, int n_classes, int top_k);
size_t stream_clean_bytes(int batch, int n_classes);
int launch_detect_stream(const Layout& L, const float* anchors_center, const LayeredF& probs, const LayeredF* loc,
                         const LayeredF* refine, const LayeredF* det, int batch, int C, int logits, int ignore_class,
                         float select_thr, float nms_thr, int top_k, int keep, const float* clip, float* out_scores,
                         float* out_boxes, int32_t* out_counts, void* ws, const unsigned** cnt_out, int* cap_out,
                         const unsigned** any_out, cudaStream_t st);

}  // namespace rod

extern "C" int rod_bboxes_nms_batch(const float* scores, const float* bboxes, int64_t rows, int n,
                                    float nms_threshold, int keep_top_k, float* out_scores,
                                    float* out_bboxes, int32_t* out_idx, void* stream) {
  using namespace rod;
  ROD_REQUIRE(scores && bboxes && out_scores && out_bboxes, "rod_bboxes_nms_batch: NULL pointer argument");
  ROD_REQUIRE(rows >= 0 && rows < 2147483647ll && n >= 1 && keep_top_k >= 1, "rod_bboxes_nms_batch: rows=%lld n=%d keep=%d invalid",
              (long long)rows, n, keep_top_k);
  if (n > ROD_MAX_TOPK) {
    set_error("rod_bboxes_nms_batch: n=%d > %d candidates per row is not supported", n, ROD_MAX_TOPK);
    return ROD_E_UNSUPPORTED;
  }
  if (rows == 0) return ROD_OK;
  // pad_axis never truncates: the output holds max(keep_top_k, #selected) rows; #selected <= keep_top_k
  const size_t smem = nms_smem_bytes(n, keep_top_k, true);
  ROD_REQUIRE(smem <= 220 * 1024, "rod_bboxes_nms_batch: n=%d keep=%d needs %zu B of shared memory", n, keep_top_k, smem);
  auto k = nms_kernel<false>;
  ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DenseBoxes d{scores, bboxes};
  GatherBoxes g{};
  k<<<(unsigned)rows, kNmsBlock, smem, (cudaStream_t)stream>>>(d, g, rows, n, nms_threshold, keep_top_k, -1, nullptr,
                                                              out_scores, out_bboxes, out_idx, nullptr);
  ROD_LAUNCH_CHECK("nms_kernel<dense>");
  return ROD_OK;
}

extern "C" size_t rod_detect_workspace_bytes(const rod_layout_t* layout, int batch, int n_classes, int top_k) {
  (void)layout;
  if (batch <= 0 || n_classes <= 0 || top_k <= 0) return 256;
  const size_t per = (size_t)batch * n_classes * top_k;
  return ((per * 4 + 255) / 256) * 256 * 2 + 256 + rod::stream_workspace_bytes(batch, n_classes, top_k);
}

extern "C" size_t rod_detect_flags_offset(const rod_layout_t* layout, int batch, int n_classes, int top_k) {
  (void)layout;
  if (batch <= 0 || n_classes <= 0 || top_k <= 0) return 0;
  const size_t per = (size_t)batch * n_classes * top_k;
  return ((per * 4 + 255) / 256) * 256 * 2;       // the streaming workspace starts with the flags (launch_detect_stream)
}

extern "C" size_t rod_detect_workspace_clean_bytes(const rod_layout_t* layout, int batch, int n_classes, int top_k) {
  (void)layout; (void)top_k;
  if (batch <= 0 || n_classes <= 0) return 0;
  return rod::stream_clean_bytes(batch, n_classes);
}

namespace rod {
static int detect_impl(const rod_layout_t* layout, const float* anchors_center,
                          const rod_layered_t* predictions, const rod_layered_t* localizations,
                          const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                          int n_classes, int ignore_class, float select_threshold, float nms_threshold,
                          int top_k, int keep_top_k, const float* clip_box, float* out_scores,
                          float* out_bboxes, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                          void* stream, int logits) {
  int rc = check_layout(layout);
  if (rc) return rc;
  const int nl = layout->n_layers;
  if ((rc = check_layered(predictions, nl, "predictions"))) return rc;
  if (localizations) {
    if ((rc = check_layered(localizations, nl, "localizations"))) return rc;
  } else {
    ROD_REQUIRE(refine_out && det_out && anchors_center, "rod_detect: need localizations, or refine_out + det_out + anchors");
    if ((rc = check_layered(refine_out, nl, "refine_out"))) return rc;
    if ((rc = check_layered(det_out, nl, "det_out"))) return rc;
  }
  ROD_REQUIRE(out_scores && out_bboxes && workspace, "rod_detect: NULL output / workspace");
  ROD_REQUIRE(batch >= 0 && n_classes >= 1 && n_classes <= ROD_MAX_CLASSES, "rod_detect: batch=%d n_classes=%d invalid", batch, n_classes);
  ROD_REQUIRE(batch <= 65535, "rod_detect: batch=%d exceeds 65535 (grid y dimension)", batch);
  ROD_REQUIRE(keep_top_k >= 1, "rod_detect: keep_top_k=%d invalid", keep_top_k);
  ROD_REQUIRE(top_k >= 1 && top_k <= layout->n_total, "rod_detect: top_k=%d must be in [1, N=%d] (tf.nn.top_k requires k <= N)", top_k, layout->n_total);
  if (top_k > ROD_MAX_TOPK) {
    set_error("rod_detect: top_k=%d > %d is not supported", top_k, ROD_MAX_TOPK);
    return ROD_E_UNSUPPORTED;
  }
  ROD_REQUIRE(workspace_bytes >= rod_detect_workspace_bytes(layout, batch, n_classes, top_k), "rod_detect: workspace too small");
  ROD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "rod_detect: workspace must be 256-byte aligned");
  if (batch == 0) return ROD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const Layout L = to_layout(layout);
  const long long rows = (long long)n_classes * batch;
  const size_t per = (((size_t)rows * top_k * 4 + 255) / 256) * 256;
  float* ws_scores = reinterpret_cast<float*>(workspace);
  int32_t* ws_idx = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + per);

  // ---- fast path (select_threshold > 0): streaming histogram select + per-segment sort/decode/NMS
  const LayeredF probs_l = to_layered_f(predictions, nl);
  const unsigned* over_cnt = nullptr;
  const unsigned* over_any = nullptr;
  int over_cap = 0;
  // (fused softmax is specialised for 11 classes; other depths take the general path below)
  if (select_threshold > 0.f && (!logits || n_classes == 11)) {
    const LayeredF loc_l = localizations ? to_layered_f(localizations, nl) : probs_l;
    const LayeredF ref_l = localizations ? probs_l : to_layered_f(refine_out, nl);
    const LayeredF det_l = localizations ? probs_l : to_layered_f(det_out, nl);
    void* ws_stream = reinterpret_cast<unsigned char*>(workspace) + 2 * per;
    if ((rc = launch_detect_stream(L, anchors_center, probs_l, localizations ? &loc_l : nullptr,
                                   localizations ? nullptr : &ref_l, localizations ? nullptr : &det_l, batch,
                                   n_classes, logits, ignore_class, select_threshold, nms_threshold, top_k, keep_top_k,
                                   clip_box, out_scores, out_bboxes, out_counts, ws_stream, &over_cnt, &over_cap, &over_any, st)))
      return rc;
  } else if (out_counts) {
    ROD_CUDA(cudaMemsetAsync(out_counts, 0, sizeof(int32_t) * rows, st));
  }

  // ---- general exact path: every segment when thr <= 0, else only segments whose list overflowed
  SelectedScores src;
  src.probs = probs_l;
  src.L = L;
  src.n_classes = n_classes;
  src.ignore_class = ignore_class;
  src.batch = batch;
  src.logits = logits;
  src.thr = select_threshold;
  src.over_cnt = over_cnt;
  src.over_any = over_any;
  src.over_cap = (unsigned)over_cap;
  // flagged segments of the fast path: the NMS kernel's CTA computes the row's top-k itself (one launch instead of two)
  const bool fuse_topk = over_cnt != nullptr;
  if (!fuse_topk && (rc = launch_topk_selected(src, rows, top_k, ws_scores, ws_idx, st))) return rc;

  GatherBoxes g;
  g.scores = ws_scores;
  g.idx = ws_idx;
  g.has_loc = localizations ? 1 : 0;
  g.loc = localizations ? to_layered_f(localizations, nl) : to_layered_f(refine_out, nl);
  g.refine = localizations ? g.loc : to_layered_f(refine_out, nl);
  g.det = localizations ? g.loc : to_layered_f(det_out, nl);
  g.center = anchors_center;
  g.L = L;
  g.batch = batch;
  g.over_cnt = over_cnt;
  g.over_any = over_any;
  g.over_cap = (unsigned)over_cap;
  g.fuse_topk = fuse_topk ? 1 : 0;
  g.topk_src = src;
  g.scores_w = ws_scores;
  g.idx_w = ws_idx;
  const size_t smem = nms_smem_bytes(top_k, keep_top_k, false);
  ROD_REQUIRE(smem <= 220 * 1024, "rod_detect: top_k=%d keep=%d needs %zu B of shared memory", top_k, keep_top_k, smem);
  auto k = nms_kernel<true>;
  ROD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DenseBoxes d{};
  const unsigned grid = over_cnt ? (unsigned)(rows < 4 * sm_count() ? rows : 4 * sm_count()) : (unsigned)rows;
  ROD_CUDA(launch_pdl(4, k, dim3(grid), dim3(kNmsBlock), smem, st, d, g, rows, top_k, nms_threshold, keep_top_k, ignore_class,
                      clip_box, out_scores, out_bboxes, (int32_t*)nullptr, out_counts));
  return ROD_OK;
}
}  // namespace rod

extern "C" int rod_detect(const rod_layout_t* layout, const float* anchors_center,
                          const rod_layered_t* predictions, const rod_layered_t* localizations,
                          const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                          int n_classes, int ignore_class, float select_threshold, float nms_threshold,
                          int top_k, int keep_top_k, const float* clip_box, float* out_scores,
                          float* out_bboxes, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                          void* stream) {
  return rod::detect_impl(layout, anchors_center, predictions, localizations, refine_out, det_out, batch, n_classes,
                          ignore_class, select_threshold, nms_threshold, top_k, keep_top_k, clip_box, out_scores,
                          out_bboxes, out_counts, workspace, workspace_bytes, stream, 0);
}

extern "C" int rod_detect_logits(const rod_layout_t* layout, const float* anchors_center,
                                 const rod_layered_t* logits, const rod_layered_t* localizations,
                                 const rod_layered_t* refine_out, const rod_layered_t* det_out, int batch,
                                 int n_classes, int ignore_class, float select_threshold, float nms_threshold,
                                 int top_k, int keep_top_k, const float* clip_box, float* out_scores,
                                 float* out_bboxes, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  return rod::detect_impl(layout, anchors_center, logits, localizations, refine_out, det_out, batch, n_classes,
                          ignore_class, select_threshold, nms_threshold, top_k, keep_top_k, clip_box, out_scores,
                          out_bboxes, out_counts, workspace, workspace_bytes, stream, 1);
}
