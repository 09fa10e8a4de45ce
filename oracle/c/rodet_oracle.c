/*
 * CPU oracle, C port ("Tier C").  TEST INFRASTRUCTURE ONLY: used by tests/ as a fast checker and by
 * bench.py as the timed CPU baseline / reference arm (all host threads via OpenMP).  Never linked
 * into or called by the product library.
 *
 * Plain-C restatement of the reference's box-level hot path, one rounding per reference op in the
 * reference's op order (build with -ffp-contract=off; no -ffast-math).  Each function cites the
 * reference lines it follows (/root/reference/...).  exp / log are evaluated in double and rounded
 * to float (see oracle/restated.py "exp / log policy").  Checked bit-for-bit against
 * oracle/restated.py (tests/test_oracle_c.py), which in turn is pinned to the reference fixtures.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float exp32(float x) { return (float)exp((double)x); }
static inline float log32(float x) { return (float)log((double)x); }

/* centerBboxes_2_cornerBboxes, utils/common_tools.py:28-31 */
static inline void c2c(const float* c, float* o) {
  const float hh = c[2] / 2.f, hw = c[3] / 2.f;
  o[0] = c[0] - hh; o[1] = c[1] - hw; o[2] = c[0] + hh; o[3] = c[1] + hw;
}

/* net_tools.jaccard, utils/net_tools.py:254-266 */
static inline float jaccard(const float* a, float vol_a, const float* g, float area_g) {
  const float iymin = a[0] > g[0] ? a[0] : g[0], ixmin = a[1] > g[1] ? a[1] : g[1];
  const float iymax = a[2] < g[2] ? a[2] : g[2], ixmax = a[3] < g[3] ? a[3] : g[3];
  float h = iymax - iymin, w = ixmax - ixmin;
  h = h > 0.f ? h : 0.f; w = w > 0.f ? w : 0.f;
  const float inter = h * w;
  const float uni = (vol_a - inter) + area_g;
  return inter / uni;
}

/* decode_locations_one_layer, utils/net_tools.py:226-229 (a = acy,acx,ah,aw) */
static inline void decode(const float* a, const float* o, float* out) {
  out[0] = o[0] * a[2] + a[0];
  out[1] = o[1] * a[3] + a[1];
  out[2] = exp32(o[2]) * a[2];
  out[3] = exp32(o[3]) * a[3];
}

/* refine_groundtruth / JACCARD_BIGGER, utils/net_tools.py:382-421 + 316-343, batched over images.
 * corner/center: [N,4] anchor tables; thr: [N] per-anchor layer threshold; boxes [B,gmax,4] centre
 * form; labels [B,gmax] int64; counts [B].  Outputs flat [B,N,*]. */
void orc_arm_match_encode(const float* corner, const float* center, const float* thr, int n, const float* boxes,
                          const int64_t* labels, const int32_t* counts, int batch, int gmax, float* gt, float* cb,
                          int32_t* lab, int32_t* pos, int32_t* idx) {
#pragma omp parallel
  {
    float* gcorner = (float*)malloc(sizeof(float) * 5 * (size_t)gmax);
    for (int b = 0; b < batch; ++b) {
      const int g_n = counts[b];
      const float* gb = boxes + (size_t)b * gmax * 4;
      for (int g = 0; g < g_n; ++g) {                       /* every thread keeps its own copy */
        c2c(gb + 4 * g, gcorner + 5 * g);
        gcorner[5 * g + 4] = (gcorner[5 * g + 2] - gcorner[5 * g]) * (gcorner[5 * g + 3] - gcorner[5 * g + 1]);
      }
#pragma omp for schedule(static) nowait
      for (int i = 0; i < n; ++i) {
        const float* a = corner + 4 * (size_t)i;
        const float vol_a = (a[3] - a[1]) * (a[2] - a[0]);
        float best = 0.f; int bi = 0;
        for (int g = 0; g < g_n; ++g) {
          const float j = jaccard(a, vol_a, gcorner + 5 * g, gcorner[5 * g + 4]);
          if (g == 0 || j > best) { best = j; bi = g; }      /* reduce_max / first argmax, :405-408 */
        }
        const size_t o = (size_t)b * n + i;
        const int p = g_n > 0 && best >= thr[i];
        idx[o] = bi; pos[o] = p;
        if (p) {
          const float* c = center + 4 * (size_t)i; const float* m = gb + 4 * bi;
          gt[4 * o + 0] = ((m[0] - c[0]) / c[2]) + 0.f;       /* :174-177; 0 + 1*v accumulate, :340-341 */
          gt[4 * o + 1] = ((m[1] - c[1]) / c[3]) + 0.f;
          gt[4 * o + 2] = log32(m[2] / c[2]) + 0.f;
          gt[4 * o + 3] = log32(m[3] / c[3]) + 0.f;
          for (int k = 0; k < 4; ++k) cb[4 * o + k] = m[k] + 0.f;
          lab[o] = (int32_t)labels[(size_t)b * gmax + bi];
        } else {
          for (int k = 0; k < 4; ++k) { gt[4 * o + k] = 0.f; cb[4 * o + k] = 0.f; }
          lab[o] = 0;
        }
      }
#pragma omp barrier
    }
    free(gcorner);
  }
}

/* det_groundtruth, utils/net_tools.py:454-473; all arrays flat [B,N,*] */
void orc_odm_target(const float* center, const float* thr, int n, int batch, const float* refine_out,
                    const float* offset_gt, const float* cbboxes, const int32_t* labels, const int32_t* posm,
                    float* det_gt, int32_t* mask, int32_t* det_labels, float* iou) {
  const long long total = (long long)batch * n;
#pragma omp parallel for schedule(static)
  for (long long o = 0; o < total; ++o) {
    const int i = (int)(o % n);
    float d[4], ra[4], gc[4];
    decode(center + 4 * (size_t)i, refine_out + 4 * o, d);
    c2c(d, ra);
    c2c(cbboxes + 4 * o, gc);
    const float vol_a = (ra[3] - ra[1]) * (ra[2] - ra[0]);
    const float j = jaccard(ra, vol_a, gc, (gc[2] - gc[0]) * (gc[3] - gc[1]));
    const int m = (j >= thr[i] ? 1 : 0) * posm[o];
    const float mf = (float)m;
    for (int k = 0; k < 4; ++k) det_gt[4 * o + k] = (offset_gt[4 * o + k] - refine_out[4 * o + k]) * mf;
    mask[o] = m; det_labels[o] = labels[o] * m; iou[o] = j;
  }
}

/* c2c(decode(anchors, refine_out + det_out)), evaluate.py:139-143 */
void orc_decode_corner(const float* center, int n, int batch, const float* refine_out, const float* det_out,
                       float* out) {
  const long long total = (long long)batch * n;
#pragma omp parallel for schedule(static)
  for (long long o = 0; o < total; ++o) {
    float s[4], d[4];
    for (int k = 0; k < 4; ++k) s[k] = refine_out[4 * o + k] + det_out[4 * o + k];
    decode(center + 4 * (size_t)(o % n), s, d);
    c2c(d, out + 4 * o);
  }
}

typedef struct { float s; int i; } cand_t;
static int cand_cmp(const void* pa, const void* pb) {     /* descending score, then ascending index */
  const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return a->i - b->i;
}

/* tf.image.non_max_suppression IoU (see oracle/tf_shim) */
static inline float nms_iou(const float* a, const float* b) {
  const float ay0 = a[0] < a[2] ? a[0] : a[2], ay1 = a[0] > a[2] ? a[0] : a[2];
  const float ax0 = a[1] < a[3] ? a[1] : a[3], ax1 = a[1] > a[3] ? a[1] : a[3];
  const float by0 = b[0] < b[2] ? b[0] : b[2], by1 = b[0] > b[2] ? b[0] : b[2];
  const float bx0 = b[1] < b[3] ? b[1] : b[3], bx1 = b[1] > b[3] ? b[1] : b[3];
  const float aa = (ay1 - ay0) * (ax1 - ax0), ab = (by1 - by0) * (bx1 - bx0);
  if (aa <= 0.f || ab <= 0.f) return 0.f;
  const float iy0 = ay0 > by0 ? ay0 : by0, ix0 = ax0 > bx0 ? ax0 : bx0;
  const float iy1 = ay1 < by1 ? ay1 : by1, ix1 = ax1 < bx1 ? ax1 : bx1;
  float h = iy1 - iy0, w = ix1 - ix0;
  h = h > 0.f ? h : 0.f; w = w > 0.f ? w : 0.f;
  const float inter = h * w;
  return inter / ((aa + ab) - inter);
}

/* detected_bboxes, utils/net_tools.py:739-758: select (:686-695) -> tf.nn.top_k (bboxes.py:86) ->
 * NMS + pad (bboxes.py:180-189) for classes 1..n_classes-1.  probs [B,N,C], boxes [B,N,4] corner.
 * out_scores [C,B,keep], out_boxes [C,B,keep,4] (class 0 untouched). */
void orc_detect(const float* probs, const float* boxes, int batch, int n, int n_classes, float select_thr,
                float nms_thr, int top_k, int keep, float* out_scores, float* out_boxes) {
  const int jobs = batch * (n_classes - 1);
#pragma omp parallel
  {
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
    float* kb = (float*)malloc(sizeof(float) * 4 * (size_t)top_k);
    int* sel = (int*)malloc(sizeof(int) * (size_t)keep);
#pragma omp for schedule(dynamic, 1)
    for (int job = 0; job < jobs; ++job) {
      const int b = job / (n_classes - 1), c = 1 + job % (n_classes - 1);
      const float* p = probs + (size_t)b * n * n_classes;
      const float* bx = boxes + (size_t)b * n * 4;
      for (int i = 0; i < n; ++i) {
        const float s = p[(size_t)i * n_classes + c];
        const float fm = s >= select_thr ? 1.f : 0.f;
        cand[i].s = s * fm; cand[i].i = i;
      }
      qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);     /* total order => deterministic top_k */
      for (int j = 0; j < top_k; ++j) {
        const int i = cand[j].i;
        const float fm = p[(size_t)i * n_classes + c] >= select_thr ? 1.f : 0.f;
        for (int k = 0; k < 4; ++k) kb[4 * j + k] = bx[4 * (size_t)i + k] * fm;
      }
      int ns = 0;
      for (int j = 0; j < top_k && ns < keep; ++j) {
        int ok = 1;
        for (int q = ns - 1; q >= 0; --q)
          if (nms_iou(kb + 4 * j, kb + 4 * sel[q]) > nms_thr) { ok = 0; break; }
        if (ok) sel[ns++] = j;
      }
      float* os = out_scores + ((size_t)c * batch + b) * keep;
      float* ob = out_boxes + ((size_t)c * batch + b) * keep * 4;
      for (int j = 0; j < keep; ++j) {
        if (j < ns) {
          os[j] = cand[sel[j]].s;
          memcpy(ob + 4 * j, kb + 4 * sel[j], sizeof(float) * 4);
        } else {
          os[j] = 0.f; ob[4 * j] = ob[4 * j + 1] = ob[4 * j + 2] = ob[4 * j + 3] = 0.f;
        }
      }
    }
    free(cand); free(kb); free(sel);
  }
}

void orc_set_threads(int n) {
#ifdef _OPENMP
  extern void omp_set_num_threads(int);
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
