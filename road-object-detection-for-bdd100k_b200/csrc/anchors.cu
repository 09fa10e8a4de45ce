// a1-a4: anchor table generation (utils/net_tools.py:21-142) and the corner /
// re-derived centre forms every consumer recomputes (utils/net_tools.py:156-171).
// Init-time only; one thread per anchor.
#include "common.cuh"

namespace rod {

struct AnchorGenParams {
  int n_layers;
  int feat_h[ROD_MAX_LAYERS], feat_w[ROD_MAX_LAYERS], n_anchor[ROD_MAX_LAYERS];
  int offset[ROD_MAX_LAYERS + 1];       // anchor offsets
  int size_offset[ROD_MAX_LAYERS + 1];  // offsets into the per-layer (h,w) size table
  int cell_offset[ROD_MAX_LAYERS + 1];  // offsets into the concatenated cell grids
  int img_h, img_w;
};
constexpr int kMaxSizes = 96;   // sum over layers of anchors per cell (reference: 6+5*9 = 51)
struct AnchorSizes {
  double hw[kMaxSizes][2];      // init_anchor() output, pixel (h, w), float64
};

__device__ __forceinline__ void write_forms(float y, float x, float h, float w, int n,
                                            float* corner, float* center, float* yxhw) {
  const float4 cr = center_to_corner(make_float4(y, x, h, w));   // net_tools.py:157-165
  // net_tools.py:168-171: centre / size re-derived from the float32 corners
  const float4 ce = make_float4(__fmul_rn(__fadd_rn(cr.z, cr.x), 0.5f), __fmul_rn(__fadd_rn(cr.w, cr.y), 0.5f),
                                __fsub_rn(cr.z, cr.x), __fsub_rn(cr.w, cr.y));
  st4(corner + 4ll * n, cr);
  st4(center + 4ll * n, ce);
  if (yxhw) st4(yxhw + 4ll * n, make_float4(y, x, h, w));
}

__global__ void anchor_table_kernel(AnchorGenParams P, AnchorSizes S,
                                    float* corner, float* center, float* yxhw) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= P.offset[P.n_layers]) return;
  int l = 0;
  for (int i = 1; i < P.n_layers; ++i) l += (n >= P.offset[i]) ? 1 : 0;
  const int r = n - P.offset[l];
  const int A = P.n_anchor[l];
  const int a = r % A, cell = r / A;
  const int fx = cell % P.feat_w[l], fy = cell / P.feat_w[l];
  // anchors_one_layer, net_tools.py:113-122: float64 math, then astype(float32)
  const float y = (float)(((double)fy + 0.5) / (double)P.feat_h[l]);
  const float x = (float)(((double)fx + 0.5) / (double)P.feat_w[l]);
  const float h = (float)(S.hw[P.size_offset[l] + a][0] / (double)P.img_h);
  const float w = (float)(S.hw[P.size_offset[l] + a][1] / (double)P.img_w);
  write_forms(y, x, h, w, n, corner, center, yxhw);
}

__global__ void anchor_table_from_grid_kernel(AnchorGenParams P, const float* __restrict__ ys,
                                              const float* __restrict__ xs, const float* __restrict__ hs,
                                              const float* __restrict__ ws, float* corner, float* center) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= P.offset[P.n_layers]) return;
  int l = 0;
  for (int i = 1; i < P.n_layers; ++i) l += (n >= P.offset[i]) ? 1 : 0;
  const int r = n - P.offset[l];
  const int A = P.n_anchor[l];
  const int a = r % A, cell = r / A;
  write_forms(ys[P.cell_offset[l] + cell], xs[P.cell_offset[l] + cell], hs[P.size_offset[l] + a],
              ws[P.size_offset[l] + a], n, corner, center, nullptr);
}

static int fill_params(AnchorGenParams& P, int n_layers, const int32_t* fh, const int32_t* fw,
                       const int32_t* na, int img_h, int img_w) {
  ROD_REQUIRE(n_layers >= 1 && n_layers <= ROD_MAX_LAYERS, "n_layers=%d not in [1,%d]", n_layers, ROD_MAX_LAYERS);
  ROD_REQUIRE(fh && fw && na, "feat_h/feat_w/n_anchor must not be NULL");
  P.n_layers = n_layers;
  P.offset[0] = P.size_offset[0] = P.cell_offset[0] = 0;
  for (int l = 0; l < n_layers; ++l) {
    ROD_REQUIRE(fh[l] > 0 && fw[l] > 0 && na[l] > 0, "layer %d has a non-positive size", l);
    P.feat_h[l] = fh[l]; P.feat_w[l] = fw[l]; P.n_anchor[l] = na[l];
    P.offset[l + 1] = P.offset[l] + fh[l] * fw[l] * na[l];
    P.size_offset[l + 1] = P.size_offset[l] + na[l];
    P.cell_offset[l + 1] = P.cell_offset[l] + fh[l] * fw[l];
  }
  P.img_h = img_h; P.img_w = img_w;
  return ROD_OK;
}

}  // namespace rod

extern "C" int rod_anchor_table(int n_layers, const int32_t* feat_h, const int32_t* feat_w,
                                const int32_t* n_anchor, const double* sizes_px, int img_h, int img_w,
                                float* corner, float* center, float* yxhw, void* stream) {
  rod::AnchorGenParams P;
  int rc = rod::fill_params(P, n_layers, feat_h, feat_w, n_anchor, img_h, img_w);
  if (rc) return rc;
  ROD_REQUIRE(sizes_px && corner && center, "sizes_px/corner/center must not be NULL");
  ROD_REQUIRE(img_h > 0 && img_w > 0, "image size must be positive");
  // sizes_px is a small HOST array: it travels in the kernel parameter block.
  const int n_sizes = P.size_offset[n_layers];
  ROD_REQUIRE(n_sizes <= rod::kMaxSizes, "too many anchor shapes (%d > %d)", n_sizes, rod::kMaxSizes);
  rod::AnchorSizes S;
  for (int i = 0; i < n_sizes; ++i) { S.hw[i][0] = sizes_px[2 * i]; S.hw[i][1] = sizes_px[2 * i + 1]; }
  const int n = P.offset[n_layers];
  rod::anchor_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, S, corner, center, yxhw);
  ROD_LAUNCH_CHECK("anchor_table_kernel");
  return ROD_OK;
}

extern "C" int rod_anchor_table_from_grid(int n_layers, const int32_t* feat_h, const int32_t* feat_w,
                                          const int32_t* n_anchor, const float* y, const float* x,
                                          const float* h, const float* w, float* corner, float* center,
                                          void* stream) {
  rod::AnchorGenParams P;
  int rc = rod::fill_params(P, n_layers, feat_h, feat_w, n_anchor, 1, 1);
  if (rc) return rc;
  ROD_REQUIRE(y && x && h && w && corner && center, "NULL pointer argument");
  const int n = P.offset[n_layers];
  rod::anchor_table_from_grid_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, y, x, h, w, corner, center);
  ROD_LAUNCH_CHECK("anchor_table_from_grid_kernel");
  return ROD_OK;
}
