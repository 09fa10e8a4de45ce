// Shared device/host helpers for the rodet_b200 kernels (sm_100a only).
//
// Numerics contract: every parity-critical float op is written with
// the round-to-nearest intrinsics __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn, which nvcc
// never contracts into FMAs, so the result equals TF/Eigen's one-rounding-per-op CPU
// evaluation.  The translation units are additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/rodet_b200.h"

namespace rod {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define ROD_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      rod::set_error(__VA_ARGS__);                  \
      return ROD_E_INVALID;                         \
    }                                               \
  } while (0)

#define ROD_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return rod::cuda_fail(_e, #expr);        \
  } while (0)

#define ROD_LAUNCH_CHECK(name)                                      \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return rod::cuda_fail(_e, name);         \
  } while (0)

int sm_count();

// Programmatic dependent launch (PDL): the kernel may be scheduled while its predecessor in the stream is still
// running (as soon as every CTA of the predecessor has executed pdl_launch_dependents() or exited); it must call
// pdl_wait() before touching anything the predecessor writes.  Hides the launch latency and the prologue of the
// short kernels of one rod_detect call; stream capture turns these launches into programmatic graph edges.
// `which` selects the launch class: 1 scan pass, 2 segment kernel, 4 general (fallback) kernels.  Default mask 6:
// measured on B200 (decode_nms, B = 64), the attribute on the segment / fallback launches is neutral to slightly
// positive (75.0 vs 75.6 us per step), on the scan launch it costs 10 us (85 us: the scan CTAs become resident
// under the sample kernel and the step gets slower, not faster).  ROD_PDL_MASK overrides it (measurement aid).
int pdl_mask();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & which) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Per-layer pointer table passed by value to kernels.
struct LayeredF {
  const float* base[ROD_MAX_LAYERS];
  long long stride[ROD_MAX_LAYERS];
};
struct LayeredI {
  const int32_t* base[ROD_MAX_LAYERS];
  long long stride[ROD_MAX_LAYERS];
};
struct Layout {
  int n_layers;
  int n_total;
  int offset[ROD_MAX_LAYERS + 1];
};

inline Layout to_layout(const rod_layout_t* l) {
  Layout o;
  o.n_layers = l->n_layers;
  o.n_total = l->n_total;
  for (int i = 0; i <= ROD_MAX_LAYERS; ++i) o.offset[i] = i <= l->n_layers ? l->offset[i] : l->n_total;
  return o;
}
int check_layout(const rod_layout_t* l);
inline LayeredF to_layered_f(const rod_layered_t* t, int n_layers) {
  LayeredF o;
  for (int i = 0; i < ROD_MAX_LAYERS; ++i) {
    o.base[i] = i < n_layers ? static_cast<const float*>(t->base[i]) : nullptr;
    o.stride[i] = i < n_layers ? t->batch_stride[i] : 0;
  }
  return o;
}
inline LayeredI to_layered_i(const rod_layered_t* t, int n_layers) {
  LayeredI o;
  for (int i = 0; i < ROD_MAX_LAYERS; ++i) {
    o.base[i] = i < n_layers ? static_cast<const int32_t*>(t->base[i]) : nullptr;
    o.stride[i] = i < n_layers ? t->batch_stride[i] : 0;
  }
  return o;
}
int check_layered(const rod_layered_t* t, int n_layers, const char* name);

struct Thresholds {
  float v[ROD_MAX_LAYERS];
};

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ld.shared on a 32-bit shared-window address
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// layer of flat anchor n (< n_total).  to_layout() pads offset[i] = n_total for i > n_layers, so the
// unused entries never count.
__device__ __forceinline__ int layer_of(const Layout& L, int n) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < ROD_MAX_LAYERS; ++i) l += (n >= L.offset[i]) ? 1 : 0;
  return l;
}

// warp-wide min / max of a float through the integer REDUX unit (order-preserving key transform)
__device__ __forceinline__ int float_order_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float order_key_float(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }
__device__ __forceinline__ float warp_min(float f) {
  return order_key_float(__reduce_min_sync(0xffffffffu, float_order_key(f)));
}
__device__ __forceinline__ float warp_max(float f) {
  return order_key_float(__reduce_max_sync(0xffffffffu, float_order_key(f)));
}

// centerBboxes_2_cornerBboxes, utils/common_tools.py:28-31 (h / 2 is exact in binary fp)
__device__ __forceinline__ float4 center_to_corner(float4 c) {
  const float hh = __fmul_rn(c.z, 0.5f), hw = __fmul_rn(c.w, 0.5f);
  return make_float4(__fsub_rn(c.x, hh), __fsub_rn(c.y, hw), __fadd_rn(c.x, hh), __fadd_rn(c.y, hw));
}
// cornerBboxes_2_centerBboxes, utils/common_tools.py:51-54
__device__ __forceinline__ float4 corner_to_center(float4 b) {
  return make_float4(__fmul_rn(__fadd_rn(b.x, b.z), 0.5f), __fmul_rn(__fadd_rn(b.y, b.w), 0.5f),
                     __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// area term of net_tools.jaccard for the anchor side: (xmax-xmin)*(ymax-ymin), :254
__device__ __forceinline__ float box_vol(float4 b) {
  return __fmul_rn(__fsub_rn(b.w, b.y), __fsub_rn(b.z, b.x));
}
// net_tools.jaccard, utils/net_tools.py:254-266: union = (vol_a - inter) + area_g, plain divide
__device__ __forceinline__ float jaccard_ref(float4 a, float vol_a, float4 g, float area_g) {
  const float h = fmaxf(__fsub_rn(fminf(a.z, g.z), fmaxf(a.x, g.x)), 0.f);
  const float w = fmaxf(__fsub_rn(fminf(a.w, g.w), fmaxf(a.y, g.y)), 0.f);
  const float inter = __fmul_rn(h, w);
  const float uni = __fadd_rn(__fsub_rn(vol_a, inter), area_g);
  return __fdiv_rn(inter, uni);
}
// correctly rounded float32 exp / log (see oracle/restated.py "exp / log policy")
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_cr(float x) { return (float)log((double)x); }

// slim.softmax over the class axis (evaluate.py:136-137, predict.py:127-128), one definition for the
// stand-alone kernel, the fused select and the general fallback so that all three produce the same
// bits: e_c = ex2((x_c - max) * log2(e)) with the hardware ex2 (2 ulp), s = e_0 + e_1 + ... in class
// order, p_c = e_c * (1 / s).  Relative error vs an exact softmax < 1e-6 (tests: tolerance 1e-5).
__device__ __forceinline__ float softmax_exp(float x, float m) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(__fsub_rn(x, m), 1.44269504088896340736f)));
  return r;
}
// probability of class c of one row of n logits (generic class count; used by the fallback path)
__device__ __forceinline__ float softmax_pick(const float* __restrict__ row, int n, int c) {
  float m = __ldg(row);
  for (int j = 1; j < n; ++j) m = fmaxf(m, __ldg(row + j));
  float s = softmax_exp(__ldg(row), m);
  for (int j = 1; j < n; ++j) s = __fadd_rn(s, softmax_exp(__ldg(row + j), m));
  return __fmul_rn(softmax_exp(__ldg(row + c), m), __frcp_rn(s));
}

// decode_locations_one_layer, utils/net_tools.py:226-229 (a = acy,acx,ah,aw)
__device__ __forceinline__ float4 decode_center(float4 a, float4 o) {
  return make_float4(__fadd_rn(__fmul_rn(o.x, a.z), a.x), __fadd_rn(__fmul_rn(o.y, a.w), a.y),
                     __fmul_rn(exp_cr(o.z), a.z), __fmul_rn(exp_cr(o.w), a.w));
}
// encode_locations_one_layer, utils/net_tools.py:174-177 (a = acy,acx,ah,aw ; g = gcy,gcx,gh,gw)
__device__ __forceinline__ float4 encode_center(float4 a, float4 g) {
  return make_float4(__fdiv_rn(__fsub_rn(g.x, a.x), a.z), __fdiv_rn(__fsub_rn(g.y, a.y), a.w),
                     log_cr(__fdiv_rn(g.z, a.z)), log_cr(__fdiv_rn(g.w, a.w)));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) store for write-once outputs
__device__ __forceinline__ void st4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// order-preserving float -> uint key (ascending); +-0 are canonicalised so they tie
__device__ __forceinline__ uint32_t float_key(float s) {
  if (s == 0.f) s = 0.f;
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

}  // namespace rod
