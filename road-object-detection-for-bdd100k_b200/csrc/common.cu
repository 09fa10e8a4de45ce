// Error slot, device info and small host helpers of the C ABI.
#include <stdlib.h>

#include "common.cuh"

#include <string.h>

namespace rod {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return ROD_E_CUDA;
}

int pdl_mask() {
  static const int mask = [] {
    const char* e = getenv("ROD_PDL_MASK");
    return e ? atoi(e) : 6;
  }();
  return mask;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

int check_layout(const rod_layout_t* l) {
  ROD_REQUIRE(l != nullptr, "layout is NULL");
  ROD_REQUIRE(l->n_layers >= 1 && l->n_layers <= ROD_MAX_LAYERS, "layout.n_layers=%d not in [1,%d]",
              l->n_layers, ROD_MAX_LAYERS);
  ROD_REQUIRE(l->offset[0] == 0, "layout.offset[0] must be 0");
  for (int i = 0; i < l->n_layers; ++i)
    ROD_REQUIRE(l->offset[i + 1] > l->offset[i], "layout.offset must be strictly increasing (layer %d)", i);
  ROD_REQUIRE(l->offset[l->n_layers] == l->n_total, "layout.n_total=%d != offset[n_layers]=%d",
              l->n_total, l->offset[l->n_layers]);
  return ROD_OK;
}

int check_layered(const rod_layered_t* t, int n_layers, const char* name) {
  ROD_REQUIRE(t != nullptr, "%s is NULL", name);
  for (int i = 0; i < n_layers; ++i) {
    ROD_REQUIRE(t->base[i] != nullptr, "%s.base[%d] is NULL", name, i);
    ROD_REQUIRE((reinterpret_cast<uintptr_t>(t->base[i]) & 3u) == 0, "%s.base[%d] not 4-byte aligned", name, i);
  }
  return ROD_OK;
}

}  // namespace rod

extern "C" {

const char* rod_last_error(void) { return rod::g_err; }

int rod_version(void) { return ROD_ABI_VERSION; }

int rod_device_info(int* sm, int* major, int* minor, size_t* smem_optin) {
  int dev = 0;
  ROD_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm) { ROD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
  if (major) { ROD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *major = v; }
  if (minor) { ROD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *minor = v; }
  if (smem_optin) { ROD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *smem_optin = (size_t)v; }
  return ROD_OK;
}

}  // extern "C"
