// Measurement helpers for bench.py: an FP32 non-FMA issue-rate probe (the ARM matching loop
// may not use FMA, so its FP32 roofline denominator is the FADD/FMUL rate) and an L2 flush.
#include "common.cuh"

namespace rod {

constexpr int kPeakBlock = 256;
constexpr int kPeakChains = 8;

__global__ void __launch_bounds__(kPeakBlock)
fp32_nofma_kernel(int iters, float* sink) {
  float v[kPeakChains];
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) v[c] = 1.0f + 1e-3f * (float)(threadIdx.x + c);
  const float m = 1.0000001f, a = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) {
      v[c] = __fmul_rn(v[c], m);
      v[c] = __fadd_rn(v[c], a);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) s += v[c];
  if (s == 123.456f) sink[0] = s;   // never true; keeps the chains alive
}

__global__ void flush_kernel(float4* buf, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace rod

extern "C" int rod_peak_fp32_nofma(int iters, float* sink, double* ops, void* stream) {
  using namespace rod;
  ROD_REQUIRE(iters > 0 && sink, "rod_peak_fp32_nofma: bad arguments");
  const int blocks = sm_count() * 8;
  fp32_nofma_kernel<<<blocks, kPeakBlock, 0, (cudaStream_t)stream>>>(iters, sink);
  ROD_LAUNCH_CHECK("fp32_nofma_kernel");
  if (ops) *ops = (double)blocks * kPeakBlock * (double)iters * kPeakChains * 2.0;
  return ROD_OK;
}

extern "C" int rod_l2_flush(void* buf, size_t bytes, void* stream) {
  using namespace rod;
  ROD_REQUIRE(buf && (bytes % 16) == 0, "rod_l2_flush: buffer must be non-NULL and a multiple of 16 bytes");
  flush_kernel<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(buf), bytes / 16);
  ROD_LAUNCH_CHECK("flush_kernel");
  return ROD_OK;
}
