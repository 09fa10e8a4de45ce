// f-1: slim.softmax over the class axis (evaluate.py:136-137, predict.py:127-128) as a stand-alone
// kernel.  rod_detect_logits fuses the same arithmetic (softmax_exp / __frcp_rn, common.cuh) into the
// select pass, so rod_softmax(x) followed by rod_detect gives bit-identical detections.
#include "common.cuh"

namespace rod {

constexpr int kSoftmaxBlock = 256;

// 256 rows per tile: the tile's 256 * C floats are copied linearly into shared memory (float4 when the
// tensor is 16 B aligned), one thread per row works on its C words in place (stride C: conflict-free
// for odd C), and the tile is copied back linearly.  HBM-bound: 8 * C bytes per row.
__global__ void __launch_bounds__(kSoftmaxBlock)
softmax_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C, int vec) {
  extern __shared__ __align__(16) float s_x[];
  const int tid = threadIdx.x;
  for (long long r0 = (long long)blockIdx.x * kSoftmaxBlock; r0 < rows; r0 += (long long)gridDim.x * kSoftmaxBlock) {
    const int nrow = (int)(rows - r0 < kSoftmaxBlock ? rows - r0 : kSoftmaxBlock), total = nrow * C;
    const float* src = in + r0 * C;
    float* dst = out + r0 * C;
    const int nv = vec ? total >> 2 : 0;               // r0 * C is a multiple of 256 floats: tiles stay aligned
    for (int i = tid; i < nv; i += kSoftmaxBlock) reinterpret_cast<float4*>(s_x)[i] = ldg4(src + 4 * i);
    for (int i = 4 * nv + tid; i < total; i += kSoftmaxBlock) s_x[i] = __ldg(src + i);
    __syncthreads();
    if (tid < nrow) {
      float* x = s_x + tid * C;
      float m = x[0];
      for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) {
        const float e = softmax_exp(x[c], m);
        x[c] = e;
        sum = c == 0 ? e : __fadd_rn(sum, e);
      }
      const float rinv = __frcp_rn(sum);
      for (int c = 0; c < C; ++c) x[c] = __fmul_rn(x[c], rinv);
    }
    __syncthreads();
    for (int i = tid; i < nv; i += kSoftmaxBlock) st4_cs(dst + 4 * i, reinterpret_cast<const float4*>(s_x)[i]);
    for (int i = 4 * nv + tid; i < total; i += kSoftmaxBlock) __stcs(dst + i, s_x[i]);
    __syncthreads();
  }
}

}  // namespace rod

extern "C" int rod_softmax(const float* logits, int64_t rows, int n_classes, float* out, void* stream) {
  using namespace rod;
  ROD_REQUIRE(logits && out, "rod_softmax: NULL pointer argument");
  ROD_REQUIRE(rows >= 0 && n_classes >= 1 && n_classes <= ROD_MAX_CLASSES, "rod_softmax: rows=%lld n_classes=%d invalid",
              (long long)rows, n_classes);
  if (rows == 0) return ROD_OK;
  const size_t smem = (size_t)kSoftmaxBlock * n_classes * sizeof(float);
  ROD_CUDA(cudaFuncSetAttribute(softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (rows + kSoftmaxBlock - 1) / kSoftmaxBlock;
  const long long cap = 8ll * sm_count();
  softmax_kernel<<<(unsigned)(tiles < cap ? tiles : cap), kSoftmaxBlock, smem, (cudaStream_t)stream>>>(logits, out, rows, n_classes,
      ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 ? 1 : 0);
  ROD_LAUNCH_CHECK("softmax_kernel");
  return ROD_OK;
}
