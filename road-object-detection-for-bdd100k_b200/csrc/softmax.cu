// f-1: slim.softmax over the class axis (evaluate.py:136-137, predict.py:127-128) as a stand-alone
// kernel.  rod_detect_logits fuses the same arithmetic (softmax_exp / __frcp_rn, common.cuh) into the
// select pass, so rod_softmax(x) followed by rod_detect gives bit-identical detections.
#include "common.cuh"

namespace rod {

constexpr int kSoftmaxBlock = 256;

// 256 rows per tile: coalesced load into shared memory (row stride C | 1: conflict-free), one thread
// per row, coalesced store.  HBM-bound: 8 * C bytes per row.
__global__ void __launch_bounds__(kSoftmaxBlock)
softmax_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C) {
  extern __shared__ float s_x[];
  const int stride = C | 1, tid = threadIdx.x;
  for (long long r0 = (long long)blockIdx.x * kSoftmaxBlock; r0 < rows; r0 += (long long)gridDim.x * kSoftmaxBlock) {
    const int nrow = (int)(rows - r0 < kSoftmaxBlock ? rows - r0 : kSoftmaxBlock), total = nrow * C;
    const float* src = in + r0 * C;
    for (int i = tid; i < total; i += kSoftmaxBlock) {
      const int row = i / C;
      s_x[row * stride + (i - row * C)] = __ldg(src + i);
    }
    __syncthreads();
    if (tid < nrow) {
      float* x = s_x + tid * stride;
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
    float* dst = out + r0 * C;
    for (int i = tid; i < total; i += kSoftmaxBlock) {
      const int row = i / C;
      __stcs(dst + i, s_x[row * stride + (i - row * C)]);
    }
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
  const size_t smem = (size_t)kSoftmaxBlock * (n_classes | 1) * sizeof(float);
  ROD_CUDA(cudaFuncSetAttribute(softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (rows + kSoftmaxBlock - 1) / kSoftmaxBlock;
  const long long cap = 8ll * sm_count();
  softmax_kernel<<<(unsigned)(tiles < cap ? tiles : cap), kSoftmaxBlock, smem, (cudaStream_t)stream>>>(logits, out, rows, n_classes);
  ROD_LAUNCH_CHECK("softmax_kernel");
  return ROD_OK;
}
