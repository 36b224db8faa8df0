// K13 / K14: the tail of the step (SURVEY.md section 8f rank 2 remainder).
//
//   K13  linear_f32_fwd / _bwd   logits = pooled . W^T + b                      (self.head, HWGATE.py:359)
//        The pooled vector is (B, 512) fp32 and the class count is 262 .. 2002: 0.1 - 1 GFLOP, far below anything a
//        tensor-core tile amortises, and the logits feed a log-softmax, so it runs on the blocked-summation FFMA GEMM
//        of the fp32 parity path (attn_f32.cu) - in fp32 also under autocast.
//   K14  smooth_ce_fwd / _bwd    loss = mean_b [(1-s) * nll_b + s * (-mean_c logp_bc)]   (SmoothCrossEntropy.py:35-39)
//        One warp per row: max, sum-exp and the class-mean in one pass over the row (registers for <= 4096 classes are
//        not needed: three shuffles-reductions over a strided loop), per-row loss to a scratch vector, then a
//        single-CTA fixed-order sum (deterministic).  Backward: dlogits = g/B * (softmax - (1-s) onehot - s/C).
#include "common.cuh"

namespace hwgat {

HW_DEV float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
HW_DEV float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per row: lse, row loss
__global__ void __launch_bounds__(256) smooth_ce_fwd_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ target,
                                                            float* __restrict__ lse, float* __restrict__ row_loss,
                                                            int rows, int classes, float smooth) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* z = logits + (size_t)row * classes;
  float m = -INFINITY, tot = 0.f;
  for (int c = lane; c < classes; c += 32) { const float v = z[c]; m = fmaxf(m, v); tot += v; }
  m = warp_max_f(m);
  tot = warp_sum_f(tot);
  float se = 0.f;
  for (int c = lane; c < classes; c += 32) se += __expf(z[c] - m);
  se = warp_sum_f(se);
  const float l = m + logf(se);
  if (lane == 0) {
    const long long t = target[row];
    const float zt = (t >= 0 && t < classes) ? z[t] : 0.f;
    const float nll = l - zt;                       // -logp[target]
    const float uni = l - tot / (float)classes;     // -mean_c logp
    lse[row] = l;
    row_loss[row] = (1.f - smooth) * nll + smooth * uni;
  }
}

// fixed-order sum of the row losses (one CTA): loss = sum / rows
__global__ void __launch_bounds__(256) smooth_ce_mean_kernel(const float* __restrict__ row_loss, float* __restrict__ loss,
                                                             int rows) {
  __shared__ float red[256];
  float a = 0.f;
  for (int i = threadIdx.x; i < rows; i += 256) a += row_loss[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] / (float)rows;
}

__global__ void __launch_bounds__(256) smooth_ce_bwd_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ target,
                                                            const float* __restrict__ lse, const float* __restrict__ g,
                                                            float* __restrict__ dlogits, int rows, int classes,
                                                            float smooth) {
  const long long total = (long long)rows * classes;
  const float gs = *g / (float)rows, sc = smooth / (float)classes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / classes), c = (int)(i - (long long)row * classes);
    const float p = __expf(logits[i] - lse[row]);
    const float hot = target[row] == c ? 1.f - smooth : 0.f;
    dlogits[i] = gs * (p - hot - sc);
  }
}

int smooth_ce_fwd(const float* logits, const long long* target, float* lse, float* row_loss, float* loss, int rows,
                  int classes, float smooth, cudaStream_t s) {
  smooth_ce_fwd_kernel<<<(rows + 7) / 8, 256, 0, s>>>(logits, target, lse, row_loss, rows, classes, smooth);
  count_launch();
  smooth_ce_mean_kernel<<<1, 256, 0, s>>>(row_loss, loss, rows);
  count_launch();
  return (int)cudaGetLastError();
}

int smooth_ce_bwd(const float* logits, const long long* target, const float* lse, const float* g, float* dlogits,
                  int rows, int classes, float smooth, cudaStream_t s) {
  const long long total = (long long)rows * classes;
  long long want = (total + 255) / 256;
  const int grid = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
  smooth_ce_bwd_kernel<<<grid, 256, 0, s>>>(logits, target, lse, g, dlogits, rows, classes, smooth);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
