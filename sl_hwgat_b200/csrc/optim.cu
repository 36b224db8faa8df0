// K11: multi-tensor AdamW, the optimizer the reference trains with (utils.py:73-75: torch.optim.AdamW(params,
// lr=cfg.lr), i.e. betas (0.9, 0.999), eps 1e-8, weight_decay 0.01, no amsgrad; stepped at utils.py:107).
//
// PyTorch's default path launches ~10 elementwise kernels per parameter tensor (or a few foreach kernels); here ONE
// launch walks every tensor of a chunk of up to 160 parameters: the device pointers travel in the kernel-argument
// struct, a block finds its (tensor, offset) by bisection over the block prefix sums, and every element makes one
// round trip: read p, g, m, v; write p, m, v (28 bytes per parameter, HBM-bound).
//
//   p <- p (1 - lr wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)              (torch/optim/adamw.py, single_tensor)
#include "common.cuh"

namespace hwgat {

constexpr int kOptMaxTensors = 160;
constexpr int kOptBlockElems = 256 * 4 * 4;  // 256 threads x 4 float4

struct AdamwChunk {
  float* p[kOptMaxTensors];
  const float* g[kOptMaxTensors];
  float* m[kOptMaxTensors];
  float* v[kOptMaxTensors];
  long long n[kOptMaxTensors];
  int first_block[kOptMaxTensors + 1];  // prefix sums of ceil(n / kOptBlockElems)
  int count;
};

struct AdamwHyper {
  float decay;      // 1 - lr * weight_decay
  float b1, b2;
  float omb1, omb2; // 1 - b1, 1 - b2, formed in double on the host as PyTorch forms them (1.f - 0.999f is off by 5e-5)
  float step_size;  // lr / (1 - b1^t)
  float inv_bc2;    // 1 / sqrt(1 - b2^t)
  float eps;
  float grad_scale; // multiplies every gradient first (1/world_size for a summed all-reduce, loss-scale inverse, ...)
};

HW_DEV void adamw_elem(float& p, float g, float& m, float& v, const AdamwHyper& h) {
  g *= h.grad_scale;
  p *= h.decay;
  m = fmaf(h.b1, m, h.omb1 * g);
  v = fmaf(h.b2, v, h.omb2 * g * g);
  const float denom = fmaf(sqrtf(v), h.inv_bc2, h.eps);
  p -= h.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamwChunk c, const AdamwHyper h) {
  int lo = 0, hi = c.count;  // first_block[lo] <= blockIdx.x < first_block[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x >= c.first_block[mid]) lo = mid; else hi = mid;
  }
  const long long n = c.n[lo];
  const long long base = (long long)(blockIdx.x - c.first_block[lo]) * kOptBlockElems;
  float* __restrict__ p = c.p[lo];
  const float* __restrict__ g = c.g[lo];
  float* __restrict__ m = c.m[lo];
  float* __restrict__ v = c.v[lo];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = base + ((long long)u * 256 + threadIdx.x) * 4;
    if (e >= n) break;
    if (vec && e + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + e), mv = *reinterpret_cast<float4*>(m + e),
             vv = *reinterpret_cast<float4*>(v + e);
      const float4 gv = *reinterpret_cast<const float4*>(g + e);
      adamw_elem(pv.x, gv.x, mv.x, vv.x, h);
      adamw_elem(pv.y, gv.y, mv.y, vv.y, h);
      adamw_elem(pv.z, gv.z, mv.z, vv.z, h);
      adamw_elem(pv.w, gv.w, mv.w, vv.w, h);
      *reinterpret_cast<float4*>(p + e) = pv;
      *reinterpret_cast<float4*>(m + e) = mv;
      *reinterpret_cast<float4*>(v + e) = vv;
    } else {
      for (long long i = e; i < n && i < e + 4; ++i) adamw_elem(p[i], g[i], m[i], v[i], h);
    }
  }
}

// pointer arrays are HOST arrays of DEVICE pointers
int adamw_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
               float* const* exp_avg_sq, const long long* sizes, double lr, double beta1, double beta2, double eps,
               double weight_decay, long long step, float grad_scale, cudaStream_t s) {
  AdamwHyper h;
  h.decay = (float)(1.0 - lr * weight_decay);
  h.b1 = (float)beta1; h.b2 = (float)beta2; h.omb1 = (float)(1.0 - beta1); h.omb2 = (float)(1.0 - beta2);
  h.eps = (float)eps; h.grad_scale = grad_scale;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  h.step_size = (float)(lr / bc1);
  h.inv_bc2 = (float)(1.0 / sqrt(bc2));
  for (int t0 = 0; t0 < n_tensors; t0 += kOptMaxTensors) {
    AdamwChunk c;
    c.count = 0;
    int blocks = 0;
    for (int t = t0; t < n_tensors && c.count < kOptMaxTensors; ++t) {
      if (sizes[t] <= 0) continue;
      const int k = c.count++;
      c.p[k] = params[t]; c.g[k] = grads[t]; c.m[k] = exp_avg[t]; c.v[k] = exp_avg_sq[t]; c.n[k] = sizes[t];
      c.first_block[k] = blocks;
      blocks += (int)((sizes[t] + kOptBlockElems - 1) / kOptBlockElems);
    }
    c.first_block[c.count] = blocks;
    if (blocks == 0) continue;
    adamw_kernel<<<blocks, 256, 0, s>>>(c, h);
    count_launch();
    const int st = (int)cudaGetLastError();
    if (st) return st;
  }
  return 0;
}

}  // namespace hwgat
