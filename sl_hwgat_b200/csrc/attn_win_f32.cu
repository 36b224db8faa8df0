// fp32 parity mode of K2b / K3b: the windowed graph attention of HWGATE for window_size 32 / 64 (N = 64 / 128 tokens,
// BASELINE configs[4], model_params.py:254) and of HGATE's 58-token blocks stored as 64 (HGATE.py:84-108), in true fp32.
// Same semantics as attn_f32.cu (N = 32), which is the north_star's "fp32 within 1e-5" path:
//   qkv  = xn . Wqkv^T + b                      (HWGATE.py:86)
//   S    = (q*scale) . k^T                      (HWGATE.py:89-91)
//   keep = !(softmax(S) > thr)     [training]   (HWGATE.py:94-100)
//   live = mask & keep & (S != 0)               (HWGATE.py:102-110)
//   P    = softmax(live ? S : -10000)           (HWGATE.py:110-111)
//   out  = P . v                                (HWGATE.py:114)
// One CTA per (window, head), one thread per token of the window: the thread is the query row in the forward and for dQ,
// and the key row for dK / dV.  k and v of the window sit in shared memory, P and dS (N x N) too; q and dO rows of
// the other tokens are read from global memory by all threads at the same address (a broadcast).  True fp32 FFMA, no
// tensor cores: a correctness mode, not a timed one.
#include "common.cuh"

namespace hwgat {

namespace winf32 {

struct Geo {
  int F, K, shift, layout, f, nW, W;
};

// global token row of row `row` of window `widx` (windows ordered (sample, temporal group, keypoint window)):
// HWGATE.py:197-201 + window_partition (30-36) without the copies
template <int N>
HW_DEV long long token_row(const Geo& g, long long widx, int row) {
  if (g.layout == HWGAT_LAYOUT_WINDOWS) return widx * N + row;
  const int per = g.f * g.nW;
  const long long b = widx / per;
  const int r = (int)(widx - b * per), fi = r / g.nW, kw = r - fi * g.nW;
  const int tp = row / g.W, k = row - tp * g.W;
  int fr = 2 * fi + tp + g.shift;
  fr = fr >= g.F ? fr - g.F : fr;
  return (b * g.F + fr) * g.K + kw * g.W + k;
}
// first mask word of that row: bits is (f * nW windows, N rows, N / 32 words)
template <int N>
HW_DEV long long mask_word(const Geo& g, long long widx, int row) {
  return ((widx % (g.f * g.nW)) * N + row) * (N / 32);
}

HW_DEV float dot64(const float (&a)[kHd], const float* __restrict__ b) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    p0 = fmaf(a[e], b[e], p0);
    p1 = fmaf(a[e + 1], b[e + 1], p1);
    p2 = fmaf(a[e + 2], b[e + 2], p2);
    p3 = fmaf(a[e + 3], b[e + 3], p3);
  }
  return (p0 + p1) + (p2 + p3);
}

// probabilities of one query row held by one thread: s[] in, p[] out (in place); live[] = the entries the gradient
// flows through.  The same steps, in the same order, as row_softmax_f32 of attn_f32.cu.
template <int N>
HW_DEV void row_softmax(float (&s)[N], uint32_t (&live)[N / 32], float threshold) {
  // (the loops over the N keys stay rolled: s[] lives in local memory, which is fine for a parity mode and keeps the
  //  code size and the compile time of the N = 128 instance sane)
  if (threshold >= 0.f) {
    float m = s[0];
#pragma unroll 1
    for (int j = 1; j < N; ++j) m = fmaxf(m, s[j]);
    float sum = 0.f;
#pragma unroll 1
    for (int j = 0; j < N; ++j) sum += expf(s[j] - m);
#pragma unroll 1
    for (int j = 0; j < N; ++j)
      if (expf(s[j] - m) / sum > threshold) live[j >> 5] &= ~(1u << (j & 31));
  }
#pragma unroll 1
  for (int j = 0; j < N; ++j)
    if (s[j] == 0.f) live[j >> 5] &= ~(1u << (j & 31));
  float m = -INFINITY;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    s[j] = ((live[j >> 5] >> (j & 31)) & 1u) ? s[j] : kNegFill;
    m = fmaxf(m, s[j]);
  }
  float sum = 0.f;
#pragma unroll 1
  for (int j = 0; j < N; ++j) { s[j] = expf(s[j] - m); sum += s[j]; }
  const float inv = 1.f / sum;
#pragma unroll 1
  for (int j = 0; j < N; ++j) s[j] *= inv;
}

template <int N>
HW_DEV void load_scaled_row(float (&q)[kHd], const float* src, float scale) {
  const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int e = 0; e < kHd / 4; ++e) {
    const float4 t = p[e];
    q[4 * e] = t.x * scale; q[4 * e + 1] = t.y * scale; q[4 * e + 2] = t.z * scale; q[4 * e + 3] = t.w * scale;
  }
}

template <int N>
__global__ void __launch_bounds__(N) win_fwd_f32_kernel(const float* __restrict__ qkv, const uint32_t* __restrict__ bits,
                                                        float threshold, float* __restrict__ out, Geo g, int d) {
  extern __shared__ __align__(16) float smf[];
  float (*sk)[kHd] = reinterpret_cast<float (*)[kHd]>(smf);
  float (*sv)[kHd] = sk + N;
  const long long widx = blockIdx.x;
  const int h = blockIdx.y, row = threadIdx.x, d3 = 3 * d;
  const long long tok = token_row<N>(g, widx, row);
  const float* src = qkv + tok * d3 + h * kHd;
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    *reinterpret_cast<float4*>(&sk[row][e]) = *reinterpret_cast<const float4*>(src + d + e);
    *reinterpret_cast<float4*>(&sv[row][e]) = *reinterpret_cast<const float4*>(src + 2 * d + e);
  }
  float q[kHd];
  load_scaled_row<N>(q, src, 0.125f);             // head_dim^-0.5, head_dim = 64
  __syncthreads();
  float s[N];
#pragma unroll 1
  for (int j = 0; j < N; ++j) s[j] = dot64(q, sk[j]);
  uint32_t live[N / 32];
  const long long mw = mask_word<N>(g, widx, row);
#pragma unroll
  for (int x = 0; x < N / 32; ++x) live[x] = bits[mw + x];
  row_softmax<N>(s, live, threshold);
  float o[kHd];
#pragma unroll
  for (int e = 0; e < kHd; ++e) o[e] = 0.f;
#pragma unroll 1
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int e = 0; e < kHd; ++e) o[e] = fmaf(s[j], sv[j][e], o[e]);
  float4* dst = reinterpret_cast<float4*>(out + tok * d + h * kHd);
#pragma unroll
  for (int e = 0; e < kHd / 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
}

template <int N>
__global__ void __launch_bounds__(N) win_bwd_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out,
                                                        const uint32_t* __restrict__ bits, float threshold,
                                                        float* __restrict__ dqkv, Geo g, int d) {
  extern __shared__ __align__(16) float smf[];
  float (*sk)[kHd] = reinterpret_cast<float (*)[kHd]>(smf);
  float (*sv)[kHd] = sk + N;
  float (*sp)[N + 1] = reinterpret_cast<float (*)[N + 1]>(smf + 2 * N * kHd);
  float (*sds)[N + 1] = sp + N;
  __shared__ long long s_tok[N];
  const long long widx = blockIdx.x;
  const int h = blockIdx.y, row = threadIdx.x, d3 = 3 * d;
  const float scale = 0.125f;
  const long long tok = token_row<N>(g, widx, row);
  s_tok[row] = tok;
  const float* src = qkv + tok * d3 + h * kHd;
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    *reinterpret_cast<float4*>(&sk[row][e]) = *reinterpret_cast<const float4*>(src + d + e);
    *reinterpret_cast<float4*>(&sv[row][e]) = *reinterpret_cast<const float4*>(src + 2 * d + e);
  }
  __syncthreads();
  // ---- thread = query i: P row, dP row, dS row, dQ row
  float s[N];
  {
    float q[kHd];
    load_scaled_row<N>(q, src, scale);
#pragma unroll 1
    for (int j = 0; j < N; ++j) s[j] = dot64(q, sk[j]);
  }
  uint32_t live[N / 32];
  const long long mw = mask_word<N>(g, widx, row);
#pragma unroll
  for (int x = 0; x < N / 32; ++x) live[x] = bits[mw + x];
  row_softmax<N>(s, live, threshold);            // s[] now holds P
  {
    float gr[kHd];
    load_scaled_row<N>(gr, d_out + tok * d + h * kHd, 1.f);
    float dsum = 0.f;
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
      const float a = dot64(gr, sv[j]);
      sds[row][j] = a;                            // dP for now
      dsum = fmaf(s[j], a, dsum);
    }
    float dq[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) dq[e] = 0.f;
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
      const float dsv = ((live[j >> 5] >> (j & 31)) & 1u) ? s[j] * (sds[row][j] - dsum) : 0.f;
      sp[row][j] = s[j];
      sds[row][j] = dsv;
#pragma unroll
      for (int e = 0; e < kHd; ++e) dq[e] = fmaf(dsv, sk[j][e], dq[e]);
    }
    float4* dst = reinterpret_cast<float4*>(dqkv + tok * d3 + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e)
      dst[e] = make_float4(dq[4 * e] * scale, dq[4 * e + 1] * scale, dq[4 * e + 2] * scale, dq[4 * e + 3] * scale);
  }
  __syncthreads();
  // ---- thread = key j: dK row = sum_i dS[i][j] (q_i * scale) ; dV row = sum_i P[i][j] dO_i
  {
    float dk[kHd], dv[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) { dk[e] = 0.f; dv[e] = 0.f; }
    for (int i = 0; i < N; ++i) {
      const float cds = sds[i][row] * scale, cp = sp[i][row];
      const float4* qi = reinterpret_cast<const float4*>(qkv + s_tok[i] * d3 + h * kHd);       // broadcast reads
      const float4* gi = reinterpret_cast<const float4*>(d_out + s_tok[i] * d + h * kHd);
#pragma unroll
      for (int e = 0; e < kHd / 4; ++e) {
        const float4 a = qi[e], b = gi[e];
        dk[4 * e] = fmaf(cds, a.x, dk[4 * e]); dk[4 * e + 1] = fmaf(cds, a.y, dk[4 * e + 1]);
        dk[4 * e + 2] = fmaf(cds, a.z, dk[4 * e + 2]); dk[4 * e + 3] = fmaf(cds, a.w, dk[4 * e + 3]);
        dv[4 * e] = fmaf(cp, b.x, dv[4 * e]); dv[4 * e + 1] = fmaf(cp, b.y, dv[4 * e + 1]);
        dv[4 * e + 2] = fmaf(cp, b.z, dv[4 * e + 2]); dv[4 * e + 3] = fmaf(cp, b.w, dv[4 * e + 3]);
      }
    }
    float4* dstk = reinterpret_cast<float4*>(dqkv + tok * d3 + d + h * kHd);
    float4* dstv = reinterpret_cast<float4*>(dqkv + tok * d3 + 2 * d + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e) {
      dstk[e] = make_float4(dk[4 * e], dk[4 * e + 1], dk[4 * e + 2], dk[4 * e + 3]);
      dstv[e] = make_float4(dv[4 * e], dv[4 * e + 1], dv[4 * e + 2], dv[4 * e + 3]);
    }
  }
}

template <int N>
static int launch_fwd(const float* qkv, const AttnArgs& a, const Geo& g, long long windows, cudaStream_t s) {
  constexpr int smem = 2 * N * kHd * (int)sizeof(float);
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(win_fwd_f32_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  win_fwd_f32_kernel<N><<<dim3((unsigned)windows, a.heads), N, smem, s>>>(qkv, a.bits, a.threshold, (float*)a.out, g, a.d);
  count_launch();
  return (int)cudaGetLastError();
}
template <int N>
static int launch_bwd(const float* qkv, float* dqkv, const AttnArgs& a, const Geo& g, long long windows, cudaStream_t s) {
  constexpr int smem = (2 * N * kHd + 2 * N * (N + 1)) * (int)sizeof(float);
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(win_bwd_f32_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  win_bwd_f32_kernel<N><<<dim3((unsigned)windows, a.heads), N, smem, s>>>(qkv, (const float*)a.d_out, a.bits, a.threshold,
                                                                         dqkv, g, a.d);
  count_launch();
  return (int)cudaGetLastError();
}

static Geo make_geo(const AttnArgs& a, int W) {
  Geo g;
  g.F = a.F; g.K = a.K; g.shift = a.shift; g.layout = a.layout; g.f = a.F / 2; g.nW = a.K / W; g.W = W;
  return g;
}

}  // namespace winf32

// forward workspace: none (qkv goes to the caller's buffer, kept for the backward); backward: dqkv [n, 3d] fp32
size_t attn2_workspace_bytes_f32(long long n, int d, int backward) {
  return backward ? sizeof(float) * (size_t)n * 3 * d + 256 : 0;
}

int attn2_fwd_f32(const AttnArgs& a, int W, float* qkv, cudaStream_t s) {
  const long long n = a.tokens();
  if (a.d / a.heads != kHd || (n + 63) / 64 > 65535 || n > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  int st = linear_f32_fwd((const float*)a.xn, (const float*)a.w_qkv, a.b_qkv, qkv, (int)n, a.d, 3 * a.d, s);
  if (st) return st;
  const winf32::Geo g = winf32::make_geo(a, W);
  const long long windows = n / (2 * W);
  switch (2 * W) {
    case 64: return winf32::launch_fwd<64>(qkv, a, g, windows, s);
    case 128: return winf32::launch_fwd<128>(qkv, a, g, windows, s);
  }
  return HWGAT_ERR_UNSUPPORTED;
}

int attn2_bwd_f32(const AttnArgs& a, int W, const float* qkv, cudaStream_t s) {
  const long long n = a.tokens();
  if (a.d / a.heads != kHd || (n + 63) / 64 > 65535 || n > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  float* dqkv = (float*)a.workspace;
  const winf32::Geo g = winf32::make_geo(a, W);
  const long long windows = n / (2 * W);
  int st;
  switch (2 * W) {
    case 64: st = winf32::launch_bwd<64>(qkv, dqkv, a, g, windows, s); break;
    case 128: st = winf32::launch_bwd<128>(qkv, dqkv, a, g, windows, s); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  if (st) return st;
  return qkv_weight_grads_f32(dqkv, (const float*)a.xn, (const float*)a.w_qkv, (float*)a.d_xn, a.d_w, a.d_b, n, a.d, s);
}

}  // namespace hwgat
