// Bandwidth-bound kernels of the rest of the block, bf16/autocast
// path only.  Each replaces a chain of PyTorch elementwise kernels of
// PartAttentionBlock.forward (HWGATE.py:189-221) and its autograd with ONE pass over HBM:
//
//   K5  ln_fwd            y(bf16) = LayerNorm(x fp32) ; saves mean, rstd        (norm1 / norm2, :203, :219)
//   K5' ln_bwd            dx(fp32) = d_residual + LayerNorm'(dy bf16) ; dgamma, dbeta
//   K6  dropout_add_fwd   out(fp32) = res(fp32) + dropout(a bf16)               (proj_drop + shortcut :116,:217; ff.drop + x :135,:219)
//   K6' dropout_add_bwd   da(bf16) = mask * scale * dout(fp32)
//   K7  gelu_dropout_fwd  g(bf16) = dropout(gelu(u bf16))                        (ff.act + ff.drop, :132-133)
//   K7' gelu_dropout_bwd  du(bf16) = dg * mask * scale * gelu'(u)
//
// Dropout masks are never stored: forward and backward regenerate them from a Philox4x32-7
// counter stream keyed by (seed, offset) taken from PyTorch's CUDA generator on the host side.
#include "common.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// Philox4x32-7 (Salmon et al. 2011): 4 random words for counter (idx, offset), key = seed
// ---------------------------------------------------------------------------
HW_DEV uint4 philox4x32(unsigned long long idx, unsigned long long offset, unsigned long long seed) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 7; ++r) {  // Philox4x32-7: the shortest variant that passes BigCrush (Salmon et al., table 2)
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep flags of 8 consecutive elements (one 16-byte bf16 vector) of vector index v: bit i = element i kept
HW_DEV uint32_t keep8(unsigned long long v, unsigned long long offset, unsigned long long seed, uint32_t thresh16) {
  const uint4 r = philox4x32(v, offset, seed);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xFFFFu) >= thresh16 ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= thresh16 ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}
// drop probability -> 16-bit threshold: P(u16 < thresh) = p
static uint32_t drop_threshold16(float p) {
  if (p <= 0.f) return 0u;
  long long t = (long long)(p * 65536.0 + 0.5);
  return (uint32_t)(t > 65535 ? 65535 : t);
}

HW_DEV float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
HW_DEV float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

HW_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// K5: one warp per row; lane holds kV float4 (kV = d/128)
// ---------------------------------------------------------------------------
template <int kV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                     long long n, float eps) {
  constexpr int d = kV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  float4 gm[kV], bt[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    gm[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    bt[i] = *reinterpret_cast<const float4*>(beta + i * 128 + lane * 4);
  }
  for (long long row = warp; row < n; row += nwarps) {
    const float* xr = x + row * d;
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = rsqrtf(warp_sum(q) * (1.f / d) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    bf16* yr = y + row * d;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      uint2 o;
      o.x = pack_bf16(v[i].x * rs * gm[i].x + bt[i].x, v[i].y * rs * gm[i].y + bt[i].y);
      o.y = pack_bf16(v[i].z * rs * gm[i].z + bt[i].z, v[i].w * rs * gm[i].w + bt[i].w);
      *reinterpret_cast<uint2*>(yr + i * 128 + lane * 4) = o;
    }
  }
}

// ---------------------------------------------------------------------------
// K5': dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//      dgamma += sum_rows dy * xhat, dbeta += sum_rows dy   (per-lane partials, smem reduce, one atomic per column per CTA)
// ---------------------------------------------------------------------------
template <int kV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ dres,
                                                     const float* __restrict__ x, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                     float* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, long long n) {
  constexpr int d = kV * 128;
  __shared__ float red[2][8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
  float4 gm[kV], dg[kV], db[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    gm[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp; row < n; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[kV], g[kV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(x + row * d + i * 128 + lane * 4);
      const uint2 dv = *reinterpret_cast<const uint2*>(dy + row * d + i * 128 + lane * 4);
      const float4 dyv = make_float4(bf16_lo(dv.x), bf16_hi(dv.x), bf16_lo(dv.y), bf16_hi(dv.y));
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dyv.x * gm[i].x, dyv.y * gm[i].y, dyv.z * gm[i].z, dyv.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
      db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
    }
    const float m1 = warp_sum(s1) * (1.f / d), m2 = warp_sum(s2) * (1.f / d);
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                             rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * d + i * 128 + lane * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dx + row * d + i * 128 + lane * 4) = o;
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    *reinterpret_cast<float4*>(&red[0][wib][i * 128 + lane * 4]) = dg[i];
    *reinterpret_cast<float4*>(&red[1][wib][i * 128 + lane * 4]) = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

// ---------------------------------------------------------------------------
// K6 / K6' / K7 / K7': one 8-element vector per thread-iteration
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dropout_add_fwd_kernel(const float* __restrict__ res, const bf16* __restrict__ a,
                                                              float* __restrict__ out, long long nvec, float scale,
                                                              uint32_t thresh, unsigned long long seed,
                                                              unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int4 av = ld_stream16(a + v * 8);
    const float4 r0 = *reinterpret_cast<const float4*>(res + v * 8), r1 = *reinterpret_cast<const float4*>(res + v * 8 + 4);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const uint32_t w[4] = {(uint32_t)av.x, (uint32_t)av.y, (uint32_t)av.z, (uint32_t)av.w};
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = ((keep >> (2 * i)) & 1u) ? bf16_lo(w[i]) * scale : 0.f;
      f[2 * i + 1] = ((keep >> (2 * i + 1)) & 1u) ? bf16_hi(w[i]) * scale : 0.f;
    }
    *reinterpret_cast<float4*>(out + v * 8) = make_float4(r0.x + f[0], r0.y + f[1], r0.z + f[2], r0.w + f[3]);
    *reinterpret_cast<float4*>(out + v * 8 + 4) = make_float4(r1.x + f[4], r1.y + f[5], r1.z + f[6], r1.w + f[7]);
  }
}

__global__ void __launch_bounds__(256) dropout_add_bwd_kernel(const float* __restrict__ dout, bf16* __restrict__ da,
                                                              long long nvec, float scale, uint32_t thresh,
                                                              unsigned long long seed, unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const float4 g0 = *reinterpret_cast<const float4*>(dout + v * 8), g1 = *reinterpret_cast<const float4*>(dout + v * 8 + 4);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16(((keep >> (2 * i)) & 1u) ? g[2 * i] * scale : 0.f,
                       ((keep >> (2 * i + 1)) & 1u) ? g[2 * i + 1] * scale : 0.f);
    st_stream16(da + v * 8, make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]));
  }
}

// Exact (erf) GELU of nn.GELU() and its derivative, sharing one exponential:
//   Phi(x) = 1 - erfc(x/sqrt2)/2,  erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1/(1 + p z), z >= 0
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: two orders below the bf16 rounding of the output;
//  erff() costs ~3x as many instructions and made these kernels compute-bound.)
HW_DEV float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
  return r;
}
HW_DEV float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
  return r;
}
HW_DEV void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));     // 1 ulp; the polynomial's own error is 1.5e-7
  const float e = ex2_approx(-z * z * 1.4426950408889634f);  // exp(-x^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float half_erfc = 0.5f * t * poly * e;
  cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
  pdf = 0.3989422804014327f * e;
}
HW_DEV float gelu_exact(float x) {
  float c, p;
  gelu_cdf_pdf(x, c, p);
  return x * c;
}
HW_DEV float gelu_grad(float x) {
  float c, p;
  gelu_cdf_pdf(x, c, p);
  return fmaf(x, p, c);
}

__global__ void __launch_bounds__(256) gelu_dropout_fwd_kernel(const bf16* __restrict__ u, bf16* __restrict__ g,
                                                               long long nvec, float scale, uint32_t thresh,
                                                               unsigned long long seed, unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int4 uv = ld_stream16(u + v * 8);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const uint32_t w[4] = {(uint32_t)uv.x, (uint32_t)uv.y, (uint32_t)uv.z, (uint32_t)uv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16(((keep >> (2 * i)) & 1u) ? gelu_exact(bf16_lo(w[i])) * scale : 0.f,
                       ((keep >> (2 * i + 1)) & 1u) ? gelu_exact(bf16_hi(w[i])) * scale : 0.f);
    st_stream16(g + v * 8, make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]));
  }
}

__global__ void __launch_bounds__(256) gelu_dropout_bwd_kernel(const bf16* __restrict__ u, const bf16* __restrict__ dg,
                                                               bf16* __restrict__ du, long long nvec, float scale,
                                                               uint32_t thresh, unsigned long long seed,
                                                               unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int4 uv = ld_stream16(u + v * 8), gv = ld_stream16(dg + v * 8);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const uint32_t w[4] = {(uint32_t)uv.x, (uint32_t)uv.y, (uint32_t)uv.z, (uint32_t)uv.w};
    const uint32_t q[4] = {(uint32_t)gv.x, (uint32_t)gv.y, (uint32_t)gv.z, (uint32_t)gv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16(((keep >> (2 * i)) & 1u) ? bf16_lo(q[i]) * scale * gelu_grad(bf16_lo(w[i])) : 0.f,
                       ((keep >> (2 * i + 1)) & 1u) ? bf16_hi(q[i]) * scale * gelu_grad(bf16_hi(w[i])) : 0.f);
    st_stream16(du + v * 8, make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]));
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int ew_grid(long long work_items) {
  long long want = (work_items + 255) / 256;
  const long long cap = 148LL * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

int launch_ln_fwd(const float* x, const float* gamma, const float* beta, bf16* y, float* mean, float* rstd,
                  long long n, int d, float eps, cudaStream_t s) {
  const int grid = ew_grid(n * 32);
  switch (d) {
    case 128: ln_fwd_kernel<1><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    case 256: ln_fwd_kernel<2><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    case 512: ln_fwd_kernel<4><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}

int launch_ln_bwd(const bf16* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                  const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d, cudaStream_t s) {
  cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s);
  cudaMemsetAsync(dbeta, 0, sizeof(float) * d, s);
  long long want = (n + 7) / 8;
  const int grid = (int)(want < 148LL * 4 ? (want < 1 ? 1 : want) : 148LL * 4);
  switch (d) {
    case 128: ln_bwd_kernel<1><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n); break;
    case 256: ln_bwd_kernel<2><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n); break;
    case 512: ln_bwd_kernel<4><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}

int launch_dropout_add(const float* res_or_dout, const bf16* a, float* out, bf16* da, long long numel, float p,
                       unsigned long long seed, unsigned long long offset, bool backward, cudaStream_t s) {
  const long long nvec = numel / 8;
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(nvec);
  if (backward)
    dropout_add_bwd_kernel<<<grid, 256, 0, s>>>(res_or_dout, da, nvec, scale, thresh, seed, offset);
  else
    dropout_add_fwd_kernel<<<grid, 256, 0, s>>>(res_or_dout, a, out, nvec, scale, thresh, seed, offset);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_gelu_dropout(const bf16* u, const bf16* dg, bf16* out, long long numel, float p, unsigned long long seed,
                        unsigned long long offset, bool backward, cudaStream_t s) {
  const long long nvec = numel / 8;
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(nvec);
  if (backward)
    gelu_dropout_bwd_kernel<<<grid, 256, 0, s>>>(u, dg, out, nvec, scale, thresh, seed, offset);
  else
    gelu_dropout_fwd_kernel<<<grid, 256, 0, s>>>(u, out, nvec, scale, thresh, seed, offset);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
