// Bandwidth-bound kernels of the rest of the block, bf16/autocast
// path only.  Each replaces a chain of PyTorch elementwise kernels of
// PartAttentionBlock.forward (HWGATE.py:189-221) and its autograd with ONE pass over HBM:
//
//   K5  ln_fwd            y(bf16) = LayerNorm(x fp32) ; saves mean, rstd        (norm1 / norm2, :203, :219)
//   K5' ln_bwd            dx(fp32) = d_residual + LayerNorm'(dy bf16) ; dgamma, dbeta
//   K6  bda_ln_fwd        x1(fp32) = res + dropout(a0 bf16 + bias); y(bf16) = LN_next(x1)   (Linear bias + proj_drop/ff.drop
//                         + shortcut add :115-116,:134-135,:217,:219, fused with the NEXT LayerNorm :203/:219)
//   K6' bda_ln_bwd        d_res = g_x1 + LN'(dy); d_a0(bf16) = mask*scale*d_res; dbias, dgamma, dbeta
//   K7  bias_gelu_dropout g(bf16) = dropout(gelu(u0 bf16 + bias))                 (fc1 bias + ff.act + ff.drop, :131-133)
//   K7' ... backward      du0(bf16) = dg * mask * scale * gelu'(u0 + bias); dbias
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
// K6: x1 = res + dropout(a0 + bias)  and, fused, the LayerNorm that follows it: y = LN(x1) as bf16.
// One warp per row, lane owns columns i*128 + 4*lane .. +3.  Dropout flags of a 4-element granule come
// from one Philox call keyed by the granule index (row*d + col)/4, identically in forward and backward.
// ---------------------------------------------------------------------------
HW_DEV uint32_t keep4(unsigned long long granule, unsigned long long offset, unsigned long long seed, uint32_t thresh16) {
  const uint4 r = philox4x32(granule, offset, seed);
  return ((r.x & 0xFFFFu) >= thresh16 ? 1u : 0u) | ((r.x >> 16) >= thresh16 ? 2u : 0u) |
         ((r.y & 0xFFFFu) >= thresh16 ? 4u : 0u) | ((r.y >> 16) >= thresh16 ? 8u : 0u);
}

template <int kV, bool kLN>
__global__ void __launch_bounds__(256) bda_ln_fwd_kernel(const float* __restrict__ res, const bf16* __restrict__ a0,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ x1,
                                                         bf16* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd, long long n, float eps, float scale,
                                                         uint32_t thresh, unsigned long long seed,
                                                         unsigned long long offset) {
  constexpr int d = kV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  float4 bs[kV], gm[kV], bt[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    bs[i] = bias ? *reinterpret_cast<const float4*>(bias + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (kLN) {
      gm[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
      bt[i] = *reinterpret_cast<const float4*>(beta + i * 128 + lane * 4);
    }
  }
  for (long long row = warp; row < n; row += nwarps) {
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const long long e = row * d + i * 128 + lane * 4;
      const float4 r = *reinterpret_cast<const float4*>(res + e);
      const uint2 av = *reinterpret_cast<const uint2*>(a0 + e);
      const uint32_t keep = thresh ? keep4((unsigned long long)e >> 2, offset, seed, thresh) : 0xFu;
      v[i].x = r.x + ((keep & 1u) ? (bf16_lo(av.x) + bs[i].x) * scale : 0.f);
      v[i].y = r.y + ((keep & 2u) ? (bf16_hi(av.x) + bs[i].y) * scale : 0.f);
      v[i].z = r.z + ((keep & 4u) ? (bf16_lo(av.y) + bs[i].z) * scale : 0.f);
      v[i].w = r.w + ((keep & 8u) ? (bf16_hi(av.y) + bs[i].w) * scale : 0.f);
      *reinterpret_cast<float4*>(x1 + e) = v[i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    if (!kLN) continue;
    const float mu = warp_sum(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = rsqrtf(warp_sum(q) * (1.f / d) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      uint2 o;
      o.x = pack_bf16(v[i].x * rs * gm[i].x + bt[i].x, v[i].y * rs * gm[i].y + bt[i].y);
      o.y = pack_bf16(v[i].z * rs * gm[i].z + bt[i].z, v[i].w * rs * gm[i].w + bt[i].w);
      *reinterpret_cast<uint2*>(y + row * d + i * 128 + lane * 4) = o;
    }
  }
}

// K6': dx = g_x1 + LN'(dy)  [kLN]  or  dx = g_x1  [!kLN];   d_res = dx (written only if kLN),
//      d_a0 = mask * scale * dx (bf16),  dbias += colsum(mask * scale * dx),  dgamma / dbeta as K5'.
template <int kV, bool kLN>
__global__ void __launch_bounds__(256) bda_ln_bwd_kernel(const float* __restrict__ g_x1, const bf16* __restrict__ dy,
                                                         const float* __restrict__ x1, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                         float* __restrict__ d_res, bf16* __restrict__ d_a0,
                                                         float* __restrict__ dbias, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, long long n, float scale,
                                                         uint32_t thresh, unsigned long long seed,
                                                         unsigned long long offset) {
  constexpr int d = kV * 128;
  __shared__ float red[kLN ? 3 : 1][8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
  float4 gm[kV], dg[kV], db[kV], dbs[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    if (kLN) gm[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    dg[i] = db[i] = dbs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp; row < n; row += nwarps) {
    float4 o[kV];
    if (kLN) {
      const float mu = mean[row], rs = rstd[row];
      float4 xh[kV], g[kV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const long long e = row * d + i * 128 + lane * 4;
        const float4 xv = *reinterpret_cast<const float4*>(x1 + e);
        const uint2 dv = *reinterpret_cast<const uint2*>(dy + e);
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
        o[i] = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                           rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
        if (g_x1) {
          const float4 r = *reinterpret_cast<const float4*>(g_x1 + row * d + i * 128 + lane * 4);
          o[i].x += r.x; o[i].y += r.y; o[i].z += r.z; o[i].w += r.w;
        }
        *reinterpret_cast<float4*>(d_res + row * d + i * 128 + lane * 4) = o[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < kV; ++i) o[i] = *reinterpret_cast<const float4*>(g_x1 + row * d + i * 128 + lane * 4);
    }
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const long long e = row * d + i * 128 + lane * 4;
      const uint32_t keep = thresh ? keep4((unsigned long long)e >> 2, offset, seed, thresh) : 0xFu;
      const float4 da = make_float4((keep & 1u) ? o[i].x * scale : 0.f, (keep & 2u) ? o[i].y * scale : 0.f,
                                    (keep & 4u) ? o[i].z * scale : 0.f, (keep & 8u) ? o[i].w * scale : 0.f);
      uint2 pk;
      pk.x = pack_bf16(da.x, da.y);
      pk.y = pack_bf16(da.z, da.w);
      *reinterpret_cast<uint2*>(d_a0 + e) = pk;
      dbs[i].x += da.x; dbs[i].y += da.y; dbs[i].z += da.z; dbs[i].w += da.w;
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    *reinterpret_cast<float4*>(&red[0][wib][i * 128 + lane * 4]) = dbs[i];
    if (kLN) {
      *reinterpret_cast<float4*>(&red[1][wib][i * 128 + lane * 4]) = dg[i];
      *reinterpret_cast<float4*>(&red[2][wib][i * 128 + lane * 4]) = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[0][w][c];
      if (kLN) { b += red[1][w][c]; cc += red[2][w][c]; }
    }
    if (dbias) atomicAdd(dbias + c, a);
    if (kLN) { atomicAdd(dgamma + c, b); atomicAdd(dbeta + c, cc); }
  }
}

// ---------------------------------------------------------------------------
// K7 / K7': g = dropout(gelu(u0 + bias)), one 8-element vector per thread-iteration; the grid stride is a
// multiple of the row length, so a thread always sees the same 8 columns (bias in registers, and
// dbias = colsum(du0) accumulates in registers in the backward).
// ---------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256) bias_gelu_dropout_fwd_kernel(const bf16* __restrict__ u0,
                                                                    const float* __restrict__ bias,
                                                                    bf16* __restrict__ g, long long nvec, int cols,
                                                                    float scale, uint32_t thresh,
                                                                    unsigned long long seed, unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float bs[8];
  {
    const int c = (int)((v0 * 8) % cols);
#pragma unroll
    for (int i = 0; i < 8; ++i) bs[i] = bias ? bias[c + i] : 0.f;
  }
  for (long long v = v0; v < nvec; v += stride) {
    const int4 uv = ld_stream16(u0 + v * 8);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const uint32_t w[4] = {(uint32_t)uv.x, (uint32_t)uv.y, (uint32_t)uv.z, (uint32_t)uv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16(((keep >> (2 * i)) & 1u) ? gelu_exact(bf16_lo(w[i]) + bs[2 * i]) * scale : 0.f,
                       ((keep >> (2 * i + 1)) & 1u) ? gelu_exact(bf16_hi(w[i]) + bs[2 * i + 1]) * scale : 0.f);
    st_stream16(g + v * 8, make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]));
  }
}

__global__ void __launch_bounds__(256) bias_gelu_dropout_bwd_kernel(const bf16* __restrict__ u0,
                                                                    const float* __restrict__ bias,
                                                                    const bf16* __restrict__ dg, bf16* __restrict__ du0,
                                                                    float* __restrict__ dbias, long long nvec, int cols,
                                                                    float scale, uint32_t thresh,
                                                                    unsigned long long seed, unsigned long long offset) {
  __shared__ float red[256][9];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)((v0 * 8) % cols);
  float bs[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { bs[i] = bias ? bias[c0 + i] : 0.f; acc[i] = 0.f; }
  for (long long v = v0; v < nvec; v += stride) {
    const int4 uv = ld_stream16(u0 + v * 8), gv = ld_stream16(dg + v * 8);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
    const uint32_t w[4] = {(uint32_t)uv.x, (uint32_t)uv.y, (uint32_t)uv.z, (uint32_t)uv.w};
    const uint32_t q[4] = {(uint32_t)gv.x, (uint32_t)gv.y, (uint32_t)gv.z, (uint32_t)gv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = ((keep >> (2 * i)) & 1u) ? bf16_lo(q[i]) * scale * gelu_grad(bf16_lo(w[i]) + bs[2 * i]) : 0.f;
      const float b = ((keep >> (2 * i + 1)) & 1u) ? bf16_hi(q[i]) * scale * gelu_grad(bf16_hi(w[i]) + bs[2 * i + 1]) : 0.f;
      o[i] = pack_bf16(a, b);
      acc[2 * i] += a;
      acc[2 * i + 1] += b;
    }
    st_stream16(du0 + v * 8, make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]));
  }
  if (!dbias) return;
  // threads t, t + G, t + 2G, ... (G = cols/8 column groups, 256 % G == 0) own the same 8 columns
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  const int G = cols / 8;
  if ((int)threadIdx.x < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float sum = 0.f;
      for (int j = threadIdx.x; j < 256; j += G) sum += red[j][i];
      atomicAdd(dbias + c0 + i, sum);
    }
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

template <bool kLN>
static int bda_fwd_dispatch(int grid, cudaStream_t s, const float* res, const bf16* a0, const float* bias,
                            const float* gamma, const float* beta, float* x1, bf16* y, float* mean, float* rstd,
                            long long n, int d, float eps, float scale, uint32_t thresh, unsigned long long seed,
                            unsigned long long offset) {
  switch (d) {
    case 128: bda_ln_fwd_kernel<1, kLN><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset); break;
    case 256: bda_ln_fwd_kernel<2, kLN><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset); break;
    case 512: bda_ln_fwd_kernel<4, kLN><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  return 0;
}

int launch_bda_ln_fwd(const float* res, const bf16* a0, const float* bias, const float* gamma, const float* beta,
                      float* x1, bf16* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                      unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(n * 32);
  int st = gamma ? bda_fwd_dispatch<true>(grid, s, res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, scale, thresh, seed, offset)
                 : bda_fwd_dispatch<false>(grid, s, res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, scale, thresh, seed, offset);
  if (st) return st;
  count_launch();
  return (int)cudaGetLastError();
}

template <bool kLN>
static int bda_bwd_dispatch(int grid, cudaStream_t s, const float* g_x1, const bf16* dy, const float* x1,
                            const float* mean, const float* rstd, const float* gamma, float* d_res, bf16* d_a0,
                            float* dbias, float* dgamma, float* dbeta, long long n, int d, float scale, uint32_t thresh,
                            unsigned long long seed, unsigned long long offset) {
  switch (d) {
    case 128: bda_ln_bwd_kernel<1, kLN><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset); break;
    case 256: bda_ln_bwd_kernel<2, kLN><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset); break;
    case 512: bda_ln_bwd_kernel<4, kLN><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  return 0;
}

int launch_bda_ln_bwd(const float* g_x1, const bf16* dy, const float* x1, const float* mean, const float* rstd,
                      const float* gamma, float* d_res, bf16* d_a0, float* dbias, float* dgamma, float* dbeta,
                      long long n, int d, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * d, s);
  if (gamma) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * d, s);
  }
  long long want = (n + 7) / 8;
  const int grid = (int)(want < 148LL * 4 ? (want < 1 ? 1 : want) : 148LL * 4);
  int st = gamma ? bda_bwd_dispatch<true>(grid, s, g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, scale, thresh, seed, offset)
                 : bda_bwd_dispatch<false>(grid, s, g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, scale, thresh, seed, offset);
  if (st) return st;
  count_launch();
  return (int)cudaGetLastError();
}

int launch_bias_gelu_dropout(const bf16* u0, const float* bias, const bf16* dg, bf16* out, float* dbias, long long n,
                             int cols, float p, unsigned long long seed, unsigned long long offset, bool backward,
                             cudaStream_t s) {
  const long long nvec = n * cols / 8;
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(nvec);
  if (backward) {
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * cols, s);
    bias_gelu_dropout_bwd_kernel<<<grid, 256, 0, s>>>(u0, bias, dg, out, dbias, nvec, cols, scale, thresh, seed, offset);
  } else {
    bias_gelu_dropout_fwd_kernel<<<grid, 256, 0, s>>>(u0, bias, out, nvec, cols, scale, thresh, seed, offset);
  }
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
