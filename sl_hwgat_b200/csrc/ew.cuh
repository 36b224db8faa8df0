// Elementwise device helpers shared by the bandwidth-bound block kernels (block_fused.cu) and the
// GEMM epilogues (ffn_tc.cu): the Philox dropout stream and the exact GELU.
#pragma once
#include "common.cuh"

namespace hwgat {

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
// keep flags of a 4-element granule (index (row*d + col)/4), identical in forward and backward
HW_DEV uint32_t keep4(unsigned long long granule, unsigned long long offset, unsigned long long seed, uint32_t thresh16) {
  const uint4 r = philox4x32(granule, offset, seed);
  return ((r.x & 0xFFFFu) >= thresh16 ? 1u : 0u) | ((r.x >> 16) >= thresh16 ? 2u : 0u) |
         ((r.y & 0xFFFFu) >= thresh16 ? 4u : 0u) | ((r.y >> 16) >= thresh16 ? 8u : 0u);
}
// drop probability -> 16-bit threshold: P(u16 < thresh) = p
inline uint32_t drop_threshold16(float p) {
  if (p <= 0.f) return 0u;
  long long t = (long long)(p * 65536.0 + 0.5);
  return (uint32_t)(t > 65535 ? 65535 : t);
}
inline float drop_scale16(uint32_t thresh) { return thresh ? 65536.f / (65536.f - (float)thresh) : 1.f; }

HW_DEV float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
HW_DEV float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ---------------------------------------------------------------------------
// Exact (erf) GELU of nn.GELU() and its derivative, sharing one exponential:
//   Phi(x) = 1 - erfc(x/sqrt2)/2,  erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1/(1 + p z), z >= 0
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: two orders below the bf16 rounding of the output;
//  erff() costs ~3x as many instructions and made these kernels compute-bound.)
// ---------------------------------------------------------------------------
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
// Constants are pre-folded to shave multiplies (these epilogues are ALU-bound): zp = |x| sqrt(log2(e)/2), so
// exp(-x^2/2) = 2^(-zp^2) and p z = 0.27273747 zp; the polynomial coefficients carry the 1/2 of erfc/2.
HW_DEV void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  const float zp = fabsf(x) * 0.84932178f;
  const float t = rcp_approx(fmaf(0.27273747f, zp, 1.f));   // 1 ulp; the polynomial's own error is 1.5e-7
  const float e = ex2_approx(-zp * zp);                      // exp(-x^2/2)
  float poly = fmaf(t, 0.5307027145f, -0.7265760135f);
  poly = fmaf(t, poly, 0.7107068705f);
  poly = fmaf(t, poly, -0.142248368f);
  poly = fmaf(t, poly, 0.127414796f);
  const float half_erfc = (t * e) * poly;
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

}  // namespace hwgat
