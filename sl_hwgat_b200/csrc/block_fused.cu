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
#include "ew.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

HW_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic mode: per-CTA partial column sums part[v][cta][col] -> out_v[col], CTAs added in index order.
__global__ void col_finish_kernel(const float* __restrict__ part, int nblocks, int cols, float* __restrict__ o0,
                                  float* __restrict__ o1, float* __restrict__ o2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float* outs[3] = {o0, o1, o2};
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    if (!outs[v]) continue;
    float a = 0.f;
    for (int b = 0; b < nblocks; ++b) a += part[((size_t)v * nblocks + b) * cols + c];
    outs[v][c] = a;
  }
}
int det_scratch(float** part, int nvec, int grid, int cols, cudaStream_t s) {
  *part = nullptr;
  if (!deterministic()) return 0;
  void* p = nullptr;
  if (cudaMallocAsync(&p, sizeof(float) * (size_t)nvec * grid * cols, s) != cudaSuccess) {
    cudaGetLastError();
    return HWGAT_ERR_WORKSPACE;   // never fall back to the atomic sums silently: the caller asked for reproducibility
  }
  *part = (float*)p;
  return 0;
}
void det_finish(float* part, int grid, int cols, float* o0, float* o1, float* o2, cudaStream_t s) {
  if (!part) return;
  col_finish_kernel<<<(cols + 127) / 128, 128, 0, s>>>(part, grid, cols, o0, o1, o2);
  count_launch();
  cudaFreeAsync(part, s);
}

// ---------------------------------------------------------------------------
// K5 / K6 thread mapping.  A row of d = 128 / 256 / 512 columns is owned by kL = d/16 = 8 / 16 / 32 lanes, each
// holding FOUR float4 (columns i*4*kL + 4*lc .. +3, i = 0..3), and a warp works on 32/kL rows at once.  Every
// width therefore has the same bytes in flight per lane and the same per-lane code, and the row reductions are
// log2(kL) shuffle steps shared by all rows of the warp (one warp per row at d = 128 spent 10 shuffles per 128
// columns and ran at 65-78 % of the HBM rate the d = 512 kernels reach).
// ---------------------------------------------------------------------------
template <int kL>
HW_DEV float group_sum(float v) {
#pragma unroll
  for (int o = kL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// per-lane column partials -> sum over the warp's row groups (lanes with equal lane % kL own the same columns)
template <int kL>
HW_DEV void fold_groups(float4 (&a)[4]) {
#pragma unroll
  for (int o = kL; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i].x += __shfl_xor_sync(0xffffffffu, a[i].x, o);
      a[i].y += __shfl_xor_sync(0xffffffffu, a[i].y, o);
      a[i].z += __shfl_xor_sync(0xffffffffu, a[i].z, o);
      a[i].w += __shfl_xor_sync(0xffffffffu, a[i].w, o);
    }
}


// Activation I/O of K5 - K7 in two element types: bf16 (the autocast path: GEMM operands / outputs) and float (the fp32
// path, whose GEMMs are the x3 tcgen05 kernels of gemm_x3.cu).  Four consecutive elements per access.
template <class T> HW_DEV float4 ld4(const T* p);
template <> HW_DEV float4 ld4<bf16>(const bf16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}
template <> HW_DEV float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <class T> HW_DEV void st4(T* p, const float4& v);
template <> HW_DEV void st4<bf16>(bf16* p, const float4& v) {
  uint2 o;
  o.x = pack_bf16(v.x, v.y);
  o.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = o;
}
template <> HW_DEV void st4<float>(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------
// TemporalMerging (HWGATE.py:55-63) folded into its neighbours.  The merged tensor (B, F/2, K, 2d) is, in units of
// d-wide rows, (B, F/2, K, 2, d): merging is a ROW permutation of the (B, F, K, d) tensor,
//   row (b*F + fr)*K + k   ->   ((b*(F/2) + fr/2)*K + k)*2 + (fr & 1).
// The last K6 of a level stores its rows through this map (forward), and the first LayerNorm-backward of the next
// level stores each half of its 2d-wide rows through the inverse (backward), so K4 and its adjoint are not launched.
// ---------------------------------------------------------------------------
// (row indices fit 31 bits: the C ABI refuses more than 2^29 tokens; 32-bit divisions)
HW_DEV long long merged_row(long long r, int F, int K) {
  const uint32_t r32 = (uint32_t)r;
  const uint32_t bf = r32 / (uint32_t)K, k = r32 - bf * (uint32_t)K;
  const uint32_t b = bf / (uint32_t)F, fr = bf - b * (uint32_t)F;
  return (long long)((((b * (uint32_t)(F >> 1)) + (fr >> 1)) * (uint32_t)K + k) << 1 | (fr & 1u));
}
// inverse, for row R of the MERGED tensor (F2 = F/2 frames): the un-merged d-wide row of its FIRST half (frame 2 fi);
// the second half (frame 2 fi + 1) is K rows further
HW_DEV long long unmerged_row0(long long R, int F2, int K) {
  const uint32_t r32 = (uint32_t)R;
  const uint32_t bf = r32 / (uint32_t)K, k = r32 - bf * (uint32_t)K;
  const uint32_t b = bf / (uint32_t)F2, fi = bf - b * (uint32_t)F2;
  return (long long)((b * (uint32_t)(2 * F2) + 2u * fi) * (uint32_t)K + k);
}

template <int kL, class T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                     long long n, float eps) {
  constexpr int d = kL * 16, kRows = 32 / kL;
  const int lane = threadIdx.x & 31, sub = lane / kL, lc = lane % kL;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  float4 gm[4], bt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    gm[i] = *reinterpret_cast<const float4*>(gamma + i * 4 * kL + lc * 4);
    bt[i] = *reinterpret_cast<const float4*>(beta + i * 4 * kL + lc * 4);
  }
  for (long long base = warp * kRows; base < n; base += nwarps * kRows) {
    const long long row = base + sub;
    const bool ok = row < n;
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = ok ? *reinterpret_cast<const float4*>(x + row * d + i * 4 * kL + lc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = group_sum<kL>(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = rsqrtf(group_sum<kL>(q) * (1.f / d) + eps);
    if (!ok) continue;
    if (lc == 0) { mean[row] = mu; rstd[row] = rs; }
    T* yr = y + row * d;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st4<T>(yr + i * 4 * kL + lc * 4, make_float4(v[i].x * rs * gm[i].x + bt[i].x, v[i].y * rs * gm[i].y + bt[i].y,
                                                    v[i].z * rs * gm[i].z + bt[i].z, v[i].w * rs * gm[i].w + bt[i].w));
  }
}

// ---------------------------------------------------------------------------
// K5': dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//      dgamma += sum_rows dy * xhat, dbeta += sum_rows dy   (per-lane partials, smem reduce, one atomic per column per CTA)
// ---------------------------------------------------------------------------
template <int kL, class T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ dres,
                                                     const float* __restrict__ x, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                     float* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, long long n, int uF2, int uK,
                                                     float* __restrict__ part) {
  constexpr int d = kL * 16, kRows = 32 / kL;
  __shared__ float red[2][8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, sub = lane / kL, lc = lane % kL;
  const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
  float4 gm[4], dg[4], db[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    gm[i] = *reinterpret_cast<const float4*>(gamma + i * 4 * kL + lc * 4);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long base = warp * kRows; base < n; base += nwarps * kRows) {
    const bool ok = base + sub < n;
    const long long row = ok ? base + sub : 0;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[4], g[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(x + row * d + i * 4 * kL + lc * 4);
      float4 dyv = ld4<T>(dy + row * d + i * 4 * kL + lc * 4);
      if (!ok) dyv = make_float4(0.f, 0.f, 0.f, 0.f);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dyv.x * gm[i].x, dyv.y * gm[i].y, dyv.z * gm[i].z, dyv.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
      db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
    }
    const float m1 = group_sum<kL>(s1) * (1.f / d), m2 = group_sum<kL>(s2) * (1.f / d);
    if (!ok) continue;
    // output row pointers of the two halves of the row (chunks i = 0, 1 are columns < d/2; i = 2, 3 the rest).
    // Un-merge: the row is a merged 2 x (d/2) row; its halves go to frames 2 fi and 2 fi + 1 of the (B, F, K, d/2)
    // tensor, K rows apart.  o1 is biased by -d/2 so that both halves index with the merged column.
    float* o0 = dx + row * d;
    float* o1 = o0;
    if (uF2) {
      o0 = dx + unmerged_row0(row, uF2, uK) * (d / 2);
      o1 = o0 + (long long)uK * (d / 2) - d / 2;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                             rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + row * d + i * 4 * kL + lc * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>((i < 2 ? o0 : o1) + i * 4 * kL + lc * 4) = o;
    }
  }
  fold_groups<kL>(dg);
  fold_groups<kL>(db);
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(&red[0][wib][i * 4 * kL + lc * 4]) = dg[i];
      *reinterpret_cast<float4*>(&red[1][wib][i * 4 * kL + lc * 4]) = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    if (part) {
      part[((size_t)0 * gridDim.x + blockIdx.x) * d + c] = a;
      part[((size_t)1 * gridDim.x + blockIdx.x) * d + c] = b;
    } else {
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, b);
    }
  }
}

// ---------------------------------------------------------------------------
// K6: x1 = res + dropout(a0 + bias)  and, fused, the LayerNorm that follows it: y = LN(x1) as bf16.
// Thread mapping as K5.  Dropout flags of an 8-element granule (two of the lane's float4) come from one Philox call
// keyed by (row, pair, lane column), identically in forward and backward.
// ---------------------------------------------------------------------------
template <int kL, bool kLN, class T>
__global__ void __launch_bounds__(256) bda_ln_fwd_kernel(const float* __restrict__ res, const T* __restrict__ a0,
                                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ x1,
                                                         T* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd, long long n, float eps, float scale,
                                                         uint32_t thresh, unsigned long long seed,
                                                         unsigned long long offset, int mF, int mK) {
  constexpr int d = kL * 16, kRows = 32 / kL;
  const int lane = threadIdx.x & 31, sub = lane / kL, lc = lane % kL;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  float4 bs[4], gm[4], bt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bs[i] = bias ? *reinterpret_cast<const float4*>(bias + i * 4 * kL + lc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (kLN) {
      gm[i] = *reinterpret_cast<const float4*>(gamma + i * 4 * kL + lc * 4);
      bt[i] = *reinterpret_cast<const float4*>(beta + i * 4 * kL + lc * 4);
    }
  }
  for (long long base = warp * kRows; base < n; base += nwarps * kRows) {
    const bool ok = base + sub < n;
    const long long row = ok ? base + sub : 0;   // rows past the end read row 0 and store nothing
    // all loads first (a non-uniform `if (ok)` around load + store made four dependent load -> store rounds)
    float4 v[4], av[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = row * d + i * 4 * kL + lc * 4;
      v[i] = *reinterpret_cast<const float4*>(res + e);
      av[i] = ld4<T>(a0 + e);
    }
    float s = 0.f;
    // one Philox call covers 8 elements: the lane's float4 pair (i, i+1); granule id = (row, pair, lane column)
    uint32_t keep2[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
      keep2[j] = thresh ? keep8(((unsigned long long)row * 2 + j) * kL + lc, offset, seed, thresh) : 0xFFu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t keep = keep2[i >> 1] >> (4 * (i & 1));
      v[i].x += (keep & 1u) ? (av[i].x + bs[i].x) * scale : 0.f;
      v[i].y += (keep & 2u) ? (av[i].y + bs[i].y) * scale : 0.f;
      v[i].z += (keep & 4u) ? (av[i].z + bs[i].z) * scale : 0.f;
      v[i].w += (keep & 8u) ? (av[i].w + bs[i].w) * scale : 0.f;
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    if (ok) {
      const long long orow = mF ? merged_row(row, mF, mK) : row;   // (x1 stored in the merged layout: see merged_row)
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(x1 + orow * d + i * 4 * kL + lc * 4) = v[i];
    }
    if (!kLN) continue;
    const float mu = group_sum<kL>(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = rsqrtf(group_sum<kL>(q) * (1.f / d) + eps);
    if (!ok) continue;
    if (lc == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st4<T>(y + row * d + i * 4 * kL + lc * 4, make_float4(v[i].x * rs * gm[i].x + bt[i].x, v[i].y * rs * gm[i].y + bt[i].y,
                                                             v[i].z * rs * gm[i].z + bt[i].z, v[i].w * rs * gm[i].w + bt[i].w));
  }
}

// K6': dx = g_x1 + LN'(dy)  [kLN]  or  dx = g_x1  [!kLN];   d_res = dx (written only if kLN),
//      d_a0 = mask * scale * dx (bf16),  dbias += colsum(mask * scale * dx),  dgamma / dbeta as K5'.
template <int kL, bool kLN, class T>
__global__ void __launch_bounds__(256) bda_ln_bwd_kernel(const float* __restrict__ g_x1, const T* __restrict__ dy,
                                                         const float* __restrict__ x1, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                         float* __restrict__ d_res, T* __restrict__ d_a0,
                                                         float* __restrict__ dbias, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, long long n, float scale,
                                                         uint32_t thresh, unsigned long long seed,
                                                         unsigned long long offset, float* __restrict__ part) {
  constexpr int d = kL * 16, kRows = 32 / kL;
  __shared__ float red[kLN ? 3 : 1][8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, sub = lane / kL, lc = lane % kL;
  const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
  float4 gm[4], dg[4], db[4], dbs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (kLN) gm[i] = *reinterpret_cast<const float4*>(gamma + i * 4 * kL + lc * 4);
    dg[i] = db[i] = dbs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long base = warp * kRows; base < n; base += nwarps * kRows) {
    const bool ok = base + sub < n;
    const long long row = ok ? base + sub : 0;
    float4 o[4];
    if (kLN) {
      const float mu = mean[row], rs = rstd[row];
      float4 xh[4], g[4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long e = row * d + i * 4 * kL + lc * 4;
        const float4 xv = *reinterpret_cast<const float4*>(x1 + e);
        float4 dyv = ld4<T>(dy + e);
        if (!ok) dyv = make_float4(0.f, 0.f, 0.f, 0.f);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(dyv.x * gm[i].x, dyv.y * gm[i].y, dyv.z * gm[i].z, dyv.w * gm[i].w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
        db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
      }
      const float m1 = group_sum<kL>(s1) * (1.f / d), m2 = group_sum<kL>(s2) * (1.f / d);
      if (!ok) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                           rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
        if (g_x1) {
          const float4 r = *reinterpret_cast<const float4*>(g_x1 + row * d + i * 4 * kL + lc * 4);
          o[i].x += r.x; o[i].y += r.y; o[i].z += r.z; o[i].w += r.w;
        }
        *reinterpret_cast<float4*>(d_res + row * d + i * 4 * kL + lc * 4) = o[i];
      }
    } else {
      if (!ok) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = *reinterpret_cast<const float4*>(g_x1 + row * d + i * 4 * kL + lc * 4);
    }
    uint32_t keep2[2];   // same granules as the forward kernel
#pragma unroll
    for (int j = 0; j < 2; ++j)
      keep2[j] = thresh ? keep8(((unsigned long long)row * 2 + j) * kL + lc, offset, seed, thresh) : 0xFFu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = row * d + i * 4 * kL + lc * 4;
      const uint32_t keep = keep2[i >> 1] >> (4 * (i & 1));
      const float4 da = make_float4((keep & 1u) ? o[i].x * scale : 0.f, (keep & 2u) ? o[i].y * scale : 0.f,
                                    (keep & 4u) ? o[i].z * scale : 0.f, (keep & 8u) ? o[i].w * scale : 0.f);
      st4<T>(d_a0 + e, da);
      dbs[i].x += da.x; dbs[i].y += da.y; dbs[i].z += da.z; dbs[i].w += da.w;
    }
  }
  fold_groups<kL>(dbs);
  if (kLN) { fold_groups<kL>(dg); fold_groups<kL>(db); }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(&red[0][wib][i * 4 * kL + lc * 4]) = dbs[i];
      if (kLN) {
        *reinterpret_cast<float4*>(&red[1][wib][i * 4 * kL + lc * 4]) = dg[i];
        *reinterpret_cast<float4*>(&red[2][wib][i * 4 * kL + lc * 4]) = db[i];
      }
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
    if (part) {
      part[((size_t)0 * gridDim.x + blockIdx.x) * d + c] = a;
      if (kLN) {
        part[((size_t)1 * gridDim.x + blockIdx.x) * d + c] = b;
        part[((size_t)2 * gridDim.x + blockIdx.x) * d + c] = cc;
      }
    } else {
      if (dbias) atomicAdd(dbias + c, a);
      if (kLN) { atomicAdd(dgamma + c, b); atomicAdd(dbeta + c, cc); }
    }
  }
}

// ---------------------------------------------------------------------------
// K7 / K7': g = dropout(gelu(u0 + bias)), one 8-element vector per thread-iteration; the grid stride is a
// multiple of the row length, so a thread always sees the same 8 columns (bias in registers, and
// dbias = colsum(du0) accumulates in registers in the backward).
// ---------------------------------------------------------------------------
// eight consecutive activation elements (streaming access: touched once)
template <class T> HW_DEV void ld8(const T* p, float (&o)[8]);
template <> HW_DEV void ld8<bf16>(const bf16* p, float (&o)[8]) {
  const int4 v = ld_stream16(p);
  const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = bf16_lo(w[i]); o[2 * i + 1] = bf16_hi(w[i]); }
}
template <> HW_DEV void ld8<float>(const float* p, float (&o)[8]) {
  const int4 a = ld_stream16(p), b = ld_stream16(p + 4);
  o[0] = __int_as_float(a.x); o[1] = __int_as_float(a.y); o[2] = __int_as_float(a.z); o[3] = __int_as_float(a.w);
  o[4] = __int_as_float(b.x); o[5] = __int_as_float(b.y); o[6] = __int_as_float(b.z); o[7] = __int_as_float(b.w);
}
template <class T> HW_DEV void st8(T* p, const float (&o)[8]);
template <> HW_DEV void st8<bf16>(bf16* p, const float (&o)[8]) {
  st_stream16(p, make_int4((int)pack_bf16(o[0], o[1]), (int)pack_bf16(o[2], o[3]), (int)pack_bf16(o[4], o[5]),
                           (int)pack_bf16(o[6], o[7])));
}
template <> HW_DEV void st8<float>(float* p, const float (&o)[8]) {
  st_stream16(p, make_int4(__float_as_int(o[0]), __float_as_int(o[1]), __float_as_int(o[2]), __float_as_int(o[3])));
  st_stream16(p + 4, make_int4(__float_as_int(o[4]), __float_as_int(o[5]), __float_as_int(o[6]), __float_as_int(o[7])));
}

template <class T>
__global__ void __launch_bounds__(256) bias_gelu_dropout_fwd_kernel(const T* __restrict__ u0,
                                                                    const float* __restrict__ bias,
                                                                    T* __restrict__ g, long long nvec, int cols,
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
    float u[8], o[8];
    ld8<T>(u0 + v * 8, u);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = ((keep >> i) & 1u) ? gelu_exact(u[i] + bs[i]) * scale : 0.f;
    st8<T>(g + v * 8, o);
  }
}

template <class T>
__global__ void __launch_bounds__(256) bias_gelu_dropout_bwd_kernel(const T* __restrict__ u0,
                                                                    const float* __restrict__ bias,
                                                                    const T* __restrict__ dg, T* __restrict__ du0,
                                                                    float* __restrict__ dbias, long long nvec, int cols,
                                                                    float scale, uint32_t thresh,
                                                                    unsigned long long seed, unsigned long long offset,
                                                                    float* __restrict__ part) {
  __shared__ float red[256][9];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)((v0 * 8) % cols);
  float bs[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { bs[i] = bias ? bias[c0 + i] : 0.f; acc[i] = 0.f; }
  for (long long v = v0; v < nvec; v += stride) {
    float u[8], q[8], o[8];
    ld8<T>(u0 + v * 8, u);
    ld8<T>(dg + v * 8, q);
    const uint32_t keep = thresh ? keep8((unsigned long long)v, offset, seed, thresh) : 0xFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = ((keep >> i) & 1u) ? q[i] * scale * gelu_grad(u[i] + bs[i]) : 0.f;
      acc[i] += o[i];
    }
    st8<T>(du0 + v * 8, o);
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
      if (part) part[(size_t)blockIdx.x * cols + c0 + i] = sum;
      else atomicAdd(dbias + c0 + i, sum);
    }
  }
}

// ---------------------------------------------------------------------------
// launchers (f32: the activation tensors y / dy / a0 / d_a0 / u0 / g are float instead of bf16)
// ---------------------------------------------------------------------------
static int ew_grid(long long work_items) {
  long long want = (work_items + 255) / 256;
  const long long cap = 148LL * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <class T>
static int ln_fwd_t(const float* x, const float* gamma, const float* beta, T* y, float* mean, float* rstd, long long n,
                    int d, float eps, cudaStream_t s) {
  const int grid = ew_grid(n * 32);
  switch (d) {
    case 128: ln_fwd_kernel<8, T><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    case 256: ln_fwd_kernel<16, T><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    case 512: ln_fwd_kernel<32, T><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, n, eps); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  return (int)cudaGetLastError();
}
int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long n, int d, float eps, cudaStream_t s, bool f32) {
  return f32 ? ln_fwd_t<float>(x, gamma, beta, (float*)y, mean, rstd, n, d, eps, s)
             : ln_fwd_t<bf16>(x, gamma, beta, (bf16*)y, mean, rstd, n, d, eps, s);
}

template <class T>
static int ln_bwd_t(const T* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                    const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d, cudaStream_t s,
                    int uF2, int uK) {
  cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s);
  cudaMemsetAsync(dbeta, 0, sizeof(float) * d, s);
  long long want = (n + 7) / 8;
  const int grid = (int)(want < 148LL * 4 ? (want < 1 ? 1 : want) : 148LL * 4);
  float* part;
  if (int st = det_scratch(&part, 2, grid, d, s)) return st;
  switch (d) {
    case 128: ln_bwd_kernel<8, T><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, uF2, uK, part); break;
    case 256: ln_bwd_kernel<16, T><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, uF2, uK, part); break;
    case 512: ln_bwd_kernel<32, T><<<grid, 256, 0, s>>>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, uF2, uK, part); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  det_finish(part, grid, d, dgamma, dbeta, nullptr, s);
  return (int)cudaGetLastError();
}
int launch_ln_bwd(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                  const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d, cudaStream_t s,
                  int unmerge_F2, int unmerge_K, bool f32) {
  return f32 ? ln_bwd_t<float>((const float*)dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, d, s, unmerge_F2, unmerge_K)
             : ln_bwd_t<bf16>((const bf16*)dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, d, s, unmerge_F2, unmerge_K);
}

template <bool kLN, class T>
static int bda_fwd_dispatch(int grid, cudaStream_t s, const float* res, const T* a0, const float* bias,
                            const float* gamma, const float* beta, float* x1, T* y, float* mean, float* rstd,
                            long long n, int d, float eps, float scale, uint32_t thresh, unsigned long long seed,
                            unsigned long long offset, int mF, int mK) {
  switch (d) {
    case 128: bda_ln_fwd_kernel<8, kLN, T><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset, mF, mK); break;
    case 256: bda_ln_fwd_kernel<16, kLN, T><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset, mF, mK); break;
    case 512: bda_ln_fwd_kernel<32, kLN, T><<<grid, 256, 0, s>>>(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, eps, scale, thresh, seed, offset, mF, mK); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  return 0;
}
template <class T>
static int bda_ln_fwd_t(const float* res, const T* a0, const float* bias, const float* gamma, const float* beta, float* x1,
                        T* y, float* mean, float* rstd, long long n, int d, float eps, float p, unsigned long long seed,
                        unsigned long long offset, cudaStream_t s, int mF, int mK) {
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(n * 32);
  int st = gamma ? bda_fwd_dispatch<true, T>(grid, s, res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, scale, thresh, seed, offset, mF, mK)
                 : bda_fwd_dispatch<false, T>(grid, s, res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, scale, thresh, seed, offset, mF, mK);
  if (st) return st;
  count_launch();
  return (int)cudaGetLastError();
}
int launch_bda_ln_fwd(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                      float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                      unsigned long long seed, unsigned long long offset, cudaStream_t s, int merge_F, int merge_K, bool f32) {
  return f32 ? bda_ln_fwd_t<float>(res, (const float*)a0, bias, gamma, beta, x1, (float*)y, mean, rstd, n, d, eps, p, seed, offset, s, merge_F, merge_K)
             : bda_ln_fwd_t<bf16>(res, (const bf16*)a0, bias, gamma, beta, x1, (bf16*)y, mean, rstd, n, d, eps, p, seed, offset, s, merge_F, merge_K);
}

template <bool kLN, class T>
static int bda_bwd_dispatch(int grid, cudaStream_t s, const float* g_x1, const T* dy, const float* x1,
                            const float* mean, const float* rstd, const float* gamma, float* d_res, T* d_a0,
                            float* dbias, float* dgamma, float* dbeta, long long n, int d, float scale, uint32_t thresh,
                            unsigned long long seed, unsigned long long offset, float* part) {
  switch (d) {
    case 128: bda_ln_bwd_kernel<8, kLN, T><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset, part); break;
    case 256: bda_ln_bwd_kernel<16, kLN, T><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset, part); break;
    case 512: bda_ln_bwd_kernel<32, kLN, T><<<grid, 256, 0, s>>>(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, scale, thresh, seed, offset, part); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  return 0;
}
template <class T>
static int bda_ln_bwd_t(const float* g_x1, const T* dy, const float* x1, const float* mean, const float* rstd,
                        const float* gamma, float* d_res, T* d_a0, float* dbias, float* dgamma, float* dbeta, long long n,
                        int d, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * d, s);
  if (gamma) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * d, s);
  }
  long long want = (n + 7) / 8;
  const int grid = (int)(want < 148LL * 4 ? (want < 1 ? 1 : want) : 148LL * 4);
  float* part;
  if (int st = det_scratch(&part, 3, grid, d, s)) return st;
  int st = gamma ? bda_bwd_dispatch<true, T>(grid, s, g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, scale, thresh, seed, offset, part)
                 : bda_bwd_dispatch<false, T>(grid, s, g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, scale, thresh, seed, offset, part);
  if (st) return st;
  count_launch();
  det_finish(part, grid, d, dbias, gamma ? dgamma : nullptr, gamma ? dbeta : nullptr, s);
  return (int)cudaGetLastError();
}
int launch_bda_ln_bwd(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                      const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                      long long n, int d, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s,
                      bool f32) {
  return f32 ? bda_ln_bwd_t<float>(g_x1, (const float*)dy, x1, mean, rstd, gamma, d_res, (float*)d_a0, dbias, dgamma, dbeta, n, d, p, seed, offset, s)
             : bda_ln_bwd_t<bf16>(g_x1, (const bf16*)dy, x1, mean, rstd, gamma, d_res, (bf16*)d_a0, dbias, dgamma, dbeta, n, d, p, seed, offset, s);
}

template <class T>
static int bias_gelu_dropout_t(const T* u0, const float* bias, const T* dg, T* out, float* dbias, long long n, int cols,
                               float p, unsigned long long seed, unsigned long long offset, bool backward, cudaStream_t s) {
  const long long nvec = n * cols / 8;
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(nvec);
  if (backward) {
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * cols, s);
    float* part = nullptr;
    if (dbias) { if (int st = det_scratch(&part, 1, grid, cols, s)) return st; }
    bias_gelu_dropout_bwd_kernel<T><<<grid, 256, 0, s>>>(u0, bias, dg, out, dbias, nvec, cols, scale, thresh, seed, offset,
                                                          part);
    if (part) { count_launch(); det_finish(part, grid, cols, dbias, nullptr, nullptr, s); return (int)cudaGetLastError(); }
  } else {
    bias_gelu_dropout_fwd_kernel<T><<<grid, 256, 0, s>>>(u0, bias, out, nvec, cols, scale, thresh, seed, offset);
  }
  count_launch();
  return (int)cudaGetLastError();
}
int launch_bias_gelu_dropout(const void* u0, const float* bias, const void* dg, void* out, float* dbias, long long n,
                             int cols, float p, unsigned long long seed, unsigned long long offset, bool backward,
                             cudaStream_t s, bool f32) {
  return f32 ? bias_gelu_dropout_t<float>((const float*)u0, bias, (const float*)dg, (float*)out, dbias, n, cols, p, seed, offset, backward, s)
             : bias_gelu_dropout_t<bf16>((const bf16*)u0, bias, (const bf16*)dg, (bf16*)out, dbias, n, cols, p, seed, offset, backward, s);
}

}  // namespace hwgat

// ===========================================================================
// Model head and tail, bf16 / autocast path
// ===========================================================================
namespace hwgat {

// K8: h(fp32) = dropout([sin(2 pi x.B^T), cos(2 pi x.B^T)] + pe[frame])   (HWGATE.py:343-347, 25-28)
// x: (n, C) keypoint coordinates, Bm: (E/2, C) frozen Fourier matrix, pe: (T, E), token i belongs to frame
// (i / K) % T.  Forward only: nothing upstream of the embedding is trainable.  One thread = 4 Fourier
// features of one token (sincosf shares the range reduction; the argument reaches ~300 rad, so the
// accurate sincosf, not the MUFU approximation).
__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Bm,
                                                        const float* __restrict__ pe, float* __restrict__ out,
                                                        long long n, int C, int E, int K, int T, float scale,
                                                        uint32_t thresh, unsigned long long seed,
                                                        unsigned long long offset) {
  const int half = E / 2, groups = half / 4;
  const long long total = n * groups, stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long tok = idx / groups;
    const int j0 = (int)(idx - tok * groups) * 4;
    const int frame = (int)((tok / K) % T);
    float arg[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      const float xc = 6.283185307179586f * x[tok * C + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) arg[i] = fmaf(xc, Bm[(j0 + i) * C + c], arg[i]);
    }
    float sn[4], cs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sincosf(arg[i], &sn[i], &cs[i]);
    const float4 p0 = *reinterpret_cast<const float4*>(pe + (long long)frame * E + j0);
    const float4 p1 = *reinterpret_cast<const float4*>(pe + (long long)frame * E + half + j0);
    const long long e0 = tok * E + j0, e1 = e0 + half;
    const uint32_t k0 = thresh ? keep4((unsigned long long)e0 >> 2, offset, seed, thresh) : 0xFu;
    const uint32_t k1 = thresh ? keep4((unsigned long long)e1 >> 2, offset, seed, thresh) : 0xFu;
    float4 o0 = make_float4(sn[0] + p0.x, sn[1] + p0.y, sn[2] + p0.z, sn[3] + p0.w);
    float4 o1 = make_float4(cs[0] + p1.x, cs[1] + p1.y, cs[2] + p1.z, cs[3] + p1.w);
    o0.x = (k0 & 1u) ? o0.x * scale : 0.f; o0.y = (k0 & 2u) ? o0.y * scale : 0.f;
    o0.z = (k0 & 4u) ? o0.z * scale : 0.f; o0.w = (k0 & 8u) ? o0.w * scale : 0.f;
    o1.x = (k1 & 1u) ? o1.x * scale : 0.f; o1.y = (k1 & 2u) ? o1.y * scale : 0.f;
    o1.z = (k1 & 4u) ? o1.z * scale : 0.f; o1.w = (k1 & 8u) ? o1.w * scale : 0.f;
    *reinterpret_cast<float4*>(out + e0) = o0;
    *reinterpret_cast<float4*>(out + e1) = o1;
  }
}

int launch_embed_fwd(const float* x, const float* Bm, const float* pe, float* out, long long n, int C, int E, int K,
                     int T, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  const uint32_t thresh = drop_threshold16(p);
  const float scale = thresh ? 65536.f / (65536.f - (float)thresh) : 1.f;
  const int grid = ew_grid(n * (E / 8));
  embed_fwd_kernel<<<grid, 256, 0, s>>>(x, Bm, pe, out, n, C, E, K, T, scale, thresh, seed, offset);
  count_launch();
  return (int)cudaGetLastError();
}

// K9: pooled[b, :] = mean over the tokens of sample b of LayerNorm(x[b, token, :])      (HWGATE.py:353-354)
// One CTA per (sample, token slice); warp per row as K5; per-lane partial sums -> smem -> atomics into the
// zeroed (B, d) output.  mean / rstd are saved for the backward.
template <int kV>
__global__ void __launch_bounds__(256) ln_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ pooled,
                                                          float* __restrict__ mean, float* __restrict__ rstd,
                                                          float* __restrict__ scratch, int tokens, int slices,
                                                          float eps, int kp_real, int kp_pad,
                                                          const float* __restrict__ tok_w) {
  // tok_w != nullptr (GATE's learned `weightedAvg`, GATE.py:207): pooled = sum_t tok_w[t] * LayerNorm(x)[t]; the
  // caller folds beta * sum(tok_w) + the pool's bias into `beta`.  nullptr: the mean (weights 1, scaled once at the end).
  constexpr int d = kV * 128;
  __shared__ float red[8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int b = blockIdx.x / slices, sl = blockIdx.x - b * slices;
  // a padded keypoint axis (HGATE: 29 keypoints stored as 32): `tokens` counts the REAL tokens of a sample; token tk
  // sits at row (tk / kp_real) * kp_pad + tk % kp_real of the sample's tokens / kp_real * kp_pad stored rows
  const int tokens_st = kp_pad ? tokens / kp_real * kp_pad : tokens;
  float4 gm[kV], acc[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    gm[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int tk = sl * 8 + wib; tk < tokens; tk += slices * 8) {
    const int tks = kp_pad ? (tk / kp_real) * kp_pad + tk % kp_real : tk;
    const long long row = (long long)b * tokens_st + tks;
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      v[i] = *reinterpret_cast<const float4*>(x + row * d + i * 128 + lane * 4);
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
    const float rw = tok_w ? rs * tok_w[tk] : rs;
#pragma unroll
    for (int i = 0; i < kV; ++i) {  // sum of xhat; gamma / beta / 1/tokens are applied once at the end
      acc[i].x += v[i].x * rw; acc[i].y += v[i].y * rw; acc[i].z += v[i].z * rw; acc[i].w += v[i].w * rw;
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) *reinterpret_cast<float4*>(&red[wib][i * 128 + lane * 4]) = acc[i];
  __syncthreads();
  const float inv = tok_w ? 1.f : 1.f / tokens;
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][c];
    if (slices == 1) pooled[(long long)b * d + c] = a * inv * gamma[c] + beta[c];
    else scratch[((long long)b * slices + sl) * d + c] = a;       // summed in a fixed order by ln_pool_finish_kernel
  }
}

// pooled[b, c] = (sum over the slices, in slice order) / tokens * gamma[c] + beta[c]: deterministic (no atomics)
__global__ void ln_pool_finish_kernel(const float* __restrict__ scratch, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ pooled, int B, int d,
                                      int slices, float inv) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * d) return;
  const int b = idx / d, c = idx - b * d;
  float a = 0.f;
  for (int sl = 0; sl < slices; ++sl) a += scratch[((long long)b * slices + sl) * d + c];
  pooled[idx] = a * inv * gamma[c] + beta[c];
}

// K9': dx[b, t, :] = LayerNorm'(g[b, :] / tokens) per row;  dgamma += sum_rows (g/tokens) * xhat ; dbeta = sum_b g
template <int kV>
__global__ void __launch_bounds__(256) ln_pool_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, float* __restrict__ dx,
                                                          float* __restrict__ dgamma, int tokens, int slices,
                                                          int kp_real, int kp_pad, float* __restrict__ part,
                                                          const float* __restrict__ tok_w, float* __restrict__ dw_part) {
  // tok_w != nullptr: the weighted pool; dw_part[b][t] = sum_c g[b,c] gamma[c] xhat[b,t,c] (the caller adds g.beta and
  // sums over the samples in a fixed order)
  constexpr int d = kV * 128;
  __shared__ float red[8][d];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int b = blockIdx.x / slices, sl = blockIdx.x - b * slices;
  const float inv = tok_w ? 1.f : 1.f / tokens;
  float4 gy[kV], gg[kV], dg[kV];
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const float4 gv = *reinterpret_cast<const float4*>(g + (long long)b * d + i * 128 + lane * 4);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);
    gy[i] = make_float4(gv.x * inv, gv.y * inv, gv.z * inv, gv.w * inv);
    gg[i] = make_float4(gy[i].x * gm.x, gy[i].y * gm.y, gy[i].z * gm.z, gy[i].w * gm.w);
    s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float m1 = warp_sum(s1) * (1.f / d);  // the same for every token of the sample
  const int tokens_st = kp_pad ? tokens / kp_real * kp_pad : tokens;   // (padded keypoint axis: see ln_pool_fwd_kernel;
  for (int tk = sl * 8 + wib; tk < tokens; tk += slices * 8) {         //  the padded rows of dx are not written)
    const int tks = kp_pad ? (tk / kp_real) * kp_pad + tk % kp_real : tk;
    const long long row = (long long)b * tokens_st + tks;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[kV];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(x + row * d + i * 128 + lane * 4);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
    }
    const float wt = tok_w ? tok_w[tk] : 1.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      dg[i].x += wt * gy[i].x * xh[i].x; dg[i].y += wt * gy[i].y * xh[i].y;
      dg[i].z += wt * gy[i].z * xh[i].z; dg[i].w += wt * gy[i].w * xh[i].w;
    }
    const float s2w = warp_sum(s2);
    if (dw_part && lane == 0) dw_part[(long long)b * tokens + tk] = s2w;
    const float m2 = s2w * (1.f / d), rw = rs * wt;
#pragma unroll
    for (int i = 0; i < kV; ++i)
      *reinterpret_cast<float4*>(dx + row * d + i * 128 + lane * 4) =
          make_float4(rw * (gg[i].x - m1 - xh[i].x * m2), rw * (gg[i].y - m1 - xh[i].y * m2),
                      rw * (gg[i].z - m1 - xh[i].z * m2), rw * (gg[i].w - m1 - xh[i].w * m2));
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) *reinterpret_cast<float4*>(&red[wib][i * 128 + lane * 4]) = dg[i];
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][c];
    if (part) part[(size_t)blockIdx.x * d + c] = a;
    else atomicAdd(dgamma + c, a);
  }
}

static int pool_slices(int B, int tokens) {
  int s = (148 * 4 + B - 1) / B;
  const int maxs = (tokens + 7) / 8;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

size_t ln_pool_scratch_bytes(int B, int tokens, int d) {
  const int slices = pool_slices(B, tokens);
  return slices > 1 ? sizeof(float) * (size_t)B * slices * d : 0;
}

int launch_ln_pool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean, float* rstd,
                       float* scratch, int B, int tokens, int d, float eps, cudaStream_t s, int kp_real, int kp_pad,
                       const float* tok_w) {
  const int slices = pool_slices(B, tokens);
  switch (d) {
    case 128: ln_pool_fwd_kernel<1><<<B * slices, 256, 0, s>>>(x, gamma, beta, pooled, mean, rstd, scratch, tokens, slices, eps, kp_real, kp_pad, tok_w); break;
    case 256: ln_pool_fwd_kernel<2><<<B * slices, 256, 0, s>>>(x, gamma, beta, pooled, mean, rstd, scratch, tokens, slices, eps, kp_real, kp_pad, tok_w); break;
    case 512: ln_pool_fwd_kernel<4><<<B * slices, 256, 0, s>>>(x, gamma, beta, pooled, mean, rstd, scratch, tokens, slices, eps, kp_real, kp_pad, tok_w); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  if (slices > 1) {
    ln_pool_finish_kernel<<<(B * d + 255) / 256, 256, 0, s>>>(scratch, gamma, beta, pooled, B, d, slices,
                                                              tok_w ? 1.f : 1.f / tokens);
    count_launch();
  }
  return (int)cudaGetLastError();
}

int launch_ln_pool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       float* dx, float* dgamma, int B, int tokens, int d, cudaStream_t s, int kp_real, int kp_pad,
                       const float* tok_w, float* dw_part, float* d_tok_w) {
  cudaMemsetAsync(dgamma, 0, sizeof(float) * d, s);
  const int slices = pool_slices(B, tokens);
  float* part;
  if (int st = det_scratch(&part, 1, B * slices, d, s)) return st;
  switch (d) {
    case 128: ln_pool_bwd_kernel<1><<<B * slices, 256, 0, s>>>(g, x, mean, rstd, gamma, dx, dgamma, tokens, slices, kp_real, kp_pad, part, tok_w, dw_part); break;
    case 256: ln_pool_bwd_kernel<2><<<B * slices, 256, 0, s>>>(g, x, mean, rstd, gamma, dx, dgamma, tokens, slices, kp_real, kp_pad, part, tok_w, dw_part); break;
    case 512: ln_pool_bwd_kernel<4><<<B * slices, 256, 0, s>>>(g, x, mean, rstd, gamma, dx, dgamma, tokens, slices, kp_real, kp_pad, part, tok_w, dw_part); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  det_finish(part, B * slices, d, dgamma, nullptr, nullptr, s);
  if (dw_part && d_tok_w) {   // d_tok_w[t] = sum over the samples, in sample order (always deterministic)
    col_finish_kernel<<<(tokens + 127) / 128, 128, 0, s>>>(dw_part, B, tokens, d_tok_w, nullptr, nullptr);
    count_launch();
  }
  return (int)cudaGetLastError();
}

}  // namespace hwgat
