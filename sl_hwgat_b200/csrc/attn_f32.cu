// fp32 parity mode of K2/K3 (the north_star's "fp32 within 1e-5" path).
// True fp32 FFMA arithmetic - no TF32, no tensor cores - so it is a
// correctness mode, not the timed one.  Same C entry points, same index
// arithmetic (roll / partition never materialised), same mask semantics as the
// bf16 kernels:
//   qkv  = xn . Wqkv^T + b                      (HWGATE.py:86)
//   S    = (q*scale) . k^T                      (HWGATE.py:89-91)
//   keep = !(softmax(S) > thr)     [training]   (HWGATE.py:94-100)
//   live = mask & keep & (S != 0)               (HWGATE.py:102-110)
//   P    = softmax(live ? S : -10000)           (HWGATE.py:110-111)
//   out  = P . v                                (HWGATE.py:114)
#include "common.cuh"

namespace hwgat {

// ---------------------------------------------------------------------------
// Tiled FFMA GEMM: C[M,N] (+)= A.B.  A is [M][K] (TA=false) or [K][M] (TA=true);
// B is [N][K] (TB=true, i.e. a Linear weight) or [K][N] (TB=false).
// kAtomic: split-K over blockIdx.z, fp32 atomics into a zeroed C.
// ---------------------------------------------------------------------------
template <bool TA, bool TB, bool kAtomic, bool kBias>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       const float* __restrict__ bias, float* __restrict__ C,
                                                       int M, int N, int K, int k_per_split) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      int kk, mm;
      if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA ? A[(size_t)gk * M + gm] : A[(size_t)gm * K + gk];
      As[kk][mm] = v;
    }
    for (int i = threadIdx.x; i < BN * BK; i += 256) {
      int kk, nn;
      if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < kend) v = TB ? Bm[(size_t)gn * K + gk] : Bm[(size_t)gk * N + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
    // blocked summation: a 16-term partial per k tile, then one add into the running sum, so
    // the rounding error grows with K/16 + 16 instead of K (the 1e-5 budget is tight at d=512)
    float part[4][4] = {};
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float v = acc[i][j] + (kBias ? bias[gn] : 0.f);
        if (kAtomic) atomicAdd(&C[(size_t)gm * N + gn], v); else C[(size_t)gm * N + gn] = v;
      }
    }
}

// column sums of a [rows][cols] matrix into a zeroed vector (d_b = sum_t dQKV[t,:])
__global__ void colsum_f32_kernel(const float* __restrict__ X, float* __restrict__ out, long long rows, int cols,
                                  long long rows_per_block, float* __restrict__ part = nullptr) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += X[r * cols + c];
  if (part) part[(size_t)blockIdx.y * cols + c] = s;   // deterministic mode: fixed-order finish (det_finish)
  else atomicAdd(&out[c], s);
}

// ---------------------------------------------------------------------------
// attention core, one warp per (window, head); lane = query token (forward and
// dQ) and = key token (dK, dV).  4 windows of one tile per CTA.
// ---------------------------------------------------------------------------
struct CoreSmemF32 {
  float k[kTok][kHd];
  float v[kTok][kHd];
};
// 24.8 KB per warp: k, v while the lanes are queries, then q, d_out in the same two tiles while they are keys
// (with all four resident, 41 KB per warp, one 4-warp CTA filled an SM: latency-bound at 4 warps per SM)
struct CoreSmemF32Bwd {
  float a[kTok][kHd];       // k rows, then q rows (scaled)
  float b[kTok][kHd];       // v rows, then d_out rows of this window/head
  float p[kTok][kTok + 1];
  float ds[kTok][kTok + 1];
};

// 64-term dot product as 4 interleaved partial sums (shorter rounding chains, more ILP)
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

// probabilities of one query row held by one lane: s[] in, p[] out (in place)
HW_DEV uint32_t row_softmax_f32(float (&s)[kTok], uint32_t mask_word, float threshold) {
  uint32_t live = mask_word;
  if (threshold >= 0.f) {
    float m = s[0];
#pragma unroll
    for (int j = 1; j < kTok; ++j) m = fmaxf(m, s[j]);
    float e[kTok], sum = 0.f;
#pragma unroll
    for (int j = 0; j < kTok; ++j) { e[j] = expf(s[j] - m); sum += e[j]; }
#pragma unroll
    for (int j = 0; j < kTok; ++j)
      if (e[j] / sum > threshold) live &= ~(1u << j);
  }
#pragma unroll
  for (int j = 0; j < kTok; ++j)
    if (s[j] == 0.f) live &= ~(1u << j);
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < kTok; ++j) { s[j] = ((live >> j) & 1u) ? s[j] : kNegFill; m = fmaxf(m, s[j]); }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) { s[j] = expf(s[j] - m); sum += s[j]; }
  float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < kTok; ++j) s[j] *= inv;
  return live;
}

// 16-byte asynchronous global -> shared copies (LDGSTS): the k / v / q / dO rows of a window go to shared memory
// without passing through registers, all 32 copies of a tile in flight at once
HW_DEV void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
HW_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// rows [0, 32) of one window / head: src_col = first of the 64 columns inside a row of `stride` floats
// (r0, r1: TileGeom::window_base - the token rows of the window's two frames)
HW_DEV void copy_window_rows(float (*dst)[kHd], const float* __restrict__ src, int stride, int src_col, long long r0,
                             long long r1, int lane) {
  const int half = lane >> 4, chunk = lane & 15;     // two rows per instruction, 16 lanes x 16 bytes per row
  const float* s0 = src + (r0 + half) * stride + src_col + chunk * 4;      // rows half, half + 2, ... of frame 0
  const float* s1 = src + (r1 + half) * stride + src_col + chunk * 4;
#pragma unroll
  for (int it = 0; it < kWin / 2; ++it) {
    cp_async16(&dst[2 * it + half][chunk * 4], s0 + (long long)(2 * it) * stride);
    cp_async16(&dst[kWin + 2 * it + half][chunk * 4], s1 + (long long)(2 * it) * stride);
  }
}

__global__ void __launch_bounds__(128) attn_core_fwd_f32_kernel(const float* __restrict__ qkv,
                                                                const uint32_t* __restrict__ bits, float threshold,
                                                                float* __restrict__ out, TileGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  CoreSmemF32& sm = reinterpret_cast<CoreSmemF32*>(smem_raw)[w];
  const int tile = blockIdx.x, h = blockIdx.y;
  const int d = g.d, d3 = 3 * d;
  const float scale = 0.125f;  // head_dim^-0.5, head_dim = 64

  // k, v rows of the window -> smem, asynchronously (the own q row is fetched meanwhile)
  long long r0, r1;
  g.window_base(tile, w, r0, r1);
  copy_window_rows(sm.k, qkv, d3, d + h * kHd, r0, r1, lane);
  copy_window_rows(sm.v, qkv, d3, 2 * d + h * kHd, r0, r1, lane);
  const long long my_row = lane < kWin ? r0 + lane : r1 + lane - kWin;
  float q[kHd];
  {
    const float4* src = reinterpret_cast<const float4*>(qkv + my_row * d3 + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e) {
      float4 t = src[e];
      q[4 * e] = t.x * scale; q[4 * e + 1] = t.y * scale; q[4 * e + 2] = t.z * scale; q[4 * e + 3] = t.w * scale;
    }
  }
  cp_async_wait_all();
  __syncwarp();
  float s[kTok];
#pragma unroll
  for (int j = 0; j < kTok; ++j) s[j] = dot64(q, sm.k[j]);
  const uint32_t mword = bits[g.mask_base(tile) + w * kTok + lane];
  row_softmax_f32(s, mword, threshold);
  float o[kHd];
#pragma unroll
  for (int e = 0; e < kHd; ++e) o[e] = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j)
#pragma unroll
    for (int e = 0; e < kHd; ++e) o[e] = fmaf(s[j], sm.v[j][e], o[e]);
  float4* dst = reinterpret_cast<float4*>(out + my_row * d + h * kHd);
#pragma unroll
  for (int e = 0; e < kHd / 4; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) attn_core_bwd_f32_kernel(const float* __restrict__ qkv,
                                                                         const float* __restrict__ d_out,
                                                                         const uint32_t* __restrict__ bits,
                                                                         float threshold, float* __restrict__ dqkv,
                                                                         TileGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  CoreSmemF32Bwd& sm = reinterpret_cast<CoreSmemF32Bwd*>(smem_raw)[wl];
  const int widx = blockIdx.x * kWarps + wl;  // window index over (tile, w)
  const int tile = widx >> 2, w = widx & 3, h = blockIdx.y;
  const int d = g.d, d3 = 3 * d;
  const float scale = 0.125f;

  long long r0, r1;
  g.window_base(tile, w, r0, r1);
  copy_window_rows(sm.a, qkv, d3, d + h * kHd, r0, r1, lane);          // k rows
  copy_window_rows(sm.b, qkv, d3, 2 * d + h * kHd, r0, r1, lane);      // v rows
  const long long my_row = lane < kWin ? r0 + lane : r1 + lane - kWin;
  // ---- lane = query i: P row, dP row, dS row, dQ row
  // (own q / d_out rows come from global into registers: reading smem row `lane`
  //  from every lane would be a 32-way bank conflict)
  float s[kTok];
  {
    float qr[kHd];
    const float4* src = reinterpret_cast<const float4*>(qkv + my_row * d3 + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e) {
      float4 t = src[e];
      qr[4 * e] = t.x * scale; qr[4 * e + 1] = t.y * scale; qr[4 * e + 2] = t.z * scale; qr[4 * e + 3] = t.w * scale;
    }
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kTok; ++j) s[j] = dot64(qr, sm.a[j]);
  }
  const uint32_t mword = bits[g.mask_base(tile) + w * kTok + lane];
  const uint32_t live = row_softmax_f32(s, mword, threshold);  // s[] now holds P
  float dp[kTok], dsum = 0.f;
  {
    float gr[kHd];
    const float4* src = reinterpret_cast<const float4*>(d_out + my_row * d + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e) {
      float4 t = src[e];
      gr[4 * e] = t.x; gr[4 * e + 1] = t.y; gr[4 * e + 2] = t.z; gr[4 * e + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < kTok; ++j) {
      float a = dot64(gr, sm.b[j]);
      dp[j] = a;
      dsum = fmaf(s[j], a, dsum);
    }
  }
  // v is no longer needed: the d_out rows of the window replace it while dS and dQ are formed
  __syncwarp();
  copy_window_rows(sm.b, d_out, d, h * kHd, r0, r1, lane);
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    float dsv = ((live >> j) & 1u) ? s[j] * (dp[j] - dsum) : 0.f;
    sm.p[lane][j] = s[j];
    sm.ds[lane][j] = dsv;
    dp[j] = dsv;
  }
  {
    float dq[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) dq[e] = 0.f;
#pragma unroll
    for (int j = 0; j < kTok; ++j)
#pragma unroll
      for (int e = 0; e < kHd; ++e) dq[e] = fmaf(dp[j], sm.a[j][e], dq[e]);
    float4* dst = reinterpret_cast<float4*>(dqkv + my_row * d3 + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e)
      dst[e] = make_float4(dq[4 * e] * scale, dq[4 * e + 1] * scale, dq[4 * e + 2] * scale, dq[4 * e + 3] * scale);
  }
  // ---- lane = key j: dK row = scale * sum_i dS[i][j] q_i ; dV row = sum_i P[i][j] g_i   (q rows replace k)
  __syncwarp();
  copy_window_rows(sm.a, qkv, d3, h * kHd, r0, r1, lane);
  cp_async_wait_all();
  __syncwarp();
  {
    float acc[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) acc[e] = 0.f;
    for (int i = 0; i < kTok; ++i) {
      float c = sm.ds[i][lane];
#pragma unroll
      for (int e = 0; e < kHd; ++e) acc[e] = fmaf(c, sm.a[i][e], acc[e]);
    }
    float4* dst = reinterpret_cast<float4*>(dqkv + my_row * d3 + d + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e)
      dst[e] = make_float4(acc[4 * e] * scale, acc[4 * e + 1] * scale, acc[4 * e + 2] * scale, acc[4 * e + 3] * scale);
#pragma unroll
    for (int e = 0; e < kHd; ++e) acc[e] = 0.f;
    for (int i = 0; i < kTok; ++i) {
      float c = sm.p[i][lane];
#pragma unroll
      for (int e = 0; e < kHd; ++e) acc[e] = fmaf(c, sm.b[i][e], acc[e]);
    }
    dst = reinterpret_cast<float4*>(dqkv + my_row * d3 + 2 * d + h * kHd);
#pragma unroll
    for (int e = 0; e < kHd / 4; ++e) dst[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
  }
}

static int check_last() { return (int)cudaGetLastError(); }

// the weight side of an fp32 attention backward: d_xn = dQKV . Wqkv, d_w = dQKV^T . xn, d_b = column sums of dQKV
int qkv_weight_grads_f32(const float* dqkv, const float* xn, const float* w_qkv, float* d_xn, float* d_w,
                                float* d_b, long long n, int d, cudaStream_t s) {
  const int d3 = 3 * d;
  int st;
  if (x3_enabled() && x3_supported(n, d, d3)) return linear_f32_bwd(dqkv, xn, w_qkv, d_xn, d_w, d_b, (int)n, d, d3, s);
  // d_xn = dQKV . Wqkv        [n, 3d] x [3d, d]
  dim3 g1((d + 63) / 64, (unsigned)((n + 63) / 64), 1);
  gemm_f32_kernel<false, false, false, false><<<g1, 256, 0, s>>>(dqkv, w_qkv, nullptr, d_xn, (int)n, d, d3, d3);
  count_launch();
  if ((st = check_last())) return st;
  // d_w = dQKV^T . xn         [3d, n] x [n, d]   (split-K over tokens, atomics into zeroed d_w)
  cudaMemsetAsync(d_w, 0, sizeof(float) * d3 * d, s);
  cudaMemsetAsync(d_b, 0, sizeof(float) * d3, s);
  int splits = (int)((n + 2047) / 2048);
  if (splits > 512) splits = 512;
  if (deterministic()) splits = 1;
  int kps = (int)(((n + splits - 1) / splits + 15) / 16 * 16);
  dim3 g2((d + 63) / 64, (d3 + 63) / 64, (unsigned)((n + kps - 1) / kps));
  gemm_f32_kernel<true, false, true, false><<<g2, 256, 0, s>>>(dqkv, xn, nullptr, d_w, d3, d, (int)n, kps);
  count_launch();
  if ((st = check_last())) return st;
  long long rpb = 512;
  dim3 g3((d3 + 127) / 128, (unsigned)((n + rpb - 1) / rpb));
  float* part;
  if ((st = det_scratch(&part, 1, (int)g3.y, d3, s))) return st;
  colsum_f32_kernel<<<g3, 128, 0, s>>>(dqkv, d_b, n, d3, rpb, part);
  count_launch();
  det_finish(part, (int)g3.y, d3, d_b, nullptr, nullptr, s);
  return check_last();
}


int attn_fwd_f32(const AttnArgs& a, cudaStream_t s) {
  const long long n = (long long)a.B * a.F * a.K;
  const int d = a.d;
  float* qkv = (float*)a.workspace;
  int st = linear_f32_fwd((const float*)a.xn, (const float*)a.w_qkv, a.b_qkv, qkv, (int)n, d, 3 * d, s);
  if (st) return st;
  TileGeom g = make_geom(a.F, a.K, d, a.shift, a.layout);
  size_t smem = 4 * sizeof(CoreSmemF32);
  cudaFuncSetAttribute(attn_core_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_core_fwd_f32_kernel<<<dim3(a.tiles(), a.heads), 128, smem, s>>>(qkv, a.bits, a.threshold, (float*)a.out, g);
  count_launch();
  return check_last();
}

// kept_qkv: the projected rows the forward left in ITS workspace (the caller kept that buffer): no re-projection, and
// the backward workspace holds dqkv only.  NULL: re-project into the first half of the backward workspace.
int attn_bwd_f32(const AttnArgs& a, cudaStream_t s, const float* kept_qkv) {
  const long long n = (long long)a.B * a.F * a.K;
  const int d = a.d, d3 = 3 * d;
  int st;
  const float* qkv = kept_qkv;
  float* dqkv = (float*)a.workspace;
  if (!kept_qkv) {
    float* q = (float*)a.workspace;
    dqkv = q + n * d3;
    if ((st = linear_f32_fwd((const float*)a.xn, (const float*)a.w_qkv, a.b_qkv, q, (int)n, d, d3, s))) return st;
    qkv = q;
  }
  TileGeom g = make_geom(a.F, a.K, d, a.shift, a.layout);
  constexpr int kWarps = 4;
  size_t smem = kWarps * sizeof(CoreSmemF32Bwd);
  cudaFuncSetAttribute(attn_core_bwd_f32_kernel<kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_core_bwd_f32_kernel<kWarps><<<dim3(a.tiles() * 4 / kWarps, a.heads), kWarps * 32, smem, s>>>(
      qkv, (const float*)a.d_out, a.bits, a.threshold, dqkv, g);
  count_launch();
  if ((st = check_last())) return st;
  return qkv_weight_grads_f32(dqkv, (const float*)a.xn, (const float*)a.w_qkv, (float*)a.d_xn, a.d_w, a.d_b, n, d, s);
}

// ---------------------------------------------------------------------------
// fp32 parity mode of K15 / K16 (band_attn.cu): the frame-banded graph attention of WGATE / GATE in true fp32.
// One thread per (token, head) walks the set bits of its three band words - the graph has a handful of edges per
// token - so nothing is staged and nothing is tiled: a correctness mode (1e-5 against the reference's fp64 outputs),
// not a timed one.  The backward is gather-only (no atomics): a thread forms dQ of its token as a query, then dK / dV
// of the same token as a key by testing, for each keypoint of the three neighbouring frames, whether that query
// attends it.
// ---------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128) band_fwd_f32_kernel(const float* __restrict__ qkv, const uint32_t* __restrict__ bits,
                                                           float* __restrict__ out, float* __restrict__ lse,
                                                           long long n, int F, int K, int d, int heads, int W, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * heads) return;
  const int h = (int)(idx % heads);
  const long long tok = idx / heads;
  const int k = (int)(tok % K), f = (int)((tok / K) % F), w = k / W, i = k - w * W;
  const int d3 = 3 * d;
  const float* qp = qkv + tok * d3 + h * HD;
  float q[HD];
#pragma unroll
  for (int e = 0; e < HD; ++e) q[e] = qp[e] * scale;                       // q * scale first, as WGATE.py:96
  const long long win0 = tok - i;                                          // first token of the window in this frame
  float m = -INFINITY;
  for (int r = 0; r < 3; ++r) {
    const int fr = f - 1 + r;
    if (fr < 0 || fr >= F) continue;
    uint32_t word = bits[(w * W + i) * 3 + r];
    while (word) {
      const int j = __ffs(word) - 1;
      word &= word - 1;
      const float* kp = qkv + (win0 + (long long)(r - 1) * K + j) * d3 + d + h * HD;
      float sdot = 0.f;
#pragma unroll
      for (int e = 0; e < HD; ++e) sdot = fmaf(q[e], kp[e], sdot);
      m = fmaxf(m, sdot);
    }
  }
  float l = 0.f, o[HD];
#pragma unroll
  for (int e = 0; e < HD; ++e) o[e] = 0.f;
  for (int r = 0; r < 3; ++r) {
    const int fr = f - 1 + r;
    if (fr < 0 || fr >= F) continue;
    uint32_t word = bits[(w * W + i) * 3 + r];
    while (word) {
      const int j = __ffs(word) - 1;
      word &= word - 1;
      const float* kp = qkv + (win0 + (long long)(r - 1) * K + j) * d3 + d + h * HD;
      float sdot = 0.f;
#pragma unroll
      for (int e = 0; e < HD; ++e) sdot = fmaf(q[e], kp[e], sdot);
      const float pr = expf(sdot - m);
      l += pr;
      const float* vp = kp + d;
#pragma unroll
      for (int e = 0; e < HD; ++e) o[e] = fmaf(pr, vp[e], o[e]);
    }
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  float* op = out + tok * d + h * HD;
#pragma unroll
  for (int e = 0; e < HD; ++e) op[e] = o[e] * inv;
  if (lse) lse[idx] = l > 0.f ? m + logf(l) : 0.f;                          // natural-log logsumexp (fp32 mode)
}

template <int HD>
__global__ void __launch_bounds__(128) band_bwd_f32_kernel(const float* __restrict__ qkv, const uint32_t* __restrict__ bits,
                                                           const float* __restrict__ out, const float* __restrict__ lse,
                                                           const float* __restrict__ d_out, float* __restrict__ dqkv,
                                                           long long n, int F, int K, int d, int heads, int W, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * heads) return;
  const int h = (int)(idx % heads);
  const long long tok = idx / heads;
  const int k = (int)(tok % K), f = (int)((tok / K) % F), w = k / W, i = k - w * W;
  const int d3 = 3 * d;
  const long long win0 = tok - i;
  // ---- as a query: dQ_i = scale * sum_j dS_ij k_j,  dS_ij = P_ij (dO_i . v_j - dO_i . O_i)
  {
    const float* qp = qkv + tok * d3 + h * HD;
    const float* gp = d_out + tok * d + h * HD;
    const float* op = out + tok * d + h * HD;
    float q[HD], g[HD], dq[HD];
    float delta = 0.f;
#pragma unroll
    for (int e = 0; e < HD; ++e) {
      q[e] = qp[e] * scale;
      g[e] = gp[e];
      delta = fmaf(g[e], op[e], delta);
      dq[e] = 0.f;
    }
    const float li = lse[idx];
    for (int r = 0; r < 3; ++r) {
      const int fr = f - 1 + r;
      if (fr < 0 || fr >= F) continue;
      uint32_t word = bits[(w * W + i) * 3 + r];
      while (word) {
        const int j = __ffs(word) - 1;
        word &= word - 1;
        const float* kp = qkv + (win0 + (long long)(r - 1) * K + j) * d3 + d + h * HD;
        const float* vp = kp + d;
        float sdot = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < HD; ++e) { sdot = fmaf(q[e], kp[e], sdot); dp = fmaf(g[e], vp[e], dp); }
        const float ds = expf(sdot - li) * (dp - delta);
#pragma unroll
        for (int e = 0; e < HD; ++e) dq[e] = fmaf(ds, kp[e], dq[e]);
      }
    }
    float* dst = dqkv + tok * d3 + h * HD;
#pragma unroll
    for (int e = 0; e < HD; ++e) dst[e] = dq[e] * scale;
  }
  // ---- as a key: dK_j = scale * sum_i dS_ij q_i,  dV_j = sum_i P_ij dO_i  over the queries i that attend j
  {
    const float* kp = qkv + tok * d3 + d + h * HD;
    const float* vp = kp + d;
    float kk[HD], vv[HD], dk[HD], dv[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) { kk[e] = kp[e]; vv[e] = vp[e]; dk[e] = 0.f; dv[e] = 0.f; }
    for (int r = 0; r < 3; ++r) {
      const int fq = f + 1 - r;                       // the query frame that sees this frame as its frame fq - 1 + r
      if (fq < 0 || fq >= F) continue;
      for (int qi = 0; qi < W; ++qi) {
        if (!((bits[(w * W + qi) * 3 + r] >> i) & 1u)) continue;
        const long long qt = win0 + (long long)(1 - r) * K + qi;
        const float* qp = qkv + qt * d3 + h * HD;
        const float* gp = d_out + qt * d + h * HD;
        const float* op = out + qt * d + h * HD;
        float sdot = 0.f, dp = 0.f, delta = 0.f;
#pragma unroll
        for (int e = 0; e < HD; ++e) {
          sdot = fmaf(qp[e] * scale, kk[e], sdot);
          dp = fmaf(gp[e], vv[e], dp);
          delta = fmaf(gp[e], op[e], delta);
        }
        const float pr = expf(sdot - lse[qt * heads + h]);
        const float ds = pr * (dp - delta) * scale;
#pragma unroll
        for (int e = 0; e < HD; ++e) { dk[e] = fmaf(ds, qp[e], dk[e]); dv[e] = fmaf(pr, gp[e], dv[e]); }
      }
    }
    float* dst = dqkv + tok * d3 + d + h * HD;
#pragma unroll
    for (int e = 0; e < HD; ++e) { dst[e] = dk[e]; dst[d + e] = dv[e]; }
  }
}

// forward: qkv (caller's buffer, kept) = xn . Wqkv^T + b ; out, lse (optional)
int band_attn_fwd_f32(const float* xn, const float* w_qkv, const float* b_qkv, const uint32_t* bits, float* out,
                      float* qkv, float* lse, int B, int F, int K, int d, int heads, int W, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  const int hd = d / heads;
  if ((n + 63) / 64 > 65535) return HWGAT_ERR_UNSUPPORTED;          // gemm_f32_kernel puts the token tiles on grid.y
  int st = linear_f32_fwd(xn, w_qkv, b_qkv, qkv, (int)n, d, 3 * d, s);
  if (st) return st;
  const float scale = 1.0f / sqrtf((float)hd);
  const unsigned grid = (unsigned)((n * heads + 127) / 128);
  switch (hd) {
    case 16: band_fwd_f32_kernel<16><<<grid, 128, 0, s>>>(qkv, bits, out, lse, n, F, K, d, heads, W, scale); break;
    case 32: band_fwd_f32_kernel<32><<<grid, 128, 0, s>>>(qkv, bits, out, lse, n, F, K, d, heads, W, scale); break;
    case 64: band_fwd_f32_kernel<64><<<grid, 128, 0, s>>>(qkv, bits, out, lse, n, F, K, d, heads, W, scale); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  return check_last();
}

// backward workspace: dqkv [n, 3d] fp32
int band_attn_bwd_f32(const float* xn, const float* w_qkv, const uint32_t* bits, const float* qkv, const float* ctx,
                      const float* lse, const float* d_out, float* d_xn, float* d_w, float* d_b, void* workspace, int B,
                      int F, int K, int d, int heads, int W, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  const int hd = d / heads;
  if ((n + 63) / 64 > 65535) return HWGAT_ERR_UNSUPPORTED;
  float* dqkv = (float*)workspace;
  const float scale = 1.0f / sqrtf((float)hd);
  const unsigned grid = (unsigned)((n * heads + 127) / 128);
  switch (hd) {
    case 16: band_bwd_f32_kernel<16><<<grid, 128, 0, s>>>(qkv, bits, ctx, lse, d_out, dqkv, n, F, K, d, heads, W, scale); break;
    case 32: band_bwd_f32_kernel<32><<<grid, 128, 0, s>>>(qkv, bits, ctx, lse, d_out, dqkv, n, F, K, d, heads, W, scale); break;
    case 64: band_bwd_f32_kernel<64><<<grid, 128, 0, s>>>(qkv, bits, ctx, lse, d_out, dqkv, n, F, K, d, heads, W, scale); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  count_launch();
  int st = check_last();
  if (st) return st;
  return qkv_weight_grads_f32(dqkv, xn, w_qkv, d_xn, d_w, d_b, n, d, s);
}

// ---------------------------------------------------------------------------
// K13: a plain fp32 Linear on the same FFMA GEMM (the classifier head, HWGATE.py:359)
// ---------------------------------------------------------------------------
int linear_f32_fwd(const float* x, const float* w, const float* bias, float* y, int n, int d_in, int d_out,
                   cudaStream_t s) {
  if (x3_enabled() && x3_supported(n, d_in, d_out)) return linear_x3_fwd(x, w, bias, y, n, d_in, d_out, s);
  dim3 g((d_out + 63) / 64, (unsigned)((n + 63) / 64), 1);
  if (bias)
    gemm_f32_kernel<false, true, false, true><<<g, 256, 0, s>>>(x, w, bias, y, n, d_out, d_in, d_in);
  else
    gemm_f32_kernel<false, true, false, false><<<g, 256, 0, s>>>(x, w, nullptr, y, n, d_out, d_in, d_in);
  count_launch();
  return check_last();
}

// dx[n, d_in] = dy . W ; dw[d_out, d_in] = dy^T . x ; db[d_out] = column sums of dy  (any of the three may be NULL)
int linear_f32_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int n, int d_in,
                   int d_out, cudaStream_t s) {
  int st;
  if (x3_enabled() && x3_supported(n, d_in, d_out) && (dx || dw)) {
    if ((st = linear_x3_bwd(dy, x, w, dx, dw, n, d_in, d_out, s))) return st;
    dx = nullptr;
    dw = nullptr;      // the column sums below stay fp32
  }
  if (dx) {
    dim3 g((d_in + 63) / 64, (unsigned)((n + 63) / 64), 1);
    gemm_f32_kernel<false, false, false, false><<<g, 256, 0, s>>>(dy, w, nullptr, dx, n, d_in, d_out, d_out);
    count_launch();
    if ((st = check_last())) return st;
  }
  if (dw) {
    // the contraction runs over the n rows: one split (deterministic plain stores) up to 4096 rows
    dim3 g((d_in + 63) / 64, (d_out + 63) / 64, 1);
    if (n <= 4096 || deterministic()) {
      gemm_f32_kernel<true, false, false, false><<<g, 256, 0, s>>>(dy, x, nullptr, dw, d_out, d_in, n, n);
    } else {
      cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d_out * d_in, s);
      const int kps = 4096;
      g.z = (unsigned)((n + kps - 1) / kps);
      gemm_f32_kernel<true, false, true, false><<<g, 256, 0, s>>>(dy, x, nullptr, dw, d_out, d_in, n, kps);
    }
    count_launch();
    if ((st = check_last())) return st;
  }
  if (db && (st = colsum_f32(dy, db, n, d_out, s))) return st;
  return 0;
}

// out[c] = sum over the n rows of x[:, c]
int colsum_f32(const float* x, float* out, long long n, int cols, cudaStream_t s) {
  int st;
  cudaMemsetAsync(out, 0, sizeof(float) * cols, s);
  if (n <= 4096) {          // the classifier head: one block per 128 columns, plain order
    dim3 g((cols + 127) / 128, 1);
    colsum_f32_kernel<<<g, 128, 0, s>>>(x, out, n, cols, n);
    count_launch();
  } else {                  // token-sized inputs (the blocks' Linears in the fp32 modes): 512-row partial sums
    const long long rpb = 512;
    dim3 g((cols + 127) / 128, (unsigned)((n + rpb - 1) / rpb));
    float* part;
    if ((st = det_scratch(&part, 1, (int)g.y, cols, s))) return st;
    colsum_f32_kernel<<<g, 128, 0, s>>>(x, out, n, cols, rpb, part);
    count_launch();
    det_finish(part, (int)g.y, cols, out, nullptr, nullptr, s);
  }
  return check_last();
}

}  // namespace hwgat
