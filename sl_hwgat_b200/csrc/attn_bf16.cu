// K2 / K3 in bf16: fused QKV projection + windowed graph attention, forward and
// backward-with-recompute.  One CTA = one tile (128 tokens = 4 windows of one
// temporal group, common.cuh) x one head.
//
//   phase 1  QKV[128 x 192] = Xtile[128 x d] . Wh^T + b     (HWGATE.py:86-89)
//            tensor cores (HMMA m16n8k16, fp32 accumulate), X and Wh streamed
//            in 64-wide k chunks through a cp.async ring; the roll and the
//            window partition are the row gather of the X loads.
//   phase 2  per window (32 tokens): S = q k^T, threshold drop, graph/shift
//            bitmask, -10000 fill, softmax, P v          (HWGATE.py:91-114)
//            Q, K, V never leave shared memory; S and P never leave registers.
//   backward recomputes phase 1 and S/P, then
//            dP = dO v^T, dS = live ? P*(dP - rowsum(P*dP)) : 0,
//            dq = dS k * scale, dk = dS^T q, dv = P^T dO
//            and writes dQKV (bf16, token order) for the two weight-side GEMMs
//            in gemm_bf16.cu (d_xn = dQKV.W, d_w = dQKV^T.xn, d_b = colsum).
#include <cstdlib>

#include "common.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

constexpr int kKC = 64;                       // k chunk: 64 bf16 = one 128-byte smem row
constexpr int kStageX = kTileTok * 128;       // 16 KB
constexpr int kStageW = 192 * 128;            // 24 KB
constexpr int kStageBytes = kStageX + kStageW;
constexpr int kStages = 2;
constexpr int kRingBytes = kStages * kStageBytes;  // 80 KB
// overlay of the ring once phase 1 is done
constexpr int kOffQ = 0, kOffK = 16384, kOffV = 32768, kOffDQ = 49152;
// backward-only regions behind the ring
constexpr int kOffDO = kRingBytes;            // 16 KB  dO[128][64]
constexpr int kOffP = kOffDO + 16384;         // 8 KB   P [4][32][32] bf16
constexpr int kOffDS = kOffP + 8192;          // 8 KB   dS[4][32][32] bf16
constexpr int kOffRowsFwd = kRingBytes;       // int[128] global token row of each tile row
constexpr int kOffRowsBwd = kOffDS + 8192;
constexpr int kSmemFwd = kOffRowsFwd + 512;
constexpr int kSmemBwd = kOffRowsBwd + 512;

// byte offset of 16-byte chunk `c` of row `r` in a tile with 128-byte rows
HW_DEV int off128(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// same for a tile with 64-byte rows (P / dS, 32 bf16 per row)
HW_DEV int off64(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

// ---------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------
HW_DEV void load_stage(unsigned char* st, const bf16* __restrict__ xn, const bf16* __restrict__ w, const int* rows,
                       int d, int h, int kc) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // X: 128 rows x 8 chunks
    int idx = tid + i * 256, r = idx >> 3, c = idx & 7;
    cp_async16(st + off128(r, c), xn + (size_t)rows[r] * d + kc * kKC + c * 8);
  }
  unsigned char* sw = st + kStageX;
#pragma unroll
  for (int i = 0; i < 6; ++i) {  // Wh: 192 rows (q|k|v of head h) x 8 chunks
    int idx = tid + i * 256, n = idx >> 3, c = idx & 7;
    int grow = (n >> 6) * d + h * kHd + (n & 63);
    cp_async16(sw + off128(n, c), w + (size_t)grow * d + kc * kKC + c * 8);
  }
}

// Leaves q*scale, k, v (bf16, +bias) of the tile in smem at kOffQ/kOffK/kOffV.
// All 256 threads; ends with a __syncthreads().
HW_DEV void qkv_phase(unsigned char* smem, const bf16* __restrict__ xn, const bf16* __restrict__ w,
                      const float* __restrict__ bias, const int* rows, int d, int h) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2;  // warp tile: rows 32*wm.., cols 96*wn..
  const int nk = d / kKC;
  float acc[2][12][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  load_stage(smem, xn, w, rows, d, h, 0);
  cp_async_commit();
  for (int kc = 0; kc < nk; ++kc) {
    if (kc + 1 < nk) {
      load_stage(smem + ((kc + 1) & 1) * kStageBytes, xn, w, rows, d, h, kc + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const unsigned char* sx = smem + (kc & 1) * kStageBytes;
    const unsigned char* sw = sx + kStageX;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(a[mt], sx + off128(32 * wm + 16 * mt + (lane & 15), 2 * ks + (lane >> 4)));
#pragma unroll
      for (int np = 0; np < 6; ++np) {
        uint32_t b[4];
        ldsm_x4(b, sw + off128(96 * wn + 16 * np + ((lane >> 4) << 3) + (lane & 7), 2 * ks + ((lane >> 3) & 1)));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * np], a[mt], b[0], b[1]);
          mma16816(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
        }
      }
    }
    __syncthreads();
  }
  // epilogue: + bias, q *= head_dim^-0.5, round to bf16, store to the overlay
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 12; ++nt) {
    const int n = 96 * wn + 8 * nt + 2 * t;
    const int s = n >> 6, e = n & 63;
    const float2 bb = *reinterpret_cast<const float2*>(bias + s * d + h * kHd + e);
    const float mul = s == 0 ? 0.125f : 1.f;
    unsigned char* dst = smem + s * 16384;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = 32 * wm + 16 * mt + g;
      *reinterpret_cast<uint32_t*>(dst + off128(r0, e >> 3) + 4 * t) =
          pack_bf16((acc[mt][nt][0] + bb.x) * mul, (acc[mt][nt][1] + bb.y) * mul);
      *reinterpret_cast<uint32_t*>(dst + off128(r0 + 8, e >> 3) + 4 * t) =
          pack_bf16((acc[mt][nt][2] + bb.x) * mul, (acc[mt][nt][3] + bb.y) * mul);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// phase 2 helpers (one warp = 16 query rows of one window)
// ---------------------------------------------------------------------------
// S[16 x 32] = Q[16 x 64] . K[32 x 64]^T for the warp's rows
HW_DEV void logits_16x32(float (&s)[4][4], const uint32_t (&qa)[4][4], const unsigned char* sK, int krow0, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ldsm_x4(b, sK + off128(krow0 + 16 * np + ((lane >> 4) << 3) + (lane & 7), 2 * ks + ((lane >> 3) & 1)));
      mma16816(s[2 * np], qa[ks], b[0], b[1]);
      mma16816(s[2 * np + 1], qa[ks], b[2], b[3]);
    }
}

// In-register masked softmax of the two rows (g, g+8) a thread shares with its
// quad.  s: logits in, probabilities out.  Returns, per row, the 8 `live` flags
// of the thread's columns (bit nt*2+x <-> column 8*nt + 2*t + x).
HW_DEV void masked_softmax(float (&s)[4][4], uint32_t mword0, uint32_t mword1, float threshold, int t,
                           uint32_t (&live)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint32_t mw = r == 0 ? mword0 : mword1;
    float v[8];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { v[2 * nt] = s[nt][2 * r]; v[2 * nt + 1] = s[nt][2 * r + 1]; }
    uint32_t lv = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = 8 * (i >> 1) + 2 * t + (i & 1);
      lv |= ((mw >> col) & 1u) << i;
    }
    if (threshold >= 0.f) {  // HWGATE.py:94-100: drop logits whose unmasked softmax exceeds the threshold
      float m0 = v[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m0 = fmaxf(m0, v[i]);
      m0 = quad_max(m0);
      float e[8], sum0 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { e[i] = __expf(v[i] - m0); sum0 += e[i]; }
      sum0 = quad_sum(sum0);
      const float inv0 = 1.f / sum0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (e[i] * inv0 > threshold) lv &= ~(1u << i);
    }
    float m1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (v[i] == 0.f) lv &= ~(1u << i);  // HWGATE.py:110 fills every exact zero
      v[i] = ((lv >> i) & 1u) ? v[i] : kNegFill;
      m1 = fmaxf(m1, v[i]);
    }
    m1 = quad_max(m1);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = __expf(v[i] - m1); sum += v[i]; }
    sum = quad_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { s[nt][2 * r] = v[2 * nt] * inv; s[nt][2 * r + 1] = v[2 * nt + 1] * inv; }
    live[r] = lv;
  }
}

// accumulator tiles of a [16 x 32] matrix -> A fragments of its two k16 steps
HW_DEV void acc_to_afrag(uint32_t (&a)[2][4], const float (&s)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    a[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    a[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    a[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    a[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

// C[16 x 64] += A[16 x 32] . Bm[32 x 64], Bm stored row-major (rows = k) in a 128-byte-row tile
HW_DEV void mma_16x64_k32(float (&c)[8][4], const uint32_t (&a)[2][4], const unsigned char* sB, int brow0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int ep = 0; ep < 4; ++ep) {
      uint32_t b[4];
      ldsm_x4_t(b, sB + off128(brow0 + 16 * kk + (lane & 7) + (((lane >> 3) & 1) << 3), 2 * ep + (lane >> 4)));
      mma16816(c[2 * ep], a[kk], b[0], b[1]);
      mma16816(c[2 * ep + 1], a[kk], b[2], b[3]);
    }
}

// C fragments [16 x 64] -> bf16 -> the warp's 16 rows of a 128-byte-row staging
// tile -> global, 128 contiguous bytes per token row (one head slice).
HW_DEV void store_16x64(const float (&c)[8][4], float mul, unsigned char* stage, int row0, bf16* __restrict__ gdst,
                        const int* rows, size_t ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + off128(row0 + g, nt) + 4 * t) = pack_bf16(c[nt][0] * mul, c[nt][1] * mul);
    *reinterpret_cast<uint32_t*>(stage + off128(row0 + g + 8, nt) + 4 * t) = pack_bf16(c[nt][2] * mul, c[nt][3] * mul);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i, r = row0 + (idx >> 3), ch = idx & 7;
    const int4 v = *reinterpret_cast<const int4*>(stage + off128(r, ch));
    *reinterpret_cast<int4*>(gdst + (size_t)rows[r] * ld + ch * 8) = v;
  }
}

HW_DEV void fill_rows(int* rows, const TileGeom& g, int tile) {
  if (threadIdx.x < kTileTok) rows[threadIdx.x] = (int)g.token_row(tile, threadIdx.x);
}

// ---------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) attn_fwd_bf16_kernel(const bf16* __restrict__ xn, const bf16* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               const uint32_t* __restrict__ bits, float threshold,
                                                               bf16* __restrict__ out, TileGeom geo, int heads) {
  extern __shared__ __align__(128) unsigned char smem[];
  int* rows = reinterpret_cast<int*>(smem + kOffRowsFwd);
  // heads of one tile are adjacent CTAs: they run together and share the X tile in L2
  const int tile = blockIdx.x / heads, h = blockIdx.x - tile * heads, d = geo.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_rows(rows, geo, tile);
  __syncthreads();
  qkv_phase(smem, xn, w, bias, rows, d, h);

  const int win = warp >> 1, row0 = 32 * win + 16 * (warp & 1);
  const int g = lane >> 2, t = lane & 3;
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], smem + kOffQ + off128(row0 + (lane & 15), 2 * ks + (lane >> 4)));
  float s[4][4];
  logits_16x32(s, qa, smem + kOffK, 32 * win, lane);
  const uint32_t* mw = bits + geo.mask_base(tile) + row0;
  uint32_t live[2];
  masked_softmax(s, mw[g], mw[g + 8], threshold, t, live);
  uint32_t pa[2][4];
  acc_to_afrag(pa, s);
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  mma_16x64_k32(o, pa, smem + kOffV, 32 * win, lane);
  // the warp's own Q rows are dead now (only this warp read them): reuse as staging
  store_16x64(o, 1.f, smem + kOffQ, row0, out + h * kHd, rows, (size_t)d, lane);
}

// ---------------------------------------------------------------------------
// K3a
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) attn_bwd_bf16_kernel(const bf16* __restrict__ xn, const bf16* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               const uint32_t* __restrict__ bits, float threshold,
                                                               const bf16* __restrict__ d_out,
                                                               bf16* __restrict__ dqkv, TileGeom geo, int heads) {
  extern __shared__ __align__(128) unsigned char smem[];
  int* rows = reinterpret_cast<int*>(smem + kOffRowsBwd);
  const int tile = blockIdx.x / heads, h = blockIdx.x - tile * heads, d = geo.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_rows(rows, geo, tile);
  __syncthreads();
  // dO head slice [128 x 64] -> smem; rides in the first cp.async group of phase 1
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int idx = threadIdx.x + i * 256, r = idx >> 3, c = idx & 7;
    cp_async16(smem + kOffDO + off128(r, c), d_out + (size_t)rows[r] * d + h * kHd + c * 8);
  }
  qkv_phase(smem, xn, w, bias, rows, d, h);

  const int win = warp >> 1, half = warp & 1, row0 = 32 * win + 16 * half;
  const int g = lane >> 2, t = lane & 3;
  const size_t ld3 = (size_t)3 * d;
  unsigned char* sP = smem + kOffP + win * 2048;
  unsigned char* sDS = smem + kOffDS + win * 2048;
  // ---- pass 1: the warp owns 16 query rows: P, dP, dS, dq
  {
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], smem + kOffQ + off128(row0 + (lane & 15), 2 * ks + (lane >> 4)));
    float p[4][4];
    logits_16x32(p, qa, smem + kOffK, 32 * win, lane);
    const uint32_t* mw = bits + geo.mask_base(tile) + row0;
    uint32_t live[2];
    masked_softmax(p, mw[g], mw[g + 8], threshold, t, live);
    // dP = dO . V^T  (same operand pattern as q . k^T)
    uint32_t ga[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(ga[ks], smem + kOffDO + off128(row0 + (lane & 15), 2 * ks + (lane >> 4)));
    float dp[4][4];
    logits_16x32(dp, ga, smem + kOffV, 32 * win, lane);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float delta = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) delta += p[nt][2 * r] * dp[nt][2 * r] + p[nt][2 * r + 1] * dp[nt][2 * r + 1];
      delta = quad_sum(delta);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const bool on = (live[r] >> (2 * nt + x)) & 1u;
          dp[nt][2 * r + x] = on ? p[nt][2 * r + x] * (dp[nt][2 * r + x] - delta) : 0.f;  // dS
        }
    }
    // P and dS (bf16) to smem for the transposed products of pass 2
    const int i0 = 16 * half + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(sP + off64(i0, nt) + 4 * t) = pack_bf16(p[nt][0], p[nt][1]);
      *reinterpret_cast<uint32_t*>(sP + off64(i0 + 8, nt) + 4 * t) = pack_bf16(p[nt][2], p[nt][3]);
      *reinterpret_cast<uint32_t*>(sDS + off64(i0, nt) + 4 * t) = pack_bf16(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(sDS + off64(i0 + 8, nt) + 4 * t) = pack_bf16(dp[nt][2], dp[nt][3]);
    }
    uint32_t dsa[2][4];
    acc_to_afrag(dsa, dp);
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
    mma_16x64_k32(dq, dsa, smem + kOffK, 32 * win, lane);  // dS . k
    store_16x64(dq, 0.125f, smem + kOffDQ, row0, dqkv + h * kHd, rows, ld3, lane);
  }
  __syncthreads();
  // ---- pass 2: the warp owns 16 key rows: dv = P^T dO, dk = dS^T q
  {
    uint32_t pt[2][4], dst[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int i = 16 * kk + ((lane >> 4) << 3) + (lane & 7), c = 2 * half + ((lane >> 3) & 1);
      ldsm_x4_t(pt[kk], sP + off64(i, c));
      ldsm_x4_t(dst[kk], sDS + off64(i, c));
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    mma_16x64_k32(acc, pt, smem + kOffDO, 32 * win, lane);  // dv
    // K and V are dead after pass 1: their rows are the staging for dk / dv
    store_16x64(acc, 1.f, smem + kOffV, row0, dqkv + 2 * d + h * kHd, rows, ld3, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    mma_16x64_k32(acc, dst, smem + kOffQ, 32 * win, lane);  // dk (q already carries the scale)
    store_16x64(acc, 1.f, smem + kOffK, row0, dqkv + d + h * kHd, rows, ld3, lane);
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
int gemm_bf16_nn(const bf16* A, const bf16* Bm, bf16* C, int M, int N, int K, cudaStream_t s);
int gemm_bf16_tn_f32(const bf16* A, const bf16* Bm, float* C, float* colsum, int M, int N, long long Kdim,
                     cudaStream_t s);

int attn_fwd_bf16(const AttnArgs& a, cudaStream_t s) {
  static const bool hmma_fwd = getenv("HWGAT_HMMA_FWD") != nullptr;  // A/B switch while the tcgen05 path is new
  if (!hmma_fwd) return attn_fwd_tc(a, s);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_fwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
    attr_done = true;
  }
  TileGeom g = make_geom(a.F, a.K, a.d, a.shift, a.layout);
  attn_fwd_bf16_kernel<<<dim3(a.tiles() * a.heads), 256, kSmemFwd, s>>>(
      (const bf16*)a.xn, (const bf16*)a.w_qkv, a.b_qkv, a.bits, a.threshold, (bf16*)a.out, g, a.heads);
  count_launch();
  return (int)cudaGetLastError();
}

int attn_bwd_bf16(const AttnArgs& a, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_bwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd);
    attr_done = true;
  }
  TileGeom g = make_geom(a.F, a.K, a.d, a.shift, a.layout);
  bf16* dqkv = (bf16*)a.workspace;
  static const bool hmma_bwd = getenv("HWGAT_HMMA_BWD") != nullptr;  // A/B switch while the tcgen05 path is new
  int st;
  if (hmma_bwd) {
    attn_bwd_bf16_kernel<<<dim3(a.tiles() * a.heads), 256, kSmemBwd, s>>>(
        (const bf16*)a.xn, (const bf16*)a.w_qkv, a.b_qkv, a.bits, a.threshold, (const bf16*)a.d_out, dqkv, g, a.heads);
    count_launch();
    st = (int)cudaGetLastError();
  } else {
    st = attn_bwd_tc(a, dqkv, s);
  }
  if (st) return st;
  const long long n = a.tokens();
  const int d = a.d, d3 = 3 * d;
  // d_xn[n, d] = dQKV[n, 3d] . Wqkv[3d, d]  (tcgen05 GEMM against Wqkv^T, K-major on both sides)
  static const bool hmma_gemm = getenv("HWGAT_HMMA_GEMM") != nullptr;  // A/B switch while the tcgen05 path is new
  if (hmma_gemm) {
    if ((st = gemm_bf16_nn(dqkv, (const bf16*)a.w_qkv, (bf16*)a.d_xn, (int)n, d, d3, s))) return st;
  } else {
    bf16* wt = dqkv + n * d3;
    if ((st = transpose_bf16((const bf16*)a.w_qkv, wt, d3, d, s))) return st;
    if ((st = gemm_tc_nt(dqkv, wt, (bf16*)a.d_xn, (int)n, d, d3, s))) return st;
  }
  // d_w[3d, d] = dQKV^T . xn ; d_b = column sums of dQKV (rides in the same kernel)
  if (hmma_gemm) return gemm_bf16_tn_f32(dqkv, (const bf16*)a.xn, a.d_w, a.d_b, d3, d, n, s);
  return gemm_tc_tn(dqkv, (const bf16*)a.xn, a.d_w, a.d_b, d3, d, n, s);
}

}  // namespace hwgat
