// K2 on tcgen05: fused QKV projection + windowed graph attention, bf16.
//
// Persistent CTA (one per SM), one tile (128 tokens = 4 windows) at a time, all
// heads of the tile inside the CTA:
//
//   warp 0      TMA producer: the X tile (resident in smem for all heads; the
//               roll / window partition is the box coordinate of 16-token TMA
//               boxes) and a ring of per-head weight chunks Wh[192 x 64].
//   warp 1      tcgen05.mma issuer: QKV_h[128 x 192] = X[128 x d] . Wh^T, fp32
//               accumulator in TMEM, two accumulators so head h+1 is multiplied
//               while head h is in the attention phase.
//   warps 2-9   attention: warp = (window, 16 query rows).  tcgen05.ld.16x256b
//               returns the accumulator directly in the HMMA C-fragment layout,
//               so q (A fragments), k (B fragments of q.k^T) and, after one
//               movmatrix.trans per 8x8 block, v (B fragments of P.v) are built
//               in registers: Q, K, V, S and P never touch shared memory.
//               S = q k^T, threshold drop, packed graph/shift mask, -10000 fill,
//               softmax, O = P v  (HWGATE.py:89-114), O written in token order.
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kEpiWarps);
constexpr int kXChunk = kTileTok * 128;   // 16 KB: 128 rows x 64 bf16
constexpr int kWStage = 192 * 128;        // 24 KB: q|k|v rows of one head x 64 bf16
constexpr int kMaxChunks = 8;             // d <= 512
constexpr int kMaxWStages = 8;
constexpr int kAccStride = 256;           // TMEM columns between the two accumulators

struct TcBars {
  uint64_t x_full[kMaxChunks], x_empty[kMaxChunks];
  uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

HW_DEV void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
HW_DEV uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}

// in-register masked softmax of attn_bf16.cu (same code path for both kernels)
HW_DEV void masked_softmax_tc(float (&s)[4][4], uint32_t mword0, uint32_t mword1, float threshold, int t) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint32_t mw = r == 0 ? mword0 : mword1;
    float v[8];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { v[2 * nt] = s[nt][2 * r]; v[2 * nt + 1] = s[nt][2 * r + 1]; }
    uint32_t lv = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) lv |= ((mw >> (8 * (i >> 1) + 2 * t + (i & 1))) & 1u) << i;
    if (threshold >= 0.f) {  // HWGATE.py:94-100
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
      if (v[i] == 0.f) lv &= ~(1u << i);  // HWGATE.py:110
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
  }
}

struct FwdTcArgs {
  const float* bias;
  const uint32_t* bits;
  bf16* out;
  float threshold;
  int heads, tiles, w_stages;
  TileGeom geo;
};

__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const FwdTcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.geo.d, nk = d / 64, heads = p.heads, S = p.w_stages;
  unsigned char* sX = smem;
  unsigned char* sW = smem + nk * kXChunk;
  TcBars* bars = reinterpret_cast<TcBars*>(sW + S * kWStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxChunks; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], 1); }
    for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], kEpiWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t wph = 0;
      int tcount = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
        const int tps = p.geo.f * p.geo.kgroups;
        const int b = tile / tps, rr = tile - b * tps, fi = rr / p.geo.kgroups, kg = rr - fi * p.geo.kgroups;
        for (int h = 0; h < heads; ++h) {
          for (int c = 0; c < nk; ++c) {
            if (h == 0) {
              mbar_wait(&bars->x_empty[c], (tcount & 1) ^ 1);
              mbar_expect_tx(&bars->x_full[c], kXChunk);
              unsigned char* dst = sX + c * kXChunk;
              if (p.geo.layout == HWGAT_LAYOUT_WINDOWS) {
                tma_load_2d(dst, &tmX, &bars->x_full[c], c * 64, tile * kTileTok);
              } else {
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                  for (int tp = 0; tp < 2; ++tp) {
                    int fr = 2 * fi + tp + p.geo.shift;
                    fr = fr >= p.geo.F ? fr - p.geo.F : fr;
                    tma_load_4d(dst + (w * 32 + tp * 16) * 128, &tmX, &bars->x_full[c], c * 64, kg * 64 + w * 16, fr, b);
                  }
              }
            }
            mbar_wait(&bars->w_empty[s], wph ^ 1);
            mbar_expect_tx(&bars->w_full[s], kWStage);
            unsigned char* dw = sW + s * kWStage;
#pragma unroll
            for (int q = 0; q < 3; ++q) tma_load_2d(dw + q * 8192, &tmW, &bars->w_full[s], c * 64, q * d + h * kHd);
            if (++s == S) { s = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 192);
      int s = 0, it = 0, tcount = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
        for (int h = 0; h < heads; ++h, ++it) {
          const int buf = it & 1;
          mbar_wait(&bars->acc_empty[buf], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < nk; ++c) {
            mbar_wait(&bars->w_full[s], wph);
            if (h == 0) mbar_wait(&bars->x_full[c], tcount & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(sX + c * kXChunk), sb = smem_u32(sW + s * kWStage);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem + buf * kAccStride, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                        (c | ks) != 0);
            umma_commit(&bars->w_empty[s]);
            if (h == heads - 1) umma_commit(&bars->x_empty[c]);
            if (++s == S) { s = 0; wph ^= 1; }
          }
          umma_commit(&bars->acc_full[buf]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ attention warps
    const int win = warp & 3;             // TMEM lane quarter of this warp == window of the tile
    const int qh = (warp - 2) >> 2;       // which 16 query rows of the window
    const int g = lane >> 2, t = lane & 3;
    const uint32_t lane_q = (uint32_t)(32 * win + 16 * qh) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int row0 = 32 * win + 16 * qh;
      const size_t orow0 = (size_t)p.geo.token_row(tile, row0 + g) * d;
      const size_t orow1 = (size_t)p.geo.token_row(tile, row0 + g + 8) * d;
      const uint32_t* mw = p.bits + p.geo.mask_base(tile) + row0;
      const uint32_t mw0 = mw[g], mw1 = mw[g + 8];
      for (int h = 0; h < heads; ++h, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->acc_full[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tb = tmem + buf * kAccStride;
        const float* bq = p.bias + h * kHd + 2 * t;
        const float* bk = bq + d;
        const float* bv = bk + d;
        uint32_t r[32];
        // ---- q: 16 rows x 64 -> A fragments (x head_dim^-0.5)
        uint32_t qa[4][4];
        tmem_ld_16x256b_x8(tb + lane_q, r);
        tmem_ld_wait();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const int nt = 2 * ks + x;
            const float2 bb = *reinterpret_cast<const float2*>(bq + 8 * nt);
            qa[ks][2 * x] = pack_bf16((__uint_as_float(r[4 * nt]) + bb.x) * 0.125f,
                                      (__uint_as_float(r[4 * nt + 1]) + bb.y) * 0.125f);
            qa[ks][2 * x + 1] = pack_bf16((__uint_as_float(r[4 * nt + 2]) + bb.x) * 0.125f,
                                          (__uint_as_float(r[4 * nt + 3]) + bb.y) * 0.125f);
          }
        // ---- S = q . k^T: k C-fragments are B fragments as they come
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {  // keys 16*mt .. 16*mt+15
          tmem_ld_16x256b_x8(tb + ((uint32_t)(32 * win + 16 * mt) << 16) + 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const float2 b0 = *reinterpret_cast<const float2*>(bk + 8 * (2 * ks));
            const float2 b1 = *reinterpret_cast<const float2*>(bk + 8 * (2 * ks + 1));
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {  // keys 8*hh + g of this 16-group -> n tile 2*mt + hh
              const uint32_t f0 = pack_bf16(__uint_as_float(r[4 * (2 * ks) + 2 * hh]) + b0.x,
                                            __uint_as_float(r[4 * (2 * ks) + 2 * hh + 1]) + b0.y);
              const uint32_t f1 = pack_bf16(__uint_as_float(r[4 * (2 * ks + 1) + 2 * hh]) + b1.x,
                                            __uint_as_float(r[4 * (2 * ks + 1) + 2 * hh + 1]) + b1.y);
              mma16816(s[2 * mt + hh], qa[ks], f0, f1);
            }
          }
        }
        masked_softmax_tc(s, mw0, mw1, p.threshold, t);
        uint32_t pa[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          pa[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
          pa[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
          pa[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        }
        // ---- O = P . v: v C-fragments -> bf16 8x8 blocks -> movmatrix.trans -> B fragments
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {  // keys 16*mt .. : k step mt of P.v
          tmem_ld_16x256b_x8(tb + ((uint32_t)(32 * win + 16 * mt) << 16) + 128, r);
          tmem_ld_wait();
          if (mt == 1) {  // last TMEM read of this head: hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
          }
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const float2 bb = *reinterpret_cast<const float2*>(bv + 8 * nt);
            const uint32_t v0 = movmatrix_trans(pack_bf16(__uint_as_float(r[4 * nt]) + bb.x,
                                                          __uint_as_float(r[4 * nt + 1]) + bb.y));
            const uint32_t v1 = movmatrix_trans(pack_bf16(__uint_as_float(r[4 * nt + 2]) + bb.x,
                                                          __uint_as_float(r[4 * nt + 3]) + bb.y));
            mma16816(o[nt], pa[mt], v0, v1);
          }
        }
        // ---- store: (row g, row g+8) x 64 columns of this head
        bf16* o0 = p.out + orow0 + h * kHd + 2 * t;
        bf16* o1 = p.out + orow1 + h * kHd + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(o0 + 8 * nt) = pack_bf16(o[nt][0], o[nt][1]);
          *reinterpret_cast<uint32_t*>(o1 + 8 * nt) = pack_bf16(o[nt][2], o[nt][3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int attn_fwd_tc(const AttnArgs& a, cudaStream_t s) {
  const int d = a.d, nk = d / 64;
  const int stages = d == 512 ? 4 : (d == 256 ? 6 : 6);
  const int smem_bytes = nk * kXChunk + stages * kWStage + (int)sizeof(TcBars) + 1024;
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    attr_smem = smem_bytes;
  }
  CUtensorMap tmX, tmW;
  int st;
  if (a.layout == HWGAT_LAYOUT_WINDOWS) {
    if ((st = make_tmap_2d(&tmX, a.xn, (uint64_t)a.tokens(), (uint64_t)d, kTileTok))) return st;
  } else {
    if ((st = make_tmap_4d(&tmX, a.xn, (uint64_t)d, (uint64_t)a.K, (uint64_t)a.F, (uint64_t)a.B, 16))) return st;
  }
  if ((st = make_tmap_2d(&tmW, a.w_qkv, (uint64_t)3 * d, (uint64_t)d, 64))) return st;
  FwdTcArgs p;
  p.bias = a.b_qkv; p.bits = a.bits; p.out = (bf16*)a.out; p.threshold = a.threshold;
  p.heads = a.heads; p.tiles = a.tiles(); p.w_stages = stages;
  p.geo = make_geom(a.F, a.K, d, a.shift, a.layout);
  const int grid = p.tiles < 148 ? p.tiles : 148;
  attn_fwd_tc_kernel<<<grid, kTcThreads, smem_bytes, s>>>(tmX, tmW, p);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
