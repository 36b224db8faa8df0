// K2b / K3b: windowed graph attention with EVERY product on tcgen05, for any window of N = 32, 64 or 128 tokens
// (window_size W = 16, 32, 64 keypoints x TP = 2 frames; HWGATE.py:30-36, 84-114, model_params.py:254).
//
// The QKV projection runs as one tcgen05 GEMM with the bias in its epilogue (ffn_tc.cu), q pre-scaled by 64^-1/2; this
// file is the attention core on the projected rows.  One work item = (tile of 128 tokens, head).  A tile is one
// temporal group (frame pair) x 64 keypoints = 128/N whole windows, so the per-window products are block-diagonal
// M = 128 UMMA tiles (at N = 128 a window IS the tile; at N = 32 three quarters of the S / P columns are off the
// diagonal - 4x of 4-14 % of the FLOPs - and the softmax threads never touch them):
//
//   warp 0     TMA producer: Q, K, V (and dO in the backward) tiles [128 x 64] straight from the projected rows in
//              global memory as 16-token boxes of a 4-D tensor map (roll + window partition = box coordinates),
//              landing in the SWIZZLE_128B K-major layout the MMAs read; 2-3 stage ring.
//   warp 1     tcgen05.mma issuer:  S = Q K^T  (M128 N128 K64),  O = P V  (M128 N64 K128, P from shared memory as
//              bf16, V as the MN-major B operand);  backward:  dP = dO V^T,  dV = P^T dO,  dQ = dS K,  dK = dS^T Q
//              with the transposed operands taken through MN-major descriptors of the same shared-memory tiles.
//              S of item i+1 is issued before the second-stage MMAs of item i, so the tensor pipe works under the
//              softmax of item i.
//   warps 4-11 softmax warps, ONE THREAD PER QUERY ROW (tcgen05.ld 32x32b): the row's N live logits in registers,
//              threshold drop, packed graph / shift mask, -10000 fill and softmax without any shuffle
//              (HWGATE.py:94-111); P (and dS) written as bf16 into the K-major tile; outputs read back from TMEM and
//              stored as full 32-byte sectors of the thread's own row.  Two warp sets alternate items.
//   warps 2-3  idle: they donate registers (setmaxnreg) so a softmax thread can hold a 128-column row.
//
// TMEM: two 256-column buffers; forward S [0,128) + O [128,192); backward S [0,128) + dP [128,256), then
// dQ [0,64) + dK [64,128) + dV [128,192) over the consumed S / dP.
#include <cstdlib>

#include "ew.cuh"
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

namespace tc2 {

constexpr int kThreads = 384;
constexpr int kFirstSoftWarp = 4;
constexpr int kRegsDonor = 40, kRegsSoft = 232;
constexpr int kTile = 128 * 128;            // bytes of a [128 x 64] bf16 tile
constexpr int kPBytes = 2 * kTile;          // [128 x 128] bf16 as two 64-column chunks
constexpr int kMaxDynSmem2 = 232448;
constexpr float kLog2e = 1.4426950408889634f;

HW_DEV void reg_dealloc_donor() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegsDonor)); }
HW_DEV void reg_alloc_soft() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsSoft)); }

// Geometry.  A sample's windows are ordered (temporal group fi, keypoint window kwin) = the reference's window order
// (HWGATE.py:34-35: index (b f + fi) nW + w); a tile is 128 / N consecutive windows of one sample.  At the reference's
// K = 64 that is one temporal group (W = 16: 4 windows, W = 32: 2, W = 64: 1); at K = 128 half of one; at K = 32 (HGATE's
// 29 keypoints padded to 32, one window of N = 64 per temporal group) two temporal groups.
struct Geo2 {
  int F, K, shift, layout, f, nWt, tps;   // nWt = K / W windows per temporal group, tps = tiles per sample
  // global token row of tile row `row` (HWGATE.py:197-201 without the copies)
  template <int N>
  HW_DEV long long token_row(int tile, int row) const {
    if (layout == HWGAT_LAYOUT_WINDOWS) return (long long)tile * 128 + row;
    constexpr int W = N / 2, wpt = 128 / N;
    const int b = tile / tps, w = row / N, widx = (tile - b * tps) * wpt + w;
    const int fi = widx / nWt, kwin = widx - fi * nWt;
    const int rr = row - w * N, tp = rr / W, k = rr - tp * W;
    int fr = 2 * fi + tp + shift;
    fr = fr >= F ? fr - F : fr;
    return (long long)(b * F + fr) * K + kwin * W + k;
  }
  // first mask word of tile row `row`: bits is (f * nWt windows, N rows, N/32 words)
  template <int N>
  HW_DEV long long mask_word(int tile, int row) const {
    constexpr int wpt = 128 / N;
    const int w = row / N, widx = (tile % tps) * wpt + w;
    return ((long long)widx * N + (row - w * N)) * (N / 32);
  }
};

// one [128 x 64] tile (columns col .. col+63 of the tile's 128 token rows) -> dst (16 KB, SWIZZLE_128B K-major).
// One TMA box per WINDOW: (64 columns, W keypoints, 2 frames) lands as the window's N rows in the reference's token
// order tp * W + k (HWGATE.py:34-35) - 4 / 2 / 1 copies per tile instead of eight 16-token boxes (the producer lane
// issued 24-32 copies per item and was the slowest role).  The shifted last temporal group wraps around (frames
// F-1 and 0): it takes two one-frame boxes per window (tm1).  Keypoints past K (a padded keypoint axis) read as zeros.
template <int N>
HW_DEV void load_tile(unsigned char* dst, const CUtensorMap* tm2, const CUtensorMap* tm1, uint64_t* bar, const Geo2& g,
                      int tile, int col) {
  if (g.layout == HWGAT_LAYOUT_WINDOWS) {
    tma_load_2d(dst, tm2, bar, col, tile * 128);
    return;
  }
  constexpr int W = N / 2, wpt = 128 / N;
  const int b = tile / g.tps, w0 = (tile - b * g.tps) * wpt;
#pragma unroll
  for (int w = 0; w < wpt; ++w) {
    const int widx = w0 + w, fi = widx / g.nWt, kwin = widx - fi * g.nWt;
    const int fr0 = 2 * fi + g.shift;                 // <= F - 1
    if (fr0 + 1 < g.F) {
      tma_load_4d(dst + w * N * 128, tm2, bar, col, kwin * W, fr0, b);
    } else {
      tma_load_4d(dst + w * N * 128, tm1, bar, col, kwin * W, fr0, b);
      tma_load_4d(dst + (w * N + W) * 128, tm1, bar, col, kwin * W, 0, b);
    }
  }
}

HW_DEV void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// row `row` of a [128 x 128] bf16 K-major SWIZZLE_128B tile pair, columns col0 .. col0+N-1 (col0 % 8 == 0) <- v[0..N)
template <int N>
HW_DEV void store_row_bf16(uint32_t tile_saddr, int row, int col0, const float (&v)[N]) {
#pragma unroll
  for (int u = 0; u < N / 8; ++u) {
    const int col = col0 + 8 * u;
    const uint32_t a = tile_saddr + (uint32_t)((col >> 6) * kTile + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
    sts128(a, pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
           pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
  }
}

// N consecutive TMEM columns of this thread's lane -> v (as floats)
template <int N>
HW_DEV void tmem_row(uint32_t taddr, float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + 32 * c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[32 * c + i] = __uint_as_float(r[i]);
  }
}

// exact two-pass softmax of one row with a live set (bit i of lv): v <- softmax(live ? v : -10000); dead = no live logit
template <int N>
HW_DEV void softmax_exact_row(float (&v)[N], const uint32_t (&lv)[N / 32], bool& dead) {
  float m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v[i] = ((lv[i >> 5] >> (i & 31)) & 1u) ? v[i] : kNegFill;
    m1 = fmaxf(m1, v[i]);
  }
  dead = m1 == kNegFill;
  const float ml = m1 * kLog2e;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) { v[i] = exp2f(fmaf(v[i], kLog2e, -ml)); sum += v[i]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] *= inv;
}

// Masked softmax of one query row held by one thread (HWGATE.py:94-111).  v: the row's N logits in, probabilities
// out.  mw: packed graph / shift mask of the row.  Training (kTrain): the exponentials of the UNMASKED softmax over the
// window's N keys decide the threshold drop and are reused for the masked softmax,
//   P_i = live_i e_i / sum_j live_j e_j,  e_i = exp(s_i - max s),
// which equals softmax(live ? s : -10000) whenever some live e_i is representable; otherwise (every logit dropped or
// masked, or all live ones underflow) the logits are read again from TMEM (`taddr`) and the exact form runs.
template <int N, bool kTrain>
HW_DEV void masked_softmax_row(float (&v)[N], const uint32_t (&mw)[N / 32], float threshold, uint32_t taddr, bool& dead) {
  uint32_t lv[N / 32];
  if (kTrain) {
    float m0 = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) m0 = fmaxf(m0, v[i]);
    const float ml = m0 * kLog2e;
    float sum0 = 0.f;
#pragma unroll
    for (int w = 0; w < N / 32; ++w) lv[w] = mw[w];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (v[i] == 0.f) lv[i >> 5] &= ~(1u << (i & 31));          // exact zeros are filled too (HWGATE.py:110)
      v[i] = exp2f(fmaf(v[i], kLog2e, -ml));
      sum0 += v[i];
    }
    const float t0 = threshold * sum0;                            // softmax_i > thr  <=>  e_i > thr * sum
    float sl = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (v[i] > t0) lv[i >> 5] &= ~(1u << (i & 31));              // HWGATE.py:97-100
      v[i] = ((lv[i >> 5] >> (i & 31)) & 1u) ? v[i] : 0.f;
      sl += v[i];
    }
    // tcgen05.ld is warp-collective (.sync.aligned): if ANY row of the warp needs the exact form, the whole warp reads
    // its logits again and takes it (for the other rows it is the same softmax, rounded slightly differently)
    if (__any_sync(0xffffffffu, !(sl > 1e-30f))) {
      tmem_row<N>(taddr, v);
      softmax_exact_row<N>(v, lv, dead);
    } else {
      const float inv = 1.f / sl;
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] *= inv;
      dead = false;
    }
  } else {
#pragma unroll
    for (int w = 0; w < N / 32; ++w) lv[w] = mw[w];
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (v[i] == 0.f) lv[i >> 5] &= ~(1u << (i & 31));
    softmax_exact_row<N>(v, lv, dead);
  }
}

struct CoreArgs {
  const uint32_t* bits;
  bf16* out;            // forward: [n, d]
  bf16* dqkv;           // backward: [n, 3d]
  float threshold;
  int d, heads, tiles, stages;
  // attention dropout (self.attn_drop on P, HWGATE.py:112): 16-bit threshold (0 = off), 1/(1-p), Philox stream
  uint32_t drop_thresh;
  float drop_scale;
  unsigned long long seed, offset;
  Geo2 geo;
};

// keep flags of one query row's N probabilities (bit i = key i kept): one Philox call per 8 keys, counter =
// ((global token row * heads + head) * N/8 + granule), identical in forward and backward
template <int N>
HW_DEV void attn_drop_flags(uint32_t (&kf)[N / 32], long long token_row, int head, const CoreArgs& p) {
  const unsigned long long base = ((unsigned long long)token_row * p.heads + head) * (N / 8);
#pragma unroll
  for (int w = 0; w < N / 32; ++w) kf[w] = 0;
#pragma unroll
  for (int u = 0; u < N / 8; ++u) kf[u >> 2] |= keep8(base + u, p.offset, p.seed, p.drop_thresh) << (8 * (u & 3));
}
// row of P' = P o keep / (1-p) as bf16 (see store_row_bf16)
template <int N>
HW_DEV void store_row_bf16_dropped(uint32_t tile_saddr, int row, int col0, const float (&v)[N], const uint32_t (&kf)[N / 32],
                                   float scale) {
#pragma unroll
  for (int u = 0; u < N / 8; ++u) {
    const int col = col0 + 8 * u;
    const uint32_t a = tile_saddr + (uint32_t)((col >> 6) * kTile + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = ((kf[u >> 2] >> (8 * (u & 3) + i)) & 1u) ? v[8 * u + i] * scale : 0.f;
    sts128(a, pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
// TMEM of the forward: THREE S buffers (columns 0, 128, 256) and two O buffers (384, 448).  An S buffer is free again
// as soon as its softmax has produced P, so S of item j+1 and j+2 are computed while item j is still being finished:
// the S MMA and its latency never sit in front of a softmax warp.
struct FwdBars {
  uint64_t in_full[3], in_empty[3];
  uint64_t s_full[3], p_ready[2], o_full[2], o_empty[2];
  uint32_t tmem_slot;
};
constexpr int kOCol = 384;

template <int N, bool kTrain>
__global__ void __launch_bounds__(kThreads, 1) attn_core_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                         const __grid_constant__ CUtensorMap tmQKV1,
                                                                         const CoreArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  unsigned char* sIn = smem;                          // S stages of [Q | K | V]
  unsigned char* sP = smem + S * 3 * kTile;           // two P tiles
  FwdBars* bars = reinterpret_cast<FwdBars*>(sP + 2 * kPBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.tiles * p.heads;

  // P tiles start as zeros: a softmax thread only ever writes its own window's columns, so the off-diagonal blocks
  // stay zero for the whole kernel
  for (int i = threadIdx.x; i < 2 * kPBytes / 16; i += blockDim.x) reinterpret_cast<int4*>(sP)[i] = make_int4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars->in_full[i], 1); mbar_init(&bars->in_empty[i], 1); mbar_init(&bars->s_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->p_ready[i], 4); mbar_init(&bars->o_full[i], 1); mbar_init(&bars->o_empty[i], 4);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmQKV1);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < kFirstSoftWarp) {
    reg_dealloc_donor();
    if (warp == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int g = blockIdx.x; g < items; g += gridDim.x) {
        const int tile = g / p.heads, h = g - tile * p.heads;
        mbar_wait(&bars->in_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->in_full[s], 3 * kTile);
          unsigned char* st = sIn + s * 3 * kTile;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            load_tile<N>(st + q * kTile, &tmQKV, &tmQKV1, &bars->in_full[s], p.geo, tile, q * p.d + h * kHd);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 128);
      constexpr uint32_t idO = umma_idesc_bf16(128, 64, false, true);
      int s = 0, ps = 0, j = 0;     // ps: stage of the previous item
      uint32_t ph = 0;
      auto issue_pv = [&](int i, int stage) {
        const int b = i & 1;
        mbar_wait(&bars->p_ready[b], (i >> 1) & 1);
        mbar_wait(&bars->o_empty[b], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sp = smem_u32(sP + b * kPBytes), sv = smem_u32(sIn + stage * 3 * kTile + 2 * kTile);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem + kOCol + b * 64, umma_desc_k_sw128(sp + (ks >> 2) * kTile + (ks & 3) * 32),
                      umma_desc_mn_sw128(sv + ks * 2048, 8192, 1024), idO, ks != 0);
          umma_commit(&bars->o_full[b]);
          umma_commit(&bars->in_empty[stage]);
        }
        __syncwarp();
      };
      int sb = 0;   // S buffer of item j = j % 3; it is free: P of item j-3 was waited for before PV(j-3) was issued
      for (int g = blockIdx.x; g < items; g += gridDim.x, ++j) {
        mbar_wait(&bars->in_full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sq = smem_u32(sIn + s * 3 * kTile), sk = sq + kTile;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem + sb * 128, umma_desc_k_sw128(sq + ks * 32), umma_desc_k_sw128(sk + ks * 32), idS, ks != 0);
          umma_commit(&bars->s_full[sb]);
        }
        __syncwarp();
        if (j > 0) issue_pv(j - 1, ps);
        ps = s;
        if (++s == S) { s = 0; ph ^= 1; }
        if (++sb == 3) sb = 0;
      }
      if (j > 0) issue_pv(j - 1, ps);
    }
  } else {
    reg_alloc_soft();
    const int q = warp & 3;                              // TMEM lane quarter
    const int set = (warp - kFirstSoftWarp) >> 2;        // items with (j & 1) == set
    const int row = 32 * q + lane;
    const int win = row / N, col0 = win * N;             // the row's window = its live columns (warp-uniform)
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    int j = 0;
    for (int g = blockIdx.x; g < items; g += gridDim.x, ++j) {
      if ((j & 1) != set) continue;
      const int tile = g / p.heads, h = g - tile * p.heads;
      const uint32_t par = (j >> 1) & 1;
      const int sb = j % 3;
      const uint32_t tq = tlane + sb * 128;
      uint32_t mw[N / 32];
      {
        const uint32_t* mp = p.bits + p.geo.mask_word<N>(tile, row);
#pragma unroll
        for (int w = 0; w < N / 32; ++w) mw[w] = mp[w];
      }
      const long long trow = p.geo.token_row<N>(tile, row);
      bf16* orow = p.out + (size_t)trow * p.d + h * kHd;
      mbar_wait(&bars->s_full[sb], (j / 3) & 1);
      tc_fence_after();
      {
        float v[N];
        tmem_row<N>(tq + col0, v);
        bool dead;
        masked_softmax_row<N, kTrain>(v, mw, p.threshold, tq + col0, dead);
        if (p.drop_thresh) {
          uint32_t kf[N / 32];
          attn_drop_flags<N>(kf, trow, h, p);
          store_row_bf16_dropped<N>(smem_u32(sP + set * kPBytes), row, col0, v, kf, p.drop_scale);
        } else {
          store_row_bf16<N>(smem_u32(sP + set * kPBytes), row, col0, v);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready[set]);
      mbar_wait(&bars->o_full[set], par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tlane + kOCol + set * 64 + 32 * c, r);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = pack_bf16(__uint_as_float(r[16 * gq + 2 * i]), __uint_as_float(r[16 * gq + 2 * i + 1]));
          st_global32(orow + 32 * c + 16 * gq, o);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward (S and P recomputed from the saved / re-projected q, k, v)
// ---------------------------------------------------------------------------------------------------------------------
// Shared memory of the backward: THREE 64 KB stages of [Q | K | V | dO] and ONE [128 x 128] bf16 tile that holds P
// first (for dV = P^T dO) and dS second (for dQ, dK) - with separate P and dS tiles only two stages fit, a stage is
// held until the item's last MMA, and the copies could not run far enough ahead of the MMAs (ncu: 1.16 ms for
// 3.7 GB at N = 32, long-scoreboard stalls 75 % of the samples).
// TMEM, per 256-column buffer: S [0,128) + dP [128,256); then dV [0,64) over the consumed S, and dQ [64,128),
// dK [128,192) once dS has been extracted from dP.
struct BwdBars {
  uint64_t in_full[3], in_empty[3];
  uint64_t sdp_full[2], dqkv_full[2], t_empty[2];
  uint64_t p_ready, pv_done, ds_ready, pd_empty;
  uint32_t tmem_slot;
};

template <int N, bool kTrain>
__global__ void __launch_bounds__(kThreads, 1) attn_core_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                         const __grid_constant__ CUtensorMap tmQKV1,
                                                                         const __grid_constant__ CUtensorMap tmDO,
                                                                         const __grid_constant__ CUtensorMap tmDO1,
                                                                         const CoreArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int S = 3;
  unsigned char* sIn = smem;                          // 3 stages of [Q | K | V | dO]
  unsigned char* sPD = smem + S * 4 * kTile;          // P, then dS
  BwdBars* bars = reinterpret_cast<BwdBars*>(sPD + kPBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.tiles * p.heads;

  // zeros once: a softmax thread only ever writes its own window's columns (P and dS are block-diagonal)
  for (int i = threadIdx.x; i < kPBytes / 16; i += blockDim.x) reinterpret_cast<int4*>(sPD)[i] = make_int4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&bars->in_full[i], 1); mbar_init(&bars->in_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->sdp_full[i], 1); mbar_init(&bars->dqkv_full[i], 1); mbar_init(&bars->t_empty[i], 4);
    }
    mbar_init(&bars->p_ready, 4);
    mbar_init(&bars->ds_ready, 4);
    mbar_init(&bars->pv_done, 1);
    mbar_init(&bars->pd_empty, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmQKV1);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDO1);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < kFirstSoftWarp) {
    reg_dealloc_donor();
    if (warp == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int g = blockIdx.x; g < items; g += gridDim.x) {
        const int tile = g / p.heads, h = g - tile * p.heads;
        mbar_wait(&bars->in_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->in_full[s], 4 * kTile);
          unsigned char* st = sIn + s * 4 * kTile;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            load_tile<N>(st + q * kTile, &tmQKV, &tmQKV1, &bars->in_full[s], p.geo, tile, q * p.d + h * kHd);
          load_tile<N>(st + 3 * kTile, &tmDO, &tmDO1, &bars->in_full[s], p.geo, tile, h * kHd);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 128);
      constexpr uint32_t idKM = umma_idesc_bf16(128, 64, false, true);   // A K-major,  B MN-major   (dQ = dS K)
      constexpr uint32_t idMM = umma_idesc_bf16(128, 64, true, true);    // A MN-major, B MN-major   (dV, dK)
      int s = 0, ps = 0, j = 0;
      uint32_t ph = 0;
      auto issue_grads = [&](int i, int stage) {
        const int b = i & 1;
        const uint32_t sq = smem_u32(sIn + stage * 4 * kTile), sk = sq + kTile, sdo = sq + 3 * kTile;
        const uint32_t spd = smem_u32(sPD);
        const uint32_t t = tmem + b * 256;
        mbar_wait(&bars->p_ready, i & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)     // dV[key] = sum_q P[q, key] dO[q]
            umma_bf16(t, umma_desc_mn_sw128(spd + ks * 2048, kTile, 1024), umma_desc_mn_sw128(sdo + ks * 2048, 8192, 1024),
                      idMM, ks != 0);
          umma_commit(&bars->pv_done);
        }
        __syncwarp();
        mbar_wait(&bars->ds_ready, i & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)     // dQ[q] = sum_key dS[q, key] K[key]
            umma_bf16(t + 64, umma_desc_k_sw128(spd + (ks >> 2) * kTile + (ks & 3) * 32),
                      umma_desc_mn_sw128(sk + ks * 2048, 8192, 1024), idKM, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)     // dK[key] = sum_q dS[q, key] Q[q]
            umma_bf16(t + 128, umma_desc_mn_sw128(spd + ks * 2048, kTile, 1024), umma_desc_mn_sw128(sq + ks * 2048, 8192, 1024),
                      idMM, ks != 0);
          umma_commit(&bars->dqkv_full[b]);
          umma_commit(&bars->in_empty[stage]);
          umma_commit(&bars->pd_empty);
        }
        __syncwarp();
      };
      for (int g = blockIdx.x; g < items; g += gridDim.x, ++j) {
        const int b = j & 1;
        mbar_wait(&bars->in_full[s], ph);
        mbar_wait(&bars->t_empty[b], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sq = smem_u32(sIn + s * 4 * kTile), sk = sq + kTile, sv = sq + 2 * kTile, sdo = sq + 3 * kTile;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem + b * 256, umma_desc_k_sw128(sq + ks * 32), umma_desc_k_sw128(sk + ks * 32), idS, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem + b * 256 + 128, umma_desc_k_sw128(sdo + ks * 32), umma_desc_k_sw128(sv + ks * 32), idS, ks != 0);
          umma_commit(&bars->sdp_full[b]);
        }
        __syncwarp();
        if (j > 0) issue_grads(j - 1, ps);
        ps = s;
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (j > 0) issue_grads(j - 1, ps);
    }
  } else {
    reg_alloc_soft();
    const int q = warp & 3;
    const int set = (warp - kFirstSoftWarp) >> 2;
    const int row = 32 * q + lane;
    const int win = row / N, col0 = win * N;
    const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16) + set * 256;
    const size_t d3 = (size_t)3 * p.d;
    int j = 0;
    for (int g = blockIdx.x; g < items; g += gridDim.x, ++j) {
      if ((j & 1) != set) continue;
      const int tile = g / p.heads, h = g - tile * p.heads;
      const uint32_t par = (j >> 1) & 1;
      uint32_t mw[N / 32];
      {
        const uint32_t* mp = p.bits + p.geo.mask_word<N>(tile, row);
#pragma unroll
        for (int w = 0; w < N / 32; ++w) mw[w] = mp[w];
      }
      const long long trow = p.geo.token_row<N>(tile, row);
      bf16* grow = p.dqkv + (size_t)trow * d3 + h * kHd;
      mbar_wait(&bars->sdp_full[set], par);
      tc_fence_after();
      {
        float v[N];
        tmem_row<N>(tq + col0, v);
        bool dead;
        masked_softmax_row<N, kTrain>(v, mw, p.threshold, tq + col0, dead);
        // attention dropout: O = (P o m) V with m = keep / (1-p), so dV uses P o m and dP enters as dP o m
        uint32_t kf[N / 32];
        const float dscale = p.drop_thresh ? p.drop_scale : 1.f;
        if (p.drop_thresh) {
          attn_drop_flags<N>(kf, trow, h, p);
        } else {
#pragma unroll
          for (int w = 0; w < N / 32; ++w) kf[w] = 0xFFFFFFFFu;
        }
        // delta = sum_j P_j dP_j: dP streamed from TMEM in 32-column chunks (twice: once for delta, once for dS), so
        // a 128-column row of P stays in registers next to one chunk
        float delta = 0.f;
#pragma unroll
        for (int c = 0; c < N / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tq + 128 + col0 + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            delta = fmaf(v[32 * c + i], ((kf[c] >> i) & 1u) ? __uint_as_float(r[i]) * dscale : 0.f, delta);
        }
        // P: the previous item's dQ / dK MMAs must have finished reading the tile (it then held that item's dS)
        mbar_wait(&bars->pd_empty, (j & 1) ^ 1);
        if (p.drop_thresh) store_row_bf16_dropped<N>(smem_u32(sPD), row, col0, v, kf, dscale);
        else store_row_bf16<N>(smem_u32(sPD), row, col0, v);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready);
        // dS = P (dP - delta), under the dV MMAs; P is exactly 0 off the live set; dead rows: 0
#pragma unroll
        for (int c = 0; c < N / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tq + 128 + col0 + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[32 * c + i] = dead ? 0.f
                                 : v[32 * c + i] * ((((kf[c] >> i) & 1u) ? __uint_as_float(r[i]) * dscale : 0.f) - delta);
        }
        mbar_wait(&bars->pv_done, j & 1);      // dV = P^T dO has read P: the tile may take dS
        store_row_bf16<N>(smem_u32(sPD), row, col0, v);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_ready);
      mbar_wait(&bars->dqkv_full[set], par);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 3; ++part) {      // dQ (x 64^-1/2: q carries the scale), dK, dV
        const float mul = part == 0 ? 0.125f : 1.f;
        const int tcol = part == 0 ? 64 : (part == 1 ? 128 : 0);      // TMEM: dV [0,64), dQ [64,128), dK [128,192)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tq + tcol + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o[i] = pack_bf16(__uint_as_float(r[16 * gq + 2 * i]) * mul, __uint_as_float(r[16 * gq + 2 * i + 1]) * mul);
            st_global32(grow + (size_t)part * p.d + 32 * c + 16 * gq, o);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->t_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// q rows of the weight copy and the q part of the bias pre-multiplied by 64^-1/2 (exact in bf16 / fp32)
__global__ void prep_qkv2_kernel(const bf16* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp,
                                 float* __restrict__ bs, int d) {
  const long long nw = (long long)3 * d * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw + 3 * d; i += stride) {
    if (i < nw) {
      const float v = __bfloat162float(w[i]);
      wp[i] = __float2bfloat16(i < (long long)d * d ? v * 0.125f : v);
    } else {
      const int c = (int)(i - nw);
      bs[c] = c < d ? b[c] * 0.125f : b[c];
    }
  }
}

static Geo2 make_geo2(const AttnArgs& a, int W) {
  Geo2 g;
  g.F = a.F; g.K = a.K; g.shift = a.shift; g.layout = a.layout; g.f = a.F / 2;
  g.nWt = a.K / W; g.tps = g.f * g.nWt / (64 / W);
  return g;
}

}  // namespace tc2

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
int gemm_tc_nt_epi_bias(const bf16* A, const bf16* Bt, const float* bias, bf16* C, long long M, int N, int K,
                        cudaStream_t s);

// workspace layout (bf16 units unless noted):
//   forward : wp [3d, d] | bias_s [3d] fp32                                      (qkv goes to the caller's qkv buffer)
//   backward: dqkv [n, 3d] | Wqkv^T [d, 3d] | (qkv [n, 3d] | wp [3d, d] | bias_s [3d] fp32   when qkv is re-projected)
size_t attn2_workspace_bytes(long long n, int d, int backward, int have_qkv) {
  const size_t prep = (size_t)3 * d * d * sizeof(bf16) + (size_t)3 * d * sizeof(float) + 256;
  if (!backward) return prep;
  size_t b = ((size_t)n * 3 * d + (size_t)3 * d * d) * sizeof(bf16) + 256;
  if (!have_qkv) b += (size_t)n * 3 * d * sizeof(bf16) + prep;
  return b;
}

static int project_qkv(const AttnArgs& a, bf16* qkv, unsigned char* prep_ws, cudaStream_t s) {
  const int d = a.d;
  bf16* wp = (bf16*)prep_ws;
  float* bs = (float*)(prep_ws + (((size_t)3 * d * d * sizeof(bf16) + 255) & ~(size_t)255));
  tc2::prep_qkv2_kernel<<<148, 256, 0, s>>>((const bf16*)a.w_qkv, a.b_qkv, wp, bs, d);
  count_launch();
  int st = (int)cudaGetLastError();
  if (st) return st;
  return gemm_tc_nt_epi_bias((const bf16*)a.xn, wp, bs, qkv, a.tokens(), 3 * d, d, s);
}

template <int N>
static int launch_core_fwd(const CUtensorMap& tm, const CUtensorMap& tm1, const tc2::CoreArgs& p, bool train, int grid,
                           int smem, cudaStream_t s) {
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute(tc2::attn_core_fwd_tc2_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kMaxDynSmem2);
    cudaFuncSetAttribute(tc2::attn_core_fwd_tc2_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kMaxDynSmem2);
  });
  if (train) tc2::attn_core_fwd_tc2_kernel<N, true><<<grid, tc2::kThreads, smem, s>>>(tm, tm1, p);
  else tc2::attn_core_fwd_tc2_kernel<N, false><<<grid, tc2::kThreads, smem, s>>>(tm, tm1, p);
  count_launch();
  return (int)cudaGetLastError();
}

template <int N>
static int launch_core_bwd(const CUtensorMap& tm, const CUtensorMap& tm1, const CUtensorMap& tmdo, const CUtensorMap& tmdo1,
                           const tc2::CoreArgs& p, bool train, int grid, int smem, cudaStream_t s) {
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute(tc2::attn_core_bwd_tc2_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kMaxDynSmem2);
    cudaFuncSetAttribute(tc2::attn_core_bwd_tc2_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kMaxDynSmem2);
  });
  if (train) tc2::attn_core_bwd_tc2_kernel<N, true><<<grid, tc2::kThreads, smem, s>>>(tm, tm1, tmdo, tmdo1, p);
  else tc2::attn_core_bwd_tc2_kernel<N, false><<<grid, tc2::kThreads, smem, s>>>(tm, tm1, tmdo, tmdo1, p);
  count_launch();
  return (int)cudaGetLastError();
}

// tensor maps over the token rows of a (B, F, K, cols) tensor: boxes of (64 columns, W keypoints, `frames` frames)
static int make_row_map(CUtensorMap* tm, const void* base, const AttnArgs& a, int cols, int W, int frames) {
  if (a.layout == HWGAT_LAYOUT_WINDOWS) return make_tmap_2d(tm, base, (uint64_t)a.tokens(), (uint64_t)cols, 128);
  return make_tmap_4d(tm, base, (uint64_t)cols, (uint64_t)a.K, (uint64_t)a.F, (uint64_t)a.B, (uint32_t)W, (uint32_t)frames);
}

// forward: qkv (caller's buffer, kept for the backward) = xn . Wp^T + b ; out = attention core
int attn2_fwd(const AttnArgs& a, int W, bf16* qkv, cudaStream_t s) {
  int st;
  if ((st = project_qkv(a, qkv, (unsigned char*)a.workspace, s))) return st;
  CUtensorMap tm, tm1;
  if ((st = make_row_map(&tm, qkv, a, 3 * a.d, W, 2))) return st;
  if ((st = make_row_map(&tm1, qkv, a, 3 * a.d, W, 1))) return st;
  tc2::CoreArgs p{};
  p.bits = a.bits; p.out = (bf16*)a.out; p.threshold = a.threshold; p.d = a.d; p.heads = a.heads; p.tiles = (int)(a.tokens() / 128);
  p.stages = 3;
  p.drop_thresh = drop_threshold16(a.attn_p); p.drop_scale = drop_scale16(p.drop_thresh); p.seed = a.seed; p.offset = a.offset;
  p.geo = tc2::make_geo2(a, W);
  const int items = p.tiles * p.heads;
  const int grid = items < 148 ? items : 148;
  const int smem = p.stages * 3 * tc2::kTile + 2 * tc2::kPBytes + (int)sizeof(tc2::FwdBars) + 1024;
  const bool train = a.threshold >= 0.f;
  switch (2 * W) {
    case 32: return launch_core_fwd<32>(tm, tm1, p, train, grid, smem, s);
    case 64: return launch_core_fwd<64>(tm, tm1, p, train, grid, smem, s);
    case 128: return launch_core_fwd<128>(tm, tm1, p, train, grid, smem, s);
  }
  return HWGAT_ERR_UNSUPPORTED;
}

int attn2_bwd(const AttnArgs& a, int W, const bf16* qkv_saved, cudaStream_t s) {
  const long long n = a.tokens();
  const int d = a.d, d3 = 3 * d;
  bf16* dqkv = (bf16*)a.workspace;
  bf16* wt = dqkv + (size_t)n * d3;
  int st;
  const bf16* qkv = qkv_saved;
  if (!qkv) {
    unsigned char* after = (unsigned char*)(wt + (size_t)d3 * d);
    after = (unsigned char*)(((uintptr_t)after + 255) & ~(uintptr_t)255);
    bf16* q2 = (bf16*)after;
    if ((st = project_qkv(a, q2, after + (size_t)n * d3 * sizeof(bf16), s))) return st;
    qkv = q2;
  }
  CUtensorMap tm, tm1, tmdo, tmdo1;
  if ((st = make_row_map(&tm, qkv, a, d3, W, 2))) return st;
  if ((st = make_row_map(&tm1, qkv, a, d3, W, 1))) return st;
  if ((st = make_row_map(&tmdo, a.d_out, a, d, W, 2))) return st;
  if ((st = make_row_map(&tmdo1, a.d_out, a, d, W, 1))) return st;
  tc2::CoreArgs p{};
  p.bits = a.bits; p.dqkv = dqkv; p.threshold = a.threshold; p.d = d; p.heads = a.heads; p.tiles = (int)(a.tokens() / 128);
  p.stages = 3;
  p.drop_thresh = drop_threshold16(a.attn_p); p.drop_scale = drop_scale16(p.drop_thresh); p.seed = a.seed; p.offset = a.offset;
  p.geo = tc2::make_geo2(a, W);
  const int items = p.tiles * p.heads;
  const int grid = items < 148 ? items : 148;
  const int smem = 3 * 4 * tc2::kTile + tc2::kPBytes + (int)sizeof(tc2::BwdBars) + 1024;
  const bool train = a.threshold >= 0.f;
  switch (2 * W) {
    case 32: st = launch_core_bwd<32>(tm, tm1, tmdo, tmdo1, p, train, grid, smem, s); break;
    case 64: st = launch_core_bwd<64>(tm, tm1, tmdo, tmdo1, p, train, grid, smem, s); break;
    case 128: st = launch_core_bwd<128>(tm, tm1, tmdo, tmdo1, p, train, grid, smem, s); break;
    default: return HWGAT_ERR_UNSUPPORTED;
  }
  if (st) return st;
  // d_xn[n, d] = dQKV[n, 3d] . Wqkv[3d, d]  (against Wqkv^T so that both operands are K-major)
  // qk_perm (q, k kept by K2 with permuted columns): dQ follows k's column order and dK follows q's, so the first 2d
  // columns of dQKV are permuted and the two weight-side GEMMs undo it as they do for K3; dV follows dO: plain
  const bool perm = a.qk_perm != 0;
  if ((st = transpose_bf16((const bf16*)a.w_qkv, wt, d3, d, s, perm, 2 * d))) return st;
  if ((st = gemm_tc_nt_epi_none(dqkv, wt, (bf16*)a.d_xn, n, d, d3, s))) return st;
  // d_w[3d, d] = dQKV^T . xn ; d_b = column sums of dQKV
  return gemm_tc_tn(dqkv, (const bf16*)a.xn, a.d_w, a.d_b, d3, d, n, s, perm, 2 * d);
}

}  // namespace hwgat
