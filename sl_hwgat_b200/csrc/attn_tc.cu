// K2 / K3 on tcgen05: fused QKV projection + windowed graph attention (forward, and backward with
// recompute), bf16.
//
// Persistent CTA (one per SM), one tile (128 tokens = 4 windows) at a time, all
// heads of the tile inside the CTA:
//
//   warp 0      TMA producer (whole warp, copies issued under elect.sync): the X tile
//               (resident in smem for all heads; the roll / window partition is the box
//               coordinate of 16-token TMA boxes) and a three-stage ring of per-head
//               weight chunks Wh[192 x 64].
//   warp 1      tcgen05.mma issuer (whole warp, MMAs under elect.sync):
//               QKV_h[128 x 192] = X[128 x d] . Wh^T (+ bias through a ones-column K step),
//               fp32 accumulator in TMEM, two accumulators so head h+1 is multiplied
//               while head h is in the attention phase.  Where the attention warps are
//               the limiter the chunks are paced (they share the tensor pipe with the
//               attention warps' HMMAs).
//   warps 2-3   idle; with warps 0-1 they donate registers (setmaxnreg) to
//   warps 4-11  attention: ONE warp per window (all 32 query rows), the two warp sets
//               (4-7, 8-11) take alternate heads.  tcgen05.ld.16x256b returns the
//               accumulator directly in the HMMA C-fragment layout, so q (A fragments),
//               k (B fragments of q.k^T) and, after one movmatrix.trans per 8x8 block,
//               v (B fragments of P.v) are built in registers: Q, K, V, S and P never
//               touch shared memory.  S = q k^T, threshold drop, packed graph/shift mask,
//               -10000 fill, softmax, O = P v  (HWGATE.py:89-114), O written in token
//               order; the backward kernel recomputes P and forms dq, dk, dv the same way.
#include <cstdlib>

#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

constexpr int kEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;  // warps 0-3: TMA producer, MMA issuer, two idle warps (one register-donor warpgroup)
constexpr int kTcThreads = 32 * (kFirstEpiWarp + kEpiWarps);
// Registers are allocated per SM sub-partition: with 3 warps on each, a thread gets at most 168.  The first warpgroup
// needs ~40, so it hands its share to the attention warps (setmaxnreg): 4*32*40 + 8*32*232 = 64512 <= 65536.
constexpr int kRegsDonor = 40, kRegsEpi = 232;
HW_DEV void reg_dealloc_donor() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegsDonor)); }
HW_DEV void reg_alloc_epi() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsEpi)); }
constexpr int kXChunk = kTileTok * 128;   // 16 KB: 128 rows x 64 bf16
constexpr int kWStage = 192 * 128;        // 24 KB: q|k|v rows of one head x 64 bf16
constexpr int kMaxChunks = 8;             // d <= 512
constexpr int kMaxWStages = 8;
constexpr int kMaxDynSmem = 232448;       // 227 KB: the opt-in limit per CTA on sm_100
constexpr int kAccStride = 256;           // TMEM columns between the two accumulators
// Bias (and the q scale) ride in the GEMM: one extra K=16 MMA step per head, A = a constant [128 x 16] tile whose
// first two columns are 1, B = a per-head [192 x 16] tile whose first two columns are bf16 hi / lo of the bias
// (q rows pre-multiplied by 64^-1/2, as are the q rows of the weight copy).  Both tiles use the swizzle-free
// K-major core-matrix layout.  The attention warps then only pack fp32 -> bf16.
constexpr int kOnesBytes = kTileTok * 16 * 2;   // 4 KB
constexpr int kBiasTile = 192 * 16 * 2;         // 6 KB per head
constexpr int kBiasElems = 192 * 16;

struct TcBars {
  uint64_t x_full[kMaxChunks], x_empty[kMaxChunks];
  uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t bias_full[2], bias_empty[2];
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
// tcgen05.wait::ld with the destination registers of the outstanding load(s) as in/out operands: every use of
// them is ordered after the wait, so loads can be issued early and waited for late.
HW_DEV void tmem_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
HW_DEV uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}

// ---------------------------------------------------------------------------
// In-register masked softmax of the two rows (g, g+8) a thread shares with its quad
// (HWGATE.py:94-111).  s: logits in, probabilities out.  mk[r][i] is 1.0 where the packed
// graph/shift mask lets row r see the thread's i-th column (column 8*(i>>1) + 2t + (i&1)),
// else 0.0; it is built once per tile.  dead[r] is set when the row has no live logit at all
// (then P is uniform over all 32 keys and the gradient w.r.t. the logits is zero).
//
// Training (threshold >= 0): the exponentials of the unmasked softmax are reused for the masked
// one:  P_i = live_i * e_i / sum_j live_j * e_j  with  e_i = exp(s_i - max s), which equals
// softmax(live ? s : -10000) whenever some live e_i is representable; otherwise (every logit
// dropped or masked, or all live ones underflow) an exact second pass runs.
// ---------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;

HW_DEV void build_row_masks(uint32_t mword0, uint32_t mword1, int t, float (&mk)[2][8]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint32_t mw = r == 0 ? mword0 : mword1;
#pragma unroll
    for (int i = 0; i < 8; ++i) mk[r][i] = ((mw >> (8 * (i >> 1) + 2 * t + (i & 1))) & 1u) ? 1.f : 0.f;
  }
}

// exact (two-pass) row: live flags given as 0/1 floats lv[i]
HW_DEV void softmax_row_exact(float (&v)[8], const float (&lv)[8], bool& dead) {
  float m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = lv[i] != 0.f ? v[i] : kNegFill;
    m1 = fmaxf(m1, v[i]);
  }
  m1 = quad_max(m1);
  dead = m1 == kNegFill;
  const float ml = m1 * kLog2e;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = exp2f(fmaf(v[i], kLog2e, -ml)); sum += v[i]; }
  sum = quad_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= inv;
}

template <bool kTrain>
HW_DEV void masked_softmax_tc(float (&s)[4][4], const float (&mk)[2][8], float threshold, bool (&dead)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float v[8];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { v[2 * nt] = s[nt][2 * r]; v[2 * nt + 1] = s[nt][2 * r + 1]; }
    if (kTrain) {
      float m0 = v[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m0 = fmaxf(m0, v[i]);
      m0 = quad_max(m0);
      const float ml = m0 * kLog2e;
      float e[8], sum0 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { e[i] = exp2f(fmaf(v[i], kLog2e, -ml)); sum0 += e[i]; }
      sum0 = quad_sum(sum0);
      const float t0 = threshold * sum0;  // softmax_i > thr  <=>  e_i > thr * sum
      float lv[8], sl = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lv[i] = (e[i] > t0 || v[i] == 0.f) ? 0.f : mk[r][i];  // HWGATE.py:100 drop, :110 exact zeros, masks :102-108
        e[i] *= lv[i];
        sl += e[i];
      }
      sl = quad_sum(sl);
      const bool slow = !(sl > 1e-30f);
      if (__any_sync(0xffffffffu, slow)) {  // warp-uniform branch; quads that do not need it keep their fast result
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = v[i];
        bool dd;
        softmax_row_exact(w, lv, dd);
        const float inv = 1.f / sl;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = slow ? w[i] : e[i] * inv;
        dead[r] = slow && dd;
      } else {
        const float inv = 1.f / sl;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = e[i] * inv;
        dead[r] = false;
      }
    } else {
      float lv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) lv[i] = v[i] == 0.f ? 0.f : mk[r][i];
      softmax_row_exact(v, lv, dead[r]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { s[nt][2 * r] = v[2 * nt]; s[nt][2 * r + 1] = v[2 * nt + 1]; }
  }
}

// ---------------------------------------------------------------------------
// warp 0: TMA producer shared by K2 and K3
// ---------------------------------------------------------------------------
HW_DEV void tc_producer(const TileGeom& geo, int heads, int tiles, int S, TcBars* bars, unsigned char* sX,
                        unsigned char* sW, unsigned char* sBias, const bf16* __restrict__ bias_tiles,
                        const CUtensorMap* tmX, const CUtensorMap* tmW) {
  // whole warp: the loops and waits are warp-uniform, the copies are issued by one elected lane
  const int d = geo.d, nk = d / 64;
  int s = 0, it = 0;
  uint32_t wph = 0;
  int tcount = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
    const int tps = geo.f * geo.kgroups;
    const int b = tile / tps, rr = tile - b * tps, fi = rr / geo.kgroups, kg = rr - fi * geo.kgroups;
    for (int h = 0; h < heads; ++h, ++it) {
      {  // bias tile of this head into slot it & 1
        const int slot = it & 1;
        mbar_wait(&bars->bias_empty[slot], ((it >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->bias_full[slot], kBiasTile);
          bulk_load_1d(sBias + slot * kBiasTile, bias_tiles + (size_t)h * kBiasElems, kBiasTile, &bars->bias_full[slot]);
        }
        __syncwarp();
      }
      for (int c = 0; c < nk; ++c) {
        if (h == 0) {
          mbar_wait(&bars->x_empty[c], (tcount & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(&bars->x_full[c], kXChunk);
            unsigned char* dst = sX + c * kXChunk;
            if (geo.layout == HWGAT_LAYOUT_WINDOWS) {
              tma_load_2d(dst, tmX, &bars->x_full[c], c * 64, tile * kTileTok);
            } else {
#pragma unroll
              for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int tp = 0; tp < 2; ++tp) {
                  int fr = 2 * fi + tp + geo.shift;
                  fr = fr >= geo.F ? fr - geo.F : fr;
                  tma_load_4d(dst + (w * 32 + tp * 16) * 128, tmX, &bars->x_full[c], c * 64, kg * 64 + w * 16, fr, b);
                }
            }
          }
          __syncwarp();
        }
        mbar_wait(&bars->w_empty[s], wph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->w_full[s], kWStage);
          unsigned char* dw = sW + s * kWStage;
#pragma unroll
          for (int q = 0; q < 3; ++q) tma_load_2d(dw + q * 8192, tmW, &bars->w_full[s], c * 64, q * d + h * kHd);
        }
        __syncwarp();
        if (++s == S) { s = 0; wph ^= 1; }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// warp 1: tcgen05.mma issuer shared by K2 and K3
// ---------------------------------------------------------------------------
HW_DEV void tc_issuer(const TileGeom& geo, int heads, int tiles, int S, TcBars* bars, unsigned char* sX,
                      unsigned char* sW, unsigned char* sOnes, unsigned char* sBias, uint32_t tmem, int pace) {
  // whole warp (see elect_one_sync): waits are warp-uniform, MMAs and commits come from the elected lane
  constexpr uint32_t idesc = umma_idesc_bf16(128, 192);
  const int nk = geo.d / 64;
  int s = 0, it = 0, tcount = 0;
  uint32_t wph = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
    for (int h = 0; h < heads; ++h, ++it) {
      const int buf = it & 1;
      mbar_wait(&bars->acc_empty[buf], ((it >> 1) & 1) ^ 1);
      mbar_wait(&bars->bias_full[buf], (it >> 1) & 1);
      tc_fence_after();
      // accumulator = 1 . bias^T  (K = 16 step against the ones tile), then += X . Wh^T
      if (elect_one_sync()) {
        umma_bf16(tmem + buf * kAccStride, umma_desc_k_none(smem_u32(sOnes), kTileTok * 16, 128),
                  umma_desc_k_none(smem_u32(sBias + buf * kBiasTile), 192 * 16, 128), idesc, 0);
        umma_commit(&bars->bias_empty[buf]);
      }
      __syncwarp();
      for (int c = 0; c < nk; ++c) {
        mbar_wait(&bars->w_full[s], wph);
        if (h == 0) mbar_wait(&bars->x_full[c], tcount & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_u32(sX + c * kXChunk), sb = smem_u32(sW + s * kWStage);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem + buf * kAccStride, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc, 1);
          umma_commit(&bars->w_empty[s]);
          if (h == heads - 1) umma_commit(&bars->x_empty[c]);
          if (c == nk - 1) umma_commit(&bars->acc_full[buf]);
        }
        __syncwarp();
        // Pacing: the attention warps' HMMAs share the tensor pipe with these MMAs and wait behind whatever is queued.
        // Where the attention warps are the limiter (forward at d = 128, backward at d <= 256) the next chunk is issued
        // only after this one has completed, so an HMMA waits for at most four MMAs (measured: forward -10 % at
        // d = 128, backward -4 % at d = 128 / 256; where the MMA is the limiter it costs 5-12 %; waiting after every
        // single MMA is slower everywhere).
        if (pace) mbar_wait(&bars->w_empty[s], wph);
        if (++s == S) { s = 0; wph ^= 1; }
      }
    }
  }
}

// (defined with the K3a helpers further down)
HW_DEV void rows_to_blocks(const uint32_t (&r)[32], uint32_t (&f)[4][4]);
HW_DEV void mma_rows_x_blocks_T(float (&acc)[4][4], const uint32_t (&a)[4][4], const uint32_t (&m0)[4][4],
                                const uint32_t (&m1)[4][4]);
HW_DEV void mma_16x64_k16_blocks(float (&acc)[8][4], const uint32_t (&a)[4], const uint32_t (&f)[4][4]);

struct FwdTcArgs {
  const bf16* bias_tiles;  // (heads, 192 x 16) core-matrix tiles made by prep_qkv_kernel
  const uint32_t* bits;
  bf16* out;
  bf16* qkv;               // optional [n, 3d]: q (scaled), k, v as the attention saw them, kept for K3b (hybrid backward)
  float threshold;
  int heads, tiles, w_stages;
  int pace;   // 1: one chunk of MMAs in flight at a time (tc_issuer)
  TileGeom geo;
};

template <bool kTrain>
__global__ void attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                   const FwdTcArgs p);

// Weight copy with the q rows pre-scaled by 64^-1/2 (exact in bf16) and the per-head bias tiles.
__global__ void prep_qkv_kernel(const bf16* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp,
                                bf16* __restrict__ bias_tiles, int d, int heads) {
  const long long nw = (long long)3 * d * d, nb = (long long)heads * kBiasElems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw + nb; i += stride) {
    if (i < nw) {
      const float v = __bfloat162float(w[i]);
      wp[i] = __float2bfloat16(i < (long long)d * d ? v * 0.125f : v);
    } else {
      const int j = (int)(i - nw), h = j / kBiasElems, e = j - h * kBiasElems;
      // e enumerates the tile in memory order: (k>>3)*1536 + (n>>3)*64 + (n&7)*8 + (k&7)
      const int kc = e / 1536, r = e - kc * 1536, n = (r >> 6) * 8 + ((r >> 3) & 7), k = kc * 8 + (r & 7);
      float val = 0.f;
      if (k < 2) {
        const float full = b[(n >> 6) * d + h * kHd + (n & 63)] * (n < 64 ? 0.125f : 1.f);
        const float hi = __bfloat162float(__float2bfloat16(full));
        val = k == 0 ? hi : full - hi;
      }
      bias_tiles[j] = __float2bfloat16(val);
    }
  }
}

static int prep_qkv(const AttnArgs& a, bf16* wp, bf16* bias_tiles, cudaStream_t s) {
  prep_qkv_kernel<<<148, 256, 0, s>>>((const bf16*)a.w_qkv, a.b_qkv, wp, bias_tiles, a.d, a.heads);
  count_launch();
  return (int)cudaGetLastError();
}

// weight-ring depth: THREE stages of one 64-column chunk (24 KB) each.  A deeper ring (up to 6 fit next to the
// resident X tile) lets the MMA warp run ahead in long bursts of back-to-back UTCHMMAs that hold the tensor pipe
// against the attention warps' HMMAs: measured forward 8-10 % slower at every width with 4-6 stages than with 3
// (0.60 / 0.63 / 0.94 ms against 0.53 / 0.58 / 0.86 ms at d = 128 / 256 / 512), backward equal; 2 stages starve the
// backward at d = 512 (2.84 against 2.62 ms).
static int w_stages_for(int d) {
  const int left = 232448 - 1024 - (int)sizeof(TcBars) - kOnesBytes - 2 * kBiasTile - (d / 64) * kXChunk;
  const int st = left / kWStage;
  return st > 3 ? 3 : st;
}

int attn_fwd_tc(const AttnArgs& a, cudaStream_t s) {
  const int d = a.d, nk = d / 64;
  const int stages = w_stages_for(d);
  const int smem_bytes = nk * kXChunk + stages * kWStage + kOnesBytes + 2 * kBiasTile + (int)sizeof(TcBars) + 1024;
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  });
  CUtensorMap tmX, tmW;
  int st;
  if (a.layout == HWGAT_LAYOUT_WINDOWS) {
    if ((st = make_tmap_2d(&tmX, a.xn, (uint64_t)a.tokens(), (uint64_t)d, kTileTok))) return st;
  } else {
    if ((st = make_tmap_4d(&tmX, a.xn, (uint64_t)d, (uint64_t)a.K, (uint64_t)a.F, (uint64_t)a.B, 16))) return st;
  }
  bf16* wp = (bf16*)a.workspace;                 // [3d, d] weight copy, q rows x 64^-1/2
  bf16* bias_tiles = wp + (size_t)3 * d * d;      // [heads][192 x 16]
  if ((st = prep_qkv(a, wp, bias_tiles, s))) return st;
  if ((st = make_tmap_2d(&tmW, wp, (uint64_t)3 * d, (uint64_t)d, 64))) return st;
  FwdTcArgs p;
  p.bias_tiles = bias_tiles; p.bits = a.bits; p.out = (bf16*)a.out; p.threshold = a.threshold;
  p.qkv = (bf16*)a.qkv_out;
  p.heads = a.heads; p.tiles = a.tiles(); p.w_stages = stages; p.pace = d == 128;   // forward: see tc_issuer
  p.geo = make_geom(a.F, a.K, d, a.shift, a.layout);
  const int grid = p.tiles < 148 ? p.tiles : 148;
  if (a.threshold >= 0.f)   // training: threshold drop (HWGATE.py:94-100) compiled in
    attn_fwd_tc_kernel<true><<<grid, kTcThreads, smem_bytes, s>>>(tmX, tmW, p);
  else
    attn_fwd_tc_kernel<false><<<grid, kTcThreads, smem_bytes, s>>>(tmX, tmW, p);
  count_launch();
  return (int)cudaGetLastError();
}

// raw 16 x 64 accumulator rows (tcgen05.ld.16x256b.x8 registers; bias and scale already in the accumulator)
// -> bf16 8x8-block registers: f[ks][0] = rows g, cols 16ks+2t..  f[ks][1] = rows g+8  f[ks][2], f[ks][3]: cols +8
HW_DEV void rows_to_blocks(const uint32_t (&r)[32], uint32_t (&f)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int nt = 2 * ks + x;
      f[ks][2 * x] = pack_bf16(__uint_as_float(r[4 * nt]), __uint_as_float(r[4 * nt + 1]));
      f[ks][2 * x + 1] = pack_bf16(__uint_as_float(r[4 * nt + 2]), __uint_as_float(r[4 * nt + 3]));
    }
}
// acc[16 x 32] += A[16 x 64] . M^T, M (32 rows x 64) given as the block registers of its two 16-row groups:
// the blocks are the B fragments as they are (the q.k^T / dO.v^T pattern)
HW_DEV void mma_rows_x_blocks_T(float (&acc)[4][4], const uint32_t (&a)[4][4], const uint32_t (&m0)[4][4],
                                const uint32_t (&m1)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mma16816(acc[hh], a[ks], m0[ks][hh], m0[ks][2 + hh]);
      mma16816(acc[2 + hh], a[ks], m1[ks][hh], m1[ks][2 + hh]);
    }
}

// acc[16 x 64] += A[16 x 16] . Bm[16 x 64], Bm given as the 8x8-block registers f[ks][..] of its 16 rows
// (layout of ld_rows_as_blocks): B fragments are movmatrix.trans of the blocks.
HW_DEV void mma_16x64_k16_blocks(float (&acc)[8][4], const uint32_t (&a)[4], const uint32_t (&f)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const uint32_t b0 = movmatrix_trans(f[nt >> 1][(nt & 1) * 2]);
    const uint32_t b1 = movmatrix_trans(f[nt >> 1][(nt & 1) * 2 + 1]);
    mma16816(acc[nt], a, b0, b1);
  }
}

// 4x4 transpose of 32-bit words across the 4 lanes of a quad: r[j] on lane t  <-  r[t] of lane j
HW_DEV void quad_transpose4(uint32_t (&r)[4], int t) {
  {
    const bool odd = t & 1;
    uint32_t a = odd ? r[0] : r[1], b = odd ? r[2] : r[3];
    a = __shfl_xor_sync(0xffffffffu, a, 1);
    b = __shfl_xor_sync(0xffffffffu, b, 1);
    if (odd) { r[0] = a; r[2] = b; } else { r[1] = a; r[3] = b; }
  }
  {
    const bool hi = t & 2;
    uint32_t a = hi ? r[0] : r[2], b = hi ? r[1] : r[3];
    a = __shfl_xor_sync(0xffffffffu, a, 2);
    b = __shfl_xor_sync(0xffffffffu, b, 2);
    if (hi) { r[0] = a; r[1] = b; } else { r[2] = a; r[3] = b; }
  }
}
// C fragments [16 x 64] -> bf16 -> global rows p0 (row g) and p1 (row g+8), 64 columns each.  A thread's
// fragment holds 2 columns of every 8-column group; the quad transposes so that lane t owns the whole group t
// (and 4+t) and stores 16 bytes: 4-byte scattered stores were 23 % of the forward attention warps' time.
HW_DEV void store_rows_16x64(const float (&c)[8][4], float mul, bf16* __restrict__ p0, bf16* __restrict__ p1, int t) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[4], r1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r0[j] = pack_bf16(c[4 * half + j][0] * mul, c[4 * half + j][1] * mul);
      r1[j] = pack_bf16(c[4 * half + j][2] * mul, c[4 * half + j][3] * mul);
    }
    quad_transpose4(r0, t);
    quad_transpose4(r1, t);
    *reinterpret_cast<int4*>(p0 + 32 * half + 8 * t) = make_int4((int)r0[0], (int)r0[1], (int)r0[2], (int)r0[3]);
    *reinterpret_cast<int4*>(p1 + 32 * half + 8 * t) = make_int4((int)r1[0], (int)r1[1], (int)r1[2], (int)r1[3]);
  }
}

// C fragments [16 x 64] -> bf16 -> global rows p0 (row g) and p1 (row g+8) in the PERMUTED column order of the dQKV
// workspace: inside each head's 64-column block, logical column 8 nt + 2 t + b is stored at column 16 t + 2 nt + b,
// i.e. the eight words a fragment thread holds for one row are contiguous: one 32-byte store per row, no shuffles.
// The two consumers of dQKV undo it for free: the d_xn GEMM contracts over these columns (its transposed weight copy
// is permuted the same way) and the d_w / d_b GEMM maps its output rows back (gemm_tc_tn, perm64).
HW_DEV void store_rows_16x64_perm(const float (&c)[8][4], float mul, bf16* __restrict__ p0, bf16* __restrict__ p1, int t) {
  uint32_t r0[8], r1[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    r0[nt] = pack_bf16(c[nt][0] * mul, c[nt][1] * mul);
    r1[nt] = pack_bf16(c[nt][2] * mul, c[nt][3] * mul);
  }
  st_global32(p0 + 16 * t, r0);
  st_global32(p1 + 16 * t, r1);
}

// bf16 block registers of 16 rows x 64 columns (rows_to_blocks layout) -> global rows p0 (row g) and p1 (row g+8).
// _perm: in the permuted column order of the dQKV workspace (one 32-byte store per row, no shuffles); the plain form
// transposes inside the quad like store_rows_16x64.
HW_DEV void store_blocks_16x64_perm(const uint32_t (&f)[4][4], bf16* __restrict__ p0, bf16* __restrict__ p1, int t) {
  uint32_t r0[8], r1[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    r0[nt] = f[nt >> 1][(nt & 1) * 2];
    r1[nt] = f[nt >> 1][(nt & 1) * 2 + 1];
  }
  st_global32(p0 + 16 * t, r0);
  st_global32(p1 + 16 * t, r1);
}
HW_DEV void store_blocks_16x64(const uint32_t (&f)[4][4], bf16* __restrict__ p0, bf16* __restrict__ p1, int t) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[4], r1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nt = 4 * half + j;
      r0[j] = f[nt >> 1][(nt & 1) * 2];
      r1[j] = f[nt >> 1][(nt & 1) * 2 + 1];
    }
    quad_transpose4(r0, t);
    quad_transpose4(r1, t);
    *reinterpret_cast<int4*>(p0 + 32 * half + 8 * t) = make_int4((int)r0[0], (int)r0[1], (int)r0[2], (int)r0[3]);
    *reinterpret_cast<int4*>(p1 + 32 * half + 8 * t) = make_int4((int)r1[0], (int)r1[1], (int)r1[2], (int)r1[3]);
  }
}

// ===========================================================================
// K2 kernel.  Attention warps: ONE warp per window (all 32 query rows); the two warp sets (warps 4-7,
// 8-11) work on alternate heads, one TMEM accumulator each.  Every q / k / v row is converted from
// TMEM exactly once, each transposed B fragment serves both 16-row m tiles, and nothing is exchanged
// between warps.
// ===========================================================================
// acc[m][16 x 64] += A[m][16 x 16] . Bm[16 x 64] for both m tiles, Bm as block registers (transposed once)
HW_DEV void mma_2x16x64_k16_blocks(float (&acc0)[8][4], float (&acc1)[8][4], const uint32_t (&a0)[4],
                                   const uint32_t (&a1)[4], const uint32_t (&f)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const uint32_t b0 = movmatrix_trans(f[nt >> 1][(nt & 1) * 2]);
    const uint32_t b1 = movmatrix_trans(f[nt >> 1][(nt & 1) * 2 + 1]);
    mma16816(acc0[nt], a0, b0, b1);
    mma16816(acc1[nt], a1, b0, b1);
  }
}
HW_DEV void zero8x4(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}
HW_DEV void zero4x4(float (&c)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}
HW_DEV void probs_to_afrag(uint32_t (&pa)[2][4], const float (&s)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    pa[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    pa[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    pa[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

template <bool kTrain>
__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                     const __grid_constant__ CUtensorMap tmW,
                                                                     const FwdTcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.geo.d, nk = d / 64, heads = p.heads, S = p.w_stages;
  unsigned char* sX = smem;
  unsigned char* sW = smem + nk * kXChunk;
  unsigned char* sOnes = sW + S * kWStage;
  unsigned char* sBias = sOnes + kOnesBytes;
  TcBars* bars = reinterpret_cast<TcBars*>(sBias + 2 * kBiasTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ones tile [128 x 16], swizzle-free K-major core matrices: element (row r, col k) at
  // (k>>3)*2048 + (r>>3)*128 + (r&7)*16 + (k&7)*2 bytes; columns 0 and 1 are 1.0 (bias hi + lo)
  for (int i = threadIdx.x; i < kOnesBytes / 2; i += blockDim.x) {
    const int k = ((i >> 10) << 3) | (i & 7);
    reinterpret_cast<uint16_t*>(sOnes)[i] = k < 2 ? (uint16_t)0x3f80 : (uint16_t)0;
  }
  fence_proxy_async();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxChunks; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], 1); }
    for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->bias_full[i], 1); mbar_init(&bars->bias_empty[i], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < kFirstEpiWarp) {
    reg_dealloc_donor();
    if (warp == 0) tc_producer(p.geo, heads, p.tiles, S, bars, sX, sW, sBias, p.bias_tiles, &tmX, &tmW);
    if (warp == 1) tc_issuer(p.geo, heads, p.tiles, S, bars, sX, sW, sOnes, sBias, tmem, p.pace);
  } else {
    reg_alloc_epi();
    const int win = warp & 3;                      // TMEM lane quarter == window of the tile
    const int set = (warp - kFirstEpiWarp) >> 2;   // this warp takes the heads whose accumulator is buffer `set`
    const int g = lane >> 2, t = lane & 3;
    const uint32_t tb = tmem + set * kAccStride + ((uint32_t)(32 * win) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      size_t orow[4], qrow[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const size_t tr = (size_t)p.geo.token_row(tile, 32 * win + 8 * r + g);
        orow[r] = tr * d;
        qrow[r] = tr * 3 * d;
      }
      const uint32_t* mw = p.bits + p.geo.mask_base(tile) + 32 * win;
      float mk0[2][8], mk1[2][8];
      build_row_masks(mw[g], mw[g + 8], t, mk0);
      build_row_masks(mw[g + 16], mw[g + 24], t, mk1);
      for (int h = 0; h < heads; ++h, ++it) {
        if ((it & 1) != set) continue;
        mbar_wait(&bars->acc_full[set], (it >> 1) & 1);
        tc_fence_after();
        uint32_t qa0[4][4], qa1[4][4], kb0[4][4], kb1[4][4];
        {
          uint32_t r0[32], r1[32], r2[32], r3[32];
          tmem_ld_16x256b_x8(tb, r0);
          tmem_ld_16x256b_x8(tb + (16u << 16), r1);
          tmem_ld_16x256b_x8(tb + 64, r2);
          tmem_ld_16x256b_x8(tb + (16u << 16) + 64, r3);
          tmem_wait_regs(r0); tmem_wait_regs(r1); tmem_wait_regs(r2); tmem_wait_regs(r3);
          rows_to_blocks(r0, qa0);
          rows_to_blocks(r1, qa1);
          rows_to_blocks(r2, kb0);
          rows_to_blocks(r3, kb1);
        }
        // v is read right away as well and kept as bf16 blocks (32 registers) under the softmax: the accumulator goes
        // back to the MMA warp BEFORE the attention math.  (Releasing it after the softmax held it for ~60 % of the
        // attention phase; at d=512 the MMA of head h+2 then waited ~2300 cycles per head: tensor pipe 54 % busy.)
        uint32_t vb0[4][4], vb1[4][4];
        {
          uint32_t v0[32], v1[32];
          tmem_ld_16x256b_x8(tb + 128, v0);
          tmem_ld_16x256b_x8(tb + (16u << 16) + 128, v1);
          tmem_wait_regs(v0); tmem_wait_regs(v1);
          rows_to_blocks(v0, vb0);
          rows_to_blocks(v1, vb1);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[set]);  // last TMEM read of this head
        if (p.qkv) {
          // keep q (scaled), k, v for the hybrid backward (K3b): q and k with the columns of the head permuted like
          // dQKV (S = q k^T does not care, and the weight-side GEMMs undo it for free), v in the plain order because
          // dP = dO v^T meets the un-permuted dO
          bf16* q0 = p.qkv + qrow[0] + h * kHd;
          bf16* q1 = p.qkv + qrow[1] + h * kHd;
          bf16* q2 = p.qkv + qrow[2] + h * kHd;
          bf16* q3 = p.qkv + qrow[3] + h * kHd;
          store_blocks_16x64_perm(qa0, q0, q1, t);
          store_blocks_16x64_perm(qa1, q2, q3, t);
          store_blocks_16x64_perm(kb0, q0 + d, q1 + d, t);
          store_blocks_16x64_perm(kb1, q2 + d, q3 + d, t);
          store_blocks_16x64(vb0, q0 + 2 * d, q1 + 2 * d, t);
          store_blocks_16x64(vb1, q2 + 2 * d, q3 + 2 * d, t);
        }
        float s0[4][4], s1[4][4];
        zero4x4(s0);
        zero4x4(s1);
        mma_rows_x_blocks_T(s0, qa0, kb0, kb1);
        mma_rows_x_blocks_T(s1, qa1, kb0, kb1);
        bool dead[2];
        masked_softmax_tc<kTrain>(s0, mk0, p.threshold, dead);
        masked_softmax_tc<kTrain>(s1, mk1, p.threshold, dead);
        uint32_t pa0[2][4], pa1[2][4];
        probs_to_afrag(pa0, s0);
        probs_to_afrag(pa1, s1);
        float o0[8][4], o1[8][4];
        zero8x4(o0);
        zero8x4(o1);
        mma_2x16x64_k16_blocks(o0, o1, pa0[0], pa1[0], vb0);
        mma_2x16x64_k16_blocks(o0, o1, pa0[1], pa1[1], vb1);
        store_rows_16x64(o0, 1.f, p.out + orow[0] + h * kHd, p.out + orow[1] + h * kHd, t);
        store_rows_16x64(o1, 1.f, p.out + orow[2] + h * kHd, p.out + orow[3] + h * kHd, t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

struct BwdTcArgs {
  const bf16* bias_tiles;
  const uint32_t* bits;
  const bf16* d_out;
  bf16* dqkv;
  float threshold;
  int heads, tiles, w_stages;
  int pace;   // 1: one chunk of MMAs in flight at a time (tc_issuer)
  TileGeom geo;
};

// compact per-row masks (bit i = the thread's i-th column allowed) -> the float form masked_softmax_tc takes;
// the backward keeps the compact form between heads to save 28 registers
HW_DEV void expand_row_masks(uint32_t c0, uint32_t c1, float (&mk)[2][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mk[0][i] = (c0 >> i) & 1u ? 1.f : 0.f;
    mk[1][i] = (c1 >> i) & 1u ? 1.f : 0.f;
  }
}
HW_DEV uint32_t compact_row_mask(uint32_t mword, int t) {
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) c |= ((mword >> (8 * (i >> 1) + 2 * t + (i & 1))) & 1u) << i;
  return c;
}
// dS = dead ? 0 : P (dP - rowsum(P dP)) in place of dP; returns nothing (P is exactly 0 off the live set)
HW_DEV void softmax_backward_rows(float (&ds)[4][4], const float (&pr)[4][4], const bool (&dead)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float delta = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) delta += pr[nt][2 * r] * ds[nt][2 * r] + pr[nt][2 * r + 1] * ds[nt][2 * r + 1];
    delta = quad_sum(delta);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int x = 0; x < 2; ++x) ds[nt][2 * r + x] = dead[r] ? 0.f : pr[nt][2 * r + x] * (ds[nt][2 * r + x] - delta);
  }
}
// [16 x 32] fp32 accumulator tiles -> 8x8 bf16 blocks b[hi][nn] (rows 8hi+g, columns 8nn+2t..)
HW_DEV void tiles_to_blocks(uint32_t (&b)[2][4], const float (&s)[4][4]) {
#pragma unroll
  for (int nn = 0; nn < 4; ++nn) {
    b[0][nn] = pack_bf16(s[nn][0], s[nn][1]);
    b[1][nn] = pack_bf16(s[nn][2], s[nn][3]);
  }
}
// dO rows (16 x 64 of one head) as A-fragment / block registers, 4-byte loads straight into the fragment layout
// (16-byte loads + quad transposes measured 4 % slower: the shuffles cost more than the narrow loads, which hit L2)
HW_DEV void load_rows_as_blocks_global(const bf16* __restrict__ g0, const bf16* __restrict__ g1, int t,
                                       uint32_t (&ga)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    ga[ks][0] = *reinterpret_cast<const uint32_t*>(g0 + 16 * ks + 2 * t);
    ga[ks][1] = *reinterpret_cast<const uint32_t*>(g1 + 16 * ks + 2 * t);
    ga[ks][2] = *reinterpret_cast<const uint32_t*>(g0 + 16 * ks + 8 + 2 * t);
    ga[ks][3] = *reinterpret_cast<const uint32_t*>(g1 + 16 * ks + 8 + 2 * t);
  }
}

// K3a, second generation: one warp per window, alternate heads per warp set (see attn_fwd_tc_kernel).
// Everything of a window's backward stays inside one warp: no mailbox, no pair barrier.
template <bool kTrain>
__global__ void __launch_bounds__(kTcThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                     const __grid_constant__ CUtensorMap tmW,
                                                                     const BwdTcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.geo.d, nk = d / 64, heads = p.heads, S = p.w_stages;
  unsigned char* sX = smem;
  unsigned char* sW = smem + nk * kXChunk;
  unsigned char* sOnes = sW + S * kWStage;
  unsigned char* sBias = sOnes + kOnesBytes;
  TcBars* bars = reinterpret_cast<TcBars*>(sBias + 2 * kBiasTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ones tile [128 x 16], swizzle-free K-major core matrices: element (row r, col k) at
  // (k>>3)*2048 + (r>>3)*128 + (r&7)*16 + (k&7)*2 bytes; columns 0 and 1 are 1.0 (bias hi + lo)
  for (int i = threadIdx.x; i < kOnesBytes / 2; i += blockDim.x) {
    const int k = ((i >> 10) << 3) | (i & 7);
    reinterpret_cast<uint16_t*>(sOnes)[i] = k < 2 ? (uint16_t)0x3f80 : (uint16_t)0;
  }
  fence_proxy_async();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxChunks; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], 1); }
    for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->bias_full[i], 1); mbar_init(&bars->bias_empty[i], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < kFirstEpiWarp) {
    reg_dealloc_donor();
    if (warp == 0) tc_producer(p.geo, heads, p.tiles, S, bars, sX, sW, sBias, p.bias_tiles, &tmX, &tmW);
    if (warp == 1) tc_issuer(p.geo, heads, p.tiles, S, bars, sX, sW, sOnes, sBias, tmem, p.pace);
  } else {
    reg_alloc_epi();
    const int win = warp & 3;
    const int set = (warp - kFirstEpiWarp) >> 2;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t tb = tmem + set * kAccStride + ((uint32_t)(32 * win) << 16);
    const size_t d3 = (size_t)3 * d;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      size_t tr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) tr[r] = (size_t)p.geo.token_row(tile, 32 * win + 8 * r + g);
      const uint32_t* mw = p.bits + p.geo.mask_base(tile) + 32 * win;
      uint32_t cm[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) cm[r] = compact_row_mask(mw[8 * r + g], t);
      for (int h = 0; h < heads; ++h, ++it) {
        if ((it & 1) != set) continue;
        if (h + 2 < heads) {  // pull this warp's next head's dO rows (32 x 128 bytes) into L2 ahead of time
          const size_t r = (size_t)p.geo.token_row(tile, 32 * win + lane);
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.d_out + r * d + (h + 2) * kHd));
        }
        mbar_wait(&bars->acc_full[set], (it >> 1) & 1);
        tc_fence_after();
        uint32_t qa0[4][4], qa1[4][4], kb0[4][4], kb1[4][4];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_16x256b_x8(tb, r0);
          tmem_ld_16x256b_x8(tb + (16u << 16), r1);
          tmem_wait_regs(r0); tmem_wait_regs(r1);
          rows_to_blocks(r0, qa0);
          rows_to_blocks(r1, qa1);
          tmem_ld_16x256b_x8(tb + 64, r0);
          tmem_ld_16x256b_x8(tb + (16u << 16) + 64, r1);
          tmem_wait_regs(r0); tmem_wait_regs(r1);
          rows_to_blocks(r0, kb0);
          rows_to_blocks(r1, kb1);
        }
        // dO rows (L2-resident thanks to the prefetch): loaded here, after the 128 raw q/k registers are dead
        uint32_t ga0[4][4], ga1[4][4];
        load_rows_as_blocks_global(p.d_out + tr[0] * d + h * kHd, p.d_out + tr[1] * d + h * kHd, t, ga0);
        load_rows_as_blocks_global(p.d_out + tr[2] * d + h * kHd, p.d_out + tr[3] * d + h * kHd, t, ga1);
        // ---- P (recomputed)
        float p0[4][4], p1[4][4];
        zero4x4(p0);
        zero4x4(p1);
        mma_rows_x_blocks_T(p0, qa0, kb0, kb1);
        mma_rows_x_blocks_T(p1, qa1, kb0, kb1);
        bool dead0[2], dead1[2];
        {
          float mk[2][8];
          expand_row_masks(cm[0], cm[1], mk);
          masked_softmax_tc<kTrain>(p0, mk, p.threshold, dead0);
          expand_row_masks(cm[2], cm[3], mk);
          masked_softmax_tc<kTrain>(p1, mk, p.threshold, dead1);
        }
        uint32_t v0[32], v1[32];   // (issued after the softmax: 64 more live registers under it made the kernel spill)
        tmem_ld_16x256b_x8(tb + 128, v0);
        tmem_ld_16x256b_x8(tb + (16u << 16) + 128, v1);
        tmem_wait_regs(v0); tmem_wait_regs(v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[set]);  // last TMEM read of this head
        // ---- dP = dO v^T, dS
        float s0[4][4], s1[4][4];
        zero4x4(s0);
        zero4x4(s1);
        {
          uint32_t vb0[4][4], vb1[4][4];
          rows_to_blocks(v0, vb0);
          rows_to_blocks(v1, vb1);
          mma_rows_x_blocks_T(s0, ga0, vb0, vb1);
          mma_rows_x_blocks_T(s1, ga1, vb0, vb1);
        }
        softmax_backward_rows(s0, p0, dead0);
        softmax_backward_rows(s1, p1, dead1);
        uint32_t pb0[2][4], pb1[2][4], db0[2][4], db1[2][4];
        tiles_to_blocks(pb0, p0);
        tiles_to_blocks(pb1, p1);
        tiles_to_blocks(db0, s0);
        tiles_to_blocks(db1, s1);
        // ---- dq = dS k * scale.  This is the point of highest register pressure (q, k, dO, P and dS blocks are all
        // live), so the two query m tiles go one after the other (32 accumulator registers instead of 64).
        {
          float dq[8][4];
          zero8x4(dq);
          {
            const uint32_t a0[4] = {db0[0][0], db0[1][0], db0[0][1], db0[1][1]};
            mma_16x64_k16_blocks(dq, a0, kb0);
            const uint32_t a1[4] = {db0[0][2], db0[1][2], db0[0][3], db0[1][3]};
            mma_16x64_k16_blocks(dq, a1, kb1);
          }
          store_rows_16x64_perm(dq, 0.125f, p.dqkv + tr[0] * d3 + h * kHd, p.dqkv + tr[1] * d3 + h * kHd, t);
          zero8x4(dq);
          {
            const uint32_t a0[4] = {db1[0][0], db1[1][0], db1[0][1], db1[1][1]};
            mma_16x64_k16_blocks(dq, a0, kb0);
            const uint32_t a1[4] = {db1[0][2], db1[1][2], db1[0][3], db1[1][3]};
            mma_16x64_k16_blocks(dq, a1, kb1);
          }
          store_rows_16x64_perm(dq, 0.125f, p.dqkv + tr[2] * d3 + h * kHd, p.dqkv + tr[3] * d3 + h * kHd, t);
        }
        // ---- dv = P^T dO : key m tiles jt = 0,1; k steps = the two query groups
        {
          float dv0[8][4], dv1[8][4];
          zero8x4(dv0);
          zero8x4(dv1);
          {
            const uint32_t a0[4] = {movmatrix_trans(pb0[0][0]), movmatrix_trans(pb0[0][1]), movmatrix_trans(pb0[1][0]),
                                    movmatrix_trans(pb0[1][1])};
            const uint32_t a1[4] = {movmatrix_trans(pb0[0][2]), movmatrix_trans(pb0[0][3]), movmatrix_trans(pb0[1][2]),
                                    movmatrix_trans(pb0[1][3])};
            mma_2x16x64_k16_blocks(dv0, dv1, a0, a1, ga0);
          }
          {
            const uint32_t a0[4] = {movmatrix_trans(pb1[0][0]), movmatrix_trans(pb1[0][1]), movmatrix_trans(pb1[1][0]),
                                    movmatrix_trans(pb1[1][1])};
            const uint32_t a1[4] = {movmatrix_trans(pb1[0][2]), movmatrix_trans(pb1[0][3]), movmatrix_trans(pb1[1][2]),
                                    movmatrix_trans(pb1[1][3])};
            mma_2x16x64_k16_blocks(dv0, dv1, a0, a1, ga1);
          }
          store_rows_16x64_perm(dv0, 1.f, p.dqkv + tr[0] * d3 + 2 * d + h * kHd, p.dqkv + tr[1] * d3 + 2 * d + h * kHd, t);
          store_rows_16x64_perm(dv1, 1.f, p.dqkv + tr[2] * d3 + 2 * d + h * kHd, p.dqkv + tr[3] * d3 + 2 * d + h * kHd, t);
        }
        // ---- dk = dS^T q (q carries the scale)
        {
          float dk0[8][4], dk1[8][4];
          zero8x4(dk0);
          zero8x4(dk1);
          {
            const uint32_t a0[4] = {movmatrix_trans(db0[0][0]), movmatrix_trans(db0[0][1]), movmatrix_trans(db0[1][0]),
                                    movmatrix_trans(db0[1][1])};
            const uint32_t a1[4] = {movmatrix_trans(db0[0][2]), movmatrix_trans(db0[0][3]), movmatrix_trans(db0[1][2]),
                                    movmatrix_trans(db0[1][3])};
            mma_2x16x64_k16_blocks(dk0, dk1, a0, a1, qa0);
          }
          {
            const uint32_t a0[4] = {movmatrix_trans(db1[0][0]), movmatrix_trans(db1[0][1]), movmatrix_trans(db1[1][0]),
                                    movmatrix_trans(db1[1][1])};
            const uint32_t a1[4] = {movmatrix_trans(db1[0][2]), movmatrix_trans(db1[0][3]), movmatrix_trans(db1[1][2]),
                                    movmatrix_trans(db1[1][3])};
            mma_2x16x64_k16_blocks(dk0, dk1, a0, a1, qa1);
          }
          store_rows_16x64_perm(dk0, 1.f, p.dqkv + tr[0] * d3 + d + h * kHd, p.dqkv + tr[1] * d3 + d + h * kHd, t);
          store_rows_16x64_perm(dk1, 1.f, p.dqkv + tr[2] * d3 + d + h * kHd, p.dqkv + tr[3] * d3 + d + h * kHd, t);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int attn_bwd_tc(const AttnArgs& a, bf16* dqkv, cudaStream_t s) {
  const int d = a.d, nk = d / 64;
  const int stages = w_stages_for(d);
  const int smem_bytes = nk * kXChunk + stages * kWStage + kOnesBytes + 2 * kBiasTile + (int)sizeof(TcBars) + 1024;
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  });
  CUtensorMap tmX, tmW;
  int st;
  if (a.layout == HWGAT_LAYOUT_WINDOWS) {
    if ((st = make_tmap_2d(&tmX, a.xn, (uint64_t)a.tokens(), (uint64_t)d, kTileTok))) return st;
  } else {
    if ((st = make_tmap_4d(&tmX, a.xn, (uint64_t)d, (uint64_t)a.K, (uint64_t)a.F, (uint64_t)a.B, 16))) return st;
  }
  bf16* wp = dqkv + (size_t)a.tokens() * 3 * d + (size_t)3 * d * d;   // behind dQKV and Wqkv^T
  bf16* bias_tiles = wp + (size_t)3 * d * d;
  if ((st = prep_qkv(a, wp, bias_tiles, s))) return st;
  if ((st = make_tmap_2d(&tmW, wp, (uint64_t)3 * d, (uint64_t)d, 64))) return st;
  BwdTcArgs p;
  p.bias_tiles = bias_tiles; p.bits = a.bits; p.d_out = (const bf16*)a.d_out; p.dqkv = dqkv; p.threshold = a.threshold;
  p.heads = a.heads; p.tiles = a.tiles(); p.w_stages = stages; p.pace = d <= 256;   // backward: see tc_issuer
  p.geo = make_geom(a.F, a.K, d, a.shift, a.layout);
  const int grid = p.tiles < 148 ? p.tiles : 148;
  if (a.threshold >= 0.f)
    attn_bwd_tc_kernel<true><<<grid, kTcThreads, smem_bytes, s>>>(tmX, tmW, p);
  else
    attn_bwd_tc_kernel<false><<<grid, kTcThreads, smem_bytes, s>>>(tmX, tmW, p);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// bf16 entry points of the C ABI (api.cu): K2 = one kernel; K3 = K3a + the two weight-side GEMMs
// ---------------------------------------------------------------------------
int attn_fwd_bf16(const AttnArgs& a, cudaStream_t s) { return attn_fwd_tc(a, s); }

int attn_bwd_bf16(const AttnArgs& a, cudaStream_t s) {
  bf16* dqkv = (bf16*)a.workspace;
  int st = attn_bwd_tc(a, dqkv, s);
  if (st) return st;
  const long long n = a.tokens();
  const int d = a.d, d3 = 3 * d;
  // d_xn[n, d] = dQKV[n, 3d] . Wqkv[3d, d]  (against Wqkv^T so that both operands are K-major)
  bf16* wt = dqkv + n * d3;
  // (dQKV's columns are permuted inside each head block, see store_rows_16x64_perm: the copy is permuted alike)
  if ((st = transpose_bf16((const bf16*)a.w_qkv, wt, d3, d, s, true))) return st;
  if ((st = gemm_tc_nt_epi_none(dqkv, wt, (bf16*)a.d_xn, n, d, d3, s))) return st;   // K10's GEMM (32-byte stores)
  // d_w[3d, d] = dQKV^T . xn ; d_b = column sums of dQKV (same kernel)
  return gemm_tc_tn(dqkv, (const bf16*)a.xn, a.d_w, a.d_b, d3, d, n, s, true);
}

}  // namespace hwgat
