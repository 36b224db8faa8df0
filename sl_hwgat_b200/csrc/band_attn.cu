// K15 / K16: frame-banded graph attention - the attention of the sibling models WGATE and GATE
// (hwgat/models/WGATE.py:68-108, hwgat/models/GATE.py:30-69).
//
// The reference forms the full (F*W)^2 logits of a keypoint window over ALL frames (WGATE: 4 windows of 16 keypoints,
// N = 1024 tokens at T = 64; GATE: one set of 29 keypoints, N = 1856) and ADDS a mask that is 0 on graph edges and
// -10000 elsewhere (WGATE.py:190, GATE.py:142).  The graphs the reference builds (model_params.py:204-229, 59-74) link
// a token to keypoints of its own frame and of the two adjacent frames only, and every query has at least one edge, so
// after the softmax's max subtraction every non-edge weighs exp(-10000 + O(logit range)) = 0 exactly in fp32: the result
// equals a softmax over the edges alone.  That is what runs here - per (sample, window, frame) the W queries against the
// 3 W keys of frames f-1, f, f+1 under a packed (W x 3W)-bit mask - 1/21 of the reference's logits at T = 64 and nothing
// of size N^2 ever exists.  The host side (ops.band_mask_pack) verifies that the adjacency really is frame-banded,
// frame-invariant, 0/1-valued and without empty rows, and refuses anything else: there is no dense fallback.
//
// Roofline: HBM.  Per token the forward reads q, k, v (3 d bf16) and writes ctx (d bf16) [+ 4 B * heads of logsumexp in
// training]; the backward reads q, k, v, dO, ctx (5 d) and writes dQ, dK, dV (3 d).  Measured alone on the WGATE step
// shape (ncu, profiles/r02o_band_wgate.md): DRAM traffic = those bytes; forward 67 %, backward 56 % of the HBM rate, the
// backward limited by instruction issue (~700 warp instructions per 16 tokens per head), not by memory.
// The tensor work (mma.sync m16n8k16, 12 - 42 instructions per 16 tokens per head) is two orders of magnitude under the
// pipe's rate, so legacy HMMA on register fragments is the right tool: a tcgen05 tile (M = 128, operands through smem
// descriptors, accumulator in TMEM) would add latency and synchronisation to blocks of 16 x 48 logits.
//
// One CTA (8 warps) owns FR consecutive frames of one (sample, window) and one 64-column slice of d (64 / HD heads):
// one thread issues a TMA box copy per tensor - (64 columns, W keypoints, FR+2 frames) of q, k, v (backward: + dO) with
// the 128-byte hardware swizzle ldmatrix wants; frames -1 and F of the halo come back zero-filled - and then every warp
// works on one (frame, 16-query tile) from fragments.  (The first version staged with cp.async from all threads: 17 %
// of the forward's instructions, in a kernel that turned out to be bound by instruction issue: ncu r02o.)
// kDiag: in the graphs the reference builds, the blocks between adjacent frames are the identity (a keypoint is linked
// to itself in the previous and the next frame).  The host passes diag = 1 when every off-frame word is 1 << i or 0,
// and the kernels then evaluate only the diagonal of those blocks (12 instead of 24 logits per thread per head at
// W = 16); each CTA verifies the promise against the words it loads and traps if it is broken.  Any other band runs the
// general path.
// The backward needs no atomics and no cross-CTA traffic: with the forward's logsumexp saved and delta = rowsum(dO * O)
// formed while dO is staged, a warp computes dQ of its 16 tokens as QUERIES (blocks (f, f-1..f+1)) and dK, dV of the same
// 16 tokens as KEYS (blocks (f-1..f+1, f)); every block is therefore evaluated twice (once per side), the price of having
// no exchange between warps.
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

namespace band {

constexpr int kRowBytes = 128;          // 64 bf16 columns of one token per CTA
constexpr float kLog2e = 1.4426950408889634f;

struct BandArgs {
  const uint32_t* bits;  // [nW][W][3] : bit j of word (w, i, r) = query keypoint i attends key keypoint j of frame f-1+r
  bf16* out;             // fwd: ctx [n, d]
  float* lse;            // fwd (optional) / bwd: [n, heads] BASE-2 logsumexp of the scaled logits over the edges
  const bf16* ctx;       // bwd: O  [n, d]
  bf16* dqkv;            // bwd: [n, 3d]
  int B, F, K, d, heads;
  float scale;
  int pf_dist;           // L2 prefetch distance in work items (0 = off): the items of the next wave of CTAs
};

// FRAMES = frames per CTA, one warp per (frame, 16-query tile).  Forward: 8 warps.  Backward at W = 16: 6 frames / 6
// warps - 69 KB of shared memory and <= 113 registers let a THIRD CTA share the SM (8 frames: 86 KB, two CTAs), and
// three CTAs overlap the load -> compute -> store phases of an item better than two.
template <int W, int FRAMES = 8 / (W / 16)>
struct Cfg {
  static constexpr int MT = W / 16;          // 16-query tiles per frame
  static constexpr int FR = FRAMES;          // frames per CTA
  static constexpr int kThreads = 32 * FR * MT;
  static constexpr int NSLOT = FR + 2;       // + one halo frame on each side
  static constexpr int kTensor = W * kRowBytes;          // one frame of one tensor
  static constexpr int kAll = NSLOT * kTensor;            // one tensor, all staged frames = one TMA box
};

// byte offset of 16-byte chunk `chunk` of staged row `row` (= slot * W + keypoint): CU_TENSOR_MAP_SWIZZLE_128B
HW_DEV uint32_t swz(int row, int chunk) { return (uint32_t)(row * kRowBytes + ((chunk ^ (row & 7)) << 4)); }

HW_DEV void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
HW_DEV void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// one MUFU each (exp2f / log2f carry range fix-ups worth ~8 instructions per element; the first version spent 29 % of
// its instructions there).  ex2(-inf) = 0: masked logits need no select after the exponential.
HW_DEV float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
  return r;
}
HW_DEV float lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
  return r;
}

// transpose of an 8 x 8 b16 block held in the fragment layout (one 32-bit register per thread)
HW_DEV uint32_t movm(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}
// A fragment of X^T from the A fragment of a 16 x 16 block X: the four 8 x 8 blocks transposed, the off-diagonal two swapped
HW_DEV void transpose_frag(uint32_t (&t)[4], const uint32_t (&a)[4]) {
  t[0] = movm(a[0]);
  t[1] = movm(a[2]);
  t[2] = movm(a[1]);
  t[3] = movm(a[3]);
}

// A fragment (16 rows x 16 columns) of staged rows row0.., columns [chunk0*8, chunk0*8 + 16)
HW_DEV void load_a(uint32_t (&a)[4], uint32_t base, int row0, int chunk0, int lane) {
  ldsm4(a, base + swz(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk0 + (lane >> 4)));
}
// B fragments of two adjacent 8-row tiles of an [n][k] tensor (K, Q, V or dO as "keys x hd"): rows row0 .. row0+15,
// k columns [chunk0*8, +16).  r[0], r[1] -> tile 0; r[2], r[3] -> tile 1.
HW_DEV void load_b_nk(uint32_t (&r)[4], uint32_t base, int row0, int chunk0, int lane) {
  ldsm4(r, base + swz(row0 + (lane & 7) + (lane >> 4) * 8, chunk0 + ((lane >> 3) & 1)));
}
// B fragments of a [k][n] tensor (V, K, Q or dO as "rows to contract x hd"): k rows row0 .. row0+15, two adjacent 8-column
// n tiles starting at chunk0.  r[0], r[1] -> n tile chunk0; r[2], r[3] -> n tile chunk0 + 1.
HW_DEV void load_b_kn(uint32_t (&r)[4], uint32_t base, int row0, int chunk0, int lane) {
  ldsm4t(r, base + swz(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk0 + (lane >> 4)));
}

struct Item {
  int b, w, f0, cc;
};
template <int W, typename C>
HW_DEV Item decode_item(const BandArgs& p, int i) {
  const int nchunks = (p.F + C::FR - 1) / C::FR, ncc = p.d / 64, nW = p.K / W;
  Item it;
  it.f0 = (i % nchunks) * C::FR; i /= nchunks;
  it.cc = i % ncc; i /= ncc;
  it.w = i % nW;
  it.b = i / nW;
  return it;
}

// the words of window w into shared memory; kDiag: hold the host to its promise
template <int W, bool kDiag, int kThreads>
HW_DEV void load_bits(uint32_t* sbits, const BandArgs& p, int w, int tid) {
  for (int i = tid; i < W * 3; i += kThreads) {
    const uint32_t v = p.bits[w * W * 3 + i];
    sbits[i] = v;
    if (kDiag && i % 3 != 1 && (v & ~(1u << (i / 3))) != 0u) __trap();
  }
}

// In the diagonal path the only live logit of row r of an off-frame block is column r.  In the accumulator layout of
// the tile pair np == mt it belongs to lane quad position qd == (g >> 1): element (g & 1) of tile 0 for row g and element
// 2 + (g & 1) of tile 1 for row g + 8.  Only that lane contributes; the quad reductions spread the result.
HW_DEV float diag_row0(const float (&t0)[4], int g) { return g & 1 ? t0[1] : t0[0]; }
HW_DEV float diag_row1(const float (&t1)[4], int g) { return g & 1 ? t1[3] : t1[2]; }
HW_DEV void diag_frag(uint32_t (&a)[4], float v0, float v1, int g) {   // A fragment that is zero off the diagonal
  a[0] = g & 1 ? pack_bf16(0.f, v0) : pack_bf16(v0, 0.f);
  a[1] = 0u;
  a[2] = 0u;
  a[3] = g & 1 ? pack_bf16(0.f, v1) : pack_bf16(v1, 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// K15 forward
// ---------------------------------------------------------------------------------------------------------------------
template <int W, int HD, bool kDiag>
__global__ void __launch_bounds__(Cfg<W>::kThreads, kDiag ? 3 : 2)
band_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const BandArgs p) {
  using C = Cfg<W>;
  constexpr int kThreads = C::kThreads;
  constexpr int KS = HD / 16, HPC = 64 / HD, CH = HD / 8, NTF = W / 8;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + 3 * C::kAll);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbits + W * 3);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Item it = decode_item<W, C>(p, blockIdx.x);

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQKV);
  }
  load_bits<W, kDiag, kThreads>(sbits, p, it.w, tid);
  __syncthreads();
  if (tid == 0) {   // q, k, v of frames f0-1 .. f0+FR: three boxes of (64 columns, W keypoints, FR+2 frames)
    mbar_expect_tx(bar, 3 * C::kAll);
#pragma unroll
    for (int t = 0; t < 3; ++t) tma_load_4d(smem + t * C::kAll, &tmQKV, bar, t * p.d + it.cc * 64, it.w * W, it.f0 - 1, it.b);
    // Pull the boxes of the item a CTA of the NEXT wave will work on into L2: with 2 - 3 resident CTAs per SM the
    // load phase that opens an item is exposed latency, and an L2 hit shortens it.
    if (p.pf_dist > 0 && blockIdx.x + p.pf_dist < gridDim.x) {
      const Item nx = decode_item<W, C>(p, blockIdx.x + p.pf_dist);
#pragma unroll
      for (int t = 0; t < 3; ++t) tma_prefetch_4d(&tmQKV, t * p.d + nx.cc * 64, nx.w * W, nx.f0 - 1, nx.b);
    }
  }
  mbar_wait(bar, 0);

  const int fi = warp / C::MT, mt = warp % C::MT;
  const int frame = it.f0 + fi;
  if (frame >= p.F) return;
  const int slot = fi + 1, g = lane >> 2, qd = lane & 3;
  const int qrow = slot * W + mt * 16;                       // first staged row of the warp's queries
  const uint32_t sq = sbase, sk = sbase + C::kAll, sv = sbase + 2 * C::kAll;
  const int r0 = mt * 16 + g, r1 = r0 + 8;                   // keypoints of the thread's two rows
  uint32_t mw[2][3];
#pragma unroll
  for (int kf = 0; kf < 3; ++kf) {
    const bool ok = (unsigned)(frame - 1 + kf) < (unsigned)p.F;
    mw[0][kf] = ok ? sbits[r0 * 3 + kf] : 0u;
    mw[1][kf] = ok ? sbits[r1 * 3 + kf] : 0u;
  }
  const bool own = qd == (g >> 1);
  const float sl2 = p.scale * kLog2e;
  const long long tok0 = ((long long)it.b * p.F + frame) * p.K + it.w * W + mt * 16;

#pragma unroll 1
  for (int hh = 0; hh < HPC; ++hh) {
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) load_a(qa[ks], sq, qrow, hh * CH + ks * 2, lane);
    constexpr int NB = kDiag ? 1 : 3;          // blocks held as full tiles: the own frame, or all three
    float s[NB][NTF][4];
    float dg[2][2];                            // kDiag: the diagonal logits of the previous / next frame, rows g, g+8
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      const bool ok = (unsigned)(frame - 1 + kf) < (unsigned)p.F;
      const int krow = (slot - 1 + kf) * W;
      if (kDiag && kf != 1) {
        float t2[2][4] = {};
        if (ok) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t kb[4];
            load_b_nk(kb, sk, krow + mt * 16, hh * CH + ks * 2, lane);
            mma16816(t2[0], qa[ks], kb[0], kb[1]);
            mma16816(t2[1], qa[ks], kb[2], kb[3]);
          }
        }
        dg[kf >> 1][0] = own && ((mw[0][kf] >> r0) & 1u) ? diag_row0(t2[0], g) : -INFINITY;
        dg[kf >> 1][1] = own && ((mw[1][kf] >> r1) & 1u) ? diag_row1(t2[1], g) : -INFINITY;
        continue;
      }
      float (&sb)[NTF][4] = s[kDiag ? 0 : kf];
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt) sb[nt][0] = sb[nt][1] = sb[nt][2] = sb[nt][3] = 0.f;
      if (!ok) continue;
#pragma unroll
      for (int np = 0; np < NTF / 2; ++np)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t kb[4];
          load_b_nk(kb, sk, krow + np * 16, hh * CH + ks * 2, lane);
          mma16816(sb[2 * np], qa[ks], kb[0], kb[1]);
          mma16816(sb[2 * np + 1], qa[ks], kb[2], kb[3]);
        }
    }
    // masked softmax over the edges of each row (rows g and g + 8 of the tile): non-edges become -inf once, so the
    // exponentials need no second mask test
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kp = nt * 8 + qd * 2 + (i & 1);
          const float v = (mw[i >> 1][kDiag ? 1 : b] >> kp) & 1u ? s[b][nt][i] : -INFINITY;
          s[b][nt][i] = v;
          if (i < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
        }
    if (kDiag) {
      m0 = fmaxf(m0, fmaxf(dg[0][0], dg[1][0]));
      m1 = fmaxf(m1, fmaxf(dg[0][1], dg[1][1]));
    }
    m0 = quad_max(m0); m1 = quad_max(m1);
    const float mm0 = m0 == -INFINITY ? 0.f : m0 * sl2, mm1 = m1 == -INFINITY ? 0.f : m1 * sl2;   // log2 units
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float e = ex2(fmaf(s[b][nt][i], sl2, -(i < 2 ? mm0 : mm1)));
          s[b][nt][i] = e;
          if (i < 2) l0 += e; else l1 += e;
        }
    if (kDiag) {
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        dg[x][0] = ex2(fmaf(dg[x][0], sl2, -mm0));
        dg[x][1] = ex2(fmaf(dg[x][1], sl2, -mm1));
        l0 += dg[x][0];
        l1 += dg[x][1];
      }
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);

    float o[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f;
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      if ((unsigned)(frame - 1 + kf) >= (unsigned)p.F) continue;
      const int vrow = (slot - 1 + kf) * W;
      if (kDiag && kf != 1) {
        uint32_t pa[4];
        diag_frag(pa, dg[kf >> 1][0], dg[kf >> 1][1], g);
#pragma unroll
        for (int hp = 0; hp < CH / 2; ++hp) {
          uint32_t vb[4];
          load_b_kn(vb, sv, vrow + mt * 16, hh * CH + hp * 2, lane);
          mma16816(o[2 * hp], pa, vb[0], vb[1]);
          mma16816(o[2 * hp + 1], pa, vb[2], vb[3]);
        }
        continue;
      }
      const float (&sb)[NTF][4] = s[kDiag ? 0 : kf];
#pragma unroll
      for (int np = 0; np < NTF / 2; ++np) {
        uint32_t pa[4];
        pa[0] = pack_bf16(sb[2 * np][0], sb[2 * np][1]);
        pa[1] = pack_bf16(sb[2 * np][2], sb[2 * np][3]);
        pa[2] = pack_bf16(sb[2 * np + 1][0], sb[2 * np + 1][1]);
        pa[3] = pack_bf16(sb[2 * np + 1][2], sb[2 * np + 1][3]);
#pragma unroll
        for (int hp = 0; hp < CH / 2; ++hp) {
          uint32_t vb[4];
          load_b_kn(vb, sv, vrow + np * 16, hh * CH + hp * 2, lane);
          mma16816(o[2 * hp], pa, vb[0], vb[1]);
          mma16816(o[2 * hp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    const float i0 = l0 > 0.f ? __fdividef(1.f, l0) : 0.f, i1 = l1 > 0.f ? __fdividef(1.f, l1) : 0.f;
    __syncwarp();
    // the head's output replaces its (consumed) q columns in the warp's own rows
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      *reinterpret_cast<uint32_t*>(smem + swz(qrow + g, hh * CH + c) + qd * 4) = pack_bf16(o[c][0] * i0, o[c][1] * i0);
      *reinterpret_cast<uint32_t*>(smem + swz(qrow + g + 8, hh * CH + c) + qd * 4) = pack_bf16(o[c][2] * i1, o[c][3] * i1);
    }
    if (p.lse && qd == 0) {
      float* lrow = p.lse + (tok0 + g) * p.heads + it.cc * HPC + hh;       // base-2 logsumexp of the scaled logits
      lrow[0] = l0 > 0.f ? mm0 + lg2(l0) : 0.f;
      lrow[8 * p.heads] = l1 > 0.f ? mm1 + lg2(l1) : 0.f;
    }
  }
  __syncwarp();
  bf16* orow = p.out + tok0 * p.d + it.cc * 64;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int piece = lane + 32 * i, row = piece >> 3, chunk = piece & 7;
    const int4 v = *reinterpret_cast<const int4*>(smem + swz(qrow + row, chunk));
    *reinterpret_cast<int4*>(orow + row * p.d + chunk * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K16 backward
// ---------------------------------------------------------------------------------------------------------------------
template <int W>
struct BwdCfg : Cfg<W, W == 16 ? 6 : 4> {
  static constexpr int kCtasPerSm = W == 16 ? 3 : 2;
};

template <int W, int HD, bool kDiag>
__global__ void __launch_bounds__(BwdCfg<W>::kThreads, BwdCfg<W>::kCtasPerSm)
band_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmO, const BandArgs p) {
  using C = BwdCfg<W>;
  constexpr int kThreads = C::kThreads;
  constexpr int KS = HD / 16, HPC = 64 / HD, CH = HD / 8, NTF = W / 8;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_lse = reinterpret_cast<float*>(smem + 4 * C::kAll);                      // [NSLOT * W][HPC]
  float* s_delta = s_lse + C::NSLOT * W * HPC;                                      // [NSLOT * W][HPC]
  uint32_t* sbits = reinterpret_cast<uint32_t*>(s_delta + C::NSLOT * W * HPC);     // [W][3]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbits + W * 3);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Item it = decode_item<W, C>(p, blockIdx.x);
  const int d3 = 3 * p.d;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
  }
  load_bits<W, kDiag, kThreads>(sbits, p, it.w, tid);
  __syncthreads();
  if (tid == 0) {   // q, k, v, dO of frames f0-1 .. f0+FR
    mbar_expect_tx(bar, 4 * C::kAll);
#pragma unroll
    for (int t = 0; t < 3; ++t) tma_load_4d(smem + t * C::kAll, &tmQKV, bar, t * p.d + it.cc * 64, it.w * W, it.f0 - 1, it.b);
    tma_load_4d(smem + 3 * C::kAll, &tmDO, bar, it.cc * 64, it.w * W, it.f0 - 1, it.b);
    if (p.pf_dist > 0 && blockIdx.x + p.pf_dist < gridDim.x) {   // next wave's boxes -> L2 (see the forward)
      const Item nx = decode_item<W, C>(p, blockIdx.x + p.pf_dist);
#pragma unroll
      for (int t = 0; t < 3; ++t) tma_prefetch_4d(&tmQKV, t * p.d + nx.cc * 64, nx.w * W, nx.f0 - 1, nx.b);
      tma_prefetch_4d(&tmDO, nx.cc * 64, nx.w * W, nx.f0 - 1, nx.b);
      tma_prefetch_4d(&tmO, nx.cc * 64, nx.w * W, nx.f0 - 1, nx.b);     // O is read through registers below
    }
  }
  // while the boxes fly: O through registers (delta = rowsum(dO * O) per head) and the saved logsumexp
  // staged rows per pass = kThreads / 8; the last pass may be partial: whole warps drop out (4 rows per warp, row
  // count a multiple of 4), so the shuffles below stay warp-uniform
  constexpr int kRows = C::NSLOT * W, kPasses = (kRows * 8 + kThreads - 1) / kThreads;
  static_assert(kRows % 4 == 0 && kThreads % 32 == 0, "warp-uniform tail");
  const long long tok_base = (long long)it.b * p.F * p.K + it.w * W;      // token (b, frame 0, first keypoint of w)
  const bf16* oo_base = p.ctx + tok_base * p.d + it.cc * 64;
  int4 oo[kPasses];
#pragma unroll
  for (int ps = 0; ps < kPasses; ++ps) {
    const int srow = ps * (kThreads / 8) + (tid >> 3), chunk = tid & 7;    // staged row = slot * W + keypoint
    const int frame = it.f0 - 1 + srow / W;
    oo[ps] = srow < kRows && (unsigned)frame < (unsigned)p.F
                 ? ld_stream16(oo_base + (frame * p.K + srow % W) * p.d + chunk * 8) : make_int4(0, 0, 0, 0);
  }
  for (int i = tid; i < C::NSLOT * W * HPC; i += kThreads) {
    const int hh = i % HPC, srow = i / HPC;
    const int frame = it.f0 - 1 + srow / W;
    s_lse[i] = (unsigned)frame < (unsigned)p.F
                   ? p.lse[(tok_base + (long long)frame * p.K + srow % W) * p.heads + it.cc * HPC + hh] : 0.f;
  }
  mbar_wait(bar, 0);
#pragma unroll
  for (int ps = 0; ps < kPasses; ++ps) {
    const int srow = ps * (kThreads / 8) + (tid >> 3), chunk = tid & 7;
    if (srow >= kRows) break;                                               // warp-uniform
    const int4 go = *reinterpret_cast<const int4*>(smem + 3 * C::kAll + swz(srow, chunk));
    const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&go);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&oo[ps]);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(a[j]), y = __bfloat1622float2(b[j]);
      dot += x.x * y.x + x.y * y.y;
    }
#pragma unroll
    for (int o = 1; o < CH; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (chunk % CH == 0) s_delta[srow * HPC + chunk / CH] = dot;
  }
  __syncthreads();

  const int fi = warp / C::MT, mt = warp % C::MT;
  const int frame = it.f0 + fi;
  const bool active = frame < p.F;
  const int slot = fi + 1, g = lane >> 2, qd = lane & 3;
  const int own_row = slot * W + mt * 16;                     // first staged row of the warp's 16 tokens
  const float sl2 = p.scale * kLog2e;
  const uint32_t sq = sbase, sk = sbase + C::kAll, sv = sbase + 2 * C::kAll, sdo = sbase + 3 * C::kAll;
  const bool own = qd == (g >> 1);
  uint32_t outq[HPC][CH][2], outk[HPC][CH][2], outv[HPC][CH][2];

  if (active) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;   // keypoints of the thread's two rows
#pragma unroll
    for (int hh = 0; hh < HPC; ++hh) {
      // P and dS of the block (own 16 queries x own 16 keys), kept as A fragments: the key side below needs exactly
      // their transposes and gets them with eight movmatrix instead of evaluating the block a second time
      uint32_t p_own[4] = {0u, 0u, 0u, 0u}, ds_own[4] = {0u, 0u, 0u, 0u};
      // ---- the warp's tokens as QUERIES: dQ = scale * sum over key frames dS . K
      {
        uint32_t qa[KS][4], da[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          load_a(qa[ks], sq, own_row, hh * CH + ks * 2, lane);
          load_a(da[ks], sdo, own_row, hh * CH + ks * 2, lane);
        }
        const float lse0 = s_lse[(own_row + g) * HPC + hh], lse1 = s_lse[(own_row + g + 8) * HPC + hh];
        const float dl0 = s_delta[(own_row + g) * HPC + hh], dl1 = s_delta[(own_row + g + 8) * HPC + hh];
        float dq[CH][4];
#pragma unroll
        for (int c = 0; c < CH; ++c) dq[c][0] = dq[c][1] = dq[c][2] = dq[c][3] = 0.f;
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          if ((unsigned)(frame - 1 + kf) >= (unsigned)p.F) continue;
          const int krow = (slot - 1 + kf) * W;
          const uint32_t w0 = sbits[r0 * 3 + kf], w1 = sbits[r1 * 3 + kf];
          const bool diag = kDiag && kf != 1;
#pragma unroll
          for (int np = 0; np < NTF / 2; ++np) {
            if (diag && np != mt) continue;
            float s[2][4] = {}, dp[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t kb[4], vb[4];
              load_b_nk(kb, sk, krow + np * 16, hh * CH + ks * 2, lane);
              load_b_nk(vb, sv, krow + np * 16, hh * CH + ks * 2, lane);
              mma16816(s[0], qa[ks], kb[0], kb[1]);
              mma16816(s[1], qa[ks], kb[2], kb[3]);
              mma16816(dp[0], da[ks], vb[0], vb[1]);
              mma16816(dp[1], da[ks], vb[2], vb[3]);
            }
            uint32_t dsa[4];
            if (diag) {
              const bool lv0 = own && ((w0 >> r0) & 1u), lv1 = own && ((w1 >> r1) & 1u);
              const float p0 = ex2(lv0 ? fmaf(diag_row0(s[0], g), sl2, -lse0) : -INFINITY);
              const float p1 = ex2(lv1 ? fmaf(diag_row1(s[1], g), sl2, -lse1) : -INFINITY);
              diag_frag(dsa, p0 * (diag_row0(dp[0], g) - dl0), p1 * (diag_row1(dp[1], g) - dl1), g);
            } else {
              float pt[2][4], ds[2][4];
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int kp = np * 16 + t * 8 + qd * 2 + (i & 1);
                  const bool live = ((i < 2 ? w0 : w1) >> kp) & 1u;
                  const float pr = ex2(live ? fmaf(s[t][i], sl2, -(i < 2 ? lse0 : lse1)) : -INFINITY);
                  pt[t][i] = pr;
                  ds[t][i] = pr * (dp[t][i] - (i < 2 ? dl0 : dl1));
                }
              dsa[0] = pack_bf16(ds[0][0], ds[0][1]); dsa[1] = pack_bf16(ds[0][2], ds[0][3]);
              dsa[2] = pack_bf16(ds[1][0], ds[1][1]); dsa[3] = pack_bf16(ds[1][2], ds[1][3]);
              if (W == 16 && kf == 1 && np == mt) {
                p_own[0] = pack_bf16(pt[0][0], pt[0][1]); p_own[1] = pack_bf16(pt[0][2], pt[0][3]);
                p_own[2] = pack_bf16(pt[1][0], pt[1][1]); p_own[3] = pack_bf16(pt[1][2], pt[1][3]);
#pragma unroll
                for (int x = 0; x < 4; ++x) ds_own[x] = dsa[x];
              }
            }
#pragma unroll
            for (int hp = 0; hp < CH / 2; ++hp) {
              uint32_t kb[4];
              load_b_kn(kb, sk, krow + np * 16, hh * CH + hp * 2, lane);
              mma16816(dq[2 * hp], dsa, kb[0], kb[1]);
              mma16816(dq[2 * hp + 1], dsa, kb[2], kb[3]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          outq[hh][c][0] = pack_bf16(dq[c][0] * p.scale, dq[c][1] * p.scale);
          outq[hh][c][1] = pack_bf16(dq[c][2] * p.scale, dq[c][3] * p.scale);
        }
      }
      // ---- the warp's tokens as KEYS: dK = scale * sum over query frames dS^T . Q ; dV = sum P^T . dO
      {
        uint32_t ka[KS][4], va[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          load_a(ka[ks], sk, own_row, hh * CH + ks * 2, lane);
          load_a(va[ks], sv, own_row, hh * CH + ks * 2, lane);
        }
        float dk[CH][4], dv[CH][4];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          dk[c][0] = dk[c][1] = dk[c][2] = dk[c][3] = 0.f;
          dv[c][0] = dv[c][1] = dv[c][2] = dv[c][3] = 0.f;
        }
#pragma unroll
        for (int qo = 0; qo < 3; ++qo) {
          if ((unsigned)(frame - 1 + qo) >= (unsigned)p.F) continue;
          const int qrow = (slot - 1 + qo) * W, kfrel = 2 - qo;   // this warp's frame seen from the query frame
          const bool diag = kDiag && qo != 1;
#pragma unroll
          for (int np = 0; np < NTF / 2; ++np) {
            if (diag && np != mt) continue;
            // (own queries x own keys): transposes of the query side's block.  W = 16 only: at W = 32 it covers one of the
            // two in-frame pairs and the eight extra live registers cost more than it saves (measured: +1.5 %)
            const bool reuse = W == 16 && qo == 1 && np == mt;
            float s[2][4] = {}, dp[2][4] = {};
            if (!reuse) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                uint32_t qb[4], gb[4];
                load_b_nk(qb, sq, qrow + np * 16, hh * CH + ks * 2, lane);
                load_b_nk(gb, sdo, qrow + np * 16, hh * CH + ks * 2, lane);
                mma16816(s[0], ka[ks], qb[0], qb[1]);
                mma16816(s[1], ka[ks], qb[2], qb[3]);
                mma16816(dp[0], va[ks], gb[0], gb[1]);
                mma16816(dp[1], va[ks], gb[2], gb[3]);
              }
            }
            uint32_t pa[4], dsa[4];
            if (reuse) {
              transpose_frag(pa, p_own);
              transpose_frag(dsa, ds_own);
            } else if (diag) {   // query keypoint == key keypoint: the thread's own two rows
              const bool lv0 = own && ((sbits[r0 * 3 + kfrel] >> r0) & 1u), lv1 = own && ((sbits[r1 * 3 + kfrel] >> r1) & 1u);
              const float p0 = ex2(lv0 ? fmaf(diag_row0(s[0], g), sl2, -s_lse[(qrow + r0) * HPC + hh]) : -INFINITY);
              const float p1 = ex2(lv1 ? fmaf(diag_row1(s[1], g), sl2, -s_lse[(qrow + r1) * HPC + hh]) : -INFINITY);
              diag_frag(pa, p0, p1, g);
              diag_frag(dsa, p0 * (diag_row0(dp[0], g) - s_delta[(qrow + r0) * HPC + hh]),
                        p1 * (diag_row1(dp[1], g) - s_delta[(qrow + r1) * HPC + hh]), g);
            } else {
              float pt[2][4], ds[2][4];
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int qc = np * 16 + t * 8 + qd * 2 + (i & 1);          // query keypoint (column)
                  const bool live = (sbits[qc * 3 + kfrel] >> (i < 2 ? r0 : r1)) & 1u;
                  const float lq = s_lse[(qrow + qc) * HPC + hh], dq_ = s_delta[(qrow + qc) * HPC + hh];
                  const float pr = ex2(live ? fmaf(s[t][i], sl2, -lq) : -INFINITY);
                  pt[t][i] = pr;
                  ds[t][i] = pr * (dp[t][i] - dq_);
                }
              pa[0] = pack_bf16(pt[0][0], pt[0][1]); pa[1] = pack_bf16(pt[0][2], pt[0][3]);
              pa[2] = pack_bf16(pt[1][0], pt[1][1]); pa[3] = pack_bf16(pt[1][2], pt[1][3]);
              dsa[0] = pack_bf16(ds[0][0], ds[0][1]); dsa[1] = pack_bf16(ds[0][2], ds[0][3]);
              dsa[2] = pack_bf16(ds[1][0], ds[1][1]); dsa[3] = pack_bf16(ds[1][2], ds[1][3]);
            }
#pragma unroll
            for (int hp = 0; hp < CH / 2; ++hp) {
              uint32_t gb[4], qb[4];
              load_b_kn(gb, sdo, qrow + np * 16, hh * CH + hp * 2, lane);
              load_b_kn(qb, sq, qrow + np * 16, hh * CH + hp * 2, lane);
              mma16816(dv[2 * hp], pa, gb[0], gb[1]);
              mma16816(dv[2 * hp + 1], pa, gb[2], gb[3]);
              mma16816(dk[2 * hp], dsa, qb[0], qb[1]);
              mma16816(dk[2 * hp + 1], dsa, qb[2], qb[3]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          outk[hh][c][0] = pack_bf16(dk[c][0] * p.scale, dk[c][1] * p.scale);
          outk[hh][c][1] = pack_bf16(dk[c][2] * p.scale, dk[c][3] * p.scale);
          outv[hh][c][0] = pack_bf16(dv[c][0], dv[c][1]);
          outv[hh][c][1] = pack_bf16(dv[c][2], dv[c][3]);
        }
      }
    }
  }
  __syncthreads();   // every warp is done reading its neighbours' rows: the staged q, k, v rows become the output stage
  if (!active) return;
#pragma unroll
  for (int hh = 0; hh < HPC; ++hh)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t o0 = swz(own_row + g, hh * CH + c) + qd * 4, o1 = swz(own_row + g + 8, hh * CH + c) + qd * 4;
      *reinterpret_cast<uint32_t*>(smem + o0) = outq[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + o1) = outq[hh][c][1];
      *reinterpret_cast<uint32_t*>(smem + C::kAll + o0) = outk[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + C::kAll + o1) = outk[hh][c][1];
      *reinterpret_cast<uint32_t*>(smem + 2 * C::kAll + o0) = outv[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + 2 * C::kAll + o1) = outv[hh][c][1];
    }
  __syncwarp();
  bf16* orow = p.dqkv + (((long long)it.b * p.F + frame) * p.K + it.w * W + mt * 16) * d3 + it.cc * 64;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int piece = lane + 32 * i, row = piece >> 3, chunk = piece & 7;
      const int4 v = *reinterpret_cast<const int4*>(smem + t * C::kAll + swz(own_row + row, chunk));
      *reinterpret_cast<int4*>(orow + row * d3 + t * p.d + chunk * 8) = v;
    }
}

// One wave of resident CTAs ahead (occupancy x SMs), per device and kernel; HWGAT_BAND_PREFETCH=0 switches it off
static int prefetch_distance(const void* kernel, int threads, int smem) {
  static const bool off = [] { const char* e = getenv("HWGAT_BAND_PREFETCH"); return e && e[0] == '0'; }();
  if (off) return 0;
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  return per_sm * sms;
}

constexpr int kMaxDevices = 64;
static int cached_prefetch_distance(int* cache, const void* kernel, int threads, int smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return prefetch_distance(kernel, threads, smem);
  int v = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
  if (v == 0) {
    v = prefetch_distance(kernel, threads, smem) + 1;     // stored + 1 so that "prefetch off" (0) is cached too
    __atomic_store_n(&cache[dev], v, __ATOMIC_RELAXED);
  }
  return v - 1;
}

template <int W, int HD, bool kDiag>
static int launch_fwd(const BandArgs& p, const bf16* qkv, cudaStream_t s) {
  using C = Cfg<W>;
  constexpr int smem = 3 * C::kAll + W * 3 * 4 + 16 + 1024;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(band_attn_fwd_kernel<W, HD, kDiag>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  const long long grid = (long long)p.B * (p.K / W) * (p.d / 64) * ((p.F + C::FR - 1) / C::FR);
  if (grid > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  CUtensorMap tm;
  int st;
  if ((st = make_tmap_4d(&tm, qkv, (uint64_t)3 * p.d, (uint64_t)p.K, (uint64_t)p.F, (uint64_t)p.B, W, C::NSLOT))) return st;
  BandArgs q = p;
  static int pf_cache[kMaxDevices];          // occupancy x SMs of this instantiation, per device (0 = not asked yet)
  q.pf_dist = cached_prefetch_distance(pf_cache, (const void*)band_attn_fwd_kernel<W, HD, kDiag>, C::kThreads, smem);
  band_attn_fwd_kernel<W, HD, kDiag><<<(unsigned)grid, C::kThreads, smem, s>>>(tm, q);
  count_launch();
  return (int)cudaGetLastError();
}
template <int W, int HD, bool kDiag>
static int launch_bwd(const BandArgs& p, const bf16* qkv, const bf16* d_out, cudaStream_t s) {
  using C = BwdCfg<W>;
  constexpr int HPC = 64 / HD;
  constexpr int smem = 4 * C::kAll + 2 * C::NSLOT * W * HPC * 4 + W * 3 * 4 + 16 + 1024;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(band_attn_bwd_kernel<W, HD, kDiag>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  const long long grid = (long long)p.B * (p.K / W) * (p.d / 64) * ((p.F + C::FR - 1) / C::FR);
  if (grid > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  CUtensorMap tm, tmdo, tmo;
  int st;
  if ((st = make_tmap_4d(&tm, qkv, (uint64_t)3 * p.d, (uint64_t)p.K, (uint64_t)p.F, (uint64_t)p.B, W, C::NSLOT))) return st;
  if ((st = make_tmap_4d(&tmdo, d_out, (uint64_t)p.d, (uint64_t)p.K, (uint64_t)p.F, (uint64_t)p.B, W, C::NSLOT))) return st;
  if ((st = make_tmap_4d(&tmo, p.ctx, (uint64_t)p.d, (uint64_t)p.K, (uint64_t)p.F, (uint64_t)p.B, W, C::NSLOT))) return st;
  BandArgs q = p;
  static int pf_cache[kMaxDevices];
  q.pf_dist = cached_prefetch_distance(pf_cache, (const void*)band_attn_bwd_kernel<W, HD, kDiag>, C::kThreads, smem);
  band_attn_bwd_kernel<W, HD, kDiag><<<(unsigned)grid, C::kThreads, smem, s>>>(tm, tmdo, tmo, q);
  count_launch();
  return (int)cudaGetLastError();
}

template <bool kBwd>
static int dispatch(const BandArgs& p, const bf16* qkv, const bf16* d_out, int W, bool diag, cudaStream_t s) {
  const int hd = p.d / p.heads;
#define BAND_CASE(WW, HH)                                                                                  \
  if (W == WW && hd == HH) {                                                                               \
    if (diag) return kBwd ? launch_bwd<WW, HH, true>(p, qkv, d_out, s) : launch_fwd<WW, HH, true>(p, qkv, s);   \
    return kBwd ? launch_bwd<WW, HH, false>(p, qkv, d_out, s) : launch_fwd<WW, HH, false>(p, qkv, s);           \
  }
  BAND_CASE(16, 16) BAND_CASE(16, 32) BAND_CASE(16, 64) BAND_CASE(32, 16) BAND_CASE(32, 32) BAND_CASE(32, 64)
#undef BAND_CASE
  return HWGAT_ERR_UNSUPPORTED;
}

}  // namespace band

int gemm_tc_nt_epi_bias(const bf16* A, const bf16* Bt, const float* bias, bf16* C, long long M, int N, int K,
                        cudaStream_t s);

bool band_attn_supported(int B, int F, int K, int d, int heads, int W) {
  if (B < 0 || F < 1 || K < 1 || d < 1 || heads < 1 || (W != 16 && W != 32) || K % W || d % heads) return false;
  const int hd = d / heads;
  const long long n = (long long)B * F * K;
  return (hd == 16 || hd == 32 || hd == 64) && d % 128 == 0 && n % 128 == 0;
}

// backward workspace: dqkv [n, 3d] | Wqkv^T [d, 3d]  (bf16)
size_t band_attn_workspace_bytes(long long n, int d, int backward) {
  if (!backward) return 0;
  return ((size_t)n * 3 * d + (size_t)3 * d * d) * sizeof(bf16) + 256;
}

// forward: qkv (caller's buffer, kept for the backward) = xn . Wqkv^T + b ; out = banded attention ; lse optional
int band_attn_fwd(const bf16* xn, const bf16* w_qkv, const float* b_qkv, const uint32_t* bits, bf16* out, bf16* qkv,
                  float* lse, int B, int F, int K, int d, int heads, int W, int diag, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  int st;
  if ((st = gemm_tc_nt_epi_bias(xn, w_qkv, b_qkv, qkv, n, 3 * d, d, s))) return st;
  band::BandArgs p{};
  p.bits = bits; p.out = out; p.lse = lse;
  p.B = B; p.F = F; p.K = K; p.d = d; p.heads = heads; p.scale = 1.0f / sqrtf((float)(d / heads));
  return band::dispatch<false>(p, qkv, nullptr, W, diag != 0, s);
}

// backward: dqkv (workspace) from the core, then d_xn = dqkv . Wqkv, d_w = dqkv^T . xn, d_b = column sums of dqkv
int band_attn_bwd(const bf16* xn, const bf16* w_qkv, const uint32_t* bits, const bf16* qkv, const bf16* ctx,
                  const float* lse, const bf16* d_out, bf16* d_xn, float* d_w, float* d_b, void* workspace, int B, int F,
                  int K, int d, int heads, int W, int diag, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  const int d3 = 3 * d;
  bf16* dqkv = (bf16*)workspace;
  bf16* wt = dqkv + (size_t)n * d3;
  band::BandArgs p{};
  p.bits = bits; p.lse = const_cast<float*>(lse); p.ctx = ctx; p.dqkv = dqkv;
  p.B = B; p.F = F; p.K = K; p.d = d; p.heads = heads; p.scale = 1.0f / sqrtf((float)(d / heads));
  int st;
  if ((st = band::dispatch<true>(p, qkv, d_out, W, diag != 0, s))) return st;
  if ((st = transpose_bf16(w_qkv, wt, d3, d, s))) return st;
  if ((st = gemm_tc_nt_epi_none(dqkv, wt, d_xn, n, d, d3, s))) return st;
  return gemm_tc_tn(dqkv, xn, d_w, d_b, d3, d, n, s);
}

}  // namespace hwgat
