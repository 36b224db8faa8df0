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
// Bound: HBM.  Per token the forward reads q, k, v (3 d bf16) and writes ctx (d bf16) [+ 4 B * heads of logsumexp in
// training]; the backward reads q, k, v, dO, ctx (5 d) and writes dQ, dK, dV (3 d).  The tensor work (mma.sync
// m16n8k16, 12 - 36 instructions per 16 queries per head) is two orders of magnitude under the pipe's rate, so legacy
// HMMA on register fragments is the right tool: a tcgen05 tile (M = 128, operands through smem descriptors, accumulator
// in TMEM) would only add latency to a kernel that waits for memory.
//
// One CTA (8 warps) owns FR consecutive frames of one (sample, window) and one 64-column slice of d (64 / HD heads):
// it stages q, k, v (backward: + dO) of frames f0-1 .. f0+FR in shared memory with 16-byte coalesced copies (128-byte
// rows, XOR-swizzled for ldmatrix), then every warp works on one (frame, 16-query tile) from fragments.
// The backward needs no atomics and no cross-CTA traffic: with the forward's logsumexp saved and delta = rowsum(dO * O)
// formed while dO is staged, a warp computes dQ of its 16 tokens as QUERIES (blocks (f, f-1..f+1)) and dK, dV of the same
// 16 tokens as KEYS (blocks (f-1..f+1, f)); the off-diagonal blocks are computed twice, which costs nothing here.
#include "common.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

namespace band {

constexpr int kThreads = 256;
constexpr int kRowBytes = 128;          // 64 bf16 columns of one token per CTA
constexpr float kLog2e = 1.4426950408889634f;

struct BandArgs {
  const bf16* qkv;       // [n, 3d]  q | k | v (unscaled; the kernels apply hd^-1/2 to the fp32 logits)
  const uint32_t* bits;  // [nW][W][3] : bit j of word (w, i, r) = query keypoint i attends key keypoint j of frame f-1+r
  bf16* out;             // fwd: ctx [n, d]
  float* lse;            // fwd (optional) / bwd: [n, heads] logsumexp of the scaled logits over the edges
  const bf16* d_out;     // bwd: dO [n, d]
  const bf16* ctx;       // bwd: O  [n, d]
  bf16* dqkv;            // bwd: [n, 3d]
  int B, F, K, d, heads;
  float scale;
};

template <int W>
struct Cfg {
  static constexpr int MT = W / 16;          // 16-query tiles per frame
  static constexpr int FR = 8 / MT;          // frames per CTA (one warp per (frame, tile))
  static constexpr int NSLOT = FR + 2;       // + one halo frame on each side
  static constexpr int kTensor = W * kRowBytes;
};

HW_DEV uint32_t swz(int row, int chunk) { return (uint32_t)(row * kRowBytes + ((chunk ^ (row & 7)) << 4)); }

HW_DEV void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
HW_DEV void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
HW_DEV void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;   // 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
HW_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// A fragment (16 rows x 16 columns) of rows row0.. of a staged tensor, columns [chunk0*8, chunk0*8 + 16)
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
template <int W>
HW_DEV Item decode_item(const BandArgs& p) {
  using C = Cfg<W>;
  const int nchunks = (p.F + C::FR - 1) / C::FR, ncc = p.d / 64, nW = p.K / W;
  int i = blockIdx.x;
  Item it;
  it.f0 = (i % nchunks) * C::FR; i /= nchunks;
  it.cc = i % ncc; i /= ncc;
  it.w = i % nW;
  it.b = i / nW;
  return it;
}

// ---------------------------------------------------------------------------------------------------------------------
// K15 forward
// ---------------------------------------------------------------------------------------------------------------------
template <int W, int HD>
__global__ void __launch_bounds__(kThreads, 2) band_attn_fwd_kernel(const BandArgs p) {
  using C = Cfg<W>;
  constexpr int KS = HD / 16, HPC = 64 / HD, CH = HD / 8, NTF = W / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + C::NSLOT * 3 * C::kTensor);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Item it = decode_item<W>(p);
  const int d3 = 3 * p.d;

  // stage q (own frames only), k, v of frames f0-1 .. f0+FR
  for (int piece = tid; piece < C::NSLOT * 3 * W * 8; piece += kThreads) {
    const int chunk = piece & 7, row = (piece >> 3) % W, t = (piece / (8 * W)) % 3, slot = piece / (8 * W * 3);
    const int frame = it.f0 - 1 + slot;
    const bool valid = frame >= 0 && frame < p.F && !(t == 0 && (slot == 0 || slot == C::NSLOT - 1));
    const long long token = ((long long)it.b * p.F + (valid ? frame : 0)) * p.K + it.w * W + row;
    cp_async16(sbase + (slot * 3 + t) * C::kTensor + swz(row, chunk),
               p.qkv + token * d3 + (size_t)t * p.d + it.cc * 64 + chunk * 8, valid);
  }
  for (int i = tid; i < W * 3; i += kThreads) sbits[i] = p.bits[it.w * W * 3 + i];
  cp_async_wait_all();
  __syncthreads();

  const int fi = warp / C::MT, mt = warp % C::MT;
  const int frame = it.f0 + fi;
  if (frame >= p.F) return;
  const int slot = fi + 1, g = lane >> 2, qd = lane & 3;
  const uint32_t sq = sbase + (slot * 3 + 0) * C::kTensor;
  uint32_t mw[2][3];
#pragma unroll
  for (int ri = 0; ri < 2; ++ri)
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      const int kfr = frame - 1 + kf;
      mw[ri][kf] = kfr >= 0 && kfr < p.F ? sbits[(mt * 16 + g + ri * 8) * 3 + kf] : 0u;
    }
  const float sl2 = p.scale * kLog2e;
  const long long tok0 = ((long long)it.b * p.F + frame) * p.K + it.w * W + mt * 16;

#pragma unroll 1
  for (int hh = 0; hh < HPC; ++hh) {
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) load_a(qa[ks], sq, mt * 16, hh * CH + ks * 2, lane);
    float s[3][NTF][4];
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt) s[kf][nt][0] = s[kf][nt][1] = s[kf][nt][2] = s[kf][nt][3] = 0.f;
      const int kfr = frame - 1 + kf;
      if (kfr < 0 || kfr >= p.F) continue;
      const uint32_t sk = sbase + ((slot - 1 + kf) * 3 + 1) * C::kTensor;
#pragma unroll
      for (int np = 0; np < NTF / 2; ++np)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t kb[4];
          load_b_nk(kb, sk, np * 16, hh * CH + ks * 2, lane);
          mma16816(s[kf][2 * np], qa[ks], kb[0], kb[1]);
          mma16816(s[kf][2 * np + 1], qa[ks], kb[2], kb[3]);
        }
    }
    // masked softmax over the edges of each row (rows g and g + 8 of the tile)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int kf = 0; kf < 3; ++kf)
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kp = nt * 8 + qd * 2 + (i & 1);
          const bool live = (mw[i >> 1][kf] >> kp) & 1u;
          if (live) { if (i < 2) m0 = fmaxf(m0, s[kf][nt][i]); else m1 = fmaxf(m1, s[kf][nt][i]); }
        }
    m0 = quad_max(m0); m1 = quad_max(m1);
    const float mm0 = m0 == -INFINITY ? 0.f : m0, mm1 = m1 == -INFINITY ? 0.f : m1;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int kf = 0; kf < 3; ++kf)
#pragma unroll
      for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kp = nt * 8 + qd * 2 + (i & 1);
          const bool live = (mw[i >> 1][kf] >> kp) & 1u;
          const float e = live ? exp2f((s[kf][nt][i] - (i < 2 ? mm0 : mm1)) * sl2) : 0.f;
          s[kf][nt][i] = e;
          if (i < 2) l0 += e; else l1 += e;
        }
    l0 = quad_sum(l0); l1 = quad_sum(l1);

    float o[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f;
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      const int kfr = frame - 1 + kf;
      if (kfr < 0 || kfr >= p.F) continue;
      const uint32_t sv = sbase + ((slot - 1 + kf) * 3 + 2) * C::kTensor;
#pragma unroll
      for (int np = 0; np < NTF / 2; ++np) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[kf][2 * np][0], s[kf][2 * np][1]);
        pa[1] = pack_bf16(s[kf][2 * np][2], s[kf][2 * np][3]);
        pa[2] = pack_bf16(s[kf][2 * np + 1][0], s[kf][2 * np + 1][1]);
        pa[3] = pack_bf16(s[kf][2 * np + 1][2], s[kf][2 * np + 1][3]);
#pragma unroll
        for (int hp = 0; hp < CH / 2; ++hp) {
          uint32_t vb[4];
          load_b_kn(vb, sv, np * 16, hh * CH + hp * 2, lane);
          mma16816(o[2 * hp], pa, vb[0], vb[1]);
          mma16816(o[2 * hp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __syncwarp();
    // the head's output replaces its (consumed) q columns in the warp's own rows
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      *reinterpret_cast<uint32_t*>(smem + (slot * 3) * C::kTensor + swz(mt * 16 + g, hh * CH + c) + qd * 4) =
          pack_bf16(o[c][0] * i0, o[c][1] * i0);
      *reinterpret_cast<uint32_t*>(smem + (slot * 3) * C::kTensor + swz(mt * 16 + g + 8, hh * CH + c) + qd * 4) =
          pack_bf16(o[c][2] * i1, o[c][3] * i1);
    }
    if (p.lse && qd == 0) {
      const int head = it.cc * HPC + hh;
      p.lse[(tok0 + g) * p.heads + head] = l0 > 0.f ? mm0 * p.scale + logf(l0) : 0.f;
      p.lse[(tok0 + g + 8) * p.heads + head] = l1 > 0.f ? mm1 * p.scale + logf(l1) : 0.f;
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int piece = lane + 32 * i, row = piece >> 3, chunk = piece & 7;
    const int4 v = *reinterpret_cast<const int4*>(smem + (slot * 3) * C::kTensor + swz(mt * 16 + row, chunk));
    *reinterpret_cast<int4*>(p.out + (tok0 + row) * p.d + it.cc * 64 + chunk * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K16 backward
// ---------------------------------------------------------------------------------------------------------------------
template <int W, int HD>
__global__ void __launch_bounds__(kThreads, 2) band_attn_bwd_kernel(const BandArgs p) {
  using C = Cfg<W>;
  constexpr int KS = HD / 16, HPC = 64 / HD, CH = HD / 8, NTF = W / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float* s_lse = reinterpret_cast<float*>(smem + C::NSLOT * 4 * C::kTensor);      // [NSLOT][W][HPC]
  float* s_delta = s_lse + C::NSLOT * W * HPC;                                      // [NSLOT][W][HPC]
  uint32_t* sbits = reinterpret_cast<uint32_t*>(s_delta + C::NSLOT * W * HPC);     // [W][3]
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Item it = decode_item<W>(p);
  const int d3 = 3 * p.d;

  // stage q, k, v of frames f0-1 .. f0+FR
  for (int piece = tid; piece < C::NSLOT * 3 * W * 8; piece += kThreads) {
    const int chunk = piece & 7, row = (piece >> 3) % W, t = (piece / (8 * W)) % 3, slot = piece / (8 * W * 3);
    const int frame = it.f0 - 1 + slot;
    const bool valid = frame >= 0 && frame < p.F;
    const long long token = ((long long)it.b * p.F + (valid ? frame : 0)) * p.K + it.w * W + row;
    cp_async16(sbase + (slot * 4 + t) * C::kTensor + swz(row, chunk),
               p.qkv + token * d3 + (size_t)t * p.d + it.cc * 64 + chunk * 8, valid);
  }
  // dO through registers: delta[row][head] = sum over the head's columns of dO * O, formed on the way
  static_assert((Cfg<W>::NSLOT * W * 8) % kThreads == 0, "uniform trip count for the shuffles below");
  for (int piece = tid; piece < C::NSLOT * W * 8; piece += kThreads) {
    const int chunk = piece & 7, row = (piece >> 3) % W, slot = piece / (8 * W);
    const int frame = it.f0 - 1 + slot;
    const bool valid = frame >= 0 && frame < p.F;
    const long long token = ((long long)it.b * p.F + (valid ? frame : 0)) * p.K + it.w * W + row;
    int4 go = make_int4(0, 0, 0, 0), oo = make_int4(0, 0, 0, 0);
    if (valid) {
      go = ld_stream16(p.d_out + token * p.d + it.cc * 64 + chunk * 8);
      oo = ld_stream16(p.ctx + token * p.d + it.cc * 64 + chunk * 8);
    }
    *reinterpret_cast<int4*>(smem + (slot * 4 + 3) * C::kTensor + swz(row, chunk)) = go;
    const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&go);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&oo);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(a[j]), y = __bfloat1622float2(b[j]);
      dot += x.x * y.x + x.y * y.y;
    }
#pragma unroll
    for (int o = 1; o < CH; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (chunk % CH == 0) s_delta[(slot * W + row) * HPC + chunk / CH] = dot;
  }
  for (int i = tid; i < C::NSLOT * W * HPC; i += kThreads) {
    const int hh = i % HPC, row = (i / HPC) % W, slot = i / (HPC * W);
    const int frame = it.f0 - 1 + slot;
    const bool valid = frame >= 0 && frame < p.F;
    const long long token = ((long long)it.b * p.F + (valid ? frame : 0)) * p.K + it.w * W + row;
    s_lse[i] = valid ? p.lse[token * p.heads + it.cc * HPC + hh] * kLog2e : 0.f;
  }
  for (int i = tid; i < W * 3; i += kThreads) sbits[i] = p.bits[it.w * W * 3 + i];
  cp_async_wait_all();
  __syncthreads();

  const int fi = warp / C::MT, mt = warp % C::MT;
  const int frame = it.f0 + fi;
  const bool active = frame < p.F;
  const int slot = fi + 1, g = lane >> 2, qd = lane & 3;
  const float sl2 = p.scale * kLog2e;
  uint32_t outq[HPC][CH][2], outk[HPC][CH][2], outv[HPC][CH][2];

  if (active) {
    const uint32_t sq = sbase + (slot * 4 + 0) * C::kTensor, sk = sbase + (slot * 4 + 1) * C::kTensor;
    const uint32_t sv = sbase + (slot * 4 + 2) * C::kTensor, sdo = sbase + (slot * 4 + 3) * C::kTensor;
    const int r0 = mt * 16 + g, r1 = r0 + 8;   // keypoints of the thread's two rows
#pragma unroll
    for (int hh = 0; hh < HPC; ++hh) {
      // ---- the warp's tokens as QUERIES: dQ = scale * sum over key frames dS . K
      {
        uint32_t qa[KS][4], da[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          load_a(qa[ks], sq, mt * 16, hh * CH + ks * 2, lane);
          load_a(da[ks], sdo, mt * 16, hh * CH + ks * 2, lane);
        }
        const float lse0 = s_lse[(slot * W + r0) * HPC + hh], lse1 = s_lse[(slot * W + r1) * HPC + hh];
        const float dl0 = s_delta[(slot * W + r0) * HPC + hh], dl1 = s_delta[(slot * W + r1) * HPC + hh];
        float dq[CH][4];
#pragma unroll
        for (int c = 0; c < CH; ++c) dq[c][0] = dq[c][1] = dq[c][2] = dq[c][3] = 0.f;
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          const int kfr = frame - 1 + kf;
          if (kfr < 0 || kfr >= p.F) continue;
          const uint32_t skk = sbase + ((slot - 1 + kf) * 4 + 1) * C::kTensor;
          const uint32_t svk = sbase + ((slot - 1 + kf) * 4 + 2) * C::kTensor;
          const uint32_t w0 = sbits[r0 * 3 + kf], w1 = sbits[r1 * 3 + kf];
#pragma unroll
          for (int np = 0; np < NTF / 2; ++np) {
            float s[2][4] = {}, dp[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t kb[4], vb[4];
              load_b_nk(kb, skk, np * 16, hh * CH + ks * 2, lane);
              load_b_nk(vb, svk, np * 16, hh * CH + ks * 2, lane);
              mma16816(s[0], qa[ks], kb[0], kb[1]);
              mma16816(s[1], qa[ks], kb[2], kb[3]);
              mma16816(dp[0], da[ks], vb[0], vb[1]);
              mma16816(dp[1], da[ks], vb[2], vb[3]);
            }
            float ds[2][4];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int kp = np * 16 + t * 8 + qd * 2 + (i & 1);
                const bool live = ((i < 2 ? w0 : w1) >> kp) & 1u;
                const float pr = live ? exp2f(s[t][i] * sl2 - (i < 2 ? lse0 : lse1)) : 0.f;
                ds[t][i] = pr * (dp[t][i] - (i < 2 ? dl0 : dl1));
              }
            uint32_t dsa[4] = {pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]),
                               pack_bf16(ds[1][0], ds[1][1]), pack_bf16(ds[1][2], ds[1][3])};
#pragma unroll
            for (int hp = 0; hp < CH / 2; ++hp) {
              uint32_t kb[4];
              load_b_kn(kb, skk, np * 16, hh * CH + hp * 2, lane);
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
          load_a(ka[ks], sk, mt * 16, hh * CH + ks * 2, lane);
          load_a(va[ks], sv, mt * 16, hh * CH + ks * 2, lane);
        }
        float dk[CH][4], dv[CH][4];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          dk[c][0] = dk[c][1] = dk[c][2] = dk[c][3] = 0.f;
          dv[c][0] = dv[c][1] = dv[c][2] = dv[c][3] = 0.f;
        }
#pragma unroll
        for (int qo = 0; qo < 3; ++qo) {
          const int qfr = frame - 1 + qo;
          if (qfr < 0 || qfr >= p.F) continue;
          const int qslot = slot - 1 + qo, kfrel = 2 - qo;   // this warp's frame seen from the query frame
          const uint32_t sqq = sbase + (qslot * 4 + 0) * C::kTensor, sdq = sbase + (qslot * 4 + 3) * C::kTensor;
#pragma unroll
          for (int np = 0; np < NTF / 2; ++np) {
            float s[2][4] = {}, dp[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t qb[4], gb[4];
              load_b_nk(qb, sqq, np * 16, hh * CH + ks * 2, lane);
              load_b_nk(gb, sdq, np * 16, hh * CH + ks * 2, lane);
              mma16816(s[0], ka[ks], qb[0], qb[1]);
              mma16816(s[1], ka[ks], qb[2], qb[3]);
              mma16816(dp[0], va[ks], gb[0], gb[1]);
              mma16816(dp[1], va[ks], gb[2], gb[3]);
            }
            float pt[2][4], ds[2][4];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int qc = np * 16 + t * 8 + qd * 2 + (i & 1);          // query keypoint (column)
                const bool live = (sbits[qc * 3 + kfrel] >> (i < 2 ? r0 : r1)) & 1u;
                const float lq = s_lse[(qslot * W + qc) * HPC + hh], dq_ = s_delta[(qslot * W + qc) * HPC + hh];
                const float pr = live ? exp2f(s[t][i] * sl2 - lq) : 0.f;
                pt[t][i] = pr;
                ds[t][i] = pr * (dp[t][i] - dq_);
              }
            uint32_t pa[4] = {pack_bf16(pt[0][0], pt[0][1]), pack_bf16(pt[0][2], pt[0][3]),
                              pack_bf16(pt[1][0], pt[1][1]), pack_bf16(pt[1][2], pt[1][3])};
            uint32_t dsa[4] = {pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]),
                               pack_bf16(ds[1][0], ds[1][1]), pack_bf16(ds[1][2], ds[1][3])};
#pragma unroll
            for (int hp = 0; hp < CH / 2; ++hp) {
              uint32_t gb[4], qb[4];
              load_b_kn(gb, sdq, np * 16, hh * CH + hp * 2, lane);
              load_b_kn(qb, sqq, np * 16, hh * CH + hp * 2, lane);
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
      const uint32_t o0 = swz(mt * 16 + g, hh * CH + c) + qd * 4, o1 = swz(mt * 16 + g + 8, hh * CH + c) + qd * 4;
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 0) * C::kTensor + o0) = outq[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 0) * C::kTensor + o1) = outq[hh][c][1];
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 1) * C::kTensor + o0) = outk[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 1) * C::kTensor + o1) = outk[hh][c][1];
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 2) * C::kTensor + o0) = outv[hh][c][0];
      *reinterpret_cast<uint32_t*>(smem + (slot * 4 + 2) * C::kTensor + o1) = outv[hh][c][1];
    }
  __syncwarp();
  const long long tok0 = ((long long)it.b * p.F + frame) * p.K + it.w * W + mt * 16;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int piece = lane + 32 * i, row = piece >> 3, chunk = piece & 7;
      const int4 v = *reinterpret_cast<const int4*>(smem + (slot * 4 + t) * C::kTensor + swz(mt * 16 + row, chunk));
      *reinterpret_cast<int4*>(p.dqkv + (tok0 + row) * d3 + (size_t)t * p.d + it.cc * 64 + chunk * 8) = v;
    }
}

template <int W, int HD>
static int launch_fwd(const BandArgs& p, cudaStream_t s) {
  using C = Cfg<W>;
  constexpr int smem = C::NSLOT * 3 * C::kTensor + W * 3 * 4 + 128;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(band_attn_fwd_kernel<W, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  const long long grid = (long long)p.B * (p.K / W) * (p.d / 64) * ((p.F + C::FR - 1) / C::FR);
  if (grid > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  band_attn_fwd_kernel<W, HD><<<(unsigned)grid, kThreads, smem, s>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}
template <int W, int HD>
static int launch_bwd(const BandArgs& p, cudaStream_t s) {
  using C = Cfg<W>;
  constexpr int HPC = 64 / HD;
  constexpr int smem = C::NSLOT * 4 * C::kTensor + 2 * C::NSLOT * W * HPC * 4 + W * 3 * 4 + 128;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(band_attn_bwd_kernel<W, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  const long long grid = (long long)p.B * (p.K / W) * (p.d / 64) * ((p.F + C::FR - 1) / C::FR);
  if (grid > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  band_attn_bwd_kernel<W, HD><<<(unsigned)grid, kThreads, smem, s>>>(p);
  count_launch();
  return (int)cudaGetLastError();
}

template <bool kBwd>
static int dispatch(const BandArgs& p, int W, cudaStream_t s) {
  const int hd = p.d / p.heads;
#define BAND_CASE(WW, HH) \
  if (W == WW && hd == HH) return kBwd ? launch_bwd<WW, HH>(p, s) : launch_fwd<WW, HH>(p, s);
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
                  float* lse, int B, int F, int K, int d, int heads, int W, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  int st;
  if ((st = gemm_tc_nt_epi_bias(xn, w_qkv, b_qkv, qkv, n, 3 * d, d, s))) return st;
  band::BandArgs p{};
  p.qkv = qkv; p.bits = bits; p.out = out; p.lse = lse;
  p.B = B; p.F = F; p.K = K; p.d = d; p.heads = heads; p.scale = 1.0f / sqrtf((float)(d / heads));
  return band::dispatch<false>(p, W, s);
}

// backward: dqkv (workspace) from the core, then d_xn = dqkv . Wqkv, d_w = dqkv^T . xn, d_b = column sums of dqkv
int band_attn_bwd(const bf16* xn, const bf16* w_qkv, const uint32_t* bits, const bf16* qkv, const bf16* ctx,
                  const float* lse, const bf16* d_out, bf16* d_xn, float* d_w, float* d_b, void* workspace, int B, int F,
                  int K, int d, int heads, int W, cudaStream_t s) {
  const long long n = (long long)B * F * K;
  const int d3 = 3 * d;
  bf16* dqkv = (bf16*)workspace;
  bf16* wt = dqkv + (size_t)n * d3;
  band::BandArgs p{};
  p.qkv = qkv; p.bits = bits; p.lse = const_cast<float*>(lse); p.d_out = d_out; p.ctx = ctx; p.dqkv = dqkv;
  p.B = B; p.F = F; p.K = K; p.d = d; p.heads = heads; p.scale = 1.0f / sqrtf((float)(d / heads));
  int st;
  if ((st = band::dispatch<true>(p, W, s))) return st;
  if ((st = transpose_bf16(w_qkv, wt, d3, d, s))) return st;
  if ((st = gemm_tc_nt_epi_none(dqkv, wt, d_xn, n, d, d3, s))) return st;
  return gemm_tc_tn(dqkv, xn, d_w, d_b, d3, d, n, s);
}

}  // namespace hwgat
