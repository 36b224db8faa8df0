// K10: the FeedForward of a block (HWGATE.py:120-136) on tcgen05, bf16 / autocast path.
//
//   forward   act = dropout(gelu(h . W1^T + b1))      one GEMM, bias + exact GELU + dropout in its epilogue; the
//             gp  = d act / d (h . W1^T + b1)          epilogue also writes the local derivative (dropout mask and
//             v0  = act . W2^T                         1/(1-p) folded in), so the backward needs no erf / Philox
//   backward  du0 = (dv0 . W2) o gp                    one GEMM with the multiply in its epilogue
//             dW2 = dv0^T . act,   dW1 = du0^T . h,  db1 = colsum(du0)       gemm_tc_tn (gemm_tc.cu)
//             dh  = du0 . W1
//
// fc2's bias, the second dropout and the residual add stay with K6 (block_fused.cu), which fuses them with the
// LayerNorm that follows.  This replaces cuBLAS GEMM + K7 (bias_gelu_dropout) pairs: the (n, 2d) hidden
// activation is written once by the GEMM that produces it instead of written, re-read and written again.
//
// gemm_nt_epi_kernel: C[M,N] = epilogue(A[M,K] . Bt[N,K]^T), persistent, warp-specialised:
//   warp 0 TMA producer, warp 1 tcgen05.mma issuer (+ TMEM owner), warps 2.. epilogue (two or four warps per TMEM
//   lane quarter, one column slice each), two TMEM accumulators so the epilogue of tile i overlaps the MMAs of
//   tile i+1.  Row-per-thread epilogues store 32 bytes (one full sector) per instruction.  The GELU epilogue is
//   ALU work (Philox + erfc, ~40 instructions per element) and gets 16 warps.
#include "ew.cuh"
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

enum { kEpiNone = 0, kEpiGelu = 1, kEpiMul = 2, kEpiBias = 3, kEpiGeluEval = 4 };   // kEpiBias: C = acc + bias (fp32 add, then bf16)
// kEpiGeluEval: C = gelu(acc + bias) only - inference (no dropout mask, no local derivative): half the instructions of kEpiGelu

struct EpiArgs {
  bf16* C;            // [M, N] output
  bf16* C2;           // kEpiGelu: local derivative (may be null: inference)
  const bf16* G;      // kEpiMul: elementwise factor [M, N]
  const float* bias;  // kEpiGelu: [N] or null
  float scale;        // dropout 1/(1-p)
  uint32_t thresh;    // dropout 16-bit threshold, 0 = no dropout
  unsigned long long seed, offset;
};

constexpr int kFM = 128, kFK = 64;
// epilogue warps: 8 for the memory-bound epilogues, 16 for the ALU-bound GELU one (4 per SM sub-partition:
// with 2 the issue slots were ~2/3 used and the fc1 GEMM ran at the speed of the standalone K7 kernel)
template <int EPI>
struct EpiWarps { static constexpr int kWarps = (EPI == 1 || EPI == 4) ? 16 : 8; static constexpr int kThreads = 32 * (2 + kWarps); };

constexpr int kMaxBiasN = 2048;  // widest bias the GELU epilogue stages in shared memory (hidden <= 2048)

template <int BN>
struct FfnCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kFM * kFK * 2;
  static constexpr int kBBytes = BN * kFK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kBiasOff = kBarOff + 256;        // fp32 bias of all N columns (GELU epilogue)
  static constexpr int kSmem = kBiasOff + kMaxBiasN * 4 + 1024;
  static constexpr int kTmemCols = 2 * BN;
};

// explicit ld.shared: the 1024-byte-aligned smem pointer is built with integer arithmetic, so the compiler no longer
// knows its address space and would emit generic LD.E loads
HW_DEV float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

template <int EPI>
HW_DEV void epilogue_chunk(const uint32_t (&r)[32], const EpiArgs& e, size_t elem, int col, const float* sbias) {
  // r: 32 consecutive columns of one row (fp32 accumulator); elem = row * N + col.  Every store is 32 bytes
  // = one full sector of the thread's own row.
  if (EPI == kEpiNone) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t p[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        p[i] = pack_bf16(__uint_as_float(r[16 * g + 2 * i]), __uint_as_float(r[16 * g + 2 * i + 1]));
      st_global32(e.C + elem + 16 * g, p);
    }
  } else if (EPI == kEpiBias) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t p[8];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 b = lds_f4(sbias + col + 16 * g + 4 * q4);
        p[2 * q4] = pack_bf16(__uint_as_float(r[16 * g + 4 * q4]) + b.x, __uint_as_float(r[16 * g + 4 * q4 + 1]) + b.y);
        p[2 * q4 + 1] = pack_bf16(__uint_as_float(r[16 * g + 4 * q4 + 2]) + b.z, __uint_as_float(r[16 * g + 4 * q4 + 3]) + b.w);
      }
      st_global32(e.C + elem + 16 * g, p);
    }
  } else if (EPI == kEpiGeluEval) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t p[8];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 b = lds_f4(sbias + col + 16 * g + 4 * q4);
        p[2 * q4] = pack_bf16(gelu_exact(__uint_as_float(r[16 * g + 4 * q4]) + b.x),
                              gelu_exact(__uint_as_float(r[16 * g + 4 * q4 + 1]) + b.y));
        p[2 * q4 + 1] = pack_bf16(gelu_exact(__uint_as_float(r[16 * g + 4 * q4 + 2]) + b.z),
                                  gelu_exact(__uint_as_float(r[16 * g + 4 * q4 + 3]) + b.w));
      }
      st_global32(e.C + elem + 16 * g, p);
    }
  } else if (EPI == kEpiMul) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t w[8], o[8];
      ld_global32(e.G + elem + 16 * g, w);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = pack_bf16(__uint_as_float(r[16 * g + 2 * i]) * bf16_lo(w[i]),
                         __uint_as_float(r[16 * g + 2 * i + 1]) * bf16_hi(w[i]));
      st_global32(e.C + elem + 16 * g, o);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t oa[8], od[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c8 = 16 * g + 8 * hh;
        // bias from shared memory (broadcast reads): read from global here, the loads sat on the long scoreboard
        // in front of every chunk (30 % of the kernel's stall samples)
        const float4 b0 = lds_f4(sbias + col + c8);
        const float4 b1 = lds_f4(sbias + col + c8 + 4);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        // same stream and granule index as K7 (bias_gelu_dropout): vector (row * N + col) / 8 of the flattened
        // tensor, element 2j <- low half of word j, element 2j+1 <- high half (keep8).  The 16-bit compares are
        // done in place: (w << 16) >= (t << 16) and w >= (t << 16), no mask word is assembled.
        uint4 rnd = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (e.thresh) rnd = philox4x32((unsigned long long)((elem + c8) >> 3), e.offset, e.seed);
        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        const uint32_t t16 = e.thresh << 16;
        float a[8], dv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = __uint_as_float(r[c8 + i]) + b[i];
          float cdf, pdf;
          gelu_cdf_pdf(x, cdf, pdf);
          const bool keep = ((i & 1) ? rw[i >> 1] : (rw[i >> 1] << 16)) >= t16;
          const float m = keep ? e.scale : 0.f;
          a[i] = x * cdf * m;
          dv[i] = fmaf(x, pdf, cdf) * m;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          oa[4 * hh + i] = pack_bf16(a[2 * i], a[2 * i + 1]);
          od[4 * hh + i] = pack_bf16(dv[2 * i], dv[2 * i + 1]);
        }
      }
      st_global32(e.C + elem + 16 * g, oa);
      if (e.C2) st_global32(e.C2 + elem + 16 * g, od);
    }
  }
}

// One epilogue warp's share of a tile: wait for the accumulator, read kHalf columns of this thread's row from TMEM
// (32 at a time), apply the epilogue, store.  taddr = TMEM address of the warp's lane quarter and first column.
template <int EPI, int kHalf>
HW_DEV void epilogue_tile(const EpiArgs& e, uint64_t* acc_full_bar, uint32_t parity, uint32_t taddr, size_t elem0, int col0,
                          const float* sbias) {
  if constexpr (EPI == kEpiMul) {
    // The factor rows are fetched BEFORE the accumulator is waited for: a row-per-thread load touches 32 lines
    // per instruction, and issued inside the chunk loop its latency was exposed once per chunk (tensor pipe
    // 39 % busy on a GEMM whose HBM time and MMA time are both ~half of what it took).
    uint32_t gq[kHalf / 16][8];
#pragma unroll
    for (int j = 0; j < kHalf / 16; ++j) ld_global32(e.G + elem0 + 16 * j, gq[j]);
    mbar_wait(acc_full_bar, parity);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < kHalf; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = pack_bf16(__uint_as_float(r[16 * g + 2 * i]) * bf16_lo(gq[c / 16 + g][i]),
                           __uint_as_float(r[16 * g + 2 * i + 1]) * bf16_hi(gq[c / 16 + g][i]));
        st_global32(e.C + elem0 + c + 16 * g, o);
      }
    }
  } else {
    mbar_wait(acc_full_bar, parity);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < kHalf; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
      epilogue_chunk<EPI>(r, e, elem0 + c, col0 + c, sbias);
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(EpiWarps<EPI>::kThreads, 1) gemm_nt_epi_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const EpiArgs e, int M, int N, int K) {
  using Cfg = FfnCfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / BN, m_blocks = M / kFM, tiles = n_blocks * m_blocks, nk = K / kFK;
  if (EPI == kEpiGelu || EPI == kEpiBias || EPI == kEpiGeluEval)
    for (int i = threadIdx.x; i < N; i += blockDim.x) sbias[i] = e.bias ? e.bias[i] : 0.f;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EpiWarps<EPI>::kWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Producer and issuer run their loops with the whole warp and issue under elect_one_sync (tc.cuh): inside
  // `if (lane == 0)` every UTMALDG / UTCHMMA was wrapped in ~20 instructions of uniform-register shuffling, and four
  // MMAs per k block took longer to issue (~720 clocks) than to execute (512).
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one_sync()) {
          unsigned char* st = smem + s * Cfg::kStage;
          mbar_expect_tx(&full[s], Cfg::kStage);
          tma_load_2d(st, &tmA, &full[s], kb * kFK, mb * kFM);
          tma_load_2d(st + Cfg::kABytes, &tmB, &full[s], kb * kFK, nb * BN);
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(kFM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + Cfg::kABytes;
#pragma unroll
          for (int ks = 0; ks < kFK / 16; ++ks)
            umma_bf16(tmem + buf * BN, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                      (kb | ks) != 0);
          umma_commit(&empty[s]);
          if (kb == nk - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // column slice of the tile
    constexpr int kHalf = BN / (EpiWarps<EPI>::kWarps / 4);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int buf = it & 1;
      const size_t row = (size_t)mb * kFM + q * 32 + lane;
      const int col0 = nb * BN + half * kHalf;
      const size_t elem0 = row * N + col0;
      epilogue_tile<EPI, kHalf>(e, acc_full + buf, (it >> 1) & 1, tmem + ((uint32_t)(q * 32) << 16) + buf * BN + half * kHalf, elem0, col0, sbias);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

template <int BN, int EPI>
static int launch_epi(const bf16* A, const bf16* Bt, const EpiArgs& e, int M, int N, int K, cudaStream_t s) {
  using Cfg = FfnCfg<BN>;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(gemm_nt_epi_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, kFM))) return st;
  if ((st = make_tmap_2d(&tmB, Bt, (uint64_t)N, (uint64_t)K, BN))) return st;
  const int tiles = (M / kFM) * (N / BN);
  const int grid = tiles < 148 ? tiles : 148;
  gemm_nt_epi_kernel<BN, EPI><<<grid, EpiWarps<EPI>::kThreads, Cfg::kSmem, s>>>(tmA, tmB, e, M, N, K);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one cluster of two CTAs (the two SMs of a TPC) computes a 256 x 256 tile.
// A single-CTA M=128, N=256 SS MMA reads 12 KB of operands from shared memory per 128 clocks while TMA writes
// 48 KB per 512: 190 B/clk against the 128 B/clk/SM shared-memory port, which is why the single-CTA kernel's
// tensor pipe stops at ~73 %.  In a pair each CTA stages its own 128 rows of A and only HALF of B's 256 rows
// (32 KB per k block instead of 48) and its tensor core reads the other half from the peer: 128 B/clk.
//   leader (cluster rank 0): `full[s]` counts the TMA bytes of BOTH CTAs; its warp 1 issues the MMAs and commits
//   with a multicast arrive to `empty[s]` / `acc_full[b]` of both CTAs; `acc_empty[b]` (leader) collects the
//   arrivals of both CTAs' epilogue warps.  The peer's warp 1 only owns its TMEM allocation.
// ---------------------------------------------------------------------------
struct PairCfg {
  static constexpr int kBN = 256;
  static constexpr int kStages = 6;
  static constexpr int kABytes = kFM * kFK * 2;          // this CTA's 128 rows of A
  static constexpr int kBBytes = (kBN / 2) * kFK * 2;    // this CTA's 128 of B's 256 rows
  static constexpr int kStage = kABytes + kBBytes;       // 32 KB
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kBiasOff = kBarOff + 256;
  static constexpr int kSmem = kBiasOff + kMaxBiasN * 4 + 1024;
  static constexpr int kTmemCols = 2 * kBN;
};

template <int EPI>
__global__ void __launch_bounds__(EpiWarps<EPI>::kThreads, 1) gemm_nt_epi_pair_kernel(
    const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiArgs e, int M, int N, int K) {
  using Cfg = PairCfg;
  constexpr int BN = Cfg::kBN;
  extern __shared__ unsigned char smem_raw[];
  // the dynamic shared-memory window starts at the same offset in both CTAs, so the aligned offsets agree
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_blocks = N / BN, m_blocks = (M + 2 * kFM - 1) / (2 * kFM), tiles = n_blocks * m_blocks, nk = K / kFK;
  if (EPI == kEpiGelu || EPI == kEpiBias || EPI == kEpiGeluEval)
    for (int i = threadIdx.x; i < N; i += blockDim.x) sbias[i] = e.bias ? e.bias[i] : 0.f;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * EpiWarps<EPI>::kWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before either touches the other's
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {   // whole warp, issue under elect_one_sync (see gemm_nt_epi_kernel)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = pair; tile < tiles; tile += npairs) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait_cluster(&empty[s], ph ^ 1);
        if (elect_one_sync()) {
          unsigned char* st = smem + s * Cfg::kStage;
          const uint32_t lead_full = mapa_shared(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::kStage);
          tma_load_2d_pair(st, &tmA, lead_full, kb * kFK, (2 * mb + (int)rank) * kFM);     // rows past M: zero fill
          tma_load_2d_pair(st + Cfg::kABytes, &tmB, lead_full, kb * kFK, nb * BN + (int)rank * (BN / 2));
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kFM, BN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < tiles; tile += npairs, ++it) {
        const int buf = it & 1;
        mbar_wait_cluster(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait_cluster(&full[s], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + Cfg::kABytes;
#pragma unroll
            for (int ks = 0; ks < kFK / 16; ++ks)
              umma_bf16_pair(tmem + buf * BN, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                             (kb | ks) != 0);
            umma_commit_pair(&empty[s]);
            if (kb == nk - 1) umma_commit_pair(&acc_full[buf]);
          }
          __syncwarp();
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kHalf = BN / (EpiWarps<EPI>::kWarps / 4);
    int it = 0;
    for (int tile = pair; tile < tiles; tile += npairs, ++it) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int buf = it & 1;
      const size_t row = (size_t)(2 * mb + (int)rank) * kFM + q * 32 + lane;
      const int col0 = nb * BN + half * kHalf;
      if (row < (size_t)M) {
        epilogue_tile<EPI, kHalf>(e, acc_full + buf, (it >> 1) & 1,
                                  tmem + ((uint32_t)(q * 32) << 16) + buf * BN + half * kHalf, row * N + col0, col0, sbias);
      } else {   // odd number of 128-row blocks: the peer's half of the last tile is padding (warp-uniform)
        mbar_wait(acc_full + buf, (it >> 1) & 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory: neither CTA may exit early
  if (warp == 1) tmem_dealloc_pair(tmem, Cfg::kTmemCols);
}

static int g_gemm_pair = -1;
bool gemm_pair_enabled() {
  if (g_gemm_pair < 0) {
    const char* s = getenv("HWGAT_GEMM_PAIR");   // A/B measurements only; default on
    g_gemm_pair = (s && s[0] == '0') ? 0 : 1;
  }
  return g_gemm_pair != 0;
}
bool set_gemm_pair(bool on) {
  const bool prev = gemm_pair_enabled();
  g_gemm_pair = on ? 1 : 0;
  return prev;
}

template <int EPI>
static int launch_epi_pair(const bf16* A, const bf16* Bt, const EpiArgs& e, int M, int N, int K, cudaStream_t s) {
  using Cfg = PairCfg;
  static PerDeviceOnce once;
  static std::atomic<int> max_clusters_s{0};
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(EpiWarps<EPI>::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  once.run([&] {
    cudaFuncSetAttribute(gemm_nt_epi_pair_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    cfg.gridDim = dim3(148);
    int mc = 0;
    if (cudaOccupancyMaxActiveClusters(&mc, gemm_nt_epi_pair_kernel<EPI>, &cfg) != cudaSuccess || mc < 1) {
      cudaGetLastError();
      mc = 74;
    }
    max_clusters_s.store(mc > 74 ? 74 : mc);
  });
  const int max_clusters = max_clusters_s.load() > 0 ? max_clusters_s.load() : 74;
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, kFM))) return st;
  if ((st = make_tmap_2d(&tmB, Bt, (uint64_t)N, (uint64_t)K, Cfg::kBN / 2))) return st;
  const int tiles = ((M + 2 * kFM - 1) / (2 * kFM)) * (N / Cfg::kBN);
  const int pairs = tiles < max_clusters ? tiles : max_clusters;
  cfg.gridDim = dim3(2 * pairs);
  cudaError_t err = cudaLaunchKernelEx(&cfg, gemm_nt_epi_pair_kernel<EPI>, tmA, tmB, e, M, N, K);
  count_launch();
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

// C[M,N] = epilogue(A[M,K] . Bt[N,K]^T); M % 128 == 0, N % 128 == 0, K % 64 == 0
template <int EPI>
static int gemm_nt_epi(const bf16* A, const bf16* Bt, const EpiArgs& e, long long M, int N, int K, cudaStream_t s) {
  if (M % kFM || K % kFK || N % 128 || M > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  if (N % 256 == 0) {
    // CTA pairs (256 x 256 tiles) pay off once the k loop is long enough to be tensor-bound (measured: +5-8 % at
    // K >= 768, -5 % at K <= 512 where the GEMM is bound by its HBM writes and the pair only adds hand-shakes)
    if (gemm_pair_enabled() && M >= 2 * kFM && K >= 768) return launch_epi_pair<EPI>(A, Bt, e, (int)M, N, K, s);
    return launch_epi<256, EPI>(A, Bt, e, (int)M, N, K, s);
  }
  return launch_epi<128, EPI>(A, Bt, e, (int)M, N, K, s);
}

int gemm_tc_nt_epi_none(const bf16* A, const bf16* Bt, bf16* C, long long M, int N, int K, cudaStream_t s) {
  EpiArgs e{};
  e.C = C;
  return gemm_nt_epi<kEpiNone>(A, Bt, e, M, N, K, s);
}

// C = A . Bt^T + bias (fp32 bias add before the bf16 rounding): the QKV projection of the general-window attention
// path (attn_core_tc2.cu).  N <= kMaxBiasN.
int gemm_tc_nt_epi_bias(const bf16* A, const bf16* Bt, const float* bias, bf16* C, long long M, int N, int K,
                        cudaStream_t s) {
  if (N > kMaxBiasN) return HWGAT_ERR_UNSUPPORTED;
  EpiArgs e{};
  e.C = C; e.bias = bias;
  return gemm_nt_epi<kEpiBias>(A, Bt, e, M, N, K, s);
}

// ---------------------------------------------------------------------------
// K10 entry points
// ---------------------------------------------------------------------------
int ffn_fwd(const bf16* h, const bf16* w1, const float* b1, const bf16* w2, bf16* act, bf16* gp, bf16* v0, long long n,
            int d, int hidden, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  EpiArgs e{};
  e.C = act; e.C2 = gp; e.bias = b1; e.thresh = drop_threshold16(p); e.scale = drop_scale16(e.thresh);
  e.seed = seed; e.offset = offset;
  if (hidden > kMaxBiasN) return HWGAT_ERR_UNSUPPORTED;
  // inference (no local derivative wanted, no dropout): the GELU-only epilogue
  int st = (!gp && e.thresh == 0) ? gemm_nt_epi<kEpiGeluEval>(h, w1, e, n, hidden, d, s)
                                  : gemm_nt_epi<kEpiGelu>(h, w1, e, n, hidden, d, s);          // act, gp  [n, hidden]
  if (st) return st;
  EpiArgs e2{};
  e2.C = v0;
  return gemm_nt_epi<kEpiNone>(act, w2, e2, n, d, hidden, s);         // v0 [n, d]
}

size_t ffn_bwd_workspace_bytes(long long n, int d, int hidden) {
  // du0 [n, hidden] + W2^T [hidden, d] + W1^T [d, hidden] (bf16) + db2 scratch [d] (fp32)
  return ((size_t)n * hidden + (size_t)2 * d * hidden) * sizeof(bf16) + (size_t)d * sizeof(float);
}

int ffn_bwd(const bf16* dv0, const bf16* h, const bf16* act, const bf16* gp, const bf16* w1, const bf16* w2, bf16* dh,
            float* dw1, float* db1, float* dw2, void* workspace, long long n, int d, int hidden, cudaStream_t s) {
  bf16* du0 = (bf16*)workspace;
  bf16* w2t = du0 + (size_t)n * hidden;       // [hidden, d]
  bf16* w1t = w2t + (size_t)d * hidden;       // [d, hidden]
  int st;
  if ((st = transpose_bf16(w2, w2t, d, hidden, s))) return st;          // W2 is [d, hidden]
  EpiArgs e{};
  e.C = du0; e.G = gp;
  if ((st = gemm_nt_epi<kEpiMul>(dv0, w2t, e, n, hidden, d, s))) return st;    // du0 = (dv0 . W2) o gp
  if ((st = gemm_tc_tn(dv0, act, dw2, nullptr, d, hidden, n, s))) return st;   // dW2 [d, hidden] = dv0^T . act (db2 comes from K6')
  if ((st = gemm_tc_tn(du0, h, dw1, db1, hidden, d, n, s))) return st;         // dW1 [hidden, d], db1 = colsum(du0)
  if ((st = transpose_bf16(w1, w1t, hidden, d, s))) return st;          // W1 is [hidden, d]
  EpiArgs e2{};
  e2.C = dh;
  return gemm_nt_epi<kEpiNone>(du0, w1t, e2, n, d, hidden, s);                 // dh = du0 . W1
}

}  // namespace hwgat
