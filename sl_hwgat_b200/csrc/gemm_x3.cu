// fp32 Linear layers on the bf16 tensor cores: the "x3" mode of the fp32 path (hwgat_set_fp32_mode).
//
// The reference's own loop runs the model in fp32 without autocast (utils.py:102, 128; inference.py:95), so a drop-in
// lands on the fp32 kernels.  A true-fp32 FFMA GEMM (attn_f32.cu, the 1e-5 parity mode) is bound by the fp32 pipe at a
// few tens of TFLOP/s.  Here every fp32 operand is split into three bf16 planes
//     x = hi + mid + lo,   hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)      (3 x 8 = 24 mantissa bits)
// and the product is formed from the six partial products whose weight is >= 2^-16,
//     x.w = hi.hi + (hi.mid + mid.hi) + (mid.mid + hi.lo + lo.hi) + O(2^-24),
// each an exact bf16 x bf16 -> fp32 tcgen05.mma.  The tensor core ADDS into its fp32 accumulator with truncation
// (measured, tools/exp_split3.py: the relative error grows linearly with the number of K = 16 steps that hit one
// accumulator, ~2e-8 per step), so
//   * the leading product hi.hi and the five corrections go to SEPARATE TMEM accumulators (the corrections are 2^-8 of
//     the sum: truncating them costs nothing, and they do not lengthen the chain of the leading accumulator), added in
//     fp32 in the epilogue;
//   * the weight-gradient GEMMs, whose contraction runs over the tokens, flush the leading accumulator to the fp32
//     output (red.global.add, round-to-nearest) every kMainKBlocks k blocks.
// Measured against fp64 (tests/test_gpu_x3.py -s): 0.9e-7 (K = 128) ... 9.5e-7 (K = 1024), 1.5e-6 (K = 1536) relative
// L2 - what is left is the truncation of the K / 16 steps of the leading chain - next to 4-8e-7 for cuBLAS' fp32 GEMM.
// Whole model, eval logits against the fp64 oracle: 7e-7 (FFMA mode and eager PyTorch fp32: 3e-7).
//
//   y  = x . W^T + b        gemm_nt_x3_kernel: TMA -> 6 planes per stage -> 24 tcgen05.mma per 64-wide k block -> TMEM
//   dx = dy . W             the same kernel against the split of W^T
//   dW = dy^T . x           gemm_tn_x3_kernel: MN-major operands straight from the row-major planes, token split
//   db = column sums of dy  colsum_f32 (fp32)
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

int g_fp32_mode = 0;   // 0: FFMA parity kernels, 1: x3 (this file) wherever the shape allows

// ---- split kernels ----------------------------------------------------------------------------------------------
HW_DEV void split3(float x, bf16& h, bf16& m, bf16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);          // exact (Sterbenz-like: the difference fits 16 bits)
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// planes[p][i] for p = hi, mid, lo; 8 elements per thread (two 16-byte loads, three 16-byte stores)
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, bf16* __restrict__ planes,
                                                     long long count8, long long plane_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count8) return;
  const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __align__(16) bf16 h[8], m[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split3(v[j], h[j], m[j], l[j]);
  reinterpret_cast<int4*>(planes)[i] = *reinterpret_cast<const int4*>(h);
  reinterpret_cast<int4*>(planes + plane_stride)[i] = *reinterpret_cast<const int4*>(m);
  reinterpret_cast<int4*>(planes + 2 * plane_stride)[i] = *reinterpret_cast<const int4*>(l);
}

// planes[p][c][r] = split(w[r][c]): the transposed split of a (small) weight matrix [R][C]
__global__ void __launch_bounds__(256) split3_t_kernel(const float* __restrict__ w, bf16* __restrict__ planes, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? w[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)R * C;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      bf16 h, m, l;
      split3(tile[threadIdx.x][j], h, m, l);
      const size_t o = (size_t)c * R + r;
      planes[o] = h; planes[plane + o] = m; planes[2 * plane + o] = l;
    }
  }
}

static int split_rows(const float* x, bf16* planes, long long count, cudaStream_t s) {
  const long long c8 = count / 8;
  split3_kernel<<<(unsigned)((c8 + 255) / 256), 256, 0, s>>>(x, planes, c8, count);
  count_launch();
  return (int)cudaGetLastError();
}
static int split_transposed(const float* w, bf16* planes, int R, int C, cudaStream_t s) {
  split3_t_kernel<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(w, planes, R, C);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- NT GEMM over split planes ------------------------------------------------------------------------------------
// C[M,N] (fp32) = sum over the six plane pairs of A_p[M,K] . B_q[N,K]^T (+ bias).  A3 = [3][M][K], B3 = [3][N][K] bf16.
// Persistent, warp-specialised like gemm_nt_epi_kernel (ffn_tc.cu): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9
// epilogue.  A stage holds the six [128 x 64] planes of one k block (96 KB), two stages; two TMEM buffers of
// (leading | corrections) x 128 columns, so the epilogue of a tile overlaps the MMAs of the next.
struct X3Cfg {
  static constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 2;
  static constexpr int kPlane = 128 * kBK * 2;            // 16 KB
  static constexpr int kStage = 6 * kPlane;               // A hi, mid, lo | B hi, mid, lo
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kBiasOff = kBarOff + 256;
  static constexpr int kMaxBiasN = 2048;
  static constexpr int kSmem = kBiasOff + kMaxBiasN * 4 + 1024;
  static constexpr int kTmemCols = 512;                   // 2 buffers x (128 leading + 128 corrections)
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 32 * (2 + kEpiWarps);
};

__global__ void __launch_bounds__(X3Cfg::kThreads, 1) gemm_nt_x3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmB,
                                                                       float* __restrict__ C, const float* __restrict__ bias,
                                                                       int M, int N, int K) {
  using Cfg = X3Cfg;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / Cfg::kBN, m_blocks = M / Cfg::kBM, tiles = n_blocks * m_blocks, nk = K / Cfg::kBK;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sbias[i] = bias ? bias[i] : 0.f;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Cfg::kEpiWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

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
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            tma_load_2d(st + p * Cfg::kPlane, &tmA, &full[s], kb * Cfg::kBK, p * M + mb * Cfg::kBM);
            tma_load_2d(st + (3 + p) * Cfg::kPlane, &tmB, &full[s], kb * Cfg::kBK, p * N + nb * Cfg::kBN);
          }
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(Cfg::kBM, Cfg::kBN);
    // (A plane, B plane) of the five corrections, smallest first; the leading product hi.hi has its own accumulator
    constexpr int kPa[5] = {2, 0, 1, 1, 0};
    constexpr int kPb[5] = {0, 2, 1, 0, 1};
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t t_main = tmem + buf * 256, t_corr = t_main + 128;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + 3 * Cfg::kPlane;
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int ks = 0; ks < Cfg::kBK / 16; ++ks)
              umma_bf16(t_corr, umma_desc_k_sw128(sa + kPa[c] * Cfg::kPlane + ks * 32),
                        umma_desc_k_sw128(sb + kPb[c] * Cfg::kPlane + ks * 32), idesc, (kb | ks | c) != 0);
#pragma unroll
          for (int ks = 0; ks < Cfg::kBK / 16; ++ks)
            umma_bf16(t_main, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc, (kb | ks) != 0);
          umma_commit(&empty[s]);
          if (kb == nk - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // 64-column slice of the tile
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int buf = it & 1;
      const size_t row = (size_t)mb * Cfg::kBM + q * 32 + lane;
      const int col0 = nb * Cfg::kBN + half * 64;
      float* crow = C + row * N + col0;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + half * 64;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32], r2[32];
        tmem_ld32(taddr + c, r);
        tmem_ld32(taddr + 128 + c, r2);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = __float_as_uint((__uint_as_float(r[8 * g + i]) + __uint_as_float(r2[8 * g + i])) + sbias[col0 + c + 8 * g + i]);
          st_global32(crow + c + 8 * g, o);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

static int launch_nt_x3(const bf16* A3, const bf16* B3, float* C, const float* bias, long long M, int N, int K, cudaStream_t s) {
  using Cfg = X3Cfg;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(gemm_nt_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A3, (uint64_t)(3 * M), (uint64_t)K, Cfg::kBM))) return st;
  if ((st = make_tmap_2d(&tmB, B3, (uint64_t)(3 * N), (uint64_t)K, Cfg::kBN))) return st;
  const long long tiles = (M / Cfg::kBM) * (N / Cfg::kBN);
  const int grid = tiles < 148 ? (int)tiles : 148;
  gemm_nt_x3_kernel<<<grid, Cfg::kThreads, Cfg::kSmem, s>>>(tmA, tmB, C, bias, (int)M, N, K);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- TN GEMM over split planes (weight gradients) -------------------------------------------------------------------
// C[M,N] (fp32) += sum over the six plane pairs of A_p[Kd,M]^T . B_q[Kd,N]: A3 = [3][Kd][M], B3 = [3][Kd][N] bf16, both
// MN-major UMMA operands straight from the row-major planes (as gemm_tc_tn, gemm_tc.cu).  The contraction runs over
// the tokens: grid = (128 x 128 tiles, token splits of kMainKBlocks k blocks), every CTA adds its partial to the zeroed
// output with red.global.add - the leading accumulator never sees more than kMainKBlocks * 4 truncating steps.
struct X3TnCfg {
  static constexpr int kStages = 2;
  static constexpr int kPlane = 64 * 128 * 2;             // [64 k][128 mn]: two 8 KB boxes
  static constexpr int kStage = 6 * kPlane;
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kSmem = kBarOff + 256 + 1024;
  static constexpr int kTmemCols = 512;                   // two sets of (128 leading + 128 corrections)
};
constexpr int kMainKBlocks = 16;   // 1024 tokens

// Persistent: a work item = (128 x 128 tile, token chain); the CTAs walk the items with a grid stride, two accumulator
// sets in TMEM so that the red.global.add epilogue of one chain runs under the MMAs of the next (one CTA per chain - the
// first version - spent a third of its time in set-up, pipeline fill and the epilogue: tensor pipe 48 % busy).
__global__ void __launch_bounds__(192, 1) gemm_tn_x3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            float* __restrict__ C, int N, int kd_rows, int k_blocks_total,
                                                            int tiles, int items) {
  using Cfg = X3TnCfg;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // item -> (tile, chain): consecutive items are the tiles of ONE token chain, so the CTAs running at the same time
  // read the same token rows of the planes (L2 hits across tiles)
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int chain = item / tiles, tile = item - chain * tiles;
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int kb0 = chain * kMainKBlocks;
      const int kb1 = kb0 + kMainKBlocks < k_blocks_total ? kb0 + kMainKBlocks : k_blocks_total;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one_sync()) {
          unsigned char* st = smem + s * Cfg::kStage;
          mbar_expect_tx(&full[s], Cfg::kStage);
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_2d(st + p * Cfg::kPlane + j * 8192, &tmA, &full[s], mb * 128 + j * 64, p * kd_rows + kb * 64);
              tma_load_2d(st + (3 + p) * Cfg::kPlane + j * 8192, &tmB, &full[s], nb * 128 + j * 64, p * kd_rows + kb * 64);
            }
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, true, true);
    constexpr int kPa[5] = {2, 0, 1, 1, 0};
    constexpr int kPb[5] = {0, 2, 1, 0, 1};
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int chain = item / tiles;
      const int kb0 = chain * kMainKBlocks;
      const int kb1 = kb0 + kMainKBlocks < k_blocks_total ? kb0 + kMainKBlocks : k_blocks_total;
      const int buf = it & 1;
      const uint32_t t_main = tmem + buf * 256, t_corr = t_main + 128;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + 3 * Cfg::kPlane;
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)   // 16 tokens per step = two 8-row atoms = 2 KB
              umma_bf16(t_corr, umma_desc_mn_sw128(sa + kPa[c] * Cfg::kPlane + ks * 2048, 8192, 1024),
                        umma_desc_mn_sw128(sb + kPb[c] * Cfg::kPlane + ks * 2048, 8192, 1024), idesc, (kb > kb0) | (ks | c));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t_main, umma_desc_mn_sw128(sa + ks * 2048, 8192, 1024), umma_desc_mn_sw128(sb + ks * 2048, 8192, 1024),
                      idesc, (kb > kb0) | ks);
          umma_commit(&empty[s]);
          if (kb == kb1 - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int chain = item / tiles, tile = item - chain * tiles;
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      float* crow = C + (size_t)(mb * 128 + q * 32 + lane) * N + (size_t)nb * 128;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * 256;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32], r2[32];
        tmem_ld32(taddr + c, r);
        tmem_ld32(taddr + 128 + c, r2);
        tmem_ld_wait();
        if (c == 96) {                       // last read of this accumulator set: hand it back before the atomics
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(crow + c + i, __uint_as_float(r[i]) + __uint_as_float(r2[i]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

static int launch_tn_x3(const bf16* A3, const bf16* B3, float* C, int M, int N, long long Kd, cudaStream_t s) {
  using Cfg = X3TnCfg;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(gemm_tn_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A3, (uint64_t)(3 * Kd), (uint64_t)M, 64))) return st;
  if ((st = make_tmap_2d(&tmB, B3, (uint64_t)(3 * Kd), (uint64_t)N, 64))) return st;
  const int tiles = (M / 128) * (N / 128);
  const int kblocks = (int)(Kd / 64);
  const long long chains = (kblocks + kMainKBlocks - 1) / kMainKBlocks;
  const long long items = chains * tiles;
  if (items > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, s);
  const int grid = items < 148 ? (int)items : 148;
  gemm_tn_x3_kernel<<<grid, 192, Cfg::kSmem, s>>>(tmA, tmB, C, N, (int)Kd, kblocks, tiles, (int)items);
  count_launch();
  return (int)cudaGetLastError();
}

// ---- the three GEMMs of a Linear layer ------------------------------------------------------------------------------
bool x3_enabled() { return __atomic_load_n(&g_fp32_mode, __ATOMIC_RELAXED) == 1 && !deterministic(); }
bool x3_supported(long long n, int d_in, int d_out) {
  return n >= 128 && n % 128 == 0 && n * 3 < 0x7fffffffLL && d_in % 128 == 0 &&
         d_out % 128 == 0 && d_in <= 4096 && d_out <= X3Cfg::kMaxBiasN;
}

struct Scratch {     // stream-ordered scratch for the bf16 planes
  void* p = nullptr;
  cudaStream_t s;
  explicit Scratch(cudaStream_t s_) : s(s_) {}
  int get(size_t bytes) {
    // keep freed scratch in the device's default pool across synchronisations: with the default release threshold
    // (0) every blocking read-back of a result hands the pages back to the driver and the next call maps them again
    // (eval forward with a per-step D2H of the predictions: 78 -> 145 ms per step)
    static PerDeviceOnce once;
    once.run([] {
      int dev = 0;
      cudaMemPool_t pool;
      cudaGetDevice(&dev);
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    });
    if (cudaMallocAsync(&p, bytes, s) != cudaSuccess) { cudaGetLastError(); p = nullptr; return HWGAT_ERR_WORKSPACE; }
    return 0;
  }
  ~Scratch() { if (p) cudaFreeAsync(p, s); }
};

// x_planes_out (optional): caller-owned [3][n][d_in] bf16 that receives the split of x, to be handed to linear_x3_bwd
// (the weight gradient needs the same planes: no second split, and the fp32 x need not be kept)
int linear_x3_fwd(const float* x, const float* w, const float* bias, float* y, long long n, int d_in, int d_out,
                  cudaStream_t s, bf16* x_planes_out) {
  Scratch sc(s);
  int st;
  const size_t xe = (size_t)n * d_in, we = (size_t)d_out * d_in;
  if ((st = sc.get(((x_planes_out ? 0 : 3 * xe) + 3 * we) * sizeof(bf16)))) return st;
  bf16* w3 = (bf16*)sc.p;
  bf16* x3 = x_planes_out ? x_planes_out : w3 + 3 * we;
  if ((st = split_rows(x, x3, (long long)xe, s))) return st;
  if ((st = split_rows(w, w3, (long long)we, s))) return st;
  return launch_nt_x3(x3, w3, y, bias, n, d_out, d_in, s);
}

// dx[n, d_in] = dy . W ; dw[d_out, d_in] = dy^T . x ; either may be NULL (db: the caller's fp32 column sums).
// x_planes (optional): the split of x kept from linear_x3_fwd; then x itself is not read.
int linear_x3_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, long long n, int d_in, int d_out,
                  cudaStream_t s, const bf16* x_planes) {
  Scratch sc(s);
  int st;
  const size_t ye = (size_t)n * d_out, xe = (dw && !x_planes) ? (size_t)n * d_in : 0, we = dx ? (size_t)d_out * d_in : 0;
  if ((st = sc.get((3 * ye + 3 * xe + 3 * we) * sizeof(bf16)))) return st;
  bf16* y3 = (bf16*)sc.p;
  bf16* x3 = y3 + 3 * ye;
  bf16* wt3 = x3 + 3 * xe;
  if ((st = split_rows(dy, y3, (long long)ye, s))) return st;
  if (dx) {
    if ((st = split_transposed(w, wt3, d_out, d_in, s))) return st;       // [3][d_in][d_out]
    if ((st = launch_nt_x3(y3, wt3, dx, nullptr, n, d_in, d_out, s))) return st;
  }
  if (dw) {
    if (!x_planes && (st = split_rows(x, x3, (long long)xe, s))) return st;
    if ((st = launch_tn_x3(y3, x_planes ? x_planes : x3, dw, d_out, d_in, n, s))) return st;
  }
  return 0;
}

}  // namespace hwgat
