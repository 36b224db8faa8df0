// Tensor maps, the bf16 transpose and the weight-gradient GEMM (gemm_tc_tn) shared by K3 (the autograd of
// self.qkv, HWGATE.py:86) and K10 (the FeedForward, ffn_tc.cu).  bf16 operands, fp32 accumulation in TMEM,
// operands staged by TMA into SWIZZLE_128B smem.  The C = A . Bt^T GEMM lives in ffn_tc.cu (gemm_nt_epi_kernel).
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * sizeof(bf16)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

int make_tmap_4d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                 uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[4] = {d0, d1, d2, d3};
  cuuint64_t gstr[3] = {d0 * sizeof(bf16), d0 * d1 * sizeof(bf16), d0 * d1 * d2 * sizeof(bf16)};
  cuuint32_t box[4] = {64, box1, box2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// bf16 transpose [R][Cc] -> [Cc][R] (Wqkv -> Wqkv^T, 3d x d: tiny)
// perm64: the input-row index r (= output column) is mapped through the dQKV column permutation of K3
// (attn_tc.cu, store_rows_16x64_perm): inside each block of 64, 8 nt + 2 t + b -> 16 t + 2 nt + b.
HW_DEV int perm64_fwd(int r) {
  const int p = r & 63, nt = p >> 3, t = (p >> 1) & 3, b = p & 1;
  return (r & ~63) + 16 * t + 2 * nt + b;
}
HW_DEV int perm64_inv(int r) {
  const int p = r & 63, t = p >> 4, nt = (p >> 1) & 7, b = p & 1;
  return (r & ~63) + 8 * nt + 2 * t + b;
}
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int Cc, bool perm64,
                                      int perm_limit) {
  __shared__ bf16 t[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < R && c < Cc) t[i][threadIdx.x] = in[(size_t)(r0 + i) * Cc + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < Cc && r < R) out[(size_t)(c0 + i) * R + (perm64 && r < perm_limit ? perm64_fwd(r) : r)] = t[threadIdx.x][i];
}

int transpose_bf16(const bf16* in, bf16* out, int R, int Cc, cudaStream_t s, bool perm64, int perm_limit) {
  if (perm64 && R % 64) return HWGAT_ERR_UNSUPPORTED;
  transpose_bf16_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(in, out, R, Cc, perm64, perm_limit);
  count_launch();
  return (int)cudaGetLastError();
}


// ---------------------------------------------------------------------------
// gemm_tc_tn : C[M,N] (fp32) += A[Kd,M]^T . B[Kd,N],  colsum[M] += sum_k A[k,m]
//   d_w = dQKV^T . xn  and  d_b = column sums of dQKV   (autograd of self.qkv, HWGATE.py:86)
// A and B are row-major with the contraction index (tokens) as the row, i.e. both are
// MN-major UMMA operands: a TMA box of [64 tokens x 64 columns] lands as 64 k-rows of 128
// swizzled bytes = eight SWIZZLE_128B MN-major atoms (LBO = next 64 columns = 8 KB, SBO = next 8
// tokens = 1 KB).  Split over tokens: each CTA accumulates its token range in TMEM and adds it
// to the zeroed fp32 output with red.global.add.  d_b comes from one extra N=16 MMA per k step
// against an all-ones B tile.
// ---------------------------------------------------------------------------
constexpr int kSumChunk = 8;   // k blocks per column-sum chunk (see gemm_tc_tn_kernel)

template <int BN>
struct GemmTnCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = 64 * 128 * 2;         // [64 k][128 m] : 2 boxes of 8 KB
  static constexpr int kBBytes = 64 * BN * 2;          // [64 k][BN n]  : BN/64 boxes of 8 KB
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kOnesOff = kStages * kStage;    // 8 KB of bf16 1.0
  static constexpr int kBarOff = kOnesOff + 8192;
  static constexpr int kSmem = kBarOff + 256 + 1024;
  static constexpr int kTmemCols = 512;                // accumulator BN columns + 16 for the column sums
  static constexpr int kOnesCol = BN;                  // TMEM column of the ones product
};

// MMA issue loop over k blocks [kb, kend).  SUM: also the column-sum MMA (N = 16 against the all-ones tile), which
// re-reads the whole A slice and costs ~1/4 of the main MMA.  The loop exists in both forms because a predicated-off
// UTCHMMA is not free (measured: a no-sum GEMM with `@!p UTCHMMA` in its loop ran 15 % slower than with a branch).
// main_first / sum_first: k block whose first MMA overwrites the accumulator instead of adding to it.
template <int BN, bool SUM>
HW_DEV void tn_issue(unsigned char* smem, uint64_t* full, uint64_t* empty, uint32_t tmem, int kb, int kend, int main_first,
                     int sum_first, int& s, uint32_t& ph) {
  using Cfg = GemmTnCfg<BN>;
  constexpr uint32_t idesc = umma_idesc_bf16(128, BN, true, true);
  constexpr uint32_t idesc1 = umma_idesc_bf16(128, 16, true, true);
  const uint32_t sones = smem_u32(smem + Cfg::kOnesOff);
  for (; kb < kend; ++kb) {   // whole warp; MMAs and commits from one elected lane
    mbar_wait(&full[s], ph);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + Cfg::kABytes;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 tokens per step = two 8-row atoms = 2 KB
        const uint64_t da = umma_desc_mn_sw128(sa + ks * 2048, 8192, 1024);
        umma_bf16(tmem, da, umma_desc_mn_sw128(sb + ks * 2048, 8192, 1024), idesc, (kb > main_first) | (ks > 0));
        if (SUM) umma_bf16(tmem + Cfg::kOnesCol, da, umma_desc_mn_sw128(sones, 8192, 1024), idesc1, (kb > sum_first) | (ks > 0));
      }
      umma_commit(&empty[s]);
    }
    __syncwarp();
    if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
  }
}

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_tc_tn_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            float* __restrict__ C, float* __restrict__ colsum, int M,
                                                            int N, int k_blocks_total, int k_blocks_per_cta,
                                                            bool perm64, int perm_limit, int sum_mod,
                                                            long long split_stride) {
  using Cfg = GemmTnCfg<BN>;
  C += (size_t)blockIdx.y * split_stride;                 // deterministic mode: one partial per token split
  if (colsum) colsum += (size_t)blockIdx.y * split_stride;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / BN;
  const int mb = blockIdx.x / n_blocks, nb = blockIdx.x - mb * n_blocks;
  const int kb0 = blockIdx.y * k_blocks_per_cta;
  const int kb1 = kb0 + k_blocks_per_cta < k_blocks_total ? kb0 + k_blocks_per_cta : k_blocks_total;
  // Column sums: the n_blocks CTAs that share an A tile split the extra MMAs between them (all of them add into
  // colsum) in INTERLEAVED chunks of kSumChunk k blocks: chunk c is summed by the CTA with nb == c % n_blocks.
  // (Contiguous 1/n_blocks shares - the first version - made the CTAs that share an A tile run at different speeds in
  // different parts of the loop: they drifted ~90 us apart, further than a line stays in L2, and A was read from HBM
  // once per CTA - 3.70 GB instead of 2.15 GB for d_w at d = 512, ncu.)
  // sum_mod == n_blocks; deterministic mode passes 1: the nb == 0 CTA alone sums every chunk (one adder per entry).
  const int sum_first = colsum && nb < sum_mod ? kb0 + nb * kSumChunk : kb1;
  const bool do_sum = sum_first < kb1;

  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + Cfg::kOnesOff)[i] = 0x3f803f80u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {   // whole warp; the copies are issued by one elected lane (elect_one_sync, tc.cuh)
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one_sync()) {
        unsigned char* st = smem + s * Cfg::kStage;
        mbar_expect_tx(&full[s], Cfg::kStage);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d(st + j * 8192, &tmA, &full[s], mb * 128 + j * 64, kb * 64);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_2d(st + Cfg::kABytes + j * 8192, &tmB, &full[s], nb * BN + j * 64, kb * 64);
      }
      __syncwarp();
      if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    {
      int s = 0;
      uint32_t ph = 0;
      for (int c0 = kb0, c = 0; c0 < kb1; c0 += kSumChunk, ++c) {
        const int c1 = c0 + kSumChunk < kb1 ? c0 + kSumChunk : kb1;
        if (colsum && c % sum_mod == nb) tn_issue<BN, true>(smem, full, empty, tmem, c0, c1, kb0, sum_first, s, ph);
        else tn_issue<BN, false>(smem, full, empty, tmem, c0, c1, kb0, sum_first, s, ph);
      }
      if (elect_one_sync()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int arow = mb * 128 + q * 32 + lane;              // row of A^T = column of A
    const int row = perm64 && arow < perm_limit ? perm64_inv(arow) : arow;   // A's columns are K3's permuted dQKV columns: undo
    float* crow = C + (size_t)row * N + (size_t)nb * BN;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(crow + c + i, __uint_as_float(r[i]));
    }
    if (do_sum) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + Cfg::kOnesCol, r);  // 16 valid columns, all equal
      tmem_ld_wait();
      atomicAdd(colsum + row, __uint_as_float(r[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

// Deterministic mode.  The token split stays (it is what fills the 148 SMs), but every split writes its own zeroed
// partial [M*N | M] - its atomics then have one adder per element - and this kernel adds the partials in split order.
__global__ void tn_finish_kernel(const float* __restrict__ part, int splits, long long mn, int M, float* __restrict__ C,
                                 float* __restrict__ colsum) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = mn + M;
  if (i >= (colsum ? stride : mn)) return;
  float a = 0.f;
  for (int sp = 0; sp < splits; ++sp) a += part[sp * stride + i];
  if (i < mn) C[i] = a; else colsum[i - mn] = a;
}
struct TnDet {
  float* part = nullptr;
  long long stride = 0;
  int begin(int splits, int M, int N, cudaStream_t s) {
    if (!deterministic() || splits <= 1) return 0;
    stride = (long long)M * N + M;
    if (cudaMallocAsync((void**)&part, sizeof(float) * (size_t)stride * splits, s) != cudaSuccess) {
      cudaGetLastError();
      part = nullptr;
      return HWGAT_ERR_WORKSPACE;
    }
    cudaMemsetAsync(part, 0, sizeof(float) * (size_t)stride * splits, s);
    return 0;
  }
  void finish(int splits, int M, int N, float* C, float* colsum, cudaStream_t s) {
    if (!part) return;
    const long long total = colsum ? stride : (long long)M * N;
    tn_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(part, splits, (long long)M * N, M, C, colsum);
    count_launch();
    cudaFreeAsync(part, s);
  }
};

template <int BN>
static int launch_tn(const bf16* A, const bf16* Bm, float* C, float* colsum, int M, int N, long long Kd,
                     cudaStream_t s, bool perm64, int perm_limit) {
  using Cfg = GemmTnCfg<BN>;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(gemm_tc_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A, (uint64_t)Kd, (uint64_t)M, 64))) return st;
  if ((st = make_tmap_2d(&tmB, Bm, (uint64_t)Kd, (uint64_t)N, 64))) return st;
  const int tiles = (M / 128) * (N / BN);
  const int kblocks = (int)(Kd / 64);
  int splits = 148 / tiles;  // one wave: tiles * splits <= 148 SMs (a 149th CTA would double the time)
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  const int per = (kblocks + splits - 1) / splits;
  splits = (kblocks + per - 1) / per;
  TnDet det;
  if ((st = det.begin(splits, M, N, s))) return st;
  if (!det.part) {
    cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, s);
    if (colsum) cudaMemsetAsync(colsum, 0, sizeof(float) * M, s);
  }
  gemm_tc_tn_kernel<BN><<<dim3(tiles, splits), 192, Cfg::kSmem, s>>>(
      tmA, tmB, det.part ? det.part : C, !colsum ? nullptr : det.part ? det.part + (size_t)M * N : colsum, M, N, kblocks,
      per, perm64, perm_limit, deterministic() ? 1 : N / BN, det.stride);
  count_launch();
  det.finish(splits, M, N, C, colsum, s);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2, see ffn_tc.cu): a cluster of two CTAs accumulates one 256 x 256 tile of C over its
// token range; each CTA stages 128 of A's columns and 128 of B's columns per 64-token k block (32 KB instead of the
// 48 KB of the single-CTA 128 x 256 tile, and half of B comes out of the peer's shared memory).
// ---------------------------------------------------------------------------
struct GemmTnPairCfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = 64 * 128 * 2;         // [64 k][128 m] of this CTA
  static constexpr int kBBytes = 64 * 128 * 2;         // [64 k][128 n] of this CTA
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kOnesOff = kStages * kStage;
  static constexpr int kBarOff = kOnesOff + 8192;
  static constexpr int kSmem = kBarOff + 256 + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kOnesCol = 256;
};

template <bool SUM>
HW_DEV void tn_pair_issue(unsigned char* smem, uint64_t* full, uint64_t* empty, uint32_t tmem, int kb, int kend,
                          int main_first, int sum_first, int& s, uint32_t& ph) {
  using Cfg = GemmTnPairCfg;
  constexpr uint32_t idesc = umma_idesc_bf16(256, 256, true, true);
  constexpr uint32_t idesc1 = umma_idesc_bf16(256, 16, true, true);
  const uint32_t sones = smem_u32(smem + Cfg::kOnesOff);
  for (; kb < kend; ++kb) {
    mbar_wait_cluster(&full[s], ph);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + Cfg::kABytes;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t da = umma_desc_mn_sw128(sa + ks * 2048, 8192, 1024);
        umma_bf16_pair(tmem, da, umma_desc_mn_sw128(sb + ks * 2048, 8192, 1024), idesc, (kb > main_first) | (ks > 0));
        if (SUM)
          umma_bf16_pair(tmem + Cfg::kOnesCol, da, umma_desc_mn_sw128(sones, 8192, 1024), idesc1, (kb > sum_first) | (ks > 0));
      }
      umma_commit_pair(&empty[s]);
    }
    __syncwarp();
    if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
  }
}

__global__ void __launch_bounds__(192, 1) gemm_tc_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 float* __restrict__ C, float* __restrict__ colsum, int M,
                                                                 int N, int k_blocks_total, int k_blocks_per_cta,
                                                                 bool perm64, int perm_limit, int sum_mod,
                                                                 long long split_stride) {
  using Cfg = GemmTnPairCfg;
  C += (size_t)blockIdx.y * split_stride;
  if (colsum) colsum += (size_t)blockIdx.y * split_stride;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_blocks = N / 256;
  const int tile = blockIdx.x >> 1;
  const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
  const int kb0 = blockIdx.y * k_blocks_per_cta;
  const int kb1 = kb0 + k_blocks_per_cta < k_blocks_total ? kb0 + k_blocks_per_cta : k_blocks_total;
  const int sum_first = colsum && nb < sum_mod ? kb0 + nb * kSumChunk : kb1;   // interleaved chunks: see gemm_tc_tn_kernel
  const bool do_sum = sum_first < kb1;

  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + Cfg::kOnesOff)[i] = 0x3f803f80u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {   // whole warp, issue under elect_one_sync
    int s = 0;
    uint32_t ph = 0;
    const int acol = (2 * mb + (int)rank) * 128, bcol = nb * 256 + (int)rank * 128;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait_cluster(&empty[s], ph ^ 1);
      if (elect_one_sync()) {
        unsigned char* st = smem + s * Cfg::kStage;
        const uint32_t lead_full = mapa_shared(smem_u32(&full[s]), 0);
        if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::kStage);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d_pair(st + j * 8192, &tmA, lead_full, acol + j * 64, kb * 64);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d_pair(st + Cfg::kABytes + j * 8192, &tmB, lead_full, bcol + j * 64, kb * 64);
      }
      __syncwarp();
      if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int c0 = kb0, c = 0; c0 < kb1; c0 += kSumChunk, ++c) {
        const int c1 = c0 + kSumChunk < kb1 ? c0 + kSumChunk : kb1;
        if (colsum && c % sum_mod == nb) tn_pair_issue<true>(smem, full, empty, tmem, c0, c1, kb0, sum_first, s, ph);
        else tn_pair_issue<false>(smem, full, empty, tmem, c0, c1, kb0, sum_first, s, ph);
      }
      if (elect_one_sync()) umma_commit_pair(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int arow = (2 * mb + (int)rank) * 128 + q * 32 + lane;
    if (arow < M) {   // warp-uniform: M % 128 == 0
      const int row = perm64 && arow < perm_limit ? perm64_inv(arow) : arow;
      float* crow = C + (size_t)row * N + (size_t)nb * 256;
#pragma unroll 1
      for (int c = 0; c < 256; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(crow + c + i, __uint_as_float(r[i]));
      }
      if (do_sum) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + Cfg::kOnesCol, r);
        tmem_ld_wait();
        atomicAdd(colsum + row, __uint_as_float(r[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem, Cfg::kTmemCols);
}

static int launch_tn_pair(const bf16* A, const bf16* Bm, float* C, float* colsum, int M, int N, long long Kd,
                          cudaStream_t s, bool perm64, int perm_limit) {
  using Cfg = GemmTnPairCfg;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(gemm_tc_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A, (uint64_t)Kd, (uint64_t)M, 64))) return st;
  if ((st = make_tmap_2d(&tmB, Bm, (uint64_t)Kd, (uint64_t)N, 64))) return st;
  const int tiles = ((M + 255) / 256) * (N / 256);
  const int kblocks = (int)(Kd / 64);
  int splits = 74 / tiles;  // one wave of CTA pairs
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  const int per = (kblocks + splits - 1) / splits;
  splits = (kblocks + per - 1) / per;
  TnDet det;
  if ((st = det.begin(splits, M, N, s))) return st;
  if (!det.part) {
    cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, s);
    if (colsum) cudaMemsetAsync(colsum, 0, sizeof(float) * M, s);
  }
  float* c_out = det.part ? det.part : C;
  float* sum_out = !colsum ? nullptr : det.part ? det.part + (size_t)M * N : colsum;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(2 * tiles, splits);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, gemm_tc_tn_pair_kernel, tmA, tmB, c_out, sum_out, M, N, kblocks, per, perm64, perm_limit,
                                       deterministic() ? 1 : N / 256, det.stride);
  count_launch();
  det.finish(splits, M, N, C, colsum, s);
  return err != cudaSuccess ? (int)err : (int)cudaGetLastError();
}

// C[M,N] = A[Kd,M]^T . B[Kd,N] (fp32 out), colsum[M] = column sums of A; M % 128 == 0, N % 128 == 0, Kd % 64 == 0
// perm64: rows of C / entries of colsum are written through the inverse of K3's dQKV column permutation
int gemm_tc_tn(const bf16* A, const bf16* Bm, float* C, float* colsum, int M, int N, long long Kd, cudaStream_t s,
               bool perm64, int perm_limit) {
  if (M % 128 || N % 128 || Kd % 64) return HWGAT_ERR_UNSUPPORTED;
  if (N % 256 == 0) {
    // CTA pairs (256 x 256 tiles): +10-15 % once there are enough tiles to split the token range over 74 pairs evenly
    // (measured at M x N = 1536 x 512, 1024 x 512, 512 x 1024, 768 x 256; slower at two tiles)
    if (gemm_pair_enabled() && M >= 256 && ((M + 255) / 256) * (N / 256) >= 3)
      return launch_tn_pair(A, Bm, C, colsum, M, N, Kd, s, perm64, perm_limit);
    return launch_tn<256>(A, Bm, C, colsum, M, N, Kd, s, perm64, perm_limit);
  }
  return launch_tn<128>(A, Bm, C, colsum, M, N, Kd, s, perm64, perm_limit);
}

}  // namespace hwgat
