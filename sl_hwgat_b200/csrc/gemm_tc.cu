// tcgen05 GEMMs of K3 (the autograd of self.qkv, HWGATE.py:86), bf16 operands,
// fp32 accumulation in TMEM, operands staged by TMA into SWIZZLE_128B smem:
//
//   gemm_tc_nt : C[M,N] (bf16) = A[M,K] . Bt[N,K]^T          d_xn = dQKV . Wqkv  (Bt = Wqkv^T)
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected thread; also owns the TMEM allocation), warps 2-5 = epilogue
// (tcgen05.ld -> bf16 -> global).  Two accumulators in TMEM so the epilogue of
// tile i overlaps the MMAs of tile i+1.
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
                 uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[4] = {d0, d1, d2, d3};
  cuuint64_t gstr[3] = {d0 * sizeof(bf16), d0 * d1 * sizeof(bf16), d0 * d1 * d2 * sizeof(bf16)};
  cuuint32_t box[4] = {64, box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// gemm_tc_nt
// ---------------------------------------------------------------------------
constexpr int kGM = 128, kGK = 64;

template <int BN>
struct GemmNtCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kGM * kGK * 2;        // 16 KB
  static constexpr int kBBytes = BN * kGK * 2;         // 16 / 32 KB
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = kStages * kStage;     // mbarriers + tmem slot
  static constexpr int kSmem = kBarOff + 256 + 1024;   // + manual 1024-byte alignment slack
  static constexpr int kTmemCols = 2 * BN;             // 256 or 512 (power of two)
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            bf16* __restrict__ C, int M, int N, int K) {
  using Cfg = GemmNtCfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / BN, m_blocks = M / kGM, tiles = n_blocks * m_blocks, nk = K / kGK;

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

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = smem + s * Cfg::kStage;
          mbar_expect_tx(&full[s], Cfg::kStage);
          tma_load_2d(st, &tmA, &full[s], kb * kGK, mb * kGM);
          tma_load_2d(st + Cfg::kABytes, &tmB, &full[s], kb * kGK, nb * BN);
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGM, BN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStage), sb = sa + Cfg::kABytes;
#pragma unroll
          for (int ks = 0; ks < kGK / 16; ++ks)
            umma_bf16(tmem + buf * BN, umma_desc_k_sw128(sa + ks * 32), umma_desc_k_sw128(sb + ks * 32), idesc,
                      (kb | ks) != 0);
          umma_commit(&empty[s]);
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int mb = tile / n_blocks, nb = tile - mb * n_blocks;
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const size_t row = (size_t)mb * kGM + q * 32 + lane;
      bf16* crow = C + row * N + (size_t)nb * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * BN + c, r);
        tmem_ld_wait();
        uint32_t p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<int4*>(crow + c + 8 * i) = make_int4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
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

// bf16 transpose [R][Cc] -> [Cc][R] (Wqkv -> Wqkv^T, 3d x d: tiny)
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int Cc) {
  __shared__ bf16 t[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < R && c < Cc) t[i][threadIdx.x] = in[(size_t)(r0 + i) * Cc + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < Cc && r < R) out[(size_t)(c0 + i) * R + r] = t[threadIdx.x][i];
}

int transpose_bf16(const bf16* in, bf16* out, int R, int Cc, cudaStream_t s) {
  transpose_bf16_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(in, out, R, Cc);
  count_launch();
  return (int)cudaGetLastError();
}

template <int BN>
static int launch_nt(const bf16* A, const bf16* Bt, bf16* C, int M, int N, int K, cudaStream_t s) {
  using Cfg = GemmNtCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_tc_nt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    attr_done = true;
  }
  CUtensorMap tmA, tmB;
  int st;
  if ((st = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, kGM))) return st;
  if ((st = make_tmap_2d(&tmB, Bt, (uint64_t)N, (uint64_t)K, BN))) return st;
  const int tiles = (M / kGM) * (N / BN);
  const int grid = tiles < 148 ? tiles : 148;
  gemm_tc_nt_kernel<BN><<<grid, 192, Cfg::kSmem, s>>>(tmA, tmB, C, M, N, K);
  count_launch();
  return (int)cudaGetLastError();
}

// C[M,N] = A[M,K] . Bt[N,K]^T ; M % 128 == 0, N in {128, 256, 512, ...}, K % 64 == 0
int gemm_tc_nt(const bf16* A, const bf16* Bt, bf16* C, int M, int N, int K, cudaStream_t s) {
  if (M % kGM || K % kGK || N % 128) return HWGAT_ERR_UNSUPPORTED;
  if (N % 256 == 0) return launch_nt<256>(A, Bt, C, M, N, K, s);
  return launch_nt<128>(A, Bt, C, M, N, K, s);
}

}  // namespace hwgat
