// Weight-side GEMMs of K3 (bf16 operands, fp32 accumulate, HMMA m16n8k16):
//   d_xn[n, d]  = dQKV[n, 3d] . Wqkv[3d, d]          (gemm_bf16_nn, bf16 out)
//   d_w [3d, d] = dQKV[n, 3d]^T . xn[n, d]           (gemm_bf16_tn_f32, split over tokens)
//   d_b [3d]    = column sums of dQKV                (extra ones-column MMA in the same kernel)
// i.e. autograd of self.qkv (HWGATE.py:86).  CTA tile 128x128, k step 32,
// 3-stage cp.async ring, 8 warps as 2(M) x 4(N), warp tile 64x32.
// All sizes that reach these kernels are multiples of the tile (api.cu).
#include "common.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

constexpr int kBM = 128, kBN = 128, kBK = 32, kGStages = 3;
constexpr int kTileBytes = 8192;                       // 128x32 or 32x128 bf16
constexpr int kGemmSmem = kGStages * 2 * kTileBytes;   // 48 KB

// [128 rows][32 k] tile, 64-byte rows
HW_DEV int offk(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }
// [32 k][128 cols] tile, 256-byte rows
HW_DEV int offr(int r, int c) { return r * 256 + ((c ^ (r & 7)) << 4); }

// load a [128][32] block of a row-major matrix (ld elements per row) -> k-contiguous tile
HW_DEV void load_kc_tile(unsigned char* st, const bf16* __restrict__ src, size_t ld) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = threadIdx.x + i * 256, r = idx >> 2, c = idx & 3;
    cp_async16(st + offk(r, c), src + (size_t)r * ld + c * 8);
  }
}
// load a [32][128] block of a row-major matrix -> row-per-k tile
HW_DEV void load_rk_tile(unsigned char* st, const bf16* __restrict__ src, size_t ld) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = threadIdx.x + i * 256, r = idx >> 4, c = idx & 15;
    cp_async16(st + offr(r, c), src + (size_t)r * ld + c * 8);
  }
}

// one k32 step of the warp tile.  kATrans: A tile is [k][m] (row-per-k) instead of [m][k].
// kOnes: also accumulate A . 1 (row sums of A^T = column sums of the stored matrix).
template <bool kATrans, bool kOnes>
HW_DEV void warp_mma_step(float (&acc)[4][4][4], float (&ones)[4][4], const unsigned char* sA, const unsigned char* sB,
                          int wm, int wn, int lane) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t a[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (kATrans) {
        const int kr = 16 * ks + ((lane >> 4) << 3) + (lane & 7);
        const int mc = ((64 * wm + 16 * mt) >> 3) + ((lane >> 3) & 1);
        ldsm_x4_t(a[mt], sA + offr(kr, mc));
      } else {
        ldsm_x4(a[mt], sA + offk(64 * wm + 16 * mt + (lane & 15), 2 * ks + (lane >> 4)));
      }
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      const int kr = 16 * ks + (((lane >> 3) & 1) << 3) + (lane & 7);
      const int nc = ((32 * wn + 16 * np) >> 3) + (lane >> 4);
      ldsm_x4_t(b, sB + offr(kr, nc));
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        mma16816(acc[mt][2 * np], a[mt], b[0], b[1]);
        mma16816(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
      }
    }
    if (kOnes) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) mma16816(ones[mt], a[mt], 0x3f803f80u, 0x3f803f80u);
    }
  }
}

// C[M,N] (bf16) = A[M,K] . B[K,N]; A, B, C row-major.
__global__ void __launch_bounds__(256) gemm_bf16_nn_kernel(const bf16* __restrict__ A, const bf16* __restrict__ Bm,
                                                           bf16* __restrict__ C, int N, int K) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 1, wn = warp >> 1;
  const size_t m0 = (size_t)blockIdx.y * kBM;
  const int n0 = blockIdx.x * kBN;
  const bf16* Ab = A + m0 * K;
  const bf16* Bb = Bm + n0;
  float acc[4][4][4] = {};
  float ones[4][4];
  const int nk = K / kBK;
  auto stage_a = [&](int s) { return smem + s * 2 * kTileBytes; };
  auto issue = [&](int kt) {
    unsigned char* st = stage_a(kt % kGStages);
    load_kc_tile(st, Ab + (size_t)kt * kBK, (size_t)K);
    load_rk_tile(st + kTileBytes, Bb + (size_t)kt * kBK * N, (size_t)N);
  };
  for (int s = 0; s < kGStages - 1; ++s) {
    if (s < nk) issue(s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kGStages - 2>();
    __syncthreads();
    if (kt + kGStages - 1 < nk) issue(kt + kGStages - 1);
    cp_async_commit();
    const unsigned char* st = stage_a(kt % kGStages);
    warp_mma_step<false, false>(acc, ones, st, st + kTileBytes, wm, wn, lane);
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const size_t r = m0 + 64 * wm + 16 * mt + g;
      const int c = n0 + 32 * wn + 8 * nt + 2 * t;
      *reinterpret_cast<uint32_t*>(C + r * N + c) = pack_bf16(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<uint32_t*>(C + (r + 8) * N + c) = pack_bf16(acc[mt][nt][2], acc[mt][nt][3]);
    }
}

// C[M,N] (fp32, zeroed by the launcher) += A[Kdim,M]^T . B[Kdim,N] over this CTA's slice of Kdim.
// colsum[M] += column sums of A over the slice (CTAs of the first N tile only).
__global__ void __launch_bounds__(256) gemm_bf16_tn_kernel(const bf16* __restrict__ A, const bf16* __restrict__ Bm,
                                                           float* __restrict__ C, float* __restrict__ colsum, int M,
                                                           int N, long long Kdim, long long k_per_cta) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 1, wn = warp >> 1;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const long long kbeg = (long long)blockIdx.z * k_per_cta;
  const long long kend = kbeg + k_per_cta < Kdim ? kbeg + k_per_cta : Kdim;
  const int nk = (int)((kend - kbeg) / kBK);
  const bf16* Ab = A + (size_t)kbeg * M + m0;
  const bf16* Bb = Bm + (size_t)kbeg * N + n0;
  const bool do_ones = blockIdx.x == 0 && wn == 0;
  float acc[4][4][4] = {};
  float ones[4][4] = {};
  auto stage_a = [&](int s) { return smem + s * 2 * kTileBytes; };
  auto issue = [&](int kt) {
    unsigned char* st = stage_a(kt % kGStages);
    load_rk_tile(st, Ab + (size_t)kt * kBK * M, (size_t)M);
    load_rk_tile(st + kTileBytes, Bb + (size_t)kt * kBK * N, (size_t)N);
  };
  for (int s = 0; s < kGStages - 1; ++s) {
    if (s < nk) issue(s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kGStages - 2>();
    __syncthreads();
    if (kt + kGStages - 1 < nk) issue(kt + kGStages - 1);
    cp_async_commit();
    const unsigned char* st = stage_a(kt % kGStages);
    if (do_ones)
      warp_mma_step<true, true>(acc, ones, st, st + kTileBytes, wm, wn, lane);
    else
      warp_mma_step<true, false>(acc, ones, st, st + kTileBytes, wm, wn, lane);
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int r = m0 + 64 * wm + 16 * mt + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = n0 + 32 * wn + 8 * nt + 2 * t;
      atomicAdd(C + (size_t)r * N + c, acc[mt][nt][0]);
      atomicAdd(C + (size_t)r * N + c + 1, acc[mt][nt][1]);
      atomicAdd(C + (size_t)(r + 8) * N + c, acc[mt][nt][2]);
      atomicAdd(C + (size_t)(r + 8) * N + c + 1, acc[mt][nt][3]);
    }
    if (do_ones && t == 0) {
      atomicAdd(colsum + r, ones[mt][0]);
      atomicAdd(colsum + r + 8, ones[mt][2]);
    }
  }
}

int gemm_bf16_nn(const bf16* A, const bf16* Bm, bf16* C, int M, int N, int K, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_bf16_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    attr_done = true;
  }
  if (M % kBM || N % kBN || K % kBK) return HWGAT_ERR_UNSUPPORTED;
  gemm_bf16_nn_kernel<<<dim3(N / kBN, M / kBM), 256, kGemmSmem, s>>>(A, Bm, C, N, K);
  count_launch();
  return (int)cudaGetLastError();
}

int gemm_bf16_tn_f32(const bf16* A, const bf16* Bm, float* C, float* colsum, int M, int N, long long Kdim,
                     cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    attr_done = true;
  }
  if (M % kBM || N % kBN || Kdim % kBK) return HWGAT_ERR_UNSUPPORTED;
  cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, s);
  cudaMemsetAsync(colsum, 0, sizeof(float) * M, s);
  const int tiles = (M / kBM) * (N / kBN);
  // split the token dimension so that the grid is ~4 waves of 148 SMs x 2 CTAs
  long long splits = (148LL * 8 + tiles - 1) / tiles;
  long long k_per_cta = ((Kdim + splits - 1) / splits + kBK - 1) / kBK * kBK;
  if (k_per_cta < 4 * kBK) k_per_cta = 4 * kBK;
  splits = (Kdim + k_per_cta - 1) / k_per_cta;
  gemm_bf16_tn_kernel<<<dim3(N / kBN, M / kBM, (unsigned)splits), 256, kGemmSmem, s>>>(A, Bm, C, colsum, M, N, Kdim,
                                                                                      k_per_cta);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
