// K10f: the FeedForward of a block in INFERENCE as one kernel (HWGATE.py:130-134 under model.eval()):
//     v0 = gelu(h . W1^T + b1) . W2^T          d = 128 / 256, hidden = 2 d
// The two-GEMM form (ffn_tc.cu) writes the (n, hidden) activation to HBM and reads it back: 3.2 GB of the 4.8 GB a
// call moves at batch 256, T = 192.  Here a 128-token tile of the activation never leaves the SM: fc1 accumulates a
// [128 x 128] chunk of it in TMEM, the epilogue warps apply bias + GELU and store it as bf16 into shared memory in the
// K-major SWIZZLE_128B layout, and fc2's MMAs read it from there as their A operand.  HBM traffic: h once, v0 once.
//
// Persistent CTA, 576 threads:
//   warp 0      TMA producer: the X tile [128 x d] (d / 64 boxes) and a ring of [128 x 64] weight boxes (16 KB), in
//               exactly the order the MMA warp consumes them: W1 chunk 0, then for every hidden chunk c: W1 chunk c+1,
//               W2 k-slices of chunk c
//   warp 1      tcgen05.mma issuer (M128 N128 K16): fc1 of chunk c+1 is issued BEFORE fc2 of chunk c, so the tensor
//               core works on the next chunk while the epilogue warps are in the GELU of this one
//   warps 2-17  epilogue: 4 per TMEM lane quarter, 32 columns each.  TMEM -> +b1 -> exact-erf GELU -> bf16 -> shared
//               memory (two [128 x 128] activation buffers); after the last chunk the [128 x d] output tile -> HBM
// TMEM: two fc1 accumulators of 128 columns + the fc2 accumulator of d columns (512 columns at d = 256).
// Training keeps the two-GEMM form: its backward needs the activation and the local derivative in HBM anyway.
#include "ew.cuh"
#include "tc.cuh"

namespace hwgat {

typedef __nv_bfloat16 bf16;

template <int D>
struct FusedCfg {
  static constexpr int kH = 2 * D;
  static constexpr int kC = kH / 128;                 // hidden chunks of 128 columns
  static constexpr int kXChunks = D / 64;
  static constexpr int kXBytes = kXChunks * 16384;
  static constexpr int kXBufs = D == 128 ? 2 : 1;
  static constexpr int kRB = D / 128;                 // 128-row blocks of W2 (output columns of fc2)
  static constexpr int kSlots = 5;                    // weight ring: [128 rows x 64 k] boxes of 16 KB
  static constexpr int kXOff = 0;
  static constexpr int kRingOff = kXBufs * kXBytes;   // 64 KB for both widths
  static constexpr int kActOff = kRingOff + kSlots * 16384;
  static constexpr int kBarOff = kActOff + 2 * 32768;
  static constexpr int kBiasOff = kBarOff + 256;
  static constexpr int kSmem = kBiasOff + kH * 4 + 1024;
  static constexpr int kAcc2Col = 256;
  static constexpr int kTmemCols = 512;
  static constexpr int kEpiWarps = 16;
  static constexpr int kThreads = 32 * (2 + kEpiWarps);
};

HW_DEV void sts128f(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(FusedCfg<D>::kThreads, 1) ffn_eval_fused_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                                 const __grid_constant__ CUtensorMap tmW1,
                                                                                 const __grid_constant__ CUtensorMap tmW2,
                                                                                 const float* __restrict__ b1,
                                                                                 bf16* __restrict__ v0, int tiles) {
  using Cfg = FusedCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* x_empty = x_full + 2;
  uint64_t* w_full = x_empty + 2;
  uint64_t* w_empty = w_full + Cfg::kSlots;
  uint64_t* acc1_full = w_empty + Cfg::kSlots;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* act_full = acc1_empty + 2;
  uint64_t* act_empty = act_full + 2;
  uint64_t* acc2_full = act_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < Cfg::kH; i += blockDim.x) sbias[i] = b1 ? b1[i] : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], Cfg::kEpiWarps);
      mbar_init(&act_full[i], Cfg::kEpiWarps); mbar_init(&act_empty[i], 1);
    }
    for (int i = 0; i < Cfg::kSlots; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, Cfg::kEpiWarps);
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    int s = 0;
    uint32_t ph = 0;
    auto load_x = [&](int it, int tile) {
      const int xb = it % Cfg::kXBufs;
      mbar_wait(&x_empty[xb], ((it / Cfg::kXBufs) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&x_full[xb], Cfg::kXBytes);
#pragma unroll
        for (int kc = 0; kc < Cfg::kXChunks; ++kc)
          tma_load_2d(smem + Cfg::kXOff + xb * Cfg::kXBytes + kc * 16384, &tmX, &x_full[xb], kc * 64, tile * 128);
      }
      __syncwarp();
    };
    auto load_w = [&](const CUtensorMap* tm, int col, int row) {
      mbar_wait(&w_empty[s], ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&w_full[s], 16384);
        tma_load_2d(smem + Cfg::kRingOff + s * 16384, tm, &w_full[s], col, row);
      }
      __syncwarp();
      if (++s == Cfg::kSlots) { s = 0; ph ^= 1; }
    };
    int it = 0;
    if ((int)blockIdx.x < tiles) load_x(0, blockIdx.x);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c <= Cfg::kC; ++c) {
        if (c < Cfg::kC) {
          for (int kc = 0; kc < Cfg::kXChunks; ++kc) load_w(&tmW1, kc * 64, c * 128);
          // the next tile's X: after the last fc1 chunk's weights are on their way (with one X buffer the copy
          // waits for this tile's fc1 to finish, and that needs only what has been issued so far)
          if (c == Cfg::kC - 1 && tile + (int)gridDim.x < tiles) load_x(it + 1, tile + gridDim.x);
        }
        if (c >= 1)
          for (int kc = 0; kc < 2; ++kc)
            for (int rb = 0; rb < Cfg::kRB; ++rb) load_w(&tmW2, (c - 1) * 128 + kc * 64, rb * 128);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int s = 0, it = 0;
    uint32_t ph = 0;
    const uint32_t ring = smem_u32(smem + Cfg::kRingOff), act = smem_u32(smem + Cfg::kActOff);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int xb = it % Cfg::kXBufs;
      const uint32_t xs = smem_u32(smem + Cfg::kXOff + xb * Cfg::kXBytes);
      mbar_wait(&x_full[xb], (it / Cfg::kXBufs) & 1);
      tc_fence_after();
      for (int c = 0; c <= Cfg::kC; ++c) {
        if (c < Cfg::kC) {                                   // fc1 of hidden chunk c -> acc1[b]
          const int g = it * Cfg::kC + c, b = g & 1;
          mbar_wait(&acc1_empty[b], ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < Cfg::kXChunks; ++kc) {
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmem + b * 128, umma_desc_k_sw128(xs + kc * 16384 + ks * 32),
                          umma_desc_k_sw128(ring + s * 16384 + ks * 32), idesc, (kc | ks) != 0);
              umma_commit(&w_empty[s]);
              if (kc == Cfg::kXChunks - 1) {
                umma_commit(&acc1_full[b]);
                if (c == Cfg::kC - 1) umma_commit(&x_empty[xb]);
              }
            }
            __syncwarp();
            if (++s == Cfg::kSlots) { s = 0; ph ^= 1; }
          }
        }
        if (c >= 1) {                                        // fc2 over hidden chunk c - 1 -> acc2
          const int cc = c - 1, g = it * Cfg::kC + cc, b = g & 1;
          if (cc == 0) {
            mbar_wait(acc2_empty, (it & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&act_full[b], (g >> 1) & 1);
          tc_fence_after();
          for (int kc = 0; kc < 2; ++kc)
            for (int rb = 0; rb < Cfg::kRB; ++rb) {
              mbar_wait(&w_full[s], ph);
              tc_fence_after();
              if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_bf16(tmem + Cfg::kAcc2Col + rb * 128, umma_desc_k_sw128(act + b * 32768 + kc * 16384 + ks * 32),
                            umma_desc_k_sw128(ring + s * 16384 + ks * 32), idesc, (cc | kc | ks) != 0);
                umma_commit(&w_empty[s]);
                if (kc == 1 && rb == Cfg::kRB - 1) {
                  umma_commit(&act_empty[b]);
                  if (cc == Cfg::kC - 1) umma_commit(acc2_full);
                }
              }
              __syncwarp();
              if (++s == Cfg::kSlots) { s = 0; ph ^= 1; }
            }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int e = warp - 2, q = warp & 3, slice = e >> 2;    // lane quarter q (= warp % 4), 32-column slice of a chunk
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t act = smem_u32(smem + Cfg::kActOff);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c < Cfg::kC; ++c) {
        const int g = it * Cfg::kC + c, b = g & 1;
        uint32_t r[32];
        mbar_wait(&acc1_full[b], (g >> 1) & 1);
        tc_fence_after();
        tmem_ld32(tlane + b * 128 + slice * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);          // the accumulator is free for fc1 of chunk g + 2
        uint32_t pk[16];
        const uint32_t bp = smem_u32(sbias + c * 128 + slice * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                        // 16-byte broadcast loads of the bias
          float4 bv;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(bv.x), "=f"(bv.y), "=f"(bv.z), "=f"(bv.w) : "r"(bp + 16 * i));
          pk[2 * i] = pack_bf16(gelu_exact(__uint_as_float(r[4 * i]) + bv.x), gelu_exact(__uint_as_float(r[4 * i + 1]) + bv.y));
          pk[2 * i + 1] = pack_bf16(gelu_exact(__uint_as_float(r[4 * i + 2]) + bv.z), gelu_exact(__uint_as_float(r[4 * i + 3]) + bv.w));
        }
        mbar_wait(&act_empty[b], ((g >> 1) & 1) ^ 1);        // fc2 of chunk g - 2 has read this buffer
        {
          const int col0 = slice * 32;                        // column inside the [128 x 128] activation buffer
          const uint32_t base = act + b * 32768 + (col0 >> 6) * 16384 + row * 128;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ch = (((col0 & 63) >> 3) + u) ^ (row & 7);
            sts128f(base + (ch << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&act_full[b]);
      }
      // output tile: this warp's D / 4 columns of its 32 rows
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      bf16* orow = v0 + ((size_t)tile * 128 + row) * D + slice * (D / 4);
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        uint32_t r[32];
        tmem_ld32(tlane + Cfg::kAcc2Col + slice * (D / 4) + 32 * j, r);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack_bf16(__uint_as_float(r[16 * gq + 2 * i]), __uint_as_float(r[16 * gq + 2 * i + 1]));
          st_global32(orow + 32 * j + 16 * gq, o);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::kTmemCols);
}

template <int D>
static int launch_fused(const bf16* h, const bf16* w1, const float* b1, const bf16* w2, bf16* v0, long long n, cudaStream_t s) {
  using Cfg = FusedCfg<D>;
  static PerDeviceOnce once;
  once.run([] { cudaFuncSetAttribute(ffn_eval_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem); });
  CUtensorMap tmX, tmW1, tmW2;
  int st;
  if ((st = make_tmap_2d(&tmX, h, (uint64_t)n, (uint64_t)D, 128))) return st;
  if ((st = make_tmap_2d(&tmW1, w1, (uint64_t)Cfg::kH, (uint64_t)D, 128))) return st;
  if ((st = make_tmap_2d(&tmW2, w2, (uint64_t)D, (uint64_t)Cfg::kH, 128))) return st;
  const int tiles = (int)(n / 128);
  const int grid = tiles < 148 ? tiles : 148;
  ffn_eval_fused_kernel<D><<<grid, Cfg::kThreads, Cfg::kSmem, s>>>(tmX, tmW1, tmW2, b1, v0, tiles);
  count_launch();
  return (int)cudaGetLastError();
}

bool ffn_eval_fused_supported(long long n, int d, int hidden) {
  return (d == 128 || d == 256) && hidden == 2 * d && n >= 128 && n % 128 == 0 && n / 128 < 0x7fffffffLL;
}

int ffn_eval_fused(const bf16* h, const bf16* w1, const float* b1, const bf16* w2, bf16* v0, long long n, int d, int hidden,
                   cudaStream_t s) {
  if (!ffn_eval_fused_supported(n, d, hidden)) return HWGAT_ERR_UNSUPPORTED;
  return d == 128 ? launch_fused<128>(h, w1, b1, w2, v0, n, s) : launch_fused<256>(h, w1, b1, w2, v0, n, s);
}

}  // namespace hwgat
