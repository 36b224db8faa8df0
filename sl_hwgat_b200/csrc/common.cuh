// Shared device helpers for the hwgat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/hwgat_b200.h"

#define HW_DEV __device__ __forceinline__

namespace hwgat {

constexpr float kNegFill = -10000.0f;  // HWGATE.py:110
constexpr int kWin = 16;               // keypoints per window (W)
constexpr int kTP = 2;                 // frames per window
constexpr int kTok = 32;               // tokens per window (N = TP*W)
constexpr int kHd = 64;                // head dim at every level
constexpr int kTileTok = 128;          // tokens per tile = one temporal group = 4 windows x 32

// Deterministic mode (hwgat_set_deterministic): parameter gradients are summed in a fixed order.  The token-split
// weight-gradient GEMMs keep the split but every split writes its own partial, and the column reductions of the
// bandwidth-bound backward kernels write per-CTA partials; a finish kernel adds the partials in index order.  The
// partials live in stream-ordered scratch (cudaMallocAsync - the one place the library allocates, only in this mode).
extern int g_deterministic;
inline bool deterministic() { return __atomic_load_n(&g_deterministic, __ATOMIC_RELAXED) != 0; }
// *part = nullptr outside deterministic mode; else [nvec][grid][cols] floats on the stream (block_fused.cu).  A failed
// allocation is an error (HWGAT_ERR_WORKSPACE), not a silent return to the atomic sums.
int det_scratch(float** part, int nvec, int grid, int cols, cudaStream_t s);
// o_v[c] = sum over grid, in index order, of part[v][.][c]; frees part on the stream; no-op for part == nullptr
void det_finish(float* part, int grid, int cols, float* o0, float* o1, float* o2, cudaStream_t s);

// Every launch made by the library is counted (hwgat_launch_count()).
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: every launcher opts its kernel in once
// per device (a process may drive several GPUs, and autograd runs backward on its own threads; setting the attribute
// twice from two racing threads is harmless).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <class Fn>
  void run(Fn&& fn) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return;
    fn();
    done.fetch_or(bit, std::memory_order_release);
  }
};

// ---- geometry of one tile -------------------------------------------------
// A tile is 4 keypoint windows (64 keypoints, group kg) of temporal group fi of
// sample b: frames (2*fi + tp + shift) mod F, tp = 0,1 - i.e.
// window_partition(torch.roll(x, -shift)) without the copy (HWGATE.py:197-201).
// Rows inside a tile are window-major:
//   row = w*32 + tp*16 + k   <->   frame(tp), keypoint kg*64 + w*16 + k
// so that rows [32w, 32w+32) are exactly window (b*f+fi)*nW + kg*4 + w with the
// reference's token order tp*W + k (HWGATE.py:34-35).  In the WINDOWS layout the
// caller has already partitioned, and tile t is simply rows [128t, 128t+128).
struct TileGeom {
  int F, K, d, shift, layout;
  int f, kgroups;  // F/2, K/64
  HW_DEV int tiles_per_sample() const { return f * kgroups; }
  // first mask word (row 0 of window w=0) of this tile: bits + mask_base(tile) + w*32 + i
  HW_DEV int mask_base(int tile) const {
    int r = tile % (f * kgroups);
    int fi = r / kgroups, kg = r - fi * kgroups;
    return (fi * (kgroups * 4) + kg * 4) * kTok;
  }
  HW_DEV long long token_row(int tile, int row) const {
    if (layout == HWGAT_LAYOUT_WINDOWS) return (long long)tile * kTileTok + row;
    int b = tile / (f * kgroups);
    int r = tile - b * (f * kgroups);
    int fi = r / kgroups, kg = r - fi * kgroups;
    int w = row >> 5, tp = (row >> 4) & 1, k = row & 15;
    int fr = 2 * fi + tp + shift;
    fr = fr >= F ? fr - F : fr;
    return ((long long)(b * F + fr) * K + kg * 64 + w * kWin + k);
  }
  // token rows of window w of a tile: rows [0, 16) of the window are tokens r0 .. r0 + 15 (frame tp = 0), rows
  // [16, 32) are r1 .. r1 + 15 (tp = 1) - the divisions of token_row once per window instead of once per row
  HW_DEV void window_base(int tile, int w, long long& r0, long long& r1) const {
    if (layout == HWGAT_LAYOUT_WINDOWS) {
      r0 = (long long)tile * kTileTok + w * kTok;
      r1 = r0 + kWin;
      return;
    }
    const int b = tile / (f * kgroups);
    const int r = tile - b * (f * kgroups);
    const int fi = r / kgroups, kg = r - fi * kgroups;
    int fr0 = 2 * fi + shift, fr1 = 2 * fi + 1 + shift;
    fr0 = fr0 >= F ? fr0 - F : fr0;
    fr1 = fr1 >= F ? fr1 - F : fr1;
    const long long col = kg * 64 + w * kWin;
    r0 = (long long)(b * F + fr0) * K + col;
    r1 = (long long)(b * F + fr1) * K + col;
  }
};
inline TileGeom make_geom(int F, int K, int d, int shift, int layout) {
  TileGeom g;
  g.F = F; g.K = K; g.d = d; g.shift = shift; g.layout = layout; g.f = F / 2; g.kgroups = K / 64;
  return g;
}

// ---- PTX wrappers -----------------------------------------------------------
HW_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
HW_DEV void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
HW_DEV uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
HW_DEV float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
HW_DEV float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// streaming 16-byte global access (data touched once: keep it out of L1)
HW_DEV int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
HW_DEV void st_stream16(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// 32-byte (one full sector) global access, sm_100+: a thread that owns a whole row segment writes full sectors
// (16-byte stores of row-per-thread epilogues hit every sector twice and halve the L2 write rate)
HW_DEV void st_global32(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
HW_DEV void ld_global32(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

// ---- launchers implemented in the .cu files --------------------------------
int launch_adjacency(const int32_t* edges, int n_edges, int nW, int W, int TP, float* adj, cudaStream_t s);
int launch_mask_bits(const float* adj, int nW, int W, int TP, int F, int shift, uint32_t* bits, cudaStream_t s);
int launch_mask_pack(const float* adj, int adj_windows, const float* mask, int n_windows, int N, uint32_t* bits,
                     cudaStream_t s);
int launch_merge(const void* src, void* dst, int B, int F, int K, int d, int elem_bytes, bool backward, cudaStream_t s);

// K5 - K7; f32: the activation tensors (y, dy, a0, d_a0, u0, g, dg, du0) are float instead of bf16 (the fp32 path)
int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long n, int d, float eps, cudaStream_t s, bool f32 = false);
int launch_ln_bwd(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                  const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d, cudaStream_t s,
                  int unmerge_F2 = 0, int unmerge_K = 0, bool f32 = false);
int launch_bda_ln_fwd(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                      float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                      unsigned long long seed, unsigned long long offset, cudaStream_t s, int merge_F = 0, int merge_K = 0,
                      bool f32 = false);
int launch_bda_ln_bwd(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                      const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                      long long n, int d, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s,
                      bool f32 = false);
int launch_bias_gelu_dropout(const void* u0, const float* bias, const void* dg, void* out, float* dbias, long long n,
                             int cols, float p, unsigned long long seed, unsigned long long offset, bool backward,
                             cudaStream_t s, bool f32 = false);

int launch_embed_fwd(const float* x, const float* Bm, const float* pe, float* out, long long n, int C, int E, int K,
                     int T, float p, unsigned long long seed, unsigned long long offset, cudaStream_t s);
size_t ln_pool_scratch_bytes(int B, int tokens, int d);
int launch_ln_pool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean, float* rstd,
                       float* scratch, int B, int tokens, int d, float eps, cudaStream_t s, int kp_real = 0,
                       int kp_pad = 0, const float* tok_w = nullptr);
int launch_ln_pool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       float* dx, float* dgamma, int B, int tokens, int d, cudaStream_t s, int kp_real = 0, int kp_pad = 0,
                       const float* tok_w = nullptr, float* dw_part = nullptr, float* d_tok_w = nullptr);

int linear_f32_fwd(const float* x, const float* w, const float* bias, float* y, int n, int d_in, int d_out,
                   cudaStream_t s);
int linear_f32_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int n, int d_in,
                   int d_out, cudaStream_t s);
// x3 mode of the fp32 path (gemm_x3.cu): fp32 GEMMs as six bf16 tcgen05 products of hi / mid / lo planes
extern int g_fp32_mode;
bool x3_enabled();
bool x3_supported(long long n, int d_in, int d_out);
int linear_x3_fwd(const float* x, const float* w, const float* bias, float* y, long long n, int d_in, int d_out,
                  cudaStream_t s, __nv_bfloat16* x_planes_out = nullptr);
int linear_x3_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, long long n, int d_in,
                  int d_out, cudaStream_t s, const __nv_bfloat16* x_planes = nullptr);
int colsum_f32(const float* x, float* out, long long n, int cols, cudaStream_t s);
int smooth_ce_fwd(const float* logits, const long long* target, float* lse, float* row_loss, float* loss, int rows,
                  int classes, float smooth, cudaStream_t s);
int smooth_ce_bwd(const float* logits, const long long* target, const float* lse, const float* g, float* dlogits,
                  int rows, int classes, float smooth, cudaStream_t s);

struct AttnArgs {
  const void* xn; const void* w_qkv; const float* b_qkv; const uint32_t* bits;
  float threshold;
  void* out;                 // fwd
  const void* d_out;         // bwd
  void* d_xn; float* d_w; float* d_b;
  void* workspace;
  void* qkv_out = nullptr;                  // K2 only: keep q, k, v for the hybrid backward (K3b)
  int qk_perm = 0;                          // K3b only: the q and k thirds of the saved qkv are column-permuted (from K2)
  float attn_p = 0.f;                       // attention dropout (K2b / K3b only)
  unsigned long long seed = 0, offset = 0;
  int B, F, K, d, heads, shift, layout;
  long long tokens() const { return (long long)B * F * K; }
  int tiles() const { return B * (F / 2) * (K / 64); }
};
int attn_fwd_f32(const AttnArgs& a, cudaStream_t s);
int attn_bwd_f32(const AttnArgs& a, cudaStream_t s, const float* kept_qkv = nullptr);
int attn_fwd_bf16(const AttnArgs& a, cudaStream_t s);
int attn_bwd_bf16(const AttnArgs& a, cudaStream_t s);
size_t attn2_workspace_bytes(long long n, int d, int backward, int have_qkv);
// fp32 parity mode of K2b / K3b (attn_win_f32.cu): windows of N = 64 / 128 tokens
size_t attn2_workspace_bytes_f32(long long n, int d, int backward);
int attn2_fwd_f32(const AttnArgs& a, int W, float* qkv, cudaStream_t s);
int attn2_bwd_f32(const AttnArgs& a, int W, const float* qkv, cudaStream_t s);
int qkv_weight_grads_f32(const float* dqkv, const float* xn, const float* w_qkv, float* d_xn, float* d_w, float* d_b,
                         long long n, int d, cudaStream_t s);
int attn2_fwd(const AttnArgs& a, int W, __nv_bfloat16* qkv, cudaStream_t s);
int attn2_bwd(const AttnArgs& a, int W, const __nv_bfloat16* qkv_saved, cudaStream_t s);
// K15 / K16 (band_attn.cu): frame-banded graph attention of WGATE / GATE
bool band_attn_supported(int B, int F, int K, int d, int heads, int W);
size_t band_attn_workspace_bytes(long long n, int d, int backward);
int band_attn_fwd(const __nv_bfloat16* xn, const __nv_bfloat16* w_qkv, const float* b_qkv, const uint32_t* bits,
                  __nv_bfloat16* out, __nv_bfloat16* qkv, float* lse, int B, int F, int K, int d, int heads, int W,
                  int diag, cudaStream_t s);
int band_attn_bwd(const __nv_bfloat16* xn, const __nv_bfloat16* w_qkv, const uint32_t* bits, const __nv_bfloat16* qkv,
                  const __nv_bfloat16* ctx, const float* lse, const __nv_bfloat16* d_out, __nv_bfloat16* d_xn,
                  float* d_w, float* d_b, void* workspace, int B, int F, int K, int d, int heads, int W, int diag,
                  cudaStream_t s);
int band_attn_fwd_f32(const float* xn, const float* w_qkv, const float* b_qkv, const uint32_t* bits, float* out,
                      float* qkv, float* lse, int B, int F, int K, int d, int heads, int W, cudaStream_t s);
int band_attn_bwd_f32(const float* xn, const float* w_qkv, const uint32_t* bits, const float* qkv, const float* ctx,
                      const float* lse, const float* d_out, float* d_xn, float* d_w, float* d_b, void* workspace, int B,
                      int F, int K, int d, int heads, int W, cudaStream_t s);
int attn_fwd_tc(const AttnArgs& a, cudaStream_t s);
int attn_bwd_tc(const AttnArgs& a, __nv_bfloat16* dqkv, cudaStream_t s);
int gemm_tc_tn(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* C, float* colsum, int M, int N, long long Kd,
               cudaStream_t s, bool perm64 = false, int perm_limit = 0x7fffffff);   // perm64 applies to rows < perm_limit
int transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int R, int Cc, cudaStream_t s, bool perm64 = false,
                   int perm_limit = 0x7fffffff);
bool gemm_pair_enabled();
bool set_gemm_pair(bool on);
int gemm_tc_nt_epi_none(const __nv_bfloat16* A, const __nv_bfloat16* Bt, __nv_bfloat16* C, long long M, int N, int K,
                        cudaStream_t s);
int ffn_fwd(const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2,
            __nv_bfloat16* act, __nv_bfloat16* gp, __nv_bfloat16* v0, long long n, int d, int hidden, float p,
            unsigned long long seed, unsigned long long offset, cudaStream_t s);
// K10f (ffn_fused.cu): inference FeedForward in one kernel, d = 128 / 256, hidden = 2 d
bool ffn_eval_fused_supported(long long n, int d, int hidden);
int ffn_eval_fused(const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2,
                   __nv_bfloat16* v0, long long n, int d, int hidden, cudaStream_t s);
int adamw_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
               float* const* exp_avg_sq, const long long* sizes, double lr, double beta1, double beta2, double eps,
               double weight_decay, long long step, float grad_scale, cudaStream_t s);
size_t ffn_bwd_workspace_bytes(long long n, int d, int hidden);
int ffn_bwd(const __nv_bfloat16* dv0, const __nv_bfloat16* h, const __nv_bfloat16* act, const __nv_bfloat16* gp,
            const __nv_bfloat16* w1, const __nv_bfloat16* w2, __nv_bfloat16* dh, float* dw1, float* db1, float* dw2,
            void* workspace, long long n, int d, int hidden, cudaStream_t s);

}  // namespace hwgat
