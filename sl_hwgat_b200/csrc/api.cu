// C ABI of libhwgat_b200 (include/hwgat_b200.h): argument checks and dispatch.
// No allocation, no synchronisation, no fallback: a geometry or dtype the
// kernels do not cover is an error, not a slower path.
#include "common.cuh"

namespace hwgat {
unsigned long long g_launches = 0;
int g_deterministic = 0;

static bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

static int check_geometry(int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout, int dtype) {
  if (B < 0 || F <= 0 || K <= 0 || d <= 0 || heads <= 0) return HWGAT_ERR_SHAPE;
  if (dtype != HWGAT_F32 && dtype != HWGAT_BF16) return HWGAT_ERR_UNSUPPORTED;
  if (layout != HWGAT_LAYOUT_BFKD && layout != HWGAT_LAYOUT_WINDOWS) return HWGAT_ERR_UNSUPPORTED;
  if (W != kWin || TP != kTP) return HWGAT_ERR_UNSUPPORTED;
  if (d != heads * kHd || d > 512) return HWGAT_ERR_UNSUPPORTED;
  if (dtype == HWGAT_BF16 && d % 128 != 0) return HWGAT_ERR_UNSUPPORTED;  // tcgen05 GEMM tiles are 128 wide
  if (F % TP != 0 || K % 64 != 0) return HWGAT_ERR_UNSUPPORTED;
  if (shift < 0 || shift >= TP) return HWGAT_ERR_SHAPE;
  if (layout == HWGAT_LAYOUT_WINDOWS && shift != 0) return HWGAT_ERR_SHAPE;
  if ((long long)B * F * K > 0x7fffffffLL / 4) return HWGAT_ERR_UNSUPPORTED;  // row index kept in 31 bits
  // the fp32 parity kernels put 64-token tiles on grid.y (limit 65535): larger calls are refused, not mis-launched
  if (dtype == HWGAT_F32 && ((long long)B * F * K + 63) / 64 > 65535) return HWGAT_ERR_UNSUPPORTED;
  return HWGAT_OK;
}
}  // namespace hwgat

using namespace hwgat;

extern "C" {

int hwgat_version(void) { return 23; }

const char* hwgat_error_string(int status) {
  switch (status) {
    case HWGAT_OK: return "ok";
    case HWGAT_ERR_NULL: return "hwgat: required pointer is NULL";
    case HWGAT_ERR_SHAPE: return "hwgat: inconsistent sizes";
    case HWGAT_ERR_UNSUPPORTED: return "hwgat: geometry or dtype not supported by the sm_100a kernels (no fallback)";
    case HWGAT_ERR_WORKSPACE: return "hwgat: workspace too small";
    case HWGAT_ERR_ALIGN: return "hwgat: pointer not 16-byte aligned";
    default: break;
  }
  if (status > 0 && status < 1000) return cudaGetErrorString((cudaError_t)status);
  return "hwgat: unknown status";
}

unsigned long long hwgat_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int hwgat_set_fp32_mode(int mode) {
  if (mode != HWGAT_FP32_FFMA && mode != HWGAT_FP32_X3) return __atomic_load_n(&g_fp32_mode, __ATOMIC_RELAXED);
  return __atomic_exchange_n(&g_fp32_mode, mode, __ATOMIC_RELAXED);
}
int hwgat_set_deterministic(int on) {
  if (on < 0) return __atomic_load_n(&g_deterministic, __ATOMIC_RELAXED);   // query
  return __atomic_exchange_n(&g_deterministic, on ? 1 : 0, __ATOMIC_RELAXED);
}

int hwgat_adjacency_build(const int32_t* edges, int n_edges, int nW, int W, int TP, float* adj,
                          hwgat_stream_t stream) {
  if (!adj || (n_edges > 0 && !edges)) return HWGAT_ERR_NULL;
  if (nW <= 0 || W <= 0 || W > 64 || TP <= 0 || n_edges < 0) return HWGAT_ERR_SHAPE;
  return launch_adjacency(edges, n_edges, nW, W, TP, adj, (cudaStream_t)stream);
}

int hwgat_mask_build(const float* adj, int nW, int W, int TP, int F, int shift, uint32_t* bits,
                     hwgat_stream_t stream) {
  if (!adj || !bits) return HWGAT_ERR_NULL;
  if (nW <= 0 || W <= 0 || TP <= 0 || F <= 0 || F % TP != 0 || shift < 0 || shift >= TP) return HWGAT_ERR_SHAPE;
  if ((TP * W) % 32 != 0) return HWGAT_ERR_UNSUPPORTED;
  return launch_mask_bits(adj, nW, W, TP, F, shift, bits, (cudaStream_t)stream);
}

int hwgat_mask_pack(const float* adj, int adj_windows, const float* mask, int n_windows, int N, uint32_t* bits,
                    hwgat_stream_t stream) {
  if (!bits) return HWGAT_ERR_NULL;
  if (n_windows < 0 || N <= 0 || (adj && adj_windows <= 0)) return HWGAT_ERR_SHAPE;
  if (N % 32 != 0) return HWGAT_ERR_UNSUPPORTED;
  return launch_mask_pack(adj, adj_windows, mask, n_windows, N, bits, (cudaStream_t)stream);
}

size_t hwgat_attn_workspace_bytes(int B, int F, int K, int d, int heads, int dtype, int backward) {
  const size_t n = (size_t)B * F * K;
  if (dtype == HWGAT_F32) return n * 3 * d * sizeof(float) * (backward ? 2 : 1);  // qkv (+ dqkv)
  // forward: scaled weight copy [3d,d] + per-head bias tiles [heads][192 x 16];
  // backward: dqkv [n,3d] + Wqkv^T [d,3d] + the same two
  const size_t prep = (size_t)3 * d * d + (size_t)heads * 192 * 16;
  return ((backward ? n * 3 * d + (size_t)3 * d * d : 0) + prep) * sizeof(__nv_bfloat16);
}

int hwgat_attn_fwd(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                   void* out, void* workspace, size_t workspace_bytes, int B, int F, int K, int d, int heads, int W,
                   int TP, int shift, int layout, int dtype, hwgat_stream_t stream) {
  int st = check_geometry(B, F, K, d, heads, W, TP, shift, layout, dtype);
  if (st) return st;
  if (B == 0) return HWGAT_OK;
  if (!xn || !w_qkv || !b_qkv || !bits || !out) return HWGAT_ERR_NULL;
  if (misaligned(xn) || misaligned(w_qkv) || misaligned(out) || misaligned(b_qkv) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  const size_t need = hwgat_attn_workspace_bytes(B, F, K, d, heads, dtype, 0);
  if (need > 0 && (!workspace || workspace_bytes < need)) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.out = out;
  a.workspace = workspace; a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return dtype == HWGAT_F32 ? attn_fwd_f32(a, (cudaStream_t)stream) : attn_fwd_bf16(a, (cudaStream_t)stream);
}

int hwgat_attn_fwd_keep(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                        void* out, void* qkv, void* workspace, size_t workspace_bytes, int B, int F, int K, int d,
                        int heads, int W, int TP, int shift, int layout, hwgat_stream_t stream) {
  int st = check_geometry(B, F, K, d, heads, W, TP, shift, layout, HWGAT_BF16);
  if (st) return st;
  if (B == 0) return HWGAT_OK;
  if (!xn || !w_qkv || !b_qkv || !bits || !out || !qkv) return HWGAT_ERR_NULL;
  if (misaligned(xn) || misaligned(w_qkv) || misaligned(out) || misaligned(b_qkv) || misaligned(workspace) ||
      misaligned(qkv))
    return HWGAT_ERR_ALIGN;
  const size_t need = hwgat_attn_workspace_bytes(B, F, K, d, heads, HWGAT_BF16, 0);
  if (!workspace || workspace_bytes < need) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.out = out;
  a.qkv_out = qkv;
  a.workspace = workspace; a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return attn_fwd_bf16(a, (cudaStream_t)stream);
}

int hwgat_attn_bwd(const void* d_out, const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits,
                   float threshold, void* d_xn, float* d_w, float* d_b, void* workspace, size_t workspace_bytes,
                   int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout, int dtype,
                   hwgat_stream_t stream) {
  int st = check_geometry(B, F, K, d, heads, W, TP, shift, layout, dtype);
  if (st) return st;
  if (!d_w || !d_b) return HWGAT_ERR_NULL;
  if (B == 0) {  // empty batch: gradients of the parameters are zero
    cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, (cudaStream_t)stream);
    cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!d_out || !xn || !w_qkv || !b_qkv || !bits || !d_xn) return HWGAT_ERR_NULL;
  if (misaligned(d_out) || misaligned(xn) || misaligned(w_qkv) || misaligned(d_xn) || misaligned(d_w) ||
      misaligned(b_qkv) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  const size_t need = hwgat_attn_workspace_bytes(B, F, K, d, heads, dtype, 1);
  if (!workspace || workspace_bytes < need) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.d_out = d_out;
  a.d_xn = d_xn; a.d_w = d_w; a.d_b = d_b; a.workspace = workspace;
  a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return dtype == HWGAT_F32 ? attn_bwd_f32(a, (cudaStream_t)stream) : attn_bwd_bf16(a, (cudaStream_t)stream);
}

static int check_geometry2(int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout) {
  if (B < 0 || F <= 0 || K <= 0 || d <= 0 || heads <= 0) return HWGAT_ERR_SHAPE;
  if (layout != HWGAT_LAYOUT_BFKD && layout != HWGAT_LAYOUT_WINDOWS) return HWGAT_ERR_UNSUPPORTED;
  if ((W != 16 && W != 32 && W != 64) || TP != kTP) return HWGAT_ERR_UNSUPPORTED;
  if (d != heads * kHd || d > 512 || d % 128 != 0) return HWGAT_ERR_UNSUPPORTED;
  // whole windows per keypoint axis and whole 128-token tiles per sample (a tile is 128 / (2 W) consecutive windows)
  if (F % TP != 0 || K % W != 0 || ((F / TP) * (K / W)) % (64 / W) != 0) return HWGAT_ERR_UNSUPPORTED;
  if (shift < 0 || shift >= TP) return HWGAT_ERR_SHAPE;
  if (layout == HWGAT_LAYOUT_WINDOWS && shift != 0) return HWGAT_ERR_SHAPE;
  if ((long long)B * F * K > 0x7fffffffLL / 4) return HWGAT_ERR_UNSUPPORTED;
  return HWGAT_OK;
}

// ---- K15 / K16: frame-banded graph attention (WGATE.py:68-108, GATE.py:30-69) ----
// fp32 (parity mode): no 128-row GEMM tiles, so only the keypoint / head geometry is constrained
static bool band_f32_supported(int B, int F, int K, int d, int heads, int W) {
  if (B < 0 || F < 1 || K < 1 || d < 1 || heads < 1 || (W != 16 && W != 32) || K % W || d % heads) return false;
  const int hd = d / heads;
  return hd == 16 || hd == 32 || hd == 64;
}
static int band_check(int dtype, int B, int F, int K, int d, int heads, int W) {
  if (dtype != HWGAT_F32 && dtype != HWGAT_BF16) return HWGAT_ERR_UNSUPPORTED;
  if (B < 0 || F < 1 || K < 1 || d < 1 || heads < 1) return HWGAT_ERR_SHAPE;
  const bool ok = dtype == HWGAT_F32 ? band_f32_supported(B, F, K, d, heads, W) : band_attn_supported(B, F, K, d, heads, W);
  return ok ? HWGAT_OK : HWGAT_ERR_UNSUPPORTED;
}

size_t hwgat_band_attn_workspace_bytes(int dtype, int B, int F, int K, int d, int backward) {
  const long long n = (long long)B * F * K;
  if (dtype == HWGAT_F32) return backward ? sizeof(float) * (size_t)n * 3 * d + 256 : 0;
  return band_attn_workspace_bytes(n, d, backward);
}

int hwgat_band_attn_fwd(int dtype, const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits,
                        void* out, void* qkv, float* lse, int B, int F, int K, int d, int heads, int W, int diag,
                        hwgat_stream_t stream) {
  if (int st = band_check(dtype, B, F, K, d, heads, W)) return st;
  if (B == 0) return HWGAT_OK;
  if (!xn || !w_qkv || !b_qkv || !bits || !out || !qkv) return HWGAT_ERR_NULL;
  if (misaligned(xn) || misaligned(w_qkv) || misaligned(b_qkv) || misaligned(out) || misaligned(qkv) || misaligned(lse))
    return HWGAT_ERR_ALIGN;
  if (dtype == HWGAT_F32)
    return band_attn_fwd_f32((const float*)xn, (const float*)w_qkv, b_qkv, bits, (float*)out, (float*)qkv, lse, B, F, K,
                             d, heads, W, (cudaStream_t)stream);
  return band_attn_fwd((const __nv_bfloat16*)xn, (const __nv_bfloat16*)w_qkv, b_qkv, bits, (__nv_bfloat16*)out,
                       (__nv_bfloat16*)qkv, lse, B, F, K, d, heads, W, diag, (cudaStream_t)stream);
}

int hwgat_band_attn_bwd(int dtype, const void* d_out, const void* xn, const void* w_qkv, const void* qkv,
                        const void* ctx, const float* lse, const uint32_t* bits, void* d_xn, float* d_w, float* d_b,
                        void* workspace, size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int diag,
                        hwgat_stream_t stream) {
  if (int st = band_check(dtype, B, F, K, d, heads, W)) return st;
  if (!d_w || !d_b) return HWGAT_ERR_NULL;
  if (B == 0) {
    cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, (cudaStream_t)stream);
    cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!d_out || !xn || !w_qkv || !qkv || !ctx || !lse || !bits || !d_xn) return HWGAT_ERR_NULL;
  if (misaligned(d_out) || misaligned(xn) || misaligned(w_qkv) || misaligned(qkv) || misaligned(ctx) || misaligned(lse) ||
      misaligned(d_xn) || misaligned(d_w) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < hwgat_band_attn_workspace_bytes(dtype, B, F, K, d, 1)) return HWGAT_ERR_WORKSPACE;
  if (dtype == HWGAT_F32)
    return band_attn_bwd_f32((const float*)xn, (const float*)w_qkv, bits, (const float*)qkv, (const float*)ctx, lse,
                             (const float*)d_out, (float*)d_xn, d_w, d_b, workspace, B, F, K, d, heads, W,
                             (cudaStream_t)stream);
  return band_attn_bwd((const __nv_bfloat16*)xn, (const __nv_bfloat16*)w_qkv, bits, (const __nv_bfloat16*)qkv,
                       (const __nv_bfloat16*)ctx, lse, (const __nv_bfloat16*)d_out, (__nv_bfloat16*)d_xn, d_w, d_b,
                       workspace, B, F, K, d, heads, W, diag, (cudaStream_t)stream);
}

size_t hwgat_attn2_workspace_bytes(int B, int F, int K, int d, int heads, int backward, int have_qkv) {
  (void)heads;
  return attn2_workspace_bytes((long long)B * F * K, d, backward, have_qkv);
}

int hwgat_attn2_fwd(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                    void* out, void* qkv, void* workspace, size_t workspace_bytes, int B, int F, int K, int d,
                    int heads, int W, int TP, int shift, int layout, float attn_p, unsigned long long seed,
                    unsigned long long offset, hwgat_stream_t stream) {
  int st = check_geometry2(B, F, K, d, heads, W, TP, shift, layout);
  if (st) return st;
  if (!(attn_p >= 0.f) || attn_p >= 1.f) return HWGAT_ERR_SHAPE;
  if (B == 0) return HWGAT_OK;
  if (!xn || !w_qkv || !b_qkv || !bits || !out || !qkv) return HWGAT_ERR_NULL;
  if (misaligned(xn) || misaligned(w_qkv) || misaligned(out) || misaligned(b_qkv) || misaligned(workspace) ||
      misaligned(qkv))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < hwgat_attn2_workspace_bytes(B, F, K, d, heads, 0, 1)) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.out = out;
  a.workspace = workspace; a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  a.attn_p = attn_p; a.seed = seed; a.offset = offset;
  return attn2_fwd(a, W, (__nv_bfloat16*)qkv, (cudaStream_t)stream);
}

int hwgat_attn2_bwd(const void* d_out, const void* xn, const void* w_qkv, const float* b_qkv, const void* qkv,
                    const uint32_t* bits, float threshold, void* d_xn, float* d_w, float* d_b, void* workspace,
                    size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                    int layout, int qk_perm, float attn_p, unsigned long long seed, unsigned long long offset,
                    hwgat_stream_t stream) {
  int st = check_geometry2(B, F, K, d, heads, W, TP, shift, layout);
  if (st) return st;
  if (!(attn_p >= 0.f) || attn_p >= 1.f) return HWGAT_ERR_SHAPE;
  if (qk_perm && !qkv) return HWGAT_ERR_NULL;      // only a kept qkv can be permuted (the re-projection is not)
  if (!d_w || !d_b) return HWGAT_ERR_NULL;
  if (B == 0) {
    cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, (cudaStream_t)stream);
    cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!d_out || !xn || !w_qkv || !b_qkv || !bits || !d_xn) return HWGAT_ERR_NULL;
  if (misaligned(d_out) || misaligned(xn) || misaligned(w_qkv) || misaligned(d_xn) || misaligned(d_w) ||
      misaligned(b_qkv) || misaligned(workspace) || misaligned(qkv))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < hwgat_attn2_workspace_bytes(B, F, K, d, heads, 1, qkv != nullptr))
    return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.d_out = d_out;
  a.d_xn = d_xn; a.d_w = d_w; a.d_b = d_b; a.workspace = workspace;
  a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  a.attn_p = attn_p; a.seed = seed; a.offset = offset; a.qk_perm = qk_perm;
  return attn2_bwd(a, W, (const __nv_bfloat16*)qkv, (cudaStream_t)stream);
}

// ---- fp32 parity mode of the general-window attention (attn_win_f32.cu): window_size 32 / 64, HGATE's blocks ----
static int check_geometry2_f32(int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout) {
  if (B < 0 || F <= 0 || K <= 0 || d <= 0 || heads <= 0) return HWGAT_ERR_SHAPE;
  if (layout != HWGAT_LAYOUT_BFKD && layout != HWGAT_LAYOUT_WINDOWS) return HWGAT_ERR_UNSUPPORTED;
  if ((W != 32 && W != 64) || TP != kTP) return HWGAT_ERR_UNSUPPORTED;     // W = 16: hwgat_attn_fwd(HWGAT_F32, ...)
  if (d != heads * kHd || d % 4 != 0) return HWGAT_ERR_UNSUPPORTED;
  if (F % TP != 0 || K % W != 0) return HWGAT_ERR_UNSUPPORTED;
  if (shift < 0 || shift >= TP) return HWGAT_ERR_SHAPE;
  if (layout == HWGAT_LAYOUT_WINDOWS && shift != 0) return HWGAT_ERR_SHAPE;
  if ((long long)B * F * K > 64LL * 65535) return HWGAT_ERR_UNSUPPORTED;   // the FFMA GEMMs put the token tiles on grid.y
  return HWGAT_OK;
}

size_t hwgat_attn2_f32_workspace_bytes(int B, int F, int K, int d, int backward) {
  return attn2_workspace_bytes_f32((long long)B * F * K, d, backward);
}

int hwgat_attn2_fwd_f32(const float* xn, const float* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                        float* out, float* qkv, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                        int layout, hwgat_stream_t stream) {
  int st = check_geometry2_f32(B, F, K, d, heads, W, TP, shift, layout);
  if (st) return st;
  if (B == 0) return HWGAT_OK;
  if (!xn || !w_qkv || !b_qkv || !bits || !out || !qkv) return HWGAT_ERR_NULL;
  if (misaligned(xn) || misaligned(w_qkv) || misaligned(out) || misaligned(b_qkv) || misaligned(qkv)) return HWGAT_ERR_ALIGN;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = b_qkv; a.bits = bits; a.threshold = threshold; a.out = out;
  a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return attn2_fwd_f32(a, W, qkv, (cudaStream_t)stream);
}

int hwgat_attn2_bwd_f32(const float* d_out, const float* xn, const float* w_qkv, const float* qkv, const uint32_t* bits,
                        float threshold, float* d_xn, float* d_w, float* d_b, void* workspace, size_t workspace_bytes,
                        int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout,
                        hwgat_stream_t stream) {
  int st = check_geometry2_f32(B, F, K, d, heads, W, TP, shift, layout);
  if (st) return st;
  if (!d_w || !d_b) return HWGAT_ERR_NULL;
  if (B == 0) {
    cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, (cudaStream_t)stream);
    cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!d_out || !xn || !w_qkv || !qkv || !bits || !d_xn) return HWGAT_ERR_NULL;
  if (misaligned(d_out) || misaligned(xn) || misaligned(w_qkv) || misaligned(qkv) || misaligned(d_xn) || misaligned(d_w) ||
      misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < hwgat_attn2_f32_workspace_bytes(B, F, K, d, 1)) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.bits = bits; a.threshold = threshold; a.d_out = d_out;
  a.d_xn = d_xn; a.d_w = d_w; a.d_b = d_b; a.workspace = workspace;
  a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return attn2_bwd_f32(a, W, qkv, (cudaStream_t)stream);
}

static int ln_fwd_impl(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                 long long n, int d, float eps, hwgat_stream_t stream, bool f32) {
  if (n < 0) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (n == 0) return HWGAT_OK;
  if (!x || !gamma || !beta || !y || !mean || !rstd) return HWGAT_ERR_NULL;
  if (misaligned(x) || misaligned(gamma) || misaligned(beta) || misaligned(y)) return HWGAT_ERR_ALIGN;
  return launch_ln_fwd(x, gamma, beta, y, mean, rstd, n, d, eps, (cudaStream_t)stream, f32);
}
int hwgat_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                 long long n, int d, float eps, hwgat_stream_t stream) {
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, n, d, eps, stream, false);
}
int hwgat_ln_fwd_f32(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                 long long n, int d, float eps, hwgat_stream_t stream) {
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, n, d, eps, stream, true);
}

static int ln_bwd_impl(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                 const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d,
                 hwgat_stream_t stream, bool f32) {
  if (n < 0) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (!dgamma || !dbeta) return HWGAT_ERR_NULL;
  if (n > 0 && (!dy || !x || !mean || !rstd || !gamma || !dx)) return HWGAT_ERR_NULL;
  if (misaligned(dy) || misaligned(dres) || misaligned(x) || misaligned(gamma) || misaligned(dx)) return HWGAT_ERR_ALIGN;
  return launch_ln_bwd(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, d, (cudaStream_t)stream, 0, 0, f32);
}
int hwgat_ln_bwd(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                 const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d,
                 hwgat_stream_t stream) {
  return ln_bwd_impl(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, d, stream, false);
}
int hwgat_ln_bwd_f32(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                 const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d,
                 hwgat_stream_t stream) {
  return ln_bwd_impl(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n, d, stream, true);
}

static bool bad_p(float p) { return !(p >= 0.f) || p >= 1.f; }

static int ln_bwd_unmerge_impl(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, float* dgamma, float* dbeta, long long n_merged, int d_merged,
                         int F_merged, int K, hwgat_stream_t stream, bool f32) {
  if (n_merged < 0 || F_merged <= 0 || K <= 0) return HWGAT_ERR_SHAPE;
  if (d_merged != 256 && d_merged != 512) return HWGAT_ERR_UNSUPPORTED;
  if (n_merged % ((long long)F_merged * K) != 0) return HWGAT_ERR_SHAPE;
  if (!dgamma || !dbeta) return HWGAT_ERR_NULL;
  if (n_merged > 0 && (!dy || !x || !mean || !rstd || !gamma || !dx)) return HWGAT_ERR_NULL;
  if (misaligned(dy) || misaligned(dres) || misaligned(x) || misaligned(gamma) || misaligned(dx)) return HWGAT_ERR_ALIGN;
  return launch_ln_bwd(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n_merged, d_merged, (cudaStream_t)stream,
                       F_merged, K, f32);
}
int hwgat_ln_bwd_unmerge(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, float* dgamma, float* dbeta, long long n_merged, int d_merged,
                         int F_merged, int K, hwgat_stream_t stream) {
  return ln_bwd_unmerge_impl(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n_merged, d_merged, F_merged, K, stream, false);
}
int hwgat_ln_bwd_unmerge_f32(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, float* dgamma, float* dbeta, long long n_merged, int d_merged,
                         int F_merged, int K, hwgat_stream_t stream) {
  return ln_bwd_unmerge_impl(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, n_merged, d_merged, F_merged, K, stream, true);
}

static int bda_merge_fwd_impl(const float* res, const void* a0, const float* bias, float* x_merged, long long n, int d, int F,
                        int K, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream, bool f32) {
  if (n < 0 || bad_p(p) || F <= 0 || K <= 0) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256) return HWGAT_ERR_UNSUPPORTED;
  if (F % 2 != 0) return HWGAT_ERR_UNSUPPORTED;
  if (n % ((long long)F * K) != 0) return HWGAT_ERR_SHAPE;
  if (n == 0) return HWGAT_OK;
  if (!res || !a0 || !x_merged) return HWGAT_ERR_NULL;
  if (misaligned(res) || misaligned(a0) || misaligned(bias) || misaligned(x_merged)) return HWGAT_ERR_ALIGN;
  return launch_bda_ln_fwd(res, a0, bias, nullptr, nullptr, x_merged, nullptr, nullptr, nullptr, n, d, 0.f, p, seed,
                           offset, (cudaStream_t)stream, F, K, f32);
}
int hwgat_bda_merge_fwd(const float* res, const void* a0, const float* bias, float* x_merged, long long n, int d, int F,
                        int K, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bda_merge_fwd_impl(res, a0, bias, x_merged, n, d, F, K, p, seed, offset, stream, false);
}
int hwgat_bda_merge_fwd_f32(const float* res, const void* a0, const float* bias, float* x_merged, long long n, int d, int F,
                        int K, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bda_merge_fwd_impl(res, a0, bias, x_merged, n, d, F, K, p, seed, offset, stream, true);
}

int hwgat_linear_f32_fwd(const float* x, const float* w, const float* bias, float* y, int n, int d_in, int d_out,
                         hwgat_stream_t stream) {
  if (n < 0 || d_in <= 0 || d_out <= 0) return HWGAT_ERR_SHAPE;
  if (n == 0) return HWGAT_OK;
  if (!x || !w || !y) return HWGAT_ERR_NULL;
  if ((n + 63) / 64 > 65535) return HWGAT_ERR_UNSUPPORTED;
  return linear_f32_fwd(x, w, bias, y, n, d_in, d_out, (cudaStream_t)stream);
}

int hwgat_linear_f32_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int n,
                         int d_in, int d_out, hwgat_stream_t stream) {
  if (n < 0 || d_in <= 0 || d_out <= 0) return HWGAT_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    if (dw) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d_in * d_out, s);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * d_out, s);
    return (int)cudaGetLastError();
  }
  if (!dy || (dx && !w) || (dw && !x)) return HWGAT_ERR_NULL;
  if ((n + 63) / 64 > 65535) return HWGAT_ERR_UNSUPPORTED;
  return linear_f32_bwd(dy, x, w, dx, dw, db, n, d_in, d_out, s);
}

/* x3 Linear with the split of x kept by the caller (see hwgat_b200.h) */
int hwgat_linear_x3_supported(long long n, int d_in, int d_out) { return x3_supported(n, d_in, d_out) ? 1 : 0; }

int hwgat_linear_x3_fwd(const float* x, const float* w, const float* bias, float* y, void* x_planes, long long n, int d_in,
                        int d_out, hwgat_stream_t stream) {
  if (n <= 0 || d_in <= 0 || d_out <= 0) return HWGAT_ERR_SHAPE;
  if (!x3_supported(n, d_in, d_out)) return HWGAT_ERR_UNSUPPORTED;
  if (!x || !w || !y) return HWGAT_ERR_NULL;
  if (misaligned(x) || misaligned(w) || misaligned(y) || misaligned(x_planes)) return HWGAT_ERR_ALIGN;
  return linear_x3_fwd(x, w, bias, y, n, d_in, d_out, (cudaStream_t)stream, (__nv_bfloat16*)x_planes);
}

int hwgat_linear_x3_bwd(const float* dy, const void* x_planes, const float* w, float* dx, float* dw, float* db,
                        long long n, int d_in, int d_out, hwgat_stream_t stream) {
  if (n <= 0 || d_in <= 0 || d_out <= 0) return HWGAT_ERR_SHAPE;
  if (!x3_supported(n, d_in, d_out)) return HWGAT_ERR_UNSUPPORTED;
  if (!dy || (dx && !w) || (dw && !x_planes)) return HWGAT_ERR_NULL;
  if (misaligned(dy) || misaligned(x_planes) || misaligned(w) || misaligned(dx) || misaligned(dw)) return HWGAT_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  int st;
  if ((dx || dw) && (st = linear_x3_bwd(dy, nullptr, w, dx, dw, n, d_in, d_out, s, (const __nv_bfloat16*)x_planes))) return st;
  if (db && (st = colsum_f32(dy, db, n, d_out, s))) return st;
  return 0;
}

int hwgat_smooth_ce_fwd(const float* logits, const long long* target, float* lse, float* row_loss, float* loss,
                        int rows, int classes, float smooth, hwgat_stream_t stream) {
  if (rows <= 0 || classes <= 0 || !(smooth >= 0.f) || smooth > 1.f) return HWGAT_ERR_SHAPE;
  if (!logits || !target || !lse || !row_loss || !loss) return HWGAT_ERR_NULL;
  return smooth_ce_fwd(logits, target, lse, row_loss, loss, rows, classes, smooth, (cudaStream_t)stream);
}

int hwgat_smooth_ce_bwd(const float* logits, const long long* target, const float* lse, const float* g,
                        float* dlogits, int rows, int classes, float smooth, hwgat_stream_t stream) {
  if (rows <= 0 || classes <= 0 || !(smooth >= 0.f) || smooth > 1.f) return HWGAT_ERR_SHAPE;
  if (!logits || !target || !lse || !g || !dlogits) return HWGAT_ERR_NULL;
  return smooth_ce_bwd(logits, target, lse, g, dlogits, rows, classes, smooth, (cudaStream_t)stream);
}

static int bda_ln_fwd_impl(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                     float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                     unsigned long long seed, unsigned long long offset, hwgat_stream_t stream, bool f32) {
  if (n < 0 || bad_p(p)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (n == 0) return HWGAT_OK;
  if (!res || !a0 || !x1) return HWGAT_ERR_NULL;
  if (gamma && (!beta || !y || !mean || !rstd)) return HWGAT_ERR_NULL;
  if (misaligned(res) || misaligned(a0) || misaligned(bias) || misaligned(gamma) || misaligned(beta) ||
      misaligned(x1) || misaligned(y))
    return HWGAT_ERR_ALIGN;
  return launch_bda_ln_fwd(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, p, seed, offset,
                           (cudaStream_t)stream, 0, 0, f32);
}
int hwgat_bda_ln_fwd(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                     float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                     unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bda_ln_fwd_impl(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, p, seed, offset, stream, false);
}
int hwgat_bda_ln_fwd_f32(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                     float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                     unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bda_ln_fwd_impl(res, a0, bias, gamma, beta, x1, y, mean, rstd, n, d, eps, p, seed, offset, stream, true);
}

static int bda_ln_bwd_impl(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                     const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                     long long n, int d, float p, unsigned long long seed, unsigned long long offset,
                     hwgat_stream_t stream, bool f32) {
  if (n < 0 || bad_p(p)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (gamma && (!dgamma || !dbeta)) return HWGAT_ERR_NULL;
  if (n > 0) {
    if (!d_a0) return HWGAT_ERR_NULL;
    if (gamma ? (!dy || !x1 || !mean || !rstd || !d_res) : !g_x1) return HWGAT_ERR_NULL;
  }
  if (misaligned(g_x1) || misaligned(dy) || misaligned(x1) || misaligned(gamma) || misaligned(d_res) ||
      misaligned(d_a0))
    return HWGAT_ERR_ALIGN;
  return launch_bda_ln_bwd(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, p, seed, offset,
                           (cudaStream_t)stream, f32);
}
int hwgat_bda_ln_bwd(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                     const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                     long long n, int d, float p, unsigned long long seed, unsigned long long offset,
                     hwgat_stream_t stream) {
  return bda_ln_bwd_impl(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, p, seed, offset, stream, false);
}
int hwgat_bda_ln_bwd_f32(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                     const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                     long long n, int d, float p, unsigned long long seed, unsigned long long offset,
                     hwgat_stream_t stream) {
  return bda_ln_bwd_impl(g_x1, dy, x1, mean, rstd, gamma, d_res, d_a0, dbias, dgamma, dbeta, n, d, p, seed, offset, stream, true);
}

static int check_gelu(long long n, int cols, float p) {
  if (n < 0 || bad_p(p)) return HWGAT_ERR_SHAPE;
  if (cols != 256 && cols != 512 && cols != 1024) return HWGAT_ERR_UNSUPPORTED;
  return HWGAT_OK;
}

static int bias_gelu_dropout_fwd_impl(const void* u0, const float* bias, void* g, long long n, int cols, float p,
                                unsigned long long seed, unsigned long long offset, hwgat_stream_t stream, bool f32) {
  int st = check_gelu(n, cols, p);
  if (st) return st;
  if (n == 0) return HWGAT_OK;
  if (!u0 || !g) return HWGAT_ERR_NULL;
  if (misaligned(u0) || misaligned(g)) return HWGAT_ERR_ALIGN;
  return launch_bias_gelu_dropout(u0, bias, nullptr, g, nullptr, n, cols, p, seed, offset, false, (cudaStream_t)stream,
                                  f32);
}
int hwgat_bias_gelu_dropout_fwd(const void* u0, const float* bias, void* g, long long n, int cols, float p,
                                unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bias_gelu_dropout_fwd_impl(u0, bias, g, n, cols, p, seed, offset, stream, false);
}
int hwgat_bias_gelu_dropout_fwd_f32(const void* u0, const float* bias, void* g, long long n, int cols, float p,
                                unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  return bias_gelu_dropout_fwd_impl(u0, bias, g, n, cols, p, seed, offset, stream, true);
}

static int bias_gelu_dropout_bwd_impl(const void* u0, const float* bias, const void* dg, void* du0, float* dbias,
                                long long n, int cols, float p, unsigned long long seed, unsigned long long offset,
                                hwgat_stream_t stream, bool f32) {
  int st = check_gelu(n, cols, p);
  if (st) return st;
  if (n == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * cols, (cudaStream_t)stream);
    return HWGAT_OK;
  }
  if (!u0 || !dg || !du0) return HWGAT_ERR_NULL;
  if (misaligned(u0) || misaligned(dg) || misaligned(du0)) return HWGAT_ERR_ALIGN;
  return launch_bias_gelu_dropout(u0, bias, dg, du0, dbias, n, cols, p, seed, offset, true, (cudaStream_t)stream, f32);
}
int hwgat_bias_gelu_dropout_bwd(const void* u0, const float* bias, const void* dg, void* du0, float* dbias,
                                long long n, int cols, float p, unsigned long long seed, unsigned long long offset,
                                hwgat_stream_t stream) {
  return bias_gelu_dropout_bwd_impl(u0, bias, dg, du0, dbias, n, cols, p, seed, offset, stream, false);
}
int hwgat_bias_gelu_dropout_bwd_f32(const void* u0, const float* bias, const void* dg, void* du0, float* dbias,
                                long long n, int cols, float p, unsigned long long seed, unsigned long long offset,
                                hwgat_stream_t stream) {
  return bias_gelu_dropout_bwd_impl(u0, bias, dg, du0, dbias, n, cols, p, seed, offset, stream, true);
}

static int check_ffn(long long n, int d, int hidden) {
  if (n < 0 || d <= 0 || hidden <= 0) return HWGAT_ERR_SHAPE;
  if (n % 128 || d % 128 || hidden % 128 || hidden > 2048 || n > 0x7fffffffLL) return HWGAT_ERR_UNSUPPORTED;
  return HWGAT_OK;
}

int hwgat_ffn_fused_supported(long long n, int d, int hidden) {
  return check_ffn(n, d, hidden) == 0 && ffn_eval_fused_supported(n, d, hidden) ? 1 : 0;
}

int hwgat_ffn_fwd(const void* h, const void* w1, const float* b1, const void* w2, void* act, void* gp, void* v0,
                  long long n, int d, int hidden, float p, unsigned long long seed, unsigned long long offset,
                  hwgat_stream_t stream) {
  int st = check_ffn(n, d, hidden);
  if (st) return st;
  if (bad_p(p)) return HWGAT_ERR_SHAPE;
  if (n == 0) return HWGAT_OK;
  if (!h || !w1 || !w2 || !v0) return HWGAT_ERR_NULL;
  if (misaligned(h) || misaligned(w1) || misaligned(b1) || misaligned(w2) || misaligned(act) || misaligned(gp) ||
      misaligned(v0))
    return HWGAT_ERR_ALIGN;
  if (!act) {   // inference, activation not wanted: the one-kernel form (ffn_fused.cu), where its shape allows
    if (gp || p != 0.f) return HWGAT_ERR_NULL;
    if (!ffn_eval_fused_supported(n, d, hidden)) return HWGAT_ERR_UNSUPPORTED;
    return ffn_eval_fused((const __nv_bfloat16*)h, (const __nv_bfloat16*)w1, b1, (const __nv_bfloat16*)w2,
                          (__nv_bfloat16*)v0, n, d, hidden, (cudaStream_t)stream);
  }
  return ffn_fwd((const __nv_bfloat16*)h, (const __nv_bfloat16*)w1, b1, (const __nv_bfloat16*)w2, (__nv_bfloat16*)act,
                 (__nv_bfloat16*)gp, (__nv_bfloat16*)v0, n, d, hidden, p, seed, offset, (cudaStream_t)stream);
}

size_t hwgat_ffn_bwd_workspace_bytes(long long n, int d, int hidden) {
  if (check_ffn(n, d, hidden)) return 0;
  return ffn_bwd_workspace_bytes(n, d, hidden);
}

int hwgat_ffn_bwd(const void* dv0, const void* h, const void* act, const void* gp, const void* w1, const void* w2,
                  void* dh, float* dw1, float* db1, float* dw2, void* workspace, size_t workspace_bytes, long long n,
                  int d, int hidden, hwgat_stream_t stream) {
  int st = check_ffn(n, d, hidden);
  if (st) return st;
  if (!dw1 || !db1 || !dw2) return HWGAT_ERR_NULL;
  if (n == 0) {  // empty batch: parameter gradients are zero
    cudaMemsetAsync(dw1, 0, sizeof(float) * (size_t)hidden * d, (cudaStream_t)stream);
    cudaMemsetAsync(dw2, 0, sizeof(float) * (size_t)hidden * d, (cudaStream_t)stream);
    cudaMemsetAsync(db1, 0, sizeof(float) * hidden, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!dv0 || !h || !act || !gp || !w1 || !w2 || !dh) return HWGAT_ERR_NULL;
  if (misaligned(dv0) || misaligned(h) || misaligned(act) || misaligned(gp) || misaligned(w1) || misaligned(w2) ||
      misaligned(dh) || misaligned(dw1) || misaligned(dw2) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < ffn_bwd_workspace_bytes(n, d, hidden)) return HWGAT_ERR_WORKSPACE;
  return ffn_bwd((const __nv_bfloat16*)dv0, (const __nv_bfloat16*)h, (const __nv_bfloat16*)act,
                 (const __nv_bfloat16*)gp, (const __nv_bfloat16*)w1, (const __nv_bfloat16*)w2, (__nv_bfloat16*)dh, dw1,
                 db1, dw2, workspace, n, d, hidden, (cudaStream_t)stream);
}

int hwgat_embed_fwd(const float* x, const float* Bm, const float* pe, float* out, long long n, int C, int E, int K,
                    int T, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream) {
  if (n < 0 || C <= 0 || E <= 0 || K <= 0 || T <= 0 || bad_p(p)) return HWGAT_ERR_SHAPE;
  if (E % 8) return HWGAT_ERR_UNSUPPORTED;
  if (n == 0) return HWGAT_OK;
  if (!x || !Bm || !pe || !out) return HWGAT_ERR_NULL;
  if (misaligned(pe) || misaligned(out)) return HWGAT_ERR_ALIGN;
  return launch_embed_fwd(x, Bm, pe, out, n, C, E, K, T, p, seed, offset, (cudaStream_t)stream);
}

size_t hwgat_ln_pool_scratch_bytes(int B, int tokens, int d) {
  if (B <= 0 || tokens <= 0 || (d != 128 && d != 256 && d != 512)) return 0;
  return ln_pool_scratch_bytes(B, tokens, d);
}

static bool bad_pad(int tokens, int kp_real, int kp_pad) {
  if (kp_real == 0 && kp_pad == 0) return false;
  return kp_real <= 0 || kp_pad < kp_real || tokens % kp_real != 0;
}

int hwgat_ln_pool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean,
                      float* rstd, void* scratch, size_t scratch_bytes, int B, int tokens, int d, float eps,
                      int kp_real, int kp_pad, hwgat_stream_t stream) {
  if (B < 0 || tokens <= 0 || bad_pad(tokens, kp_real, kp_pad)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (B == 0) return HWGAT_OK;
  if (!x || !gamma || !beta || !pooled || !mean || !rstd) return HWGAT_ERR_NULL;
  if (misaligned(x) || misaligned(gamma)) return HWGAT_ERR_ALIGN;
  const size_t need = ln_pool_scratch_bytes(B, tokens, d);
  if (need > 0 && (!scratch || scratch_bytes < need)) return HWGAT_ERR_WORKSPACE;
  return launch_ln_pool_fwd(x, gamma, beta, pooled, mean, rstd, (float*)scratch, B, tokens, d, eps, (cudaStream_t)stream,
                            kp_real, kp_pad);
}

int hwgat_ln_pool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                      float* dx, float* dgamma, int B, int tokens, int d, int kp_real, int kp_pad,
                      hwgat_stream_t stream) {
  if (B < 0 || tokens <= 0 || bad_pad(tokens, kp_real, kp_pad)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (!dgamma) return HWGAT_ERR_NULL;
  if (B > 0 && (!g || !x || !mean || !rstd || !gamma || !dx)) return HWGAT_ERR_NULL;
  if (misaligned(g) || misaligned(x) || misaligned(gamma) || misaligned(dx)) return HWGAT_ERR_ALIGN;
  return launch_ln_pool_bwd(g, x, mean, rstd, gamma, dx, dgamma, B, tokens, d, (cudaStream_t)stream, kp_real, kp_pad);
}

// the weighted pool of GATE (GATE.py:207): pooled = sum_t tok_w[t] * LayerNorm(x)[t]  (+ the caller's folded bias in beta)
int hwgat_ln_wpool_fwd(const float* x, const float* gamma, const float* beta, const float* tok_w, float* pooled,
                       float* mean, float* rstd, void* scratch, size_t scratch_bytes, int B, int tokens, int d,
                       float eps, int kp_real, int kp_pad, hwgat_stream_t stream) {
  if (!tok_w) return HWGAT_ERR_NULL;
  if (B < 0 || tokens <= 0 || bad_pad(tokens, kp_real, kp_pad)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (B == 0) return HWGAT_OK;
  if (!x || !gamma || !beta || !pooled || !mean || !rstd) return HWGAT_ERR_NULL;
  if (misaligned(x) || misaligned(gamma)) return HWGAT_ERR_ALIGN;
  const size_t need = ln_pool_scratch_bytes(B, tokens, d);
  if (need > 0 && (!scratch || scratch_bytes < need)) return HWGAT_ERR_WORKSPACE;
  return launch_ln_pool_fwd(x, gamma, beta, pooled, mean, rstd, (float*)scratch, B, tokens, d, eps, (cudaStream_t)stream,
                            kp_real, kp_pad, tok_w);
}

int hwgat_ln_wpool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* tok_w, float* dx, float* dgamma, float* d_tok_w, float* dw_part, int B, int tokens,
                       int d, int kp_real, int kp_pad, hwgat_stream_t stream) {
  if (B < 0 || tokens <= 0 || bad_pad(tokens, kp_real, kp_pad)) return HWGAT_ERR_SHAPE;
  if (d != 128 && d != 256 && d != 512) return HWGAT_ERR_UNSUPPORTED;
  if (!dgamma || !d_tok_w || !tok_w) return HWGAT_ERR_NULL;
  if (B > 0 && (!g || !x || !mean || !rstd || !gamma || !dx || !dw_part)) return HWGAT_ERR_NULL;
  if (misaligned(g) || misaligned(x) || misaligned(gamma) || misaligned(dx)) return HWGAT_ERR_ALIGN;
  return launch_ln_pool_bwd(g, x, mean, rstd, gamma, dx, dgamma, B, tokens, d, (cudaStream_t)stream, kp_real, kp_pad,
                            tok_w, dw_part, d_tok_w);
}

int hwgat_adamw_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const long long* sizes, double lr, double beta1, double beta2, double eps,
                     double weight_decay, long long step, float grad_scale, hwgat_stream_t stream) {
  if (n_tensors < 0 || step < 1) return HWGAT_ERR_SHAPE;
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0) || !(lr >= 0.0) ||
      !(weight_decay >= 0.0))
    return HWGAT_ERR_SHAPE;
  if (n_tensors == 0) return HWGAT_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !sizes) return HWGAT_ERR_NULL;
  for (int t = 0; t < n_tensors; ++t) {
    if (sizes[t] < 0) return HWGAT_ERR_SHAPE;
    if (sizes[t] > 0 && (!params[t] || !grads[t] || !exp_avg[t] || !exp_avg_sq[t])) return HWGAT_ERR_NULL;
  }
  return adamw_step(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, lr, beta1, beta2, eps, weight_decay, step,
                    grad_scale, (cudaStream_t)stream);
}

static int check_proj(long long n, int d_in, int d_out, bool backward) {
  if (n < 0 || d_in <= 0 || d_out <= 0) return HWGAT_ERR_SHAPE;
  if (n % 128 || d_out % 128 || d_in % 64 || (backward && d_in % 128)) return HWGAT_ERR_UNSUPPORTED;
  return HWGAT_OK;
}

int hwgat_proj_fwd(const void* ctx, const void* w, void* y, long long n, int d_in, int d_out, hwgat_stream_t stream) {
  int st = check_proj(n, d_in, d_out, false);
  if (st) return st;
  if (n == 0) return HWGAT_OK;
  if (!ctx || !w || !y) return HWGAT_ERR_NULL;
  if (misaligned(ctx) || misaligned(w) || misaligned(y)) return HWGAT_ERR_ALIGN;
  return gemm_tc_nt_epi_none((const __nv_bfloat16*)ctx, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, n, d_out, d_in,
                             (cudaStream_t)stream);
}

size_t hwgat_proj_bwd_workspace_bytes(int d_in, int d_out) {
  if (d_in <= 0 || d_out <= 0) return 0;
  return (size_t)d_in * d_out * sizeof(__nv_bfloat16);
}

int hwgat_proj_bwd(const void* dy, const void* ctx, const void* w, void* d_ctx, float* dw, void* workspace,
                   size_t workspace_bytes, long long n, int d_in, int d_out, hwgat_stream_t stream) {
  int st = check_proj(n, d_in, d_out, true);
  if (st) return st;
  if (!dw) return HWGAT_ERR_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d_in * d_out, s);
    return (int)cudaGetLastError();
  }
  if (!dy || !ctx || !w || !d_ctx) return HWGAT_ERR_NULL;
  if (misaligned(dy) || misaligned(ctx) || misaligned(w) || misaligned(d_ctx) || misaligned(dw) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  if (!workspace || workspace_bytes < hwgat_proj_bwd_workspace_bytes(d_in, d_out)) return HWGAT_ERR_WORKSPACE;
  __nv_bfloat16* wt = (__nv_bfloat16*)workspace;   // W^T [d_in, d_out]: both operands of d_ctx = dy . W K-major
  if ((st = transpose_bf16((const __nv_bfloat16*)w, wt, d_out, d_in, s))) return st;
  if ((st = gemm_tc_nt_epi_none((const __nv_bfloat16*)dy, wt, (__nv_bfloat16*)d_ctx, n, d_in, d_out, s))) return st;
  return gemm_tc_tn((const __nv_bfloat16*)dy, (const __nv_bfloat16*)ctx, dw, nullptr, d_out, d_in, n, s);
}

int hwgat_debug_gemm_nt(const void* A, const void* Bt, void* C, long long M, int N, int K, hwgat_stream_t stream) {
  if (!A || !Bt || !C) return HWGAT_ERR_NULL;
  if (misaligned(A) || misaligned(Bt) || misaligned(C)) return HWGAT_ERR_ALIGN;
  if (M <= 0 || N <= 0 || K <= 0) return HWGAT_ERR_SHAPE;
  return gemm_tc_nt_epi_none((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bt, (__nv_bfloat16*)C, M, N, K,
                             (cudaStream_t)stream);
}

int hwgat_debug_gemm_tn(const void* A, const void* B, float* C, float* colsum, int M, int N, long long Kd,
                        hwgat_stream_t stream) {
  if (!A || !B || !C) return HWGAT_ERR_NULL;   /* colsum may be NULL: no column sums */
  if (misaligned(A) || misaligned(B) || misaligned(C)) return HWGAT_ERR_ALIGN;
  if (M <= 0 || N <= 0 || Kd <= 0) return HWGAT_ERR_SHAPE;
  return gemm_tc_tn((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, colsum, M, N, Kd, (cudaStream_t)stream);
}

int hwgat_debug_set_gemm_pair(int on) { return set_gemm_pair(on != 0) ? 1 : 0; }

static int merge_common(const void* src, void* dst, int B, int F, int K, int d, int TP, int dtype, bool backward,
                        hwgat_stream_t stream) {
  if (B < 0 || F <= 0 || K <= 0 || d <= 0) return HWGAT_ERR_SHAPE;
  if (TP != kTP || F % TP != 0) return HWGAT_ERR_UNSUPPORTED;
  if (dtype != HWGAT_F32 && dtype != HWGAT_BF16) return HWGAT_ERR_UNSUPPORTED;
  const int eb = dtype == HWGAT_F32 ? 4 : 2;
  if ((d * eb) % 16 != 0) return HWGAT_ERR_UNSUPPORTED;
  if (B == 0) return HWGAT_OK;
  if (!src || !dst) return HWGAT_ERR_NULL;
  if (misaligned(src) || misaligned(dst)) return HWGAT_ERR_ALIGN;
  return launch_merge(src, dst, B, F, K, d, eb, backward, (cudaStream_t)stream);
}

int hwgat_attn_bwd_f32_kept(const float* d_out, const float* xn, const float* w_qkv, const float* qkv,
                            const uint32_t* bits, float threshold, float* d_xn, float* d_w, float* d_b, void* workspace,
                            size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                            int layout, hwgat_stream_t stream) {
  int st = check_geometry(B, F, K, d, heads, W, TP, shift, layout, HWGAT_F32);
  if (st) return st;
  if (!d_w || !d_b) return HWGAT_ERR_NULL;
  if (B == 0) {
    cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, (cudaStream_t)stream);
    cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, (cudaStream_t)stream);
    return (int)cudaGetLastError();
  }
  if (!d_out || !xn || !w_qkv || !qkv || !bits || !d_xn) return HWGAT_ERR_NULL;
  if (misaligned(d_out) || misaligned(xn) || misaligned(w_qkv) || misaligned(qkv) || misaligned(d_xn) ||
      misaligned(d_w) || misaligned(workspace))
    return HWGAT_ERR_ALIGN;
  const size_t need = hwgat_attn_workspace_bytes(B, F, K, d, heads, HWGAT_F32, 0);     // dqkv: as large as qkv
  if (!workspace || workspace_bytes < need) return HWGAT_ERR_WORKSPACE;
  AttnArgs a{};
  a.xn = xn; a.w_qkv = w_qkv; a.b_qkv = nullptr; a.bits = bits; a.threshold = threshold; a.d_out = d_out;
  a.d_xn = d_xn; a.d_w = d_w; a.d_b = d_b; a.workspace = workspace;
  a.B = B; a.F = F; a.K = K; a.d = d; a.heads = heads; a.shift = shift; a.layout = layout;
  return attn_bwd_f32(a, (cudaStream_t)stream, qkv);
}

int hwgat_merge_fwd(const void* x, void* out, int B, int F, int K, int d, int TP, int dtype, hwgat_stream_t stream) {
  return merge_common(x, out, B, F, K, d, TP, dtype, false, stream);
}
int hwgat_merge_bwd(const void* d_out, void* d_x, int B, int F, int K, int d, int TP, int dtype,
                    hwgat_stream_t stream) {
  return merge_common(d_out, d_x, B, F, K, d, TP, dtype, true, stream);
}

}  // extern "C"
