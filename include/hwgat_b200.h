/*
 * hwgat_b200 - C ABI of the B200-native HWGATE windowed-graph-attention path.
 *
 * Drop-in boundary for the hot path of suvajit-patra/sl-hwgat
 * (hwgat/models/HWGATE.py + the adjacency builder of
 * hwgat/models/model_params.py).  The reference is pure PyTorch and has no FFI
 * of its own; each entry point below names the reference code it replaces.
 * The Python side (sl_hwgat_b200/_lib.py) binds these with ctypes and passes
 * tensor.data_ptr() values and the current CUDA stream.
 *
 * Conventions
 *  - plain C symbols, no C++/torch types; all pointers are DEVICE pointers owned
 *    by the caller; the library never allocates, frees or synchronises;
 *  - every call is asynchronous on `stream` and re-entrant across streams;
 *  - return value: 0 = ok; 1..999 = cudaError_t of the failed launch;
 *    >= 1000 = argument error (see hwgat_error_string);
 *  - activations: layout HWGAT_LAYOUT_BFKD is the reference's un-partitioned,
 *    un-rolled (B, F, K, d) row-major tensor: the cyclic frame shift
 *    (HWGATE.py:197-200, 210-215) and window_partition / window_reverse
 *    (HWGATE.py:30-47) are index arithmetic inside the kernels, never copies.
 *    HWGAT_LAYOUT_WINDOWS is the already-partitioned (B*f*nW, TP*W, d) tensor
 *    that MSA.forward receives (HWGATE.py:84); shift must be 0 there because
 *    the caller has already rolled;
 *  - dtype: HWGAT_F32 = every tensor float32 (the 1e-5 parity mode);
 *    HWGAT_BF16 = activations and weights bfloat16, biases / reductions /
 *    weight gradients float32 (the timed mode);
 *  - supported geometry: temporal_patch TP = 2, window W = 16 (N = 32 tokens per
 *    window), head_dim 64, d = heads*64 <= 512 (bf16: d a multiple of 128), F even, K a multiple of 64.
 *    Anything else returns HWGAT_ERR_UNSUPPORTED - there is no fallback path.
 */
#ifndef HWGAT_B200_H
#define HWGAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hwgat_stream_t; /* cudaStream_t */

enum { HWGAT_F32 = 0, HWGAT_BF16 = 1 };
enum { HWGAT_LAYOUT_BFKD = 0, HWGAT_LAYOUT_WINDOWS = 1 };

enum {
  HWGAT_OK = 0,
  HWGAT_ERR_NULL = 1000,        /* required pointer is NULL              */
  HWGAT_ERR_SHAPE = 1001,       /* inconsistent sizes                    */
  HWGAT_ERR_UNSUPPORTED = 1002, /* geometry / dtype outside the kernels  */
  HWGAT_ERR_WORKSPACE = 1003,   /* workspace smaller than required       */
  HWGAT_ERR_ALIGN = 1004        /* pointer not 16-byte aligned           */
};

/* ABI version; bumped on any signature change. */
int hwgat_version(void);
const char* hwgat_error_string(int status);

/* K1a - skeleton adjacency of every keypoint window.
 * Replaces HWGATEParams.get_adj / get_adj_mat (model_params.py:373-400).
 * edges: int32 (nW, n_edges, 2) undirected pairs inside one W-keypoint window.
 * adj  : float32 (nW, TP*W, TP*W), 1.0 / 0.0, self loops included; a joint is
 *        linked to itself in the adjacent frame, nothing further apart.        */
int hwgat_adjacency_build(const int32_t* edges, int n_edges, int nW, int W, int TP,
                          float* adj, hwgat_stream_t stream);

/* K1b - packed attention mask of one block.
 * Replaces the adjacency replication (HWGATE.py:309), the shifted-window mask
 * (HWGATE.py:169-187) and their two multiplies (HWGATE.py:102-108).
 * adj : float32 (nW, N, N), N = TP*W, any non-zero = edge.
 * bits: uint32 (F/TP * nW, N, N/32): bit j of row i of window fi*nW+w = key j
 *       visible to query i.  shift = 0 for even blocks, TP/2 for odd ones.     */
int hwgat_mask_build(const float* adj, int nW, int W, int TP, int F, int shift,
                     uint32_t* bits, hwgat_stream_t stream);

/* K1c - pack arbitrary float masks, for callers of MSA.forward that hand in
 * their own `mask` / `adj_mat` tensors (HWGATE.py:84, 102-108).
 * adj : float32 (adj_windows, N, N) or NULL; window `win` uses adj[win % adj_windows].
 * mask: float32 (n_windows, N, N) or NULL.
 * bits: uint32 (n_windows, N, N/32) = (adj != 0) AND (mask != 0).                */
int hwgat_mask_pack(const float* adj, int adj_windows, const float* mask, int n_windows, int N,
                    uint32_t* bits, hwgat_stream_t stream);

/* Bytes of scratch hwgat_attn_fwd / hwgat_attn_bwd need for these sizes. */
size_t hwgat_attn_workspace_bytes(int B, int F, int K, int d, int heads, int dtype, int backward);

/* K2 - fused windowed graph attention, forward.
 * Replaces torch.roll + window_partition + MSA.forward up to (not including)
 * self.proj + window_reverse + roll back: HWGATE.py:197-201, 86-114, 207-215.
 * xn    : (B,F,K,d) normalised residual stream (output of norm1), dtype.
 * w_qkv : (3d, d) dtype, rows [q | k | v], head-major inside each (HWGATE.py:76,86).
 * b_qkv : (3d) float32.
 * bits  : mask of hwgat_mask_build for this F and shift.
 * threshold: < 0 = eval mode.  >= 0 = training mode: logits whose unmasked
 *         softmax exceeds it are zeroed first (HWGATE.py:94-100).
 * out   : (B,F,K,d) dtype, head-merged context, same token order as xn.
 * workspace: hwgat_attn_workspace_bytes(..., backward=0) bytes.                */
int hwgat_attn_fwd(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits,
                   float threshold, void* out, void* workspace, size_t workspace_bytes,
                   int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout,
                   int dtype, hwgat_stream_t stream);

/* K3 - fused backward of K2.  Recomputes Q, K, V, the logits and the
 * probabilities from xn (nothing but xn is saved by the forward).
 * d_out : (B,F,K,d) dtype, gradient of `out`.
 * d_xn  : (B,F,K,d) dtype, overwritten.
 * d_w   : (3d, d) float32, overwritten.   d_b: (3d) float32, overwritten.
 * The gradient w.r.t. the logits is  live ? P*(dP - sum_j P*dP) : 0  where
 * live = mask AND keep AND logit != 0 (autograd of HWGATE.py:100-111).          */
int hwgat_attn_bwd(const void* d_out, const void* xn, const void* w_qkv, const float* b_qkv,
                   const uint32_t* bits, float threshold, void* d_xn, float* d_w, float* d_b,
                   void* workspace, size_t workspace_bytes,
                   int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout,
                   int dtype, hwgat_stream_t stream);


/* fp32 only: the backward of hwgat_attn_fwd(..., HWGAT_F32, ...) when the caller KEPT the forward's workspace, which
 * holds the projected rows qkv (n, 3d) fp32: no re-projection (one x3 GEMM and its splits less per block), and the
 * workspace here holds dqkv only: hwgat_attn_workspace_bytes(B, F, K, d, heads, HWGAT_F32, 0) bytes.               */
int hwgat_attn_bwd_f32_kept(const float* d_out, const float* xn, const float* w_qkv, const float* qkv,
                            const uint32_t* bits, float threshold, float* d_xn, float* d_w, float* d_b, void* workspace,
                            size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                            int layout, hwgat_stream_t stream);

/* K4 - stage transition.  Replaces TemporalMerging.forward (HWGATE.py:55-63):
 * out[b,fi,k,tp*d+e] = x[b,fi*TP+tp,k,e];  x: (B,F,K,d) -> out: (B,F/TP,K,TP*d). */
int hwgat_merge_fwd(const void* x, void* out, int B, int F, int K, int d, int TP,
                    int dtype, hwgat_stream_t stream);
/* Adjoint of K4: d_out (B,F/TP,K,TP*d) -> d_x (B,F,K,d). */
int hwgat_merge_bwd(const void* d_out, void* d_x, int B, int F, int K, int d, int TP,
                    int dtype, hwgat_stream_t stream);

/* ---- K2b / K3b: the same attention for ANY window of W in {16, 32, 64} keypoints x TP = 2 frames (N = 32, 64, 128
 * tokens; the reference takes window_size at model_params.py:254 and runs W = 32 / 64 unchanged, HWGATE.py:30-36,
 * 290-291), bf16 only.  The QKV projection is one tcgen05 GEMM with the bias in its epilogue; the attention core
 * runs S = QK^T, O = PV (and dP, dV, dQ, dK in the backward) as block-diagonal M = 128 tcgen05 tiles with one thread
 * per query row for the masked softmax.  `qkv` (n x 3d bf16, n = B*F*K) is written by the forward and may be handed
 * back to the backward (NULL: the backward re-projects it into its workspace).  bits: hwgat_mask_build /
 * hwgat_mask_pack for this W (N/32 words per row).  attn_p > 0: attention dropout, self.attn_drop on the
 * probabilities (HWGATE.py:112), from the Philox stream (seed, offset); the backward regenerates the same mask.  */
size_t hwgat_attn2_workspace_bytes(int B, int F, int K, int d, int heads, int backward, int have_qkv);
int hwgat_attn2_fwd(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                    void* out, void* qkv, void* workspace, size_t workspace_bytes, int B, int F, int K, int d,
                    int heads, int W, int TP, int shift, int layout, float attn_p, unsigned long long seed,
                    unsigned long long offset, hwgat_stream_t stream);
int hwgat_attn2_bwd(const void* d_out, const void* xn, const void* w_qkv, const float* b_qkv, const void* qkv,
                    const uint32_t* bits, float threshold, void* d_xn, float* d_w, float* d_b, void* workspace,
                    size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                    int layout, int qk_perm, float attn_p, unsigned long long seed, unsigned long long offset,
                    hwgat_stream_t stream);
/* Hybrid for the reference window: K2 (hwgat_attn_fwd, bf16) that also KEEPS the q (scaled), k, v rows it formed in
 * `qkv` (n x 3d bf16; q and k with the columns of every head permuted like K3's dQKV, v plain), so that the backward
 * can be hwgat_attn2_bwd(..., qkv, ..., qk_perm = 1, ...) - K3b's tcgen05 core without the QKV recompute of K3 and
 * without a projection GEMM.  Workspace as hwgat_attn_fwd.                                                        */
int hwgat_attn_fwd_keep(const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                        void* out, void* qkv, void* workspace, size_t workspace_bytes, int B, int F, int K, int d,
                        int heads, int W, int TP, int shift, int layout, hwgat_stream_t stream);

/* ---- rest of the block, bf16 / autocast path --------------------------
 * Bandwidth-bound fusions of the PyTorch elementwise chains of PartAttentionBlock.forward
 * (HWGATE.py:189-221) and their autograd.  Dropout masks are regenerated from a Philox4x32-7
 * stream (seed, offset) in forward and backward; nothing is stored.                             */

/* K5: y(bf16) = LayerNorm(x fp32; gamma, beta, eps) per row of d in {128,256,512}; mean/rstd (n) saved.
 * Replaces self.norm1 / self.norm2 (HWGATE.py:203, 219) + the autocast cast of their output.   */
int hwgat_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                 long long n, int d, float eps, hwgat_stream_t stream);
/* K5': dx(fp32) = (dres ? dres : 0) + LayerNorm'(dy bf16); dgamma, dbeta (d) overwritten.
 * dres is the gradient arriving on the residual path (HWGATE.py:217, 219), fused into the same pass. */
int hwgat_ln_bwd(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                 const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d,
                 hwgat_stream_t stream);
/* K6: x1(fp32) = res(fp32) + dropout_p(a0(bf16) + bias) and, when gamma != NULL, the LayerNorm that consumes
 * x1 next: y(bf16) = LayerNorm(x1; gamma, beta, eps), mean / rstd (n) saved.  a0 is a Linear output WITHOUT its
 * bias (bias may be NULL).  Replaces proj bias + proj_drop + shortcut (HWGATE.py:115-116, 217) fused with norm2
 * (:219), and fc2 bias + ff.drop + residual (:134-135, :219) fused with the next block's norm1 (:203).         */
int hwgat_bda_ln_fwd(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                     float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                     unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
/* K6': gamma != NULL: d_res(fp32) = (g_x1 ? g_x1 : 0) + LayerNorm'(dy bf16), dgamma, dbeta overwritten;
 *      gamma == NULL: the gradient w.r.t. res is g_x1 itself and d_res is not written.
 *      d_a0(bf16) = mask * that / (1-p);  dbias (d, may be NULL) = column sums of d_a0.                          */
int hwgat_bda_ln_bwd(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                     const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                     long long n, int d, float p, unsigned long long seed, unsigned long long offset,
                     hwgat_stream_t stream);
/* K6 with TemporalMerging.forward (HWGATE.py:55-63) folded in: the last K6 of a level (no LayerNorm follows inside the
 * level) stores x1 = res + dropout_p(a0 + bias) directly in the merged layout (B, F/2, K, 2d); res and a0 are
 * (B, F, K, d) = n rows of d in {128, 256}.  K4 is then not launched.                                              */
int hwgat_bda_merge_fwd(const float* res, const void* a0, const float* bias, float* x_merged, long long n, int d, int F,
                        int K, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
/* K5' with the adjoint of TemporalMerging folded in: the LayerNorm-backward of the first norm1 of a level, run on the
 * merged rows (n_merged rows of d_merged in {256, 512} columns; x, dy, dres in the merged layout (B, F_merged, K,
 * d_merged)), stores dx UN-merged as (B, 2 F_merged, K, d_merged/2).  K4's adjoint is then not launched.            */
int hwgat_ln_bwd_unmerge(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, float* dgamma, float* dbeta, long long n_merged, int d_merged,
                         int F_merged, int K, hwgat_stream_t stream);

/* The fp32 path's forms of K5 - K7 (the reference's unmodified loop runs without autocast): identical arithmetic and
 * arguments, but every ACTIVATION tensor that is bf16 above - y, dy, a0, d_a0, u0, g, dg, du0 - is float32 here, so
 * that the chain LayerNorm -> Linear (x3 tcgen05 GEMM, hwgat_linear_f32_*) -> bias / dropout / residual / LayerNorm
 * stays in fp32 end to end (HWGATE.py:203-219, 130-134 as written, no autocast).                                   */
int hwgat_ln_fwd_f32(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                 long long n, int d, float eps, hwgat_stream_t stream);
int hwgat_ln_bwd_f32(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                 const float* gamma, float* dx, float* dgamma, float* dbeta, long long n, int d,
                 hwgat_stream_t stream);
int hwgat_ln_bwd_unmerge_f32(const void* dy, const float* dres, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, float* dgamma, float* dbeta, long long n_merged, int d_merged,
                         int F_merged, int K, hwgat_stream_t stream);
int hwgat_bda_merge_fwd_f32(const float* res, const void* a0, const float* bias, float* x_merged, long long n, int d, int F,
                        int K, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
int hwgat_bda_ln_fwd_f32(const float* res, const void* a0, const float* bias, const float* gamma, const float* beta,
                     float* x1, void* y, float* mean, float* rstd, long long n, int d, float eps, float p,
                     unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
int hwgat_bda_ln_bwd_f32(const float* g_x1, const void* dy, const float* x1, const float* mean, const float* rstd,
                     const float* gamma, float* d_res, void* d_a0, float* dbias, float* dgamma, float* dbeta,
                     long long n, int d, float p, unsigned long long seed, unsigned long long offset,
                     hwgat_stream_t stream);
int hwgat_bias_gelu_dropout_fwd_f32(const void* u0, const float* bias, void* g, long long n, int cols, float p,
                                unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
int hwgat_bias_gelu_dropout_bwd_f32(const void* u0, const float* bias, const void* dg, void* du0, float* dbias,
                                long long n, int cols, float p, unsigned long long seed, unsigned long long offset,
                                hwgat_stream_t stream);

/* K13: y(fp32) = x(fp32) . w^T + bias: the classifier head self.head (HWGATE.py:359) on the fp32 FFMA GEMM
 * (0.1 - 1 GFLOP; its output feeds a log-softmax, so it stays fp32 also under autocast).  bias may be NULL.       */
int hwgat_linear_f32_fwd(const float* x, const float* w, const float* bias, float* y, int n, int d_in, int d_out,
                         hwgat_stream_t stream);
/* K13': dx = dy . w, dw = dy^T . x, db = column sums of dy; each output may be NULL (skipped).                    */
int hwgat_linear_f32_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int n,
                         int d_in, int d_out, hwgat_stream_t stream);

/* The x3 form of the two entries above with the split of x kept between them: hwgat_linear_x3_fwd also writes the three
 * bf16 planes of x (hi, mid, lo: [3][n][d_in], caller-owned) and hwgat_linear_x3_bwd takes them instead of x - the
 * weight gradient contracts the same planes, so x is neither split twice nor kept in fp32.  Shapes for which
 * hwgat_linear_x3_supported() is 0 return HWGAT_ERR_UNSUPPORTED (use hwgat_linear_f32_*).  dx, dw, db may be NULL.  */
int hwgat_linear_x3_supported(long long n, int d_in, int d_out);
int hwgat_linear_x3_fwd(const float* x, const float* w, const float* bias, float* y, void* x_planes, long long n, int d_in,
                        int d_out, hwgat_stream_t stream);
int hwgat_linear_x3_bwd(const float* dy, const void* x_planes, const float* w, float* dx, float* dw, float* db,
                        long long n, int d_in, int d_out, hwgat_stream_t stream);
/* K14: label-smoothed cross entropy, SmoothedCrossEntropyLoss.forward (losses/SmoothCrossEntropy.py:35-39):
 * loss = mean_b [(1-smooth) * (-logp[b, target_b]) + smooth * (-mean_c logp[b, c])]; logits (rows, classes) fp32,
 * target int64; lse and row_loss (rows) are saved / scratch; the mean is a fixed-order sum (deterministic).        */
int hwgat_smooth_ce_fwd(const float* logits, const long long* target, float* lse, float* row_loss, float* loss,
                        int rows, int classes, float smooth, hwgat_stream_t stream);
/* K14': dlogits = (*g / rows) * (softmax(logits) - (1-smooth) * onehot(target) - smooth / classes); g: device scalar. */
int hwgat_smooth_ce_bwd(const float* logits, const long long* target, const float* lse, const float* g,
                        float* dlogits, int rows, int classes, float smooth, hwgat_stream_t stream);
/* K7: g(bf16) = dropout_p(gelu(u0(bf16) + bias)), exact erf GELU, u0: (n, cols) a Linear output without bias.
 * Replaces fc1 bias + ff.act + ff.drop (HWGATE.py:131-133).  cols in {256, 512, 1024}.                          */
int hwgat_bias_gelu_dropout_fwd(const void* u0, const float* bias, void* g, long long n, int cols, float p,
                                unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
/* K7': du0(bf16) = dg * mask / (1-p) * gelu'(u0 + bias);  dbias (cols, may be NULL) = column sums of du0.        */
int hwgat_bias_gelu_dropout_bwd(const void* u0, const float* bias, const void* dg, void* du0, float* dbias,
                                long long n, int cols, float p, unsigned long long seed, unsigned long long offset,
                                hwgat_stream_t stream);

/* K10: the FeedForward of a block without its last bias / dropout (those belong to K6), on tcgen05:
 *   act(bf16, (n, hidden)) = dropout_p(gelu(h . W1^T + b1))   bias + exact erf GELU + dropout in the GEMM epilogue
 *   gp (bf16, (n, hidden)) = d act / d (h . W1^T + b1)         dropout mask and 1/(1-p) folded in; NULL = do not save
 *   v0 (bf16, (n, d))      = act . W2^T
 * Replaces ff.fc1 + ff.act + ff.drop + ff.fc2's matmul (HWGATE.py:130-134).  h: (n, d) bf16, W1: (hidden, d) bf16,
 * b1: (hidden) fp32 or NULL, W2: (d, hidden) bf16.  n % 128 == 0, d % 128 == 0, hidden % 128 == 0, hidden <= 2048.
 * Inference: gp == NULL and p == 0 take a GELU-only epilogue; with act == NULL as well (activation not wanted) and
 * hwgat_ffn_fused_supported(n, d, hidden) != 0 (d = 128 / 256, hidden = 2 d) the whole FeedForward is ONE kernel
 * whose activation tile stays in shared memory (K10f): h is read once, v0 written once.                          */
int hwgat_ffn_fused_supported(long long n, int d, int hidden);
int hwgat_ffn_fwd(const void* h, const void* w1, const float* b1, const void* w2, void* act, void* gp, void* v0,
                  long long n, int d, int hidden, float p, unsigned long long seed, unsigned long long offset,
                  hwgat_stream_t stream);
/* K10': autograd of the above for dv0 (bf16, (n, d)): dh (bf16, (n, d)), dw1 (fp32, (hidden, d)), db1 (fp32,
 * (hidden)), dw2 (fp32, (d, hidden)), all overwritten.  du0 = (dv0 . W2) o gp is formed in the epilogue of its GEMM.
 * workspace: hwgat_ffn_bwd_workspace_bytes(n, d, hidden) bytes.                                                */
size_t hwgat_ffn_bwd_workspace_bytes(long long n, int d, int hidden);
int hwgat_ffn_bwd(const void* dv0, const void* h, const void* act, const void* gp, const void* w1, const void* w2,
                  void* dh, float* dw1, float* db1, float* dw2, void* workspace, size_t workspace_bytes, long long n,
                  int d, int hidden, hwgat_stream_t stream);

/* ---- model head and tail ----------------------------------------------- */

/* K8: out(fp32, (n, E)) = dropout_p([sin(2 pi x.Bm^T), cos(2 pi x.Bm^T)] + pe[frame]); x: (n, C) fp32 keypoints,
 * Bm: (E/2, C) the frozen Fourier matrix `B`, pe: (T, E) the sinusoid table, frame of token i = (i / K) % T.
 * Replaces Model.forward_features' embedding and PositionalEncoding.forward (HWGATE.py:343-347, 25-28).
 * Forward only (nothing upstream is trainable).  E % 8 == 0.                                              */
int hwgat_embed_fwd(const float* x, const float* Bm, const float* pe, float* out, long long n, int C, int E, int K,
                    int T, float p, unsigned long long seed, unsigned long long offset, hwgat_stream_t stream);
/* K9: pooled(fp32, (B, d)) = mean over the `tokens` rows of each sample of LayerNorm(x; gamma, beta, eps);
 * mean / rstd (B*tokens) saved.  Replaces self.norm + self.avgpool (HWGATE.py:353-354).  d in {128,256,512}.
 * Small batches are split over several CTAs per sample; their partial sums go through `scratch`
 * (hwgat_ln_pool_scratch_bytes, may be 0) and are added in a fixed order: the result is deterministic.
 * kp_real / kp_pad (0, 0 = none): a padded keypoint axis - the sibling model HGATE has 29 keypoints, stored as 32
 * (HGATE.py:341): `tokens` counts the real tokens (frames * kp_real), the rows are stored frames * kp_pad per sample,
 * mean / rstd have B * frames * kp_pad entries and only the real rows are pooled.                                */
size_t hwgat_ln_pool_scratch_bytes(int B, int tokens, int d);
int hwgat_ln_pool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean,
                      float* rstd, void* scratch, size_t scratch_bytes, int B, int tokens, int d, float eps,
                      int kp_real, int kp_pad, hwgat_stream_t stream);
/* K9': dx(fp32) of the above for g = d pooled (B, d); dgamma (d) overwritten (dbeta = column sums of g: caller).
 * With a padded keypoint axis the padded rows of dx are NOT written (the caller zero-fills dx).                  */
int hwgat_ln_pool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                      float* dx, float* dgamma, int B, int tokens, int d, int kp_real, int kp_pad,
                      hwgat_stream_t stream);

/* ---- training loop ------------------------------------------------------ */

/* K12: the output projection of MSA without its bias (the bias, proj_drop and the shortcut add belong to K6):
 *   y (bf16, (n, d_out)) = ctx (bf16, (n, d_in)) . W^T,   W: (d_out, d_in) bf16
 * Replaces self.proj's matmul (HWGATE.py:115) with the library's own TMA + tcgen05 GEMM, so the whole of MSA.forward
 * runs on this library.  n % 128 == 0, d_out % 128 == 0, d_in % 64 == 0.                                          */
int hwgat_proj_fwd(const void* ctx, const void* w, void* y, long long n, int d_in, int d_out, hwgat_stream_t stream);
/* K12': d_ctx (bf16, (n, d_in)) = dy . W,  dw (fp32, (d_out, d_in)) = dy^T . ctx, both overwritten (the bias gradient
 * is K6's).  Needs d_in % 128 == 0 as well.  workspace: hwgat_proj_bwd_workspace_bytes(d_in, d_out) bytes (W^T).  */
size_t hwgat_proj_bwd_workspace_bytes(int d_in, int d_out);
int hwgat_proj_bwd(const void* dy, const void* ctx, const void* w, void* d_ctx, float* dw, void* workspace,
                   size_t workspace_bytes, long long n, int d_in, int d_out, hwgat_stream_t stream);

/* K11: one AdamW step over n_tensors fp32 parameter tensors in ONE launch per 160 tensors (multi-tensor apply).
 * Replaces optimizer.step() of torch.optim.AdamW(model.parameters(), lr) (utils.py:73-75, stepped at utils.py:107):
 *   p *= 1 - lr*weight_decay;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps),   g = grad_scale * grads[t]
 * params / grads / exp_avg / exp_avg_sq: HOST arrays of n_tensors DEVICE pointers (fp32), sizes: HOST array of
 * element counts.  step >= 1 is the step number after the increment (torch's state['step']).                 */
int hwgat_adamw_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const long long* sizes, double lr, double beta1, double beta2, double eps,
                     double weight_decay, long long step, float grad_scale, hwgat_stream_t stream);

/* Diagnostic: the plain-epilogue form of K10's GEMM, which K3 also uses for d_xn: C[M,N] = A[M,K] . Bt[N,K]^T
 * (bf16 operands, fp32 accumulate, TMA + tcgen05, 32-byte row stores).  M % 128 == 0, N % 128 == 0, K % 64 == 0;
 * all row-major bf16 device pointers. */
int hwgat_debug_gemm_nt(const void* A, const void* Bt, void* C, long long M, int N, int K, hwgat_stream_t stream);
/* Diagnostic: the GEMM K3 and K10 use for weight and bias gradients: C[M,N] (fp32) = A[Kd,M]^T . B[Kd,N], colsum[M] = column
 * sums of A (colsum may be NULL: not computed).  M % 128 == 0, N % 128 == 0, Kd % 64 == 0; A, B bf16 row-major device
 * pointers.  N % 256 == 0 and M >= 256 run on CTA pairs (cta_group::2) unless hwgat_debug_set_gemm_pair(0). */
int hwgat_debug_gemm_tn(const void* A, const void* B, float* C, float* colsum, int M, int N, long long Kd,
                        hwgat_stream_t stream);

/* Diagnostic (A/B measurements): wide GEMMs (N % 256 == 0, M >= 256) run on CTA pairs (cta_group::2, 256 x 256
 * tiles) when on != 0 (the default), on the single-CTA 128 x 256 kernel otherwise.  Returns the previous setting. */
int hwgat_debug_set_gemm_pair(int on);

/* Number of kernel launches issued through this library since load (all
 * streams, this process) - what bench.py reports as "gpu_launches". */
/* fp32 parity mode of hwgat_attn2_fwd / hwgat_attn2_bwd: window_size 32 / 64 (N = 64 / 128 tokens) and HGATE's blocks in
 * true fp32 (FFMA, one thread per token of a window; the 1e-5 mode of the north_star - a correctness mode, not timed).
 * Every tensor float32; same semantics as hwgat_attn_fwd(HWGAT_F32, ...) (threshold drop, packed mask, -10000 fill);
 * no attention dropout; no 128-token tile constraint (any B, even F, K % W == 0).  qkv (B*F*K, 3d) is written by the
 * forward and read by the backward; backward workspace from hwgat_attn2_f32_workspace_bytes(.., 1).                */
size_t hwgat_attn2_f32_workspace_bytes(int B, int F, int K, int d, int backward);
int hwgat_attn2_fwd_f32(const float* xn, const float* w_qkv, const float* b_qkv, const uint32_t* bits, float threshold,
                        float* out, float* qkv, int B, int F, int K, int d, int heads, int W, int TP, int shift,
                        int layout, hwgat_stream_t stream);
int hwgat_attn2_bwd_f32(const float* d_out, const float* xn, const float* w_qkv, const float* qkv, const uint32_t* bits,
                        float threshold, float* d_xn, float* d_w, float* d_b, void* workspace, size_t workspace_bytes,
                        int B, int F, int K, int d, int heads, int W, int TP, int shift, int layout,
                        hwgat_stream_t stream);

/* K9 with learned token weights: GATE's `weightedAvg` = Linear(F*K, 1) over the token axis (hwgat/models/GATE.py:185,
 * 207) fused with the final LayerNorm (GATE.py:205).  pooled[b, c] = sum_t tok_w[t] * (xhat[b,t,c] gamma[c]) + beta[c];
 * the caller passes beta * sum(tok_w) + the Linear's bias as `beta`.  tok_w f32 (tokens) indexes REAL tokens
 * (frame * kp_real + keypoint).  Backward: dx, dgamma as hwgat_ln_pool_bwd; d_tok_w[t] = sum_b sum_c g[b,c] gamma[c]
 * xhat[b,t,c], summed over the samples in a fixed order through dw_part (B * tokens floats of scratch); the caller
 * adds the g.beta term.                                                                                              */
int hwgat_ln_wpool_fwd(const float* x, const float* gamma, const float* beta, const float* tok_w, float* pooled,
                       float* mean, float* rstd, void* scratch, size_t scratch_bytes, int B, int tokens, int d,
                       float eps, int kp_real, int kp_pad, hwgat_stream_t stream);
int hwgat_ln_wpool_bwd(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* tok_w, float* dx, float* dgamma, float* d_tok_w, float* dw_part, int B, int tokens,
                       int d, int kp_real, int kp_pad, hwgat_stream_t stream);

/* ---- K15 / K16: frame-banded graph attention of the sibling models WGATE and GATE --------------------------------
 * Replaces MSA.forward of hwgat/models/WGATE.py:87-108 (windows of W = 16 keypoints over ALL frames, additive
 * -10000 mask, called from PartAttentionBlock.forward, WGATE.py:150-158) and of hwgat/models/GATE.py:49-69 (all 29
 * keypoints x all frames, stored here as one window of W = 32 with three padded keypoints), up to the output
 * projection.  The adjacency those models build (model_params.py:204-229, 59-74) is frame-banded - a token is linked to
 * keypoints of its own and of the two adjacent frames - and after the softmax every non-edge weighs exactly 0 in fp32,
 * so the kernels evaluate only the band: W queries x 3W keys per (sample, window, frame).
 *   xn   bf16 (B, F, K, d)      LayerNorm-ed stream, K % W == 0, B*F*K % 128 == 0, d % 128 == 0, d/heads in {16,32,64}
 *   bits u32  (K/W, W, 3)       bit j of word (w, i, r): keypoint i of window w attends keypoint j of frame f-1+r
 *                               (r = 0 previous, 1 same, 2 next frame); a row without any bit yields a zero output row
 *   out  bf16 (B, F, K, d)      head-merged context; qkv bf16 (B*F*K, 3d) is written and must be kept for the backward
 *   lse  f32  (B*F*K, heads)    base-2 logsumexp of the scaled logits over the edges, for the backward (NULL in inference)
 * Backward: d_xn bf16, d_w f32 (3d, d), d_b f32 (3d); workspace from hwgat_band_attn_workspace_bytes(.., 1).
 * dtype HWGAT_F32: the 1e-5 parity mode - xn, w_qkv, out, qkv, ctx, d_out, d_xn float32, true fp32 FFMA, one thread per
 * (token, head) walking the set bits of its band words, gather-only backward; no B*F*K % 128 / d % 128 constraint; the
 * saved lse is a natural-log logsumexp; diag is ignored.  HWGAT_BF16: the timed kernels described above.
 * diag != 0: the caller promises that every previous- / next-frame word (r = 0, 2) of row i is 1 << i or 0 - the
 * identity the reference's graphs have between adjacent frames - and the kernels evaluate only the diagonal of those
 * blocks; every CTA checks the words it loads and traps if the promise is broken.  diag = 0: any band.
 * No attention dropout (both models default to attn_drop_rate = 0; the host side refuses p > 0).                  */
size_t hwgat_band_attn_workspace_bytes(int dtype, int B, int F, int K, int d, int backward);
int hwgat_band_attn_fwd(int dtype, const void* xn, const void* w_qkv, const float* b_qkv, const uint32_t* bits,
                        void* out, void* qkv, float* lse, int B, int F, int K, int d, int heads, int W, int diag,
                        hwgat_stream_t stream);
int hwgat_band_attn_bwd(int dtype, const void* d_out, const void* xn, const void* w_qkv, const void* qkv,
                        const void* ctx, const float* lse, const uint32_t* bits, void* d_xn, float* d_w, float* d_b,
                        void* workspace, size_t workspace_bytes, int B, int F, int K, int d, int heads, int W, int diag,
                        hwgat_stream_t stream);

unsigned long long hwgat_launch_count(void);
/* Deterministic mode (process-wide; returns the previous setting).  Outputs and input gradients are always
 * bit-reproducible; PARAMETER gradients are by default summed with fp32 atomics over a token split (order varies run
 * to run, ~1e-6 relative).  on != 0: every token split of the weight-gradient GEMMs and every CTA of the bias /
 * LayerNorm column sums writes its own partial and a finish kernel adds them in index order: bit-reproducible.
 * In this mode (only) the library takes stream-ordered scratch from cudaMallocAsync / cudaFreeAsync.  on < 0 only
 * queries.                                                                                                       */
int hwgat_set_deterministic(int on);

/* How the fp32 path (dtype HWGAT_F32, hwgat_linear_f32_*, the *_f32 attention entries) multiplies matrices
 * (process-wide; returns the previous mode; any other value only queries).  The reference's own loop runs the model in
 * fp32 without autocast (utils.py:102, 128; inference.py:95), so this is the mode a drop-in lands on.
 *   HWGAT_FP32_FFMA  true-fp32 FFMA GEMMs: the 1e-5 parity mode (default of the library).
 *   HWGAT_FP32_X3    every fp32 GEMM with n % 128 == 0, d_in % 128 == 0, d_out % 128 == 0 runs on tcgen05: the operands
 *                    are split into three bf16 planes (hi + mid + lo = 24 mantissa bits) and the six partial products of
 *                    weight >= 2^-16 are accumulated in fp32 in TMEM, the leading product in an accumulator of its own
 *                    (gemm_x3.cu).  ~2e-7 relative against fp64, the level of an fp32 GEMM.  Takes stream-ordered
 *                    scratch for the planes (cudaMallocAsync).  Other shapes, and deterministic mode, use FFMA.     */
#define HWGAT_FP32_FFMA 0
#define HWGAT_FP32_X3 1
int hwgat_set_fp32_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* HWGAT_B200_H */
