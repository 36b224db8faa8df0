// K1a/K1b: skeleton adjacency and packed window masks.  K4: temporal merge.
// All three are tiny or pure data movement; K4 is the HBM-bound stage
// transition.
#include "common.cuh"

namespace hwgat {

// ---------------------------------------------------------------------------
// K1a: one CTA per keypoint window.  Replaces model_params.py:373-400.
// ---------------------------------------------------------------------------
__global__ void adjacency_kernel(const int32_t* __restrict__ edges, int n_edges, int W, int TP,
                                 float* __restrict__ adj) {
  extern __shared__ unsigned char skel[];  // W*W, 1 = linked inside one frame
  const int w = blockIdx.x, N = TP * W;
  for (int i = threadIdx.x; i < W * W; i += blockDim.x) skel[i] = (i / W == i % W);
  __syncthreads();
  for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
    int a = edges[(w * n_edges + e) * 2], b = edges[(w * n_edges + e) * 2 + 1];
    if (a >= 0 && a < W && b >= 0 && b < W) { skel[a * W + b] = 1; skel[b * W + a] = 1; }
  }
  __syncthreads();
  float* out = adj + (size_t)w * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    int i = idx / N, j = idx % N;
    int ti = i / W, ki = i % W, tj = j / W, kj = j % W;
    int dt = ti > tj ? ti - tj : tj - ti;
    float v = dt == 0 ? (float)skel[ki * W + kj] : (dt == 1 ? (ki == kj ? 1.f : 0.f) : 0.f);
    out[idx] = v;
  }
}

int launch_adjacency(const int32_t* edges, int n_edges, int nW, int W, int TP, float* adj, cudaStream_t s) {
  adjacency_kernel<<<nW, 256, W * W, s>>>(edges, n_edges, W, TP, adj);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K1b: one thread per (window, query row, 32-key word).  Replaces
// HWGATE.py:169-187 (shift mask), :309 (replication) and the two multiplies at
// :102-108.  Group id of a (rolled) frame g: 0 if g < F-TP, 1 if g < F-shift,
// else 2; a pair is visible iff adjacency != 0 and both frames share the id.
// ---------------------------------------------------------------------------
__global__ void mask_bits_kernel(const float* __restrict__ adj, int nW, int W, int TP, int F, int shift,
                                 uint32_t* __restrict__ bits, int total) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int N = TP * W, words = N / 32;
  int word = idx % words, i = (idx / words) % N, win = idx / (words * N);
  int fi = win / nW, w = win % nW;
  auto gid = [&](int tok) {
    if (shift <= 0) return 0;
    int g = fi * TP + tok / W;
    return g < F - TP ? 0 : (g < F - shift ? 1 : 2);
  };
  const float* row = adj + ((size_t)w * N + i) * N + word * 32;
  int gi = gid(i);
  uint32_t v = 0;
#pragma unroll 8
  for (int j = 0; j < 32; ++j)
    if (row[j] != 0.f && gid(word * 32 + j) == gi) v |= (1u << j);
  bits[idx] = v;
}

int launch_mask_bits(const float* adj, int nW, int W, int TP, int F, int shift, uint32_t* bits, cudaStream_t s) {
  int N = TP * W, total = (F / TP) * nW * N * (N / 32);
  mask_bits_kernel<<<(total + 127) / 128, 128, 0, s>>>(adj, nW, W, TP, F, shift, bits, total);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K1c: generic pack of float masks handed to MSA.forward (HWGATE.py:84, 102-108):
// bit = (adj != 0) AND (mask != 0); either factor may be absent.
// ---------------------------------------------------------------------------
__global__ void mask_pack_kernel(const float* __restrict__ adj, int adj_windows, const float* __restrict__ mask,
                                 int N, uint32_t* __restrict__ bits, int total) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int words = N / 32;
  int word = idx % words, i = (idx / words) % N, win = idx / (words * N);
  const float* arow = adj ? adj + ((size_t)(win % adj_windows) * N + i) * N + word * 32 : nullptr;
  const float* mrow = mask ? mask + ((size_t)win * N + i) * N + word * 32 : nullptr;
  uint32_t v = 0;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    bool on = true;
    if (arow) on = on && arow[j] != 0.f;
    if (mrow) on = on && mrow[j] != 0.f;
    if (on) v |= (1u << j);
  }
  bits[idx] = v;
}

int launch_mask_pack(const float* adj, int adj_windows, const float* mask, int n_windows, int N, uint32_t* bits,
                     cudaStream_t s) {
  int total = n_windows * N * (N / 32);
  if (total == 0) return 0;
  mask_pack_kernel<<<(total + 127) / 128, 128, 0, s>>>(adj, adj_windows, mask, N, bits, total);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K4: out[b,fi,k,tp*d+e] = x[b,2*fi+tp,k,e]  (HWGATE.py:55-63) and its adjoint.
// One 16-byte vector per thread-iteration, indexed in the order of the MERGED
// tensor so that the merged side is perfectly contiguous and the other side is
// contiguous in runs of d elements (>= 256 B).  Grid is a multiple of the SM
// count; 4 independent vectors are in flight per thread.
// ---------------------------------------------------------------------------
template <bool kBackward>
__global__ void __launch_bounds__(256) merge_kernel(const int4* __restrict__ src, int4* __restrict__ dst,
                                                    long long n_vec, int K, int dv /* d in vectors */) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto other = [&](long long m) {  // merged vector index -> un-merged vector index
    int e = (int)(m % dv);
    long long t = m / dv;
    int tp = (int)(t & 1);
    t >>= 1;
    int k = (int)(t % K);
    long long bf = t / K;  // b*f + fi
    return ((bf * 2 + tp) * K + k) * dv + e;
  };
  for (; v + 3 * stride < n_vec; v += 4 * stride) {
    int4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = ld_stream16(src + (kBackward ? v + u * stride : other(v + u * stride)));
#pragma unroll
    for (int u = 0; u < 4; ++u) st_stream16(dst + (kBackward ? other(v + u * stride) : v + u * stride), r[u]);
  }
  for (; v < n_vec; v += stride) {
    int4 r = ld_stream16(src + (kBackward ? v : other(v)));
    st_stream16(dst + (kBackward ? other(v) : v), r);
  }
}

int launch_merge(const void* src, void* dst, int B, int F, int K, int d, int elem_bytes, bool backward,
                 cudaStream_t s) {
  const int per_vec = 16 / elem_bytes;
  const int dv = d / per_vec;
  const long long n_vec = (long long)B * F * K * dv;
  if (n_vec == 0) return 0;
  int sms = 148;
  long long want = (n_vec + 256LL * 4 - 1) / (256LL * 4);
  int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  if (grid < 1) grid = 1;
  if (backward)
    merge_kernel<true><<<grid, 256, 0, s>>>((const int4*)src, (int4*)dst, n_vec, K, dv);
  else
    merge_kernel<false><<<grid, 256, 0, s>>>((const int4*)src, (int4*)dst, n_vec, K, dv);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace hwgat
