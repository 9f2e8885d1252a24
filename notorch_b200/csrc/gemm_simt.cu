// Strict-fp32 (FFMA) version of the three W contractions of a message-passing depth:
// K2 forward, K4a dgrad, K4b wgrad. This is NT_GEMM_FP32: the arithmetic closest to the reference's
// CPU SGEMM, used as the exact-precision mode and for hidden sizes that are not a multiple of 4.
// The tensor-core path (gemm_tc.cu) is the fast one. Both are hand-written; neither calls a library.
//
// One generic 64x64x16 tile kernel; the problem is described by functors:
//   a(m, k), b(n, k) -> operand elements (zero outside the problem), store(m, n, acc).
#include "common.cuh"

namespace nt {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256, PAD = 4;

template <class P>
__global__ void __launch_bounds__(GEMM_THREADS) simt_gemm_kernel(P p) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;  // M tiles on grid.x (no 65535 limit)
  const int n0 = blockIdx.y * BN;
  int64_t k_begin = 0, k_end = p.K;
  if (P::SPLIT_K) {
    k_begin = (int64_t)blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage the two operand tiles (coalesced along whichever dimension is contiguous) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (P::A_K_CONTIG) { kk = tid & 15; mm = (tid >> 4) + 16 * i; } else { mm = tid & 63; kk = (tid >> 6) + 4 * i; }
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < k_end) ? p.a(m, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (P::B_K_CONTIG) { kk = tid & 15; nn = (tid >> 4) + 16 * i; } else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      int n = n0 + nn;
      int64_t k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < k_end) ? p.b(n, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.N) p.store(m, n, acc[i][j], blockIdx.z);
    }
  }
}

// m[e,k] = n[src[e],k] - act(h[rev[e],k])   (chemprop.py:40 with the activation of :37 folded in)
struct MessageOperand {
  const float* __restrict__ n;
  const float* __restrict__ h;
  const int32_t* __restrict__ src;
  const int32_t* __restrict__ rev;
  int d, act;
  float act_param;
  __device__ __forceinline__ float operator()(int64_t e, int64_t k) const {
    return __ldg(n + (int64_t)__ldg(src + e) * d + k) - act_fwd(__ldg(h + (int64_t)__ldg(rev + e) * d + k), act, act_param);
  }
};

struct DropoutCfg {
  float p, inv_keep;
  uint32_t thr;
  uint64_t seed, offset;
  __device__ __forceinline__ float scale(int64_t e, int d, int c) const {
    return p > 0.f ? dropout_scale1(seed, offset, (uint64_t)e * (uint64_t)d + (uint64_t)c, thr, inv_keep) : 1.f;
  }
};

struct ForwardProblem {  // K2: out = [h +] dropout(m W^T + bias)
  static constexpr bool SPLIT_K = false, A_K_CONTIG = true, B_K_CONTIG = true;
  int64_t M, K, k_chunk;
  int N;
  MessageOperand msg;
  const float* __restrict__ W;
  const float* __restrict__ bias;
  const float* __restrict__ h;
  float* __restrict__ out;
  DropoutCfg drop;
  int residual;
  __device__ __forceinline__ float a(int64_t m, int64_t k) const { return msg(m, k); }
  __device__ __forceinline__ float b(int n, int64_t k) const { return __ldg(W + (int64_t)n * N + k); }
  __device__ __forceinline__ void store(int64_t m, int n, float acc, int) const {
    float u = acc + (bias ? __ldg(bias + n) : 0.f);
    u *= drop.scale(m, N, n);
    if (residual) u = __ldg(h + m * N + n) + u;
    out[m * N + n] = u;
  }
};

struct DgradProblem {  // K4a: g_m = (mask . g / (1-p)) W
  static constexpr bool SPLIT_K = false, A_K_CONTIG = true, B_K_CONTIG = false;
  int64_t M, K, k_chunk;
  int N;
  const float* __restrict__ g;
  const float* __restrict__ W;
  float* __restrict__ g_m;
  DropoutCfg drop;
  __device__ __forceinline__ float a(int64_t m, int64_t k) const { return __ldg(g + m * N + k) * drop.scale(m, N, (int)k); }
  __device__ __forceinline__ float b(int n, int64_t k) const { return __ldg(W + k * N + n); }
  __device__ __forceinline__ void store(int64_t m, int n, float acc, int) const { g_m[m * N + n] = acc; }
};

struct WgradProblem {  // K4b: partial[z][o][i] = sum_{e in chunk z} g_u[e,o] m[e,i]; column i == d carries the bias gradient
  static constexpr bool SPLIT_K = true, A_K_CONTIG = false, B_K_CONTIG = false;
  int64_t M, K, k_chunk;  // M = d (o), K = E
  int N;                  // d + 1
  int d;
  const float* __restrict__ g;
  MessageOperand msg;
  DropoutCfg drop;
  float* __restrict__ partial;
  __device__ __forceinline__ float a(int64_t o, int64_t e) const { return __ldg(g + e * d + o) * drop.scale(e, d, (int)o); }
  __device__ __forceinline__ float b(int i, int64_t e) const { return i < d ? msg(e, i) : 1.f; }
  __device__ __forceinline__ void store(int64_t o, int i, float acc, int z) const { partial[((int64_t)z * d + o) * N + i] = acc; }
};

// gW[o,i] = sum_z partial[z][o][i] (ascending z: deterministic); gb[o] = sum_z partial[z][o][d]
__global__ void __launch_bounds__(256) wgrad_reduce(const float* __restrict__ partial, int splits, int d, float* __restrict__ gW, float* __restrict__ gb) {
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int N = d + 1;
  if (t >= (int64_t)d * N) return;
  int o = (int)(t / N), i = (int)(t - (int64_t)o * N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += __ldg(partial + ((int64_t)z * d + o) * N + i);
  if (i < d) gW[(int64_t)o * d + i] = s;
  else if (gb) gb[o] = s;
}

static DropoutCfg make_drop(float p, uint64_t seed, uint64_t offset) {
  DropoutCfg c;
  c.p = p;
  c.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  double t = (double)p * 4294967296.0;
  c.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  c.seed = seed;
  c.offset = offset;
  return c;
}

int simt_wgrad_splits(int64_t E, int64_t d) {
  int64_t tiles = cdiv(d, BM) * cdiv(d + 1, BN);
  int64_t s = (4 * (int64_t)148) / tiles;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  int64_t max_s = cdiv(E > 0 ? E : 1, 4 * BK);
  if (s > max_s) s = max_s;
  return (int)s;
}

int simt_layer_forward(const float* h, const float* n, const int32_t* src, const int32_t* rev, const float* W, const float* bias, int64_t E, int64_t d,
                       int act, float act_param, int residual, float p, uint64_t seed, uint64_t offset, float* out, cudaStream_t st) {
  ForwardProblem pr;
  pr.M = E; pr.K = d; pr.k_chunk = d; pr.N = (int)d;
  pr.msg = MessageOperand{n, h, src, rev, (int)d, act, act_param};
  pr.W = W; pr.bias = bias; pr.h = h; pr.out = out; pr.drop = make_drop(p, seed, offset); pr.residual = residual;
  dim3 grid((unsigned)cdiv(E, BM), (unsigned)cdiv(d, BN), 1);
  simt_gemm_kernel<ForwardProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  NT_LAUNCH_CHECK("simt_layer_forward", 1);
  return NT_OK;
}

int simt_layer_dgrad(const float* g, const float* W, int64_t E, int64_t d, float p, uint64_t seed, uint64_t offset, float* g_m, cudaStream_t st) {
  DgradProblem pr;
  pr.M = E; pr.K = d; pr.k_chunk = d; pr.N = (int)d;
  pr.g = g; pr.W = W; pr.g_m = g_m; pr.drop = make_drop(p, seed, offset);
  dim3 grid((unsigned)cdiv(E, BM), (unsigned)cdiv(d, BN), 1);
  simt_gemm_kernel<DgradProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  NT_LAUNCH_CHECK("simt_layer_dgrad", 1);
  return NT_OK;
}

int simt_layer_wgrad(const float* g, const float* h, const float* n, const int32_t* src, const int32_t* rev, int64_t E, int64_t d, int act,
                     float act_param, float p, uint64_t seed, uint64_t offset, float* gW, float* gb, float* partial, cudaStream_t st) {
  int splits = simt_wgrad_splits(E, d);
  WgradProblem pr;
  pr.M = d; pr.K = E; pr.N = (int)d + 1; pr.d = (int)d;
  pr.k_chunk = cdiv(cdiv(E, splits), BK) * BK;
  if (pr.k_chunk == 0) pr.k_chunk = BK;
  pr.g = g;
  pr.msg = MessageOperand{n, h, src, rev, (int)d, act, act_param};
  pr.drop = make_drop(p, seed, offset);
  pr.partial = partial;
  dim3 grid((unsigned)cdiv(d, BM), (unsigned)cdiv(d + 1, BN), (unsigned)splits);
  simt_gemm_kernel<WgradProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  int64_t total = d * (d + 1);
  wgrad_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(partial, splits, (int)d, gW, gb);
  NT_LAUNCH_CHECK("simt_layer_wgrad", 2);
  return NT_OK;
}

// ---- rectangular Linear of the prediction head (notorch/nn/mlp.py:58-62: nn.Linear(d1, d2) on [B, d1] molecule vectors) ----
struct LinearForwardProblem {  // out[m,n] = sum_k x[m,k] W[n,k] + bias[n]
  static constexpr bool SPLIT_K = false, A_K_CONTIG = true, B_K_CONTIG = true;
  int64_t M, K, k_chunk;
  int N;
  const float* __restrict__ x;
  const float* __restrict__ W;
  const float* __restrict__ bias;
  float* __restrict__ out;
  __device__ __forceinline__ float a(int64_t m, int64_t k) const { return __ldg(x + m * K + k); }
  __device__ __forceinline__ float b(int n, int64_t k) const { return __ldg(W + (int64_t)n * K + k); }
  __device__ __forceinline__ void store(int64_t m, int n, float acc, int) const { out[m * N + n] = acc + (bias ? __ldg(bias + n) : 0.f); }
};

struct LinearDgradProblem {  // gx[m,k] = sum_n g[m,n] W[n,k]        (GEMM "N" = K_in, reduction over N_out)
  static constexpr bool SPLIT_K = false, A_K_CONTIG = true, B_K_CONTIG = false;
  int64_t M, K, k_chunk;  // K = N_out
  int N;                  // K_in
  const float* __restrict__ g;
  const float* __restrict__ W;
  float* __restrict__ gx;
  __device__ __forceinline__ float a(int64_t m, int64_t n_out) const { return __ldg(g + m * K + n_out); }
  __device__ __forceinline__ float b(int k_in, int64_t n_out) const { return __ldg(W + n_out * N + k_in); }
  __device__ __forceinline__ void store(int64_t m, int k_in, float acc, int) const { gx[m * N + k_in] = acc; }
};

struct LinearWgradProblem {  // partial[z][n][k] = sum_{m in chunk z} g[m,n] x[m,k]; column k == K_in carries the bias gradient
  static constexpr bool SPLIT_K = true, A_K_CONTIG = false, B_K_CONTIG = false;
  int64_t M, K, k_chunk;  // M = N_out, K = rows
  int N;                  // K_in + 1
  int n_out, k_in;
  const float* __restrict__ g;
  const float* __restrict__ x;
  float* __restrict__ partial;
  __device__ __forceinline__ float a(int64_t n, int64_t m) const { return __ldg(g + m * n_out + n); }
  __device__ __forceinline__ float b(int k, int64_t m) const { return k < k_in ? __ldg(x + m * k_in + k) : 1.f; }
  __device__ __forceinline__ void store(int64_t n, int k, float acc, int z) const { partial[((int64_t)z * n_out + n) * N + k] = acc; }
};

__global__ void __launch_bounds__(256) linear_wgrad_reduce(const float* __restrict__ partial, int splits, int n_out, int k_in, float* __restrict__ gW,
                                                           float* __restrict__ gb) {
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int N = k_in + 1;
  if (t >= (int64_t)n_out * N) return;
  const int n = (int)(t / N), k = (int)(t - (int64_t)n * N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += __ldg(partial + ((int64_t)z * n_out + n) * N + k);  // ascending z: deterministic
  if (k < k_in) gW[(int64_t)n * k_in + k] = s;
  else if (gb) gb[n] = s;
}

static int linear_wgrad_splits(int64_t rows, int64_t n_out, int64_t k_in) {
  int64_t tiles = cdiv(n_out, BM) * cdiv(k_in + 1, BN);
  int64_t s = (4 * (int64_t)148) / tiles;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  int64_t max_s = cdiv(rows > 0 ? rows : 1, 4 * BK);
  if (s > max_s) s = max_s;
  return (int)s;
}

size_t simt_linear_wgrad_workspace_bytes(int64_t rows, int64_t n_out, int64_t k_in) {
  return (size_t)linear_wgrad_splits(rows, n_out, k_in) * n_out * (k_in + 1) * sizeof(float);
}

int simt_linear_forward(const float* x, const float* W, const float* bias, int64_t rows, int64_t n_out, int64_t k_in, float* out, cudaStream_t st) {
  LinearForwardProblem pr;
  pr.M = rows; pr.K = k_in; pr.k_chunk = k_in; pr.N = (int)n_out;
  pr.x = x; pr.W = W; pr.bias = bias; pr.out = out;
  dim3 grid((unsigned)cdiv(rows, BM), (unsigned)cdiv(n_out, BN), 1);
  simt_gemm_kernel<LinearForwardProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  NT_LAUNCH_CHECK("simt_linear_forward", 1);
  return NT_OK;
}

int simt_linear_dgrad(const float* g, const float* W, int64_t rows, int64_t n_out, int64_t k_in, float* gx, cudaStream_t st) {
  LinearDgradProblem pr;
  pr.M = rows; pr.K = n_out; pr.k_chunk = n_out; pr.N = (int)k_in;
  pr.g = g; pr.W = W; pr.gx = gx;
  dim3 grid((unsigned)cdiv(rows, BM), (unsigned)cdiv(k_in, BN), 1);
  simt_gemm_kernel<LinearDgradProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  NT_LAUNCH_CHECK("simt_linear_dgrad", 1);
  return NT_OK;
}

int simt_linear_wgrad(const float* g, const float* x, int64_t rows, int64_t n_out, int64_t k_in, float* gW, float* gb, float* partial, cudaStream_t st) {
  const int splits = linear_wgrad_splits(rows, n_out, k_in);
  LinearWgradProblem pr;
  pr.M = n_out; pr.K = rows; pr.N = (int)k_in + 1; pr.n_out = (int)n_out; pr.k_in = (int)k_in;
  pr.k_chunk = cdiv(cdiv(rows, splits), BK) * BK;
  if (pr.k_chunk == 0) pr.k_chunk = BK;
  pr.g = g; pr.x = x; pr.partial = partial;
  dim3 grid((unsigned)cdiv(n_out, BM), (unsigned)cdiv(k_in + 1, BN), (unsigned)splits);
  simt_gemm_kernel<LinearWgradProblem><<<grid, GEMM_THREADS, 0, st>>>(pr);
  const int64_t total = n_out * (k_in + 1);
  linear_wgrad_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(partial, splits, (int)n_out, (int)k_in, gW, gb);
  NT_LAUNCH_CHECK("simt_linear_wgrad", 2);
  return NT_OK;
}

}  // namespace nt
