// Remaining atom -> molecule read-outs of notorch/nn/gnn/agg.py:
//   Max          (agg.py:41-47)  torch_scatter.scatter_max  -> nt_seg_max (+ argmax for the backward)
//   Gated        (agg.py:50-63)  softmax(Linear(x)) weighted sum
//   SDPAttention (agg.py:66-86)  softmax(<Q[batch], x> / sqrt(d)) weighted sum
// The softmax read-outs are composed from four deterministic primitives, each with a hand-written
// backward in ops.py: row dot product, segment softmax, weighted segment sum, scaled row gather.
// All segment loops run sequentially in ascending row order (no atomics).
#include <math.h>

#include "common.cuh"

namespace nt {

constexpr int RO_THREADS = 256;

// out[s,c] = max_j (sign * act(x[perm[j],c])) * sign, arg[s,c] = first row attaining it; empty segment -> 0 / -1 (torch_scatter: 0 / dim_size).
// sign = +1: scatter_max, sign = -1: scatter_min (= -scatter_max(-x), sign flips are exact). VEC: four channels per thread (d % 4 == 0).
template <bool VEC>
__global__ void __launch_bounds__(RO_THREADS) seg_extreme_kernel(const float* __restrict__ x, int d, int chunks, const int32_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ perm, int64_t total, int act, float act_param,
                                                                 float sign, float* __restrict__ out, int32_t* __restrict__ arg) {
  int64_t t = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (t >= total) return;
  const int s = (int)(t / chunks), c = (int)(t - (int64_t)s * chunks) * (VEC ? 4 : 1);
  const int lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
  constexpr int W = VEC ? 4 : 1;
  float best[W];
  int where[W];
#pragma unroll
  for (int u = 0; u < W; ++u) { best[u] = 0.f; where[u] = -1; }
  for (int j = lo; j < hi; ++j) {
    const int r = perm ? __ldg(perm + j) : j;
    float v[W];
    if constexpr (VEC) {
      const float4 q = ldg4(x + (int64_t)r * d + c);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      v[0] = __ldg(x + (int64_t)r * d + c);
    }
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const float a = act_fwd(v[u], act, act_param) * sign;
      if (where[u] < 0 || a > best[u]) { best[u] = a; where[u] = r; }  // strict '>' : the first extreme wins, like torch_scatter's CPU kernel
    }
  }
  const int64_t o = (int64_t)s * d + c;
#pragma unroll
  for (int u = 0; u < W; ++u) {
    out[o + u] = where[u] < 0 ? 0.f : best[u] * sign;
    arg[o + u] = where[u];
  }
}

// gx[r,c] = (arg[seg[r],c] == r) ? g[seg[r],c] : 0
__global__ void __launch_bounds__(RO_THREADS) seg_max_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ arg, const int32_t* __restrict__ seg,
                                                                 int d, int64_t total, float* __restrict__ gx) {
  int64_t t = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (t >= total) return;
  const int r = (int)(t / d), c = (int)(t - (int64_t)r * d);
  const int64_t o = (int64_t)__ldg(seg + r) * d + c;
  gx[t] = (__ldg(arg + o) == r) ? __ldg(g + o) : 0.f;
}

// out[i] = scale * <x[i,:], y[yidx ? yidx[i] : (y_rows == 1 ? 0 : i), :]> + bias        one warp per row
__global__ void __launch_bounds__(RO_THREADS) row_dot_kernel(const float* __restrict__ x, const float* __restrict__ y, const int32_t* __restrict__ yidx,
                                                             int64_t y_rows, int64_t n, int d, float scale, float bias, float* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * RO_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int64_t yr = yidx ? (int64_t)__ldg(yidx + row) : (y_rows == 1 ? 0 : row);
  const float* xp = x + row * d;
  const float* yp = y + yr * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s = fmaf(__ldg(xp + c), __ldg(yp + c), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // fixed butterfly: deterministic
  if (lane == 0) out[row] = s * scale + bias;
}

// alpha[r] = exp(s[r] - max_seg) / sum_seg exp(...)        one thread per segment (segments are molecules: ~25 rows)
__global__ void __launch_bounds__(RO_THREADS) seg_softmax_kernel(const float* __restrict__ s, const int32_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ perm, int64_t S, float* __restrict__ alpha) {
  int64_t seg = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (seg >= S) return;
  const int lo = __ldg(rowptr + seg), hi = __ldg(rowptr + seg + 1);
  float mx = -INFINITY;
  for (int j = lo; j < hi; ++j) mx = fmaxf(mx, __ldg(s + (perm ? __ldg(perm + j) : j)));
  float sum = 0.f;
  for (int j = lo; j < hi; ++j) sum += expf(__ldg(s + (perm ? __ldg(perm + j) : j)) - mx);
  for (int j = lo; j < hi; ++j) {
    const int r = perm ? __ldg(perm + j) : j;
    alpha[r] = expf(__ldg(s + r) - mx) / sum;
  }
}

// g_s[r] = alpha[r] * (g_alpha[r] - sum_seg alpha * g_alpha)
__global__ void __launch_bounds__(RO_THREADS) seg_softmax_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ g_alpha,
                                                                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int64_t S,
                                                                     float* __restrict__ g_s) {
  int64_t seg = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (seg >= S) return;
  const int lo = __ldg(rowptr + seg), hi = __ldg(rowptr + seg + 1);
  float dot = 0.f;
  for (int j = lo; j < hi; ++j) {
    const int r = perm ? __ldg(perm + j) : j;
    dot = fmaf(__ldg(alpha + r), __ldg(g_alpha + r), dot);
  }
  for (int j = lo; j < hi; ++j) {
    const int r = perm ? __ldg(perm + j) : j;
    g_s[r] = __ldg(alpha + r) * (__ldg(g_alpha + r) - dot);
  }
}

// out[s,c] = scale * sum_j w[perm[j]] * x[perm[j],c]
__global__ void __launch_bounds__(RO_THREADS) seg_weighted_sum_kernel(const float* __restrict__ x, const float* __restrict__ w, int d,
                                                                      const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int64_t total,
                                                                      float scale, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (t >= total) return;
  const int s = (int)(t / d), c = (int)(t - (int64_t)s * d);
  const int lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
  float acc = 0.f;
  for (int j = lo; j < hi; ++j) {
    const int r = perm ? __ldg(perm + j) : j;
    acc = fmaf(__ldg(w + r), __ldg(x + (int64_t)r * d + c), acc);
  }
  out[t] = acc * scale;
}

// out[i,c] = scale * w[i] * y[yidx ? yidx[i] : 0, c]
__global__ void __launch_bounds__(RO_THREADS) row_scale_gather_kernel(const float* __restrict__ y, const float* __restrict__ w, const int32_t* __restrict__ yidx,
                                                                      int d, int64_t total, float scale, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
  if (t >= total) return;
  const int64_t i = t / d;
  const int c = (int)(t - i * d);
  const int64_t yr = yidx ? (int64_t)__ldg(yidx + i) : 0;
  out[t] = scale * __ldg(w + i) * __ldg(y + yr * d + c);
}

// column sums of w[i] * x[i,c] over all rows: two fixed-order stages
constexpr int WC_ROWS = 512;
__global__ void __launch_bounds__(RO_THREADS) weighted_colsum_partial(const float* __restrict__ x, const float* __restrict__ w, int64_t n, int d,
                                                                      float* __restrict__ part) {
  const int c = blockIdx.x * RO_THREADS + threadIdx.x;
  if (c >= d) return;
  const int64_t r0 = (int64_t)blockIdx.y * WC_ROWS, r1 = r0 + WC_ROWS < n ? r0 + WC_ROWS : n;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s = fmaf(w ? __ldg(w + r) : 1.f, __ldg(x + r * d + c), s);
  part[(int64_t)blockIdx.y * d + c] = s;
}
__global__ void __launch_bounds__(RO_THREADS) weighted_colsum_final(const float* __restrict__ part, int64_t nblk, int d, float scale, float* __restrict__ out) {
  const int c = blockIdx.x * RO_THREADS + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  for (int64_t b = 0; b < nblk; ++b) s += __ldg(part + b * d + c);
  out[c] = s * scale;
}

}  // namespace nt

using namespace nt;

#define NT_RO_F32(fn) if (dtype != NT_F32) { set_error(fn ": only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }

static int seg_extreme_impl(const char* fn, const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int act,
                            float act_param, int is_min, void* out, int32_t* arg, nt_stream_t stream) {
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && num_segments >= 0 && num_segments * d < ((int64_t)1 << 40), "%s: bad sizes", fn);
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "%s: bad activation", fn);
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(rowptr && out && arg, "%s: null pointer", fn);
  const float sign = is_min ? -1.f : 1.f;
  if (d % 4 == 0 && aligned16(x) && aligned16(out) && aligned16(arg)) {
    const int chunks = (int)(d / 4);
    const int64_t total = num_segments * chunks;
    seg_extreme_kernel<true><<<(unsigned)cdiv(total, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(
        static_cast<const float*>(x), (int)d, chunks, rowptr, perm, total, act, act_param, sign, static_cast<float*>(out), arg);
  } else {
    const int64_t total = num_segments * d;
    seg_extreme_kernel<false><<<(unsigned)cdiv(total, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(
        static_cast<const float*>(x), (int)d, (int)d, rowptr, perm, total, act, act_param, sign, static_cast<float*>(out), arg);
  }
  NT_LAUNCH_CHECK(fn, 1);
  return NT_OK;
}

extern "C" int nt_seg_max(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, void* out, int32_t* arg, int dtype,
                          nt_stream_t stream) {
  NT_RO_F32("nt_seg_max");
  return seg_extreme_impl("nt_seg_max", x, d, rowptr, perm, num_segments, NT_ACT_IDENTITY, 0.f, 0, out, arg, stream);
}

extern "C" int nt_seg_extreme(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int act, float act_param,
                              int is_min, void* out, int32_t* arg, int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_seg_extreme");
  return seg_extreme_impl("nt_seg_extreme", x, d, rowptr, perm, num_segments, act, act_param, is_min, out, arg, stream);
}

extern "C" int nt_seg_max_backward(const void* g, const int32_t* arg, const int32_t* seg_of_row, int64_t n, int64_t d, void* gx, int dtype,
                                   nt_stream_t stream) {
  NT_RO_F32("nt_seg_max_backward");
  NT_CHECK_ARG(d > 0 && n >= 0, "nt_seg_max_backward: bad sizes");
  if (n == 0) return NT_OK;
  NT_CHECK_ARG(g && arg && seg_of_row && gx, "nt_seg_max_backward: null pointer");
  const int64_t total = n * d;
  seg_max_bwd_kernel<<<(unsigned)cdiv(total, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(g), arg, seg_of_row, (int)d, total,
                                                                                             static_cast<float*>(gx));
  NT_LAUNCH_CHECK("nt_seg_max_backward", 1);
  return NT_OK;
}

extern "C" int nt_row_dot(const void* x, const void* y, const int32_t* y_index, int64_t y_rows, int64_t n, int64_t d, float scale, float bias, void* out,
                          int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_row_dot");
  NT_CHECK_ARG(d > 0 && n >= 0 && y_rows > 0, "nt_row_dot: bad sizes");
  NT_CHECK_ARG(y_index || y_rows == 1 || y_rows == n, "nt_row_dot: y needs an index, a single row, or one row per x row");
  if (n == 0) return NT_OK;
  NT_CHECK_ARG(x && y && out, "nt_row_dot: null pointer");
  row_dot_kernel<<<(unsigned)cdiv(n * 32, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<const float*>(y), y_index,
                                                                                          y_rows, n, (int)d, scale, bias, static_cast<float*>(out));
  NT_LAUNCH_CHECK("nt_row_dot", 1);
  return NT_OK;
}

extern "C" int nt_seg_softmax(const void* s, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, void* alpha, int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_seg_softmax");
  NT_CHECK_ARG(num_segments >= 0, "nt_seg_softmax: bad sizes");
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(s && rowptr && alpha, "nt_seg_softmax: null pointer");
  seg_softmax_kernel<<<(unsigned)cdiv(num_segments, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(s), rowptr, perm, num_segments,
                                                                                                    static_cast<float*>(alpha));
  NT_LAUNCH_CHECK("nt_seg_softmax", 1);
  return NT_OK;
}

extern "C" int nt_seg_softmax_backward(const void* alpha, const void* g_alpha, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, void* g_s,
                                       int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_seg_softmax_backward");
  NT_CHECK_ARG(num_segments >= 0, "nt_seg_softmax_backward: bad sizes");
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(alpha && g_alpha && rowptr && g_s, "nt_seg_softmax_backward: null pointer");
  seg_softmax_bwd_kernel<<<(unsigned)cdiv(num_segments, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(
      static_cast<const float*>(alpha), static_cast<const float*>(g_alpha), rowptr, perm, num_segments, static_cast<float*>(g_s));
  NT_LAUNCH_CHECK("nt_seg_softmax_backward", 1);
  return NT_OK;
}

extern "C" int nt_seg_weighted_sum(const void* x, const void* w, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, float scale,
                                   void* out, int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_seg_weighted_sum");
  NT_CHECK_ARG(d > 0 && num_segments >= 0, "nt_seg_weighted_sum: bad sizes");
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(x && w && rowptr && out, "nt_seg_weighted_sum: null pointer");
  const int64_t total = num_segments * d;
  seg_weighted_sum_kernel<<<(unsigned)cdiv(total, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<const float*>(w),
                                                                                                  (int)d, rowptr, perm, total, scale, static_cast<float*>(out));
  NT_LAUNCH_CHECK("nt_seg_weighted_sum", 1);
  return NT_OK;
}

extern "C" int nt_row_scale_gather(const void* y, const void* w, const int32_t* y_index, int64_t n, int64_t d, float scale, void* out, int dtype,
                                   nt_stream_t stream) {
  NT_RO_F32("nt_row_scale_gather");
  NT_CHECK_ARG(d > 0 && n >= 0, "nt_row_scale_gather: bad sizes");
  if (n == 0) return NT_OK;
  NT_CHECK_ARG(y && w && out, "nt_row_scale_gather: null pointer");
  const int64_t total = n * d;
  row_scale_gather_kernel<<<(unsigned)cdiv(total, RO_THREADS), RO_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(y), static_cast<const float*>(w),
                                                                                                  y_index, (int)d, total, scale, static_cast<float*>(out));
  NT_LAUNCH_CHECK("nt_row_scale_gather", 1);
  return NT_OK;
}

extern "C" size_t nt_weighted_col_sum_workspace_bytes(int64_t n, int64_t d) {
  return (size_t)cdiv(n > 0 ? n : 1, WC_ROWS) * (size_t)(d > 0 ? d : 1) * sizeof(float) + 256;
}

extern "C" int nt_weighted_col_sum(const void* x, const void* w, int64_t n, int64_t d, float scale, void* out, void* workspace, size_t workspace_bytes,
                                   int dtype, nt_stream_t stream) {
  NT_RO_F32("nt_weighted_col_sum");
  NT_CHECK_ARG(d > 0 && n >= 0 && out, "nt_weighted_col_sum: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    NT_CUDA(cudaMemsetAsync(out, 0, (size_t)d * sizeof(float), st));
    return NT_OK;
  }
  NT_CHECK_ARG(x, "nt_weighted_col_sum: null pointer");
  if (!workspace || workspace_bytes < nt_weighted_col_sum_workspace_bytes(n, d)) {
    set_error("nt_weighted_col_sum: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  const int64_t nblk = cdiv(n, WC_ROWS);
  float* part = static_cast<float*>(workspace);
  weighted_colsum_partial<<<dim3((unsigned)cdiv(d, RO_THREADS), (unsigned)nblk), RO_THREADS, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(w),
                                                                                                     n, (int)d, part);
  weighted_colsum_final<<<(unsigned)cdiv(d, RO_THREADS), RO_THREADS, 0, st>>>(part, nblk, (int)d, scale, static_cast<float*>(out));
  NT_LAUNCH_CHECK("nt_weighted_col_sum", 2);
  return NT_OK;
}
