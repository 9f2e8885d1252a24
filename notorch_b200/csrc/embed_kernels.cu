// GraphEmbedding: the step immediately before the message-passing block.
// Reference: notorch/nn/gnn/embed.py:20-24 — two nn.EmbeddingBag(mode="sum") over 2-D index tensors
// ([V, t_v] atom type ids, [E, t_e] bond type ids): out[i,:] = sum_j table[idx[i,j],:].
// Forward is a gather-sum (tables are tiny and stay in L1/L2); backward is a deterministic two-stage
// histogram-style reduction (per-CTA partial tables in shared memory, then a fixed-order sum) — the stock
// path uses atomics (embedding_bag backward), this one does not.
#include "common.cuh"

namespace nt {

constexpr int EMB_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(EMB_THREADS) embedding_bag_sum_kernel(const float* __restrict__ table, int64_t num_types, const int64_t* __restrict__ idx,
                                                                       int bag, int d, int chunks, int64_t total, float* __restrict__ out,
                                                                       int32_t* __restrict__ status) {
  int64_t t = (int64_t)blockIdx.x * EMB_THREADS + threadIdx.x;
  if (t >= total) return;
  const int64_t i = t / chunks;
  const int c = (int)(t - i * chunks) * (VEC ? 4 : 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < bag; ++j) {  // sequential in bag order, like EmbeddingBag(mode="sum")
    int64_t k = __ldg(idx + i * bag + j);
    if (k < 0 || k >= num_types) {
      if (c == 0) atomicOr(status, 1);
      k = 0;
    }
    if (VEC) {
      float4 v = ldg4(table + k * d + c);
      acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
    } else {
      acc.x += __ldg(table + k * d + c);
    }
  }
  if (VEC) stg4(out + i * d + c, acc);
  else out[i * d + c] = acc.x;
}

// stage 1: CTA b owns rows [b*rows_per_cta, ...); thread = feature column; partial[b][type][col].
// The CTA's index tile is staged in shared memory (coalesced), gradient rows are fetched 8 at a time so that
// eight independent global loads are in flight per thread before the (serial, per-column) accumulation.
__global__ void __launch_bounds__(1024) embedding_bag_bwd_partial(const float* __restrict__ g, const int64_t* __restrict__ idx, int64_t n, int bag,
                                                                 int num_types, int d, int rows_per_cta, int col0, int cols,
                                                                 float* __restrict__ partial) {
  extern __shared__ float acc[];  // [num_types][cols] then int32 idx_s[rows_per_cta * bag]
  int* idx_s = reinterpret_cast<int*>(acc + (size_t)num_types * cols);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int rows = (int)((r0 + rows_per_cta < n ? r0 + rows_per_cta : n) - r0);
  for (int i = threadIdx.x; i < num_types * cols; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < rows * bag; i += blockDim.x) {
    int64_t k = __ldg(idx + r0 * bag + i);
    idx_s[i] = (k < 0 || k >= num_types) ? 0 : (int)k;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float* gp = g + r0 * d + col0 + c;
    for (int rb = 0; rb < rows; rb += 8) {
      float gv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) gv[u] = (rb + u < rows) ? __ldg(gp + (int64_t)(rb + u) * d) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (rb + u < rows) {
          const int* ip = idx_s + (rb + u) * bag;
          for (int j = 0; j < bag; ++j) acc[ip[j] * cols + c] += gv[u];  // only this thread touches column c: sequential, deterministic
        }
      }
    }
  }
  __syncthreads();
  float* dst = partial + (int64_t)blockIdx.x * num_types * d;
  for (int i = threadIdx.x; i < num_types * cols; i += blockDim.x) {
    const int k = i / cols, c = i - k * cols;
    dst[(int64_t)k * d + col0 + c] = acc[i];
  }
}

// stage 2 (used twice, fixed order => deterministic): out[grp][t] = sum of in[grp*group .. min(nblk, (grp+1)*group))[t]
__global__ void __launch_bounds__(EMB_THREADS) embedding_bag_bwd_reduce(const float* __restrict__ in, int64_t nblk, int group, int64_t elems,
                                                                       float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * EMB_THREADS + threadIdx.x;
  if (t >= elems) return;
  const int64_t b0 = (int64_t)blockIdx.y * group;
  const int64_t b1 = b0 + group < nblk ? b0 + group : nblk;
  float s = 0.f;
  int64_t b = b0;
  for (; b + 8 <= b1; b += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(in + (b + u) * elems + t);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; b < b1; ++b) s += __ldg(in + b * elems + t);
  out[(int64_t)blockIdx.y * elems + t] = s;
}

constexpr int EMB_ROWS_PER_CTA = 128;
constexpr int EMB_SMEM_BUDGET = 96 * 1024;
constexpr int EMB_GROUP = 32;

}  // namespace nt

using namespace nt;

extern "C" int nt_embedding_bag_sum(const void* table, int64_t num_types, const int64_t* idx, int64_t n, int64_t bag, int64_t d, void* out,
                                    int32_t* status, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_embedding_bag_sum: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(num_types > 0 && n >= 0 && n < INT32_MAX && bag > 0 && bag < 4096 && d > 0 && d < (1 << 20), "nt_embedding_bag_sum: bad sizes");
  if (n == 0) return NT_OK;
  NT_CHECK_ARG(table && idx && out && status, "nt_embedding_bag_sum: null pointer");
  cudaStream_t st = as_stream(stream);
  const bool vec = d % 4 == 0 && aligned16(table) && aligned16(out);
  const int chunks = vec ? (int)(d / 4) : (int)d;
  const int64_t total = n * chunks;
  if (vec) embedding_bag_sum_kernel<true><<<(unsigned)cdiv(total, EMB_THREADS), EMB_THREADS, 0, st>>>(static_cast<const float*>(table), num_types, idx, (int)bag, (int)d, chunks, total, static_cast<float*>(out), status);
  else embedding_bag_sum_kernel<false><<<(unsigned)cdiv(total, EMB_THREADS), EMB_THREADS, 0, st>>>(static_cast<const float*>(table), num_types, idx, (int)bag, (int)d, chunks, total, static_cast<float*>(out), status);
  NT_LAUNCH_CHECK("nt_embedding_bag_sum", 1);
  return NT_OK;
}

extern "C" size_t nt_embedding_bag_backward_workspace_bytes(int64_t n, int64_t num_types, int64_t d) {
  if (n <= 0 || num_types <= 0 || d <= 0) return 256;
  const size_t nblk = (size_t)cdiv(n, EMB_ROWS_PER_CTA);
  return (nblk + (size_t)cdiv((int64_t)nblk, EMB_GROUP)) * (size_t)num_types * (size_t)d * sizeof(float) + 256;
}

extern "C" int nt_embedding_bag_backward(const void* g, const int64_t* idx, int64_t n, int64_t bag, int64_t num_types, int64_t d, void* g_table,
                                         void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_embedding_bag_backward: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(num_types > 0 && num_types < (1 << 20) && n >= 0 && n < INT32_MAX && bag > 0 && bag < 4096 && d > 0 && d < (1 << 20),
               "nt_embedding_bag_backward: bad sizes");
  NT_CHECK_ARG(g_table, "nt_embedding_bag_backward: null g_table");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    NT_CUDA(cudaMemsetAsync(g_table, 0, (size_t)num_types * d * sizeof(float), st));
    return NT_OK;
  }
  NT_CHECK_ARG(g && idx, "nt_embedding_bag_backward: null pointer");
  if (!workspace || workspace_bytes < nt_embedding_bag_backward_workspace_bytes(n, num_types, d)) {
    set_error("nt_embedding_bag_backward: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  // columns per pass so that the per-CTA partial table [num_types][cols] fits the shared-memory budget
  const int64_t idx_bytes = (int64_t)EMB_ROWS_PER_CTA * bag * sizeof(int);
  int64_t cols = (EMB_SMEM_BUDGET - idx_bytes) / (int64_t)(num_types * sizeof(float));
  if (cols < 1) { set_error("nt_embedding_bag_backward: vocabulary too large (%lld types)", (long long)num_types); return NT_ERR_UNSUPPORTED; }
  if (cols > d) cols = d;
  static bool attr_set = false;
  if (!attr_set) {
    NT_CUDA(cudaFuncSetAttribute(embedding_bag_bwd_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, EMB_SMEM_BUDGET));
    attr_set = true;
  }
  const int64_t nblk = cdiv(n, EMB_ROWS_PER_CTA);
  float* partial = static_cast<float*>(workspace);
  int launches = 0;
  for (int64_t col0 = 0; col0 < d; col0 += cols) {
    const int64_t w = col0 + cols <= d ? cols : d - col0;
    embedding_bag_bwd_partial<<<(unsigned)nblk, (unsigned)(w >= 1024 ? 1024 : (w + 31) / 32 * 32), (size_t)(num_types * w * sizeof(float) + idx_bytes), st>>>(
        static_cast<const float*>(g), idx, n, (int)bag, (int)num_types, (int)d, EMB_ROWS_PER_CTA, (int)col0, (int)w, partial);
    ++launches;
  }
  const int64_t elems = num_types * d;
  const int64_t ngrp = cdiv(nblk, EMB_GROUP);
  float* partial2 = partial + nblk * elems;
  embedding_bag_bwd_reduce<<<dim3((unsigned)cdiv(elems, EMB_THREADS), (unsigned)ngrp), EMB_THREADS, 0, st>>>(partial, nblk, EMB_GROUP, elems, partial2);
  embedding_bag_bwd_reduce<<<dim3((unsigned)cdiv(elems, EMB_THREADS), 1), EMB_THREADS, 0, st>>>(partial2, ngrp, (int)ngrp, elems,
                                                                                                static_cast<float*>(g_table));
  NT_LAUNCH_CHECK("nt_embedding_bag_backward", launches + 2);
  return NT_OK;
}
