// GraphEmbedding: the step immediately before the message-passing block.
// Reference: notorch/nn/gnn/embed.py:20-24 — two nn.EmbeddingBag(mode="sum") over 2-D index tensors
// ([V, t_v] atom type ids, [E, t_e] bond type ids): out[i,:] = sum_j table[idx[i,j],:].
// Forward is a gather-sum (tables are tiny and stay in L1/L2); backward is a deterministic two-stage
// histogram-style reduction (per-CTA partial tables in shared memory, then a fixed-order sum) — the stock
// path uses atomics (embedding_bag backward), this one does not.
#include <stdlib.h>

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

// Forward with the whole table in shared memory (tables are tiny: 45 x d and 13 x d): persistent CTAs, each walks row blocks
// of EMBF_ROWS rows; the block's ids are staged (validated, int32) and every (row, float4 chunk) output is `bag` shared-memory
// reads and one streaming store. The generic kernel above reads the table through L1 for every output chunk and re-reads the
// int64 ids once per chunk (37 % of HBM peak).
constexpr int EMBF_ROWS = 64;

__global__ void __launch_bounds__(EMB_THREADS) embedding_bag_sum_smem(const float* __restrict__ table, int num_types, const int64_t* __restrict__ idx,
                                                                     int64_t n, int bag, int d, float* __restrict__ out, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunks = d / 4;
  float4* tab = reinterpret_cast<float4*>(smem_raw);                       // [num_types][chunks]
  int* ids = reinterpret_cast<int*>(tab + (size_t)num_types * chunks);     // [EMBF_ROWS * bag]
  for (int i = threadIdx.x; i < num_types * chunks; i += EMB_THREADS) tab[i] = ldg4(table + 4 * (int64_t)i);
  const int64_t nblocks = (n + EMBF_ROWS - 1) / EMBF_ROWS;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t r0 = blk * EMBF_ROWS;
    const int rows = (int)((r0 + EMBF_ROWS < n ? r0 + EMBF_ROWS : n) - r0);
    __syncthreads();  // table staged (first trip) / previous block's ids no longer in use
    for (int i = threadIdx.x; i < rows * bag; i += EMB_THREADS) {
      int64_t k = __ldg(idx + r0 * bag + i);
      if (k < 0 || k >= num_types) {
        atomicOr(status, 1);
        k = 0;
      }
      ids[i] = (int)k * chunks;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < rows * chunks; u += EMB_THREADS) {
      const int r = u / chunks, c = u - r * chunks;
      const int* ip = ids + r * bag;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < bag; ++j) {  // sequential in bag order, like EmbeddingBag(mode="sum")
        const float4 v = tab[ip[j] + c];
        acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
      }
      stg4_stream(out + (r0 + r) * d + 4 * c, acc);
    }
  }
}

// stage 1 (persistent CTAs): CTA b walks the row blocks b, b + gridDim.x, ... of R rows and accumulates its own partial
// table [num_types][cols] in shared memory.
//   per block:  (1) the R x cols slice of g is copied into shared memory with cp.async (double buffered: the next block
//                   streams in while this one is reduced; zero-filled past n);
//               (2) the R * bag (row, slot) pairs are counting-sorted by type into per-type row lists, STABLY (ascending
//                   (row, slot) order) with warp match_any + a scan over warps - no atomics, so the order is fixed;
//               (3) work item = (type, float4 column chunk): the thread sums g_s[row][chunk] over that type's list in a
//                   REGISTER (independent shared-memory loads, one FADD chain) and adds the result to the table once.
// The first version kept one thread per column and did one shared-memory read-modify-write per (row, slot): a serial
// chain of ~900 dependent LDS/FADD/STS per thread and block (17 % of HBM peak). Item -> thread is a fixed mapping, so every
// table entry is only ever touched by one thread, and two runs add the same numbers in the same order.
constexpr int EMBB_THREADS = 512;
constexpr int EMBB_MAX_PAIRS = 1024;  // R * bag

__global__ void __launch_bounds__(EMBB_THREADS, 1)
embedding_bag_bwd_partial(const float* __restrict__ g, const int64_t* __restrict__ idx, int64_t n, int bag, int num_types, int d, int R,
                          int col0, int cols, float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunks = cols / 4;                 // float4 chunks per row slice (cols % 4 == 0)
  float4* gbuf[2];
  gbuf[0] = reinterpret_cast<float4*>(smem_raw);
  gbuf[1] = gbuf[0] + (size_t)R * chunks;
  float4* table = gbuf[1] + (size_t)R * chunks;                       // [num_types][chunks]
  int* offsets = reinterpret_cast<int*>(table + (size_t)num_types * chunks);  // [num_types + 1]
  int* warp_cnt = offsets + num_types + 1;                             // [EMBB_THREADS / 32 * passes][num_types] -> sized [32][num_types]
  int* list = warp_cnt + 32 * num_types;   // [R * bag] (row within the block) * chunks of every sorted pair
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nblocks = (n + R - 1) / R;
  const int pairs_per_block = R * bag;

  for (int i = tid; i < num_types * chunks; i += EMBB_THREADS) table[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto issue_copy = [&](int64_t blk, int buf) {
    if (blk < nblocks) {
      const int64_t r0 = blk * R;
      for (int u = tid; u < R * chunks; u += EMBB_THREADS) {
        const int r = u / chunks, c = u - r * chunks;
        const bool ok = r0 + r < n;
        const float* src = ok ? g + (r0 + r) * d + col0 + 4 * c : g;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(gbuf[buf] + u);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  issue_copy(blockIdx.x, 0);
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, buf ^= 1) {
    issue_copy(blk + gridDim.x, buf ^ 1);
    const int64_t r0 = blk * R;
    const int rows = (int)((r0 + R < n ? r0 + R : n) - r0);
    const int pairs = rows * bag;
    // ---- (2) stable counting sort of the (row, slot) pairs by type ----
    for (int i = tid; i < 32 * num_types; i += EMBB_THREADS) warp_cnt[i] = 0;
    __syncthreads();
    // pass structure: pair t is handled by thread t % 256 in pass t / 256; (pass, warp) = a virtual warp index vw < 32
    int key[EMBB_MAX_PAIRS / EMBB_THREADS], rank[EMBB_MAX_PAIRS / EMBB_THREADS];
#pragma unroll
    for (int ps = 0; ps < EMBB_MAX_PAIRS / EMBB_THREADS; ++ps) {
      const int t = ps * EMBB_THREADS + tid;
      key[ps] = -1;
      rank[ps] = 0;
      if (ps * EMBB_THREADS < pairs_per_block) {  // uniform per pass
        int k = -1;
        if (t < pairs) {
          const int64_t kk = __ldg(idx + r0 * bag + t);
          k = (kk < 0 || kk >= num_types) ? 0 : (int)kk;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (k >= 0) {
          key[ps] = k;
          rank[ps] = __popc(peers & ((1u << lane) - 1u));
          if (rank[ps] == 0) warp_cnt[(ps * (EMBB_THREADS / 32) + warp) * num_types + k] = __popc(peers);
        }
      }
    }
    __syncthreads();
    // exclusive scan over virtual warps for each type (thread k), then over types (thread 0)
    if (tid < num_types) {
      int run = 0;
      for (int vw = 0; vw < 32; ++vw) {
        const int c = warp_cnt[vw * num_types + tid];
        warp_cnt[vw * num_types + tid] = run;
        run += c;
      }
      offsets[tid + 1] = run;  // total of this type, scanned below
    }
    __syncthreads();
    if (tid == 0) {
      offsets[0] = 0;
      for (int k = 0; k < num_types; ++k) offsets[k + 1] += offsets[k];
    }
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < EMBB_MAX_PAIRS / EMBB_THREADS; ++ps) {
      if (key[ps] >= 0) {
        const int t = ps * EMBB_THREADS + tid;
        const int pos = offsets[key[ps]] + warp_cnt[(ps * (EMBB_THREADS / 32) + warp) * num_types + key[ps]] + rank[ps];
        list[pos] = (t / bag) * chunks;
      }
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this block's slice has landed (the next one may still be in flight)
    __syncthreads();
    // ---- (3) per (type, chunk) register reduction ----
    const float4* gs = gbuf[buf];
    for (int item = tid; item < num_types * chunks; item += EMBB_THREADS) {
      const int k = item / chunks, c = item - k * chunks;
      const int lo = offsets[k], hi = offsets[k + 1];
      if (lo == hi) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int pidx = lo;
      for (; pidx + 4 <= hi; pidx += 4) {
        const float4 v0 = gs[list[pidx] + c], v1 = gs[list[pidx + 1] + c];
        const float4 v2 = gs[list[pidx + 2] + c], v3 = gs[list[pidx + 3] + c];
        acc = make_float4(acc.x + v0.x, acc.y + v0.y, acc.z + v0.z, acc.w + v0.w);
        acc = make_float4(acc.x + v1.x, acc.y + v1.y, acc.z + v1.z, acc.w + v1.w);
        acc = make_float4(acc.x + v2.x, acc.y + v2.y, acc.z + v2.z, acc.w + v2.w);
        acc = make_float4(acc.x + v3.x, acc.y + v3.y, acc.z + v3.z, acc.w + v3.w);
      }
      for (; pidx < hi; ++pidx) {
        const float4 v = gs[list[pidx] + c];
        acc = make_float4(acc.x + v.x, acc.y + v.y, acc.z + v.z, acc.w + v.w);
      }
      float4 tv = table[item];
      table[item] = make_float4(tv.x + acc.x, tv.y + acc.y, tv.z + acc.z, tv.w + acc.w);
    }
    __syncthreads();  // the buffer, the lists and the offsets are rewritten by the next iteration
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* dst = partial + (int64_t)blockIdx.x * num_types * d;
  for (int i = tid; i < num_types * chunks; i += EMBB_THREADS) {
    const int k = i / chunks, c = i - k * chunks;
    *reinterpret_cast<float4*>(dst + (int64_t)k * d + col0 + 4 * c) = table[i];
  }
}

// scalar variant for d % 4 != 0 or unaligned bases: thread = column, one shared-memory read-modify-write per (row, slot)
__global__ void __launch_bounds__(1024) embedding_bag_bwd_partial_s(const float* __restrict__ g, const int64_t* __restrict__ idx, int64_t n, int bag,
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
    for (int rb = 0; rb < rows; ++rb) {
      const float gv = __ldg(gp + (int64_t)rb * d);
      const int* ip = idx_s + rb * bag;
      for (int j = 0; j < bag; ++j) acc[ip[j] * cols + c] += gv;
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

constexpr int EMB_ROWS_PER_CTA = 128;   // scalar variant
constexpr int EMB_SMEM_BUDGET = 96 * 1024;
constexpr int EMB_GROUP = 32;
constexpr int EMBB_SMEM_BUDGET = 220 * 1024;
constexpr int EMBB_MAX_COLS = 320;      // columns per pass of the vectorised variant

struct EmbbPlan {
  int R, cols;          // rows per block, columns per pass
  size_t smem;
  int64_t nblk;         // partial tables (= CTAs) of the vectorised variant
};

static bool embb_plan(int64_t n, int64_t bag, int64_t num_types, int64_t d, EmbbPlan* plan) {
  if (d % 4 != 0 || bag > 64) return false;
  int cols = (int)(d < EMBB_MAX_COLS ? d : EMBB_MAX_COLS);
  int R = 64;
  while (R * bag > EMBB_MAX_PAIRS) R /= 2;
  for (;;) {
    size_t sm = 2 * (size_t)R * cols * 4 + (size_t)num_types * cols * 4 + (size_t)(num_types + 1) * 4 + 32 * (size_t)num_types * 4 + (size_t)R * bag * 4 + 64;
    if (sm <= (size_t)EMBB_SMEM_BUDGET) { plan->smem = sm; break; }
    if (R > 8) R /= 2;
    else if (cols > 64) cols = cols / 2 / 4 * 4;
    else return false;
  }
  plan->R = R;
  plan->cols = cols;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int64_t blocks = cdiv(n, R);
  plan->nblk = blocks < sms ? blocks : sms;
  return true;
}

}  // namespace nt

namespace nt {
// embed_fused.cu: the table gradient as a skinny tensor-core GEMM against the on-the-fly count matrix (at most 128 types, d % 4 == 0)
int embed_bwd_mma(const float* g, const int64_t* node_types, int64_t bv, const int64_t* edge_types, int64_t be, const int32_t* src, int64_t n_rows, int64_t V,
                  int64_t Tv, int64_t Te, int64_t d, float* g_tab_v, float* g_tab_e, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t embed_bwd_mma_workspace_bytes(int64_t n_rows, int64_t T, int64_t d);
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
  const size_t smem_need = (size_t)num_types * d * sizeof(float) + (size_t)EMBF_ROWS * bag * sizeof(int);
  if (vec && smem_need <= 200 * 1024) {
    static PerDeviceOnce once;
    NT_CUDA(once.run([] { return cudaFuncSetAttribute(embedding_bag_sum_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
    int sms = num_sms();
    if (sms <= 0) sms = 148;
    int per_sm = (int)((220 * 1024) / (smem_need + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t nblocks = cdiv(n, EMBF_ROWS);
    if (grid > nblocks) grid = nblocks;
    embedding_bag_sum_smem<<<(unsigned)grid, EMB_THREADS, smem_need, st>>>(static_cast<const float*>(table), (int)num_types, idx, n, (int)bag, (int)d,
                                                                          static_cast<float*>(out), status);
    NT_LAUNCH_CHECK("nt_embedding_bag_sum", 1);
    return NT_OK;
  }
  const int chunks = vec ? (int)(d / 4) : (int)d;
  const int64_t total = n * chunks;
  if (vec) embedding_bag_sum_kernel<true><<<(unsigned)cdiv(total, EMB_THREADS), EMB_THREADS, 0, st>>>(static_cast<const float*>(table), num_types, idx, (int)bag, (int)d, chunks, total, static_cast<float*>(out), status);
  else embedding_bag_sum_kernel<false><<<(unsigned)cdiv(total, EMB_THREADS), EMB_THREADS, 0, st>>>(static_cast<const float*>(table), num_types, idx, (int)bag, (int)d, chunks, total, static_cast<float*>(out), status);
  NT_LAUNCH_CHECK("nt_embedding_bag_sum", 1);
  return NT_OK;
}

extern "C" size_t nt_embedding_bag_backward_workspace_bytes(int64_t n, int64_t num_types, int64_t d) {
  if (n <= 0 || num_types <= 0 || d <= 0) return 256;
  size_t nblk = (size_t)cdiv(n, EMB_ROWS_PER_CTA);  // scalar variant; the vectorised one never needs more than max(this, #SMs)
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  if (nblk < (size_t)sms) nblk = (size_t)sms;
  const size_t classic = (nblk + (size_t)cdiv((int64_t)nblk, EMB_GROUP)) * (size_t)num_types * (size_t)d * sizeof(float) + 256;
  const size_t mma = embed_bwd_mma_workspace_bytes(n, num_types, d);
  return classic > mma ? classic : mma;
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
  if (bag <= 64 && aligned16(g_table) && aligned16(workspace) && embed_bwd_mma_workspace_bytes(n, num_types, d) > 0) {
    // the common case (small vocabulary): tensor-core path shared with nt_embed_edge_init_backward (one id source, identity row map)
    static const bool classic = getenv("NOTORCH_B200_EMBBWD_CLASSIC") != nullptr && atoi(getenv("NOTORCH_B200_EMBBWD_CLASSIC")) != 0;
    if (!classic) {
      const int rc = embed_bwd_mma(static_cast<const float*>(g), idx, bag, nullptr, 0, nullptr, n, n, num_types, 0, d, static_cast<float*>(g_table),
                                   nullptr, workspace, workspace_bytes, st);
      if (rc != NT_ERR_UNSUPPORTED) return rc;
    }
  }
  float* partial = static_cast<float*>(workspace);
  int launches = 0;
  int64_t nblk = 0;
  EmbbPlan plan;
  if (aligned16(g) && embb_plan(n, bag, num_types, d, &plan)) {
    static PerDeviceOnce once_v;
    NT_CUDA(once_v.run([] { return cudaFuncSetAttribute(embedding_bag_bwd_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, EMBB_SMEM_BUDGET); }));
    nblk = plan.nblk;
    for (int64_t col0 = 0; col0 < d; col0 += plan.cols) {
      const int64_t w = col0 + plan.cols <= d ? plan.cols : d - col0;
      embedding_bag_bwd_partial<<<(unsigned)nblk, EMBB_THREADS, plan.smem, st>>>(static_cast<const float*>(g), idx, n, (int)bag, (int)num_types, (int)d,
                                                                                 plan.R, (int)col0, (int)w, partial);
      ++launches;
    }
  } else {
    // columns per pass so that the per-CTA partial table [num_types][cols] fits the shared-memory budget
    const int64_t idx_bytes = (int64_t)EMB_ROWS_PER_CTA * bag * sizeof(int);
    int64_t cols = (EMB_SMEM_BUDGET - idx_bytes) / (int64_t)(num_types * sizeof(float));
    if (cols < 1) { set_error("nt_embedding_bag_backward: vocabulary too large (%lld types)", (long long)num_types); return NT_ERR_UNSUPPORTED; }
    if (cols > d) cols = d;
    static PerDeviceOnce once_s;
    NT_CUDA(once_s.run([] { return cudaFuncSetAttribute(embedding_bag_bwd_partial_s, cudaFuncAttributeMaxDynamicSharedMemorySize, EMB_SMEM_BUDGET); }));
    nblk = cdiv(n, EMB_ROWS_PER_CTA);
    for (int64_t col0 = 0; col0 < d; col0 += cols) {
      const int64_t w = col0 + cols <= d ? cols : d - col0;
      embedding_bag_bwd_partial_s<<<(unsigned)nblk, (unsigned)(w >= 1024 ? 1024 : (w + 31) / 32 * 32), (size_t)(num_types * w * sizeof(float) + idx_bytes), st>>>(
          static_cast<const float*>(g), idx, n, (int)bag, (int)num_types, (int)d, EMB_ROWS_PER_CTA, (int)col0, (int)w, partial);
      ++launches;
    }
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
