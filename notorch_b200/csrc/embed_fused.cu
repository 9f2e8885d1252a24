// GraphEmbedding fused into the edge initialisation ("fuse into edge_init").
//
// Reference: notorch/nn/gnn/embed.py:20-24 (two nn.EmbeddingBag(mode="sum"): x_v = bag(node.weight, node_types [V, t_v]),
// x_e = bag(edge.weight, edge_types [E, t_e])) followed by notorch/nn/gnn/chemprop.py:83 (h0 = x_v[src] + x_e).
//
//   forward   h0[e, :] = (sum_j Tv[node_types[src[e], j], :]) + (sum_k Te[edge_types[e, k], :])
//             Both tables live in shared memory for the whole kernel (58 x d fp32 = 70 KB at d = 300); the only HBM traffic is
//             the integer ids and ONE store of h0. x_v [V, d] and x_e [E, d] are never materialised. The bag sums run in slot
//             order from a zero accumulator and the two sums are added last, i.e. exactly the arithmetic of
//             nt_embedding_bag_sum + nt_gather_add: the fused result is bit-identical to the unfused one.
//   backward  gTe[t, :] = sum_{(e,k): edge_types[e,k]=t} g[e, :]      gTv[t, :] = sum_{(e,j): node_types[src[e],j]=t} g[e, :]
//             ONE pass over g = g_{h0} [E, d] (the unfused path reads it twice and makes a [V, d] intermediate with K5).
//             A CTA holds G private copies of the combined gradient table [Tv + Te, cols] in shared memory, one per "row group"
//             of ceil(cols / 128) warps; a row group walks ITS contiguous range of edges in ascending order, every thread owns
//             one 16-byte column chunk and does one shared-memory read-modify-write per (edge, slot) — no atomics, no barrier in
//             the loop, no two threads ever touch the same word. The G copies, then the CTAs' tables, are added in fixed order:
//             run-to-run deterministic. Bound by the shared-memory pipe: (t_v + t_e) x 2 x d x 4 bytes per edge.
#include "common.cuh"

namespace nt {

constexpr int EF_THREADS = 256;
constexpr int EF_ROWS = 128;          // edges per staged id block
constexpr int EF_SMEM_MAX = 200 * 1024;

__global__ void __launch_bounds__(EF_THREADS)
embed_edge_init_kernel(const float* __restrict__ tab_v, int Tv, const float* __restrict__ tab_e, int Te, const int64_t* __restrict__ node_types, int bv,
                       const int64_t* __restrict__ edge_types, int be, const int32_t* __restrict__ src, int64_t E, int64_t V, int d, int col0, int cols,
                       float* __restrict__ h0, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunks = cols / 4, T = Tv + Te, S = bv + be;
  float4* tab = reinterpret_cast<float4*>(smem_raw);                    // [T][chunks]: node table rows, then edge table rows
  int* ids = reinterpret_cast<int*>(tab + (size_t)T * chunks);           // [EF_ROWS][S] row offsets into tab
  for (int i = threadIdx.x; i < T * chunks; i += EF_THREADS) {
    const int t = i / chunks, c = i - t * chunks;
    tab[i] = t < Tv ? ldg4(tab_v + (int64_t)t * d + col0 + 4 * c) : ldg4(tab_e + (int64_t)(t - Tv) * d + col0 + 4 * c);
  }
  const int64_t nblocks = (E + EF_ROWS - 1) / EF_ROWS;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t r0 = blk * EF_ROWS;
    const int rows = (int)((r0 + EF_ROWS < E ? r0 + EF_ROWS : E) - r0);
    __syncthreads();  // tables staged (first trip) / the previous block's ids are no longer read
    for (int i = threadIdx.x; i < rows * S; i += EF_THREADS) {
      const int r = i / S, j = i - r * S;
      int64_t k;
      int base, limit;
      if (j < bv) {
        int64_t s = __ldg(src + r0 + r);
        if (s < 0 || s >= V) { atomicOr(status, 1); s = 0; }
        k = __ldg(node_types + s * bv + j);
        base = 0; limit = Tv;
      } else {
        k = __ldg(edge_types + (r0 + r) * be + (j - bv));
        base = Tv; limit = Te;
      }
      if (k < 0 || k >= limit) { atomicOr(status, 1); k = 0; }
      ids[i] = (base + (int)k) * chunks;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < rows * chunks; u += EF_THREADS) {
      const int r = u / chunks, c = u - r * chunks;
      const int* ip = ids + r * S;
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), ae = av;
      for (int j = 0; j < bv; ++j) {  // slot order from zero, like EmbeddingBag(mode="sum")
        const float4 v = tab[ip[j] + c];
        av = make_float4(av.x + v.x, av.y + v.y, av.z + v.z, av.w + v.w);
      }
      for (int j = bv; j < S; ++j) {
        const float4 v = tab[ip[j] + c];
        ae = make_float4(ae.x + v.x, ae.y + v.y, ae.z + v.z, ae.w + v.w);
      }
      stg4_stream(h0 + (r0 + r) * d + col0 + 4 * c, make_float4(av.x + ae.x, av.y + ae.y, av.z + ae.z, av.w + ae.w));  // x_v[src] + x_e
    }
  }
}

constexpr int EFB_UNROLL = 4;  // edges whose ids and gradient rows are in flight per thread

__global__ void __launch_bounds__(1024)
embed_edge_init_bwd_kernel(const float* __restrict__ g, const int64_t* __restrict__ node_types, int bv, const int64_t* __restrict__ edge_types, int be,
                           const int32_t* __restrict__ src, int64_t E, int64_t V, int Tv, int Te, int d, int col0, int cols, int G, int warps_per_group,
                           int64_t rows_per_group, float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int chunks = cols / 4, T = Tv + Te, S = bv + be;
  float4* tab = reinterpret_cast<float4*>(smem_raw);  // [G][T][chunks]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = warp / warps_per_group;
  const int c = (warp - grp * warps_per_group) * 32 + lane;  // this thread's 16-byte column chunk inside the window
  const bool active = c < chunks;
  for (int i = tid; i < G * T * chunks; i += blockDim.x) tab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float4* my = tab + (size_t)grp * T * chunks + c;
  const int64_t q = (int64_t)blockIdx.x * G + grp;
  const int64_t e0 = q * rows_per_group;
  int64_t e1 = e0 + rows_per_group;
  if (e1 > E) e1 = E;
  for (int64_t e = e0; e < e1; e += EFB_UNROLL) {  // bounds are uniform over the row group: every warp takes the shuffles together
    int off[EFB_UNROLL];
    float4 gv[EFB_UNROLL];
#pragma unroll
    for (int u = 0; u < EFB_UNROLL; ++u) {
      off[u] = 0;
      gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e + u < e1) {
        if (lane < S) {  // lane j holds the table row of slot j of this edge (node slots first)
          int64_t k;
          int base, limit;
          if (lane < bv) {
            int64_t s = __ldg(src + e + u);
            if (s < 0 || s >= V) s = 0;
            k = __ldg(node_types + s * bv + lane);
            base = 0; limit = Tv;
          } else {
            k = __ldg(edge_types + (e + u) * be + (lane - bv));
            base = Tv; limit = Te;
          }
          if (k < 0 || k >= limit) k = 0;  // the forward pass has reported it (status flag); stay in bounds here
          off[u] = (base + (int)k) * chunks;
        }
        if (active) gv[u] = ldg4_stream(g + (e + u) * d + col0 + 4 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < EFB_UNROLL; ++u) {
      if (e + u < e1) {
        for (int j = 0; j < S; ++j) {
          const int o = __shfl_sync(0xffffffffu, off[u], j);
          if (active) {
            float4 a = my[o];
            my[o] = make_float4(a.x + gv[u].x, a.y + gv[u].y, a.z + gv[u].z, a.w + gv[u].w);
          }
        }
      }
    }
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * T * chunks;
  for (int i = tid; i < T * chunks; i += blockDim.x) {
    float4 s = tab[i];
    for (int k = 1; k < G; ++k) {  // ascending group order
      const float4 v = tab[(size_t)k * T * chunks + i];
      s = make_float4(s.x + v.x, s.y + v.y, s.z + v.z, s.w + v.w);
    }
    dst[i] = s;
  }
}

// out tables [Tv, d] / [Te, d], columns [col0, col0 + cols): sum of the CTAs' tables in ascending CTA order
__global__ void __launch_bounds__(256) embed_edge_init_bwd_reduce(const float* __restrict__ partial, int nblk, int Tv, int Te, int d, int col0, int cols,
                                                                  float* __restrict__ g_tab_v, float* __restrict__ g_tab_e) {
  const int chunks = cols / 4, T = Tv + Te;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= T * chunks) return;
  const float4* p = reinterpret_cast<const float4*>(partial) + i;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int b = 0;
  for (; b + 4 <= nblk; b += 4) {
    const float4 v0 = __ldg(p + (size_t)b * T * chunks), v1 = __ldg(p + (size_t)(b + 1) * T * chunks);
    const float4 v2 = __ldg(p + (size_t)(b + 2) * T * chunks), v3 = __ldg(p + (size_t)(b + 3) * T * chunks);
    s = make_float4(s.x + v0.x, s.y + v0.y, s.z + v0.z, s.w + v0.w);
    s = make_float4(s.x + v1.x, s.y + v1.y, s.z + v1.z, s.w + v1.w);
    s = make_float4(s.x + v2.x, s.y + v2.y, s.z + v2.z, s.w + v2.w);
    s = make_float4(s.x + v3.x, s.y + v3.y, s.z + v3.z, s.w + v3.w);
  }
  for (; b < nblk; ++b) {
    const float4 v = __ldg(p + (size_t)b * T * chunks);
    s = make_float4(s.x + v.x, s.y + v.y, s.z + v.z, s.w + v.w);
  }
  const int t = i / chunks, c = i - t * chunks;
  float* out = t < Tv ? g_tab_v + (int64_t)t * d : g_tab_e + (int64_t)(t - Tv) * d;
  stg4(out + col0 + 4 * c, s);
}

// ---- launch geometry (shared by the workspace query and the launcher) ----------------------------------------------------
struct EfbPlan {
  int cols;             // columns per pass (multiple of 4)
  int G;                // private tables (= row groups) per CTA
  int warps_per_group;
  int grid;             // CTAs = partial tables
  size_t smem;
};

static bool efb_plan(int64_t E, int64_t T, int64_t d, EfbPlan* p) {
  const size_t budget = 208 * 1024;
  if (d % 4 != 0 || T <= 0 || T * 16 > (int64_t)budget) return false;
  int64_t cols = (int64_t)(budget / 3 / (size_t)(T * 4)) / 4 * 4;  // room for three private tables if the width allows it
  if (cols < 128) cols = (int64_t)(budget / (size_t)(T * 4)) / 4 * 4;  // wide vocabulary: fewer, wider-than-nothing tables
  if (cols > d) cols = d;
  if (cols < 4) return false;
  if (cols > 4096) cols = 4096;  // 1024 threads x 16 bytes
  p->cols = (int)cols;
  const int chunks = (int)cols / 4;
  p->warps_per_group = (chunks + 31) / 32;
  int G = (int)(budget / ((size_t)T * cols * 4));
  if (G > 4) G = 4;
  while (G > 1 && G * p->warps_per_group > 32) --G;
  if (G < 1) return false;
  p->G = G;
  p->smem = (size_t)G * T * cols * 4;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int64_t want = cdiv(E, 64 * G);  // at least ~64 edges per row group
  p->grid = (int)(want < sms ? (want < 1 ? 1 : want) : sms);
  return true;
}

}  // namespace nt

using namespace nt;

extern "C" int nt_embed_edge_init(const void* table_v, int64_t num_node_types, const void* table_e, int64_t num_edge_types, const int64_t* node_types,
                                  int64_t bag_v, const int64_t* edge_types, int64_t bag_e, const int32_t* src, int64_t E, int64_t V, int64_t d, void* h0,
                                  int32_t* status, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_embed_edge_init: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(num_node_types > 0 && num_edge_types > 0 && bag_v > 0 && bag_e > 0 && bag_v + bag_e <= 32 && E >= 0 && E < INT32_MAX && V >= 0 &&
                   V < INT32_MAX && d > 0 && d < (1 << 20) && num_node_types + num_edge_types < (1 << 20),
               "nt_embed_edge_init: bad sizes");
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(table_v && table_e && node_types && edge_types && src && h0 && status, "nt_embed_edge_init: null pointer");
  if (d % 4 != 0 || !aligned16(table_v) || !aligned16(table_e) || !aligned16(h0)) {
    set_error("nt_embed_edge_init: needs d %% 4 == 0 and 16-byte aligned tables / output (use nt_embedding_bag_sum + nt_gather_add)");
    return NT_ERR_UNSUPPORTED;
  }
  const int64_t T = num_node_types + num_edge_types;
  const size_t id_bytes = (size_t)EF_ROWS * (bag_v + bag_e) * sizeof(int);
  if ((size_t)T * 16 + id_bytes > (size_t)EF_SMEM_MAX) { set_error("nt_embed_edge_init: vocabulary too large for shared memory"); return NT_ERR_UNSUPPORTED; }
  int64_t cols = (int64_t)(((size_t)EF_SMEM_MAX - id_bytes) / ((size_t)T * 4)) / 4 * 4;
  if (cols > d) cols = d;
  static PerDeviceOnce once;
  NT_CUDA(once.run([] { return cudaFuncSetAttribute(embed_edge_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EF_SMEM_MAX); }));
  cudaStream_t st = as_stream(stream);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  int launches = 0;
  for (int64_t col0 = 0; col0 < d; col0 += cols) {
    const int64_t w = col0 + cols <= d ? cols : d - col0;
    const size_t smem = (size_t)T * w * 4 + id_bytes;
    int per_sm = (int)((228 * 1024) / (smem + 1024));  // 228 KiB per SM, 1 KiB reserved per resident CTA (d = 300: three CTAs, 70 KB of tables each)
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t nblocks = cdiv(E, EF_ROWS);
    if (grid > nblocks) grid = nblocks;
    embed_edge_init_kernel<<<(unsigned)grid, EF_THREADS, smem, st>>>(static_cast<const float*>(table_v), (int)num_node_types,
                                                                     static_cast<const float*>(table_e), (int)num_edge_types, node_types, (int)bag_v,
                                                                     edge_types, (int)bag_e, src, E, V, (int)d, (int)col0, (int)w,
                                                                     static_cast<float*>(h0), status);
    ++launches;
  }
  NT_LAUNCH_CHECK("nt_embed_edge_init", launches);
  return NT_OK;
}

extern "C" size_t nt_embed_edge_init_backward_workspace_bytes(int64_t E, int64_t num_node_types, int64_t num_edge_types, int64_t d) {
  EfbPlan plan;
  if (E <= 0 || d <= 0 || !efb_plan(E, num_node_types + num_edge_types, d, &plan)) return 0;
  return (size_t)plan.grid * (size_t)(num_node_types + num_edge_types) * plan.cols * sizeof(float) + 256;
}

extern "C" int nt_embed_edge_init_backward(const void* g, const int64_t* node_types, int64_t bag_v, const int64_t* edge_types, int64_t bag_e,
                                           const int32_t* src, int64_t E, int64_t V, int64_t num_node_types, int64_t num_edge_types, int64_t d,
                                           void* g_table_v, void* g_table_e, void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_embed_edge_init_backward: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(num_node_types > 0 && num_edge_types > 0 && bag_v > 0 && bag_e > 0 && bag_v + bag_e <= 32 && E >= 0 && E < INT32_MAX && V >= 0 &&
                   V < INT32_MAX && d > 0 && d < (1 << 20) && num_node_types + num_edge_types < (1 << 20),
               "nt_embed_edge_init_backward: bad sizes");
  NT_CHECK_ARG(g_table_v && g_table_e, "nt_embed_edge_init_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  if (E == 0) {
    NT_CUDA(cudaMemsetAsync(g_table_v, 0, (size_t)num_node_types * d * sizeof(float), st));
    NT_CUDA(cudaMemsetAsync(g_table_e, 0, (size_t)num_edge_types * d * sizeof(float), st));
    return NT_OK;
  }
  NT_CHECK_ARG(g && node_types && edge_types && src, "nt_embed_edge_init_backward: null pointer");
  EfbPlan plan;
  if (!aligned16(g) || !aligned16(g_table_v) || !aligned16(g_table_e) || !efb_plan(E, num_node_types + num_edge_types, d, &plan)) {
    set_error("nt_embed_edge_init_backward: needs d %% 4 == 0, 16-byte aligned rows and a vocabulary that fits shared memory");
    return NT_ERR_UNSUPPORTED;
  }
  if (!workspace || !aligned16(workspace) || workspace_bytes < nt_embed_edge_init_backward_workspace_bytes(E, num_node_types, num_edge_types, d)) {
    set_error("nt_embed_edge_init_backward: workspace too small (nt_embed_edge_init_backward_workspace_bytes)");
    return NT_ERR_WORKSPACE;
  }
  static PerDeviceOnce once;
  NT_CUDA(once.run([] { return cudaFuncSetAttribute(embed_edge_init_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); }));
  const int T = (int)(num_node_types + num_edge_types);
  const int64_t rows_per_group = cdiv(E, (int64_t)plan.grid * plan.G);
  int launches = 0;
  for (int64_t col0 = 0; col0 < d; col0 += plan.cols) {
    const int w = (int)(col0 + plan.cols <= d ? plan.cols : d - col0);
    embed_edge_init_bwd_kernel<<<plan.grid, plan.G * plan.warps_per_group * 32, (size_t)plan.G * T * w * 4, st>>>(
        static_cast<const float*>(g), node_types, (int)bag_v, edge_types, (int)bag_e, src, E, V, (int)num_node_types, (int)num_edge_types, (int)d,
        (int)col0, w, plan.G, plan.warps_per_group, rows_per_group, static_cast<float*>(workspace));
    embed_edge_init_bwd_reduce<<<(unsigned)cdiv((int64_t)T * (w / 4), 256), 256, 0, st>>>(static_cast<const float*>(workspace), plan.grid,
                                                                                         (int)num_node_types, (int)num_edge_types, (int)d, (int)col0, w,
                                                                                         static_cast<float*>(g_table_v), static_cast<float*>(g_table_e));
    launches += 2;
  }
  NT_LAUNCH_CHECK("nt_embed_edge_init_backward", launches);
  return NT_OK;
}
