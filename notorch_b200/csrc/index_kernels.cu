// Integer kernels: batch collation (K-l) and stable CSR construction.
// Replaces BatchedGraph.from_graphs (reference notorch/data/models/graph.py:186-223) and the index
// handling inside torch_scatter.scatter / aten::index (chemprop.py:39-40, agg.py:27,36).
// All results are bit-exact functions of the inputs: integer atomics only feed a histogram
// (order-independent) and the within-segment order is restored by a rank sort.
#include "common.cuh"

namespace nt {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;                          // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 elements per block

// ---- three-phase exclusive scan over int32 (deterministic) ----------------------------------
// phase 1: per-tile sums
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t j = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (j < n) s += in[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += warp_sums[w];
    tile_sums[blockIdx.x] = t;
  }
}

// phase 2: exclusive scan of the tile sums by ONE block (tiles <= a few thousand)
__global__ void __launch_bounds__(1024) scan_tile_offsets(int32_t* __restrict__ tile_sums, int64_t num_tiles) {
  __shared__ int32_t sh[1024];
  __shared__ int32_t carry_sh;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  for (int64_t start = 0; start < num_tiles; start += 1024) {
    int64_t j = start + threadIdx.x;
    int32_t v = j < num_tiles ? tile_sums[j] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      int32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    int32_t carry = carry_sh;
    if (j < num_tiles) tile_sums[j] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_sh = carry + sh[1023];
    __syncthreads();
  }
}

// phase 3: per-tile exclusive scan + tile offset; writes out[0..n) and out[n] = total
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const int32_t* __restrict__ in, int64_t n,
                                                           const int32_t* __restrict__ tile_offsets, int32_t* __restrict__ out) {
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;  // blocked arrangement
  int32_t v[SCAN_ITEMS];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  // inclusive warp scan of the per-thread sums
  int32_t incl = s;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int32_t warp_off = 0;
  for (int w = 0; w < warp; ++w) warp_off += warp_sums[w];
  int32_t run = tile_offsets[blockIdx.x] + warp_off + incl - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
    if (base + i == n - 1) out[n] = run;
  }
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// exclusive scan: out[0..n] (n+1 entries, out[n] = total). scratch: cdiv(n, SCAN_TILE) int32.
static int exclusive_scan(const int32_t* in, int64_t n, int32_t* out, int32_t* scratch, cudaStream_t st) {
  int64_t tiles = cdiv(n, SCAN_TILE);
  if (tiles == 0) tiles = 1;
  scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, scratch);
  scan_tile_offsets<<<1, 1024, 0, st>>>(scratch, tiles);
  scan_apply<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, scratch, out);
  NT_LAUNCH_CHECK("exclusive_scan", 3);
  return NT_OK;
}

// ---- collation -------------------------------------------------------------------------------
// Largest m with ptr[m] <= x (ptr non-decreasing, ptr[0] = 0, ptr[B] > x).
__device__ __forceinline__ int upper_segment(const int32_t* __restrict__ ptr, int B, int x) {
  int lo = 0, hi = B;  // invariant: ptr[lo] <= x < ptr[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= x) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) collate_atoms(const int32_t* __restrict__ mol_atom_ptr, int B, int64_t V, int64_t* __restrict__ batch_node_index) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) batch_node_index[v] = upper_segment(mol_atom_ptr, B, (int)v);
}

__global__ void __launch_bounds__(256) collate_edges(const int32_t* __restrict__ mol_atom_ptr, const int32_t* __restrict__ mol_edge_ptr, int B,
                                                      const int32_t* __restrict__ lei, const int32_t* __restrict__ lrev, int64_t E,
                                                      int rev_offset_mode, int64_t* __restrict__ edge_index, int64_t* __restrict__ rev_index,
                                                      int64_t* __restrict__ batch_edge_index) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int m = upper_segment(mol_edge_ptr, B, (int)e);
  int64_t aoff = mol_atom_ptr[m];
  // graph.py:199-200: the cumulative ATOM count is added to edge_index AND rev_index
  edge_index[e] = (int64_t)lei[e] + aoff;
  edge_index[E + e] = (int64_t)lei[E + e] + aoff;
  rev_index[e] = (int64_t)lrev[e] + (rev_offset_mode == 0 ? aoff : (int64_t)mol_edge_ptr[m]);
  batch_edge_index[e] = m;
}

// ---- CSR ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_histogram(const int64_t* __restrict__ keys, int64_t n, int64_t S, int32_t* __restrict__ keys32,
                                                      int32_t* __restrict__ counts, int32_t* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = keys[i];
  bool ok = k >= 0 && k < S;
  if (!ok) {
    atomicOr(status, 1);
    k = 0;  // keep every later kernel in range; the host raises on status
  }
  if (keys32) keys32[i] = (int32_t)k;
  atomicAdd(counts + k, 1);
}

__global__ void __launch_bounds__(256) csr_fill(const int64_t* __restrict__ keys, int64_t n, int64_t S, const int32_t* __restrict__ rowptr,
                                                 int32_t* __restrict__ cursor, int32_t* __restrict__ perm_unsorted) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = keys[i];
  if (k < 0 || k >= S) k = 0;
  int32_t slot = atomicAdd(cursor + k, 1);
  perm_unsorted[rowptr[k] + slot] = (int32_t)i;
}

// Restore ascending-id order inside every segment: slot j goes to rowptr[s] + rank(j).
// Segments are short for molecular graphs (in-degree <= ~6); cost is O(len) per slot.
__global__ void __launch_bounds__(256) csr_rank_sort(const int64_t* __restrict__ keys, int64_t n, int64_t S, const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ perm_unsorted, int32_t* __restrict__ perm) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int32_t id = perm_unsorted[j];
  int64_t k = keys[id];
  if (k < 0 || k >= S) k = 0;
  int32_t lo = rowptr[k], hi = rowptr[k + 1];
  int32_t rank = 0;
  for (int32_t t = lo; t < hi; ++t) rank += (perm_unsorted[t] < id) ? 1 : 0;
  perm[lo + rank] = id;
}

}  // namespace nt

using namespace nt;

extern "C" size_t nt_collate_workspace_bytes(int64_t B) {
  return (size_t)(cdiv(B > 0 ? B : 1, SCAN_TILE) + 1) * sizeof(int32_t) + 256;
}

extern "C" int nt_collate(const int32_t* num_atoms, const int32_t* num_edges, int64_t B,
                          const int32_t* local_edge_index, const int32_t* local_rev_index, int64_t V, int64_t E,
                          int rev_offset_mode, int64_t* edge_index, int64_t* rev_index, int64_t* batch_node_index,
                          int64_t* batch_edge_index, int32_t* mol_atom_ptr, int32_t* mol_edge_ptr,
                          void* workspace, size_t workspace_bytes, nt_stream_t stream) {
  NT_CHECK_ARG(B >= 0 && V >= 0 && E >= 0 && B < INT32_MAX && V < INT32_MAX && E < INT32_MAX, "nt_collate: sizes out of range");
  NT_CHECK_ARG(mol_atom_ptr && mol_edge_ptr, "nt_collate: null row-pointer output");
  NT_CHECK_ARG(B == 0 || (num_atoms && num_edges), "nt_collate: null counts");
  NT_CHECK_ARG(rev_offset_mode == 0 || rev_offset_mode == 1, "nt_collate: bad rev_offset_mode");
  if (workspace_bytes < nt_collate_workspace_bytes(B) || !workspace) {
    set_error("nt_collate: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int32_t* scratch = static_cast<int32_t*>(workspace);
  int rc = exclusive_scan(num_atoms, B, mol_atom_ptr, scratch, st);
  if (rc) return rc;
  rc = exclusive_scan(num_edges, B, mol_edge_ptr, scratch, st);
  if (rc) return rc;
  if (V > 0) {
    NT_CHECK_ARG(batch_node_index, "nt_collate: null batch_node_index");
    collate_atoms<<<(unsigned)cdiv(V, 256), 256, 0, st>>>(mol_atom_ptr, (int)B, V, batch_node_index);
  }
  if (E > 0) {
    NT_CHECK_ARG(local_edge_index && local_rev_index && edge_index && rev_index && batch_edge_index, "nt_collate: null edge tensor");
    collate_edges<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(mol_atom_ptr, mol_edge_ptr, (int)B, local_edge_index, local_rev_index, E,
                                                          rev_offset_mode, edge_index, rev_index, batch_edge_index);
  }
  NT_LAUNCH_CHECK("nt_collate", (V > 0) + (E > 0));
  return NT_OK;
}

extern "C" size_t nt_build_csr_workspace_bytes(int64_t n, int64_t num_segments) {
  // counts [S] + cursor [S] + perm_unsorted [n] + scan scratch
  size_t s = (size_t)(num_segments > 0 ? num_segments : 1), m = (size_t)(n > 0 ? n : 1);
  return (2 * s + m + (size_t)cdiv((int64_t)s, SCAN_TILE) + 8) * sizeof(int32_t) + 256;
}

extern "C" int nt_build_csr(const int64_t* keys, int64_t n, int64_t num_segments, int32_t* keys32, int32_t* rowptr, int32_t* perm,
                            int32_t* status, void* workspace, size_t workspace_bytes, nt_stream_t stream) {
  NT_CHECK_ARG(n >= 0 && num_segments >= 0 && n < INT32_MAX && num_segments < INT32_MAX, "nt_build_csr: sizes out of range");
  NT_CHECK_ARG(rowptr && status, "nt_build_csr: null output");
  NT_CHECK_ARG(n == 0 || (keys && perm), "nt_build_csr: null keys/perm");
  NT_CHECK_ARG(n == 0 || num_segments > 0, "nt_build_csr: items but no segments");
  if (workspace_bytes < nt_build_csr_workspace_bytes(n, num_segments) || !workspace) {
    set_error("nt_build_csr: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  size_t s = (size_t)(num_segments > 0 ? num_segments : 1), m = (size_t)(n > 0 ? n : 1);
  int32_t* counts = static_cast<int32_t*>(workspace);
  int32_t* cursor = counts + s;
  int32_t* perm_unsorted = cursor + s;
  int32_t* scratch = perm_unsorted + m;
  NT_CUDA(cudaMemsetAsync(counts, 0, 2 * s * sizeof(int32_t), st));
  if (n > 0) csr_histogram<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(keys, n, num_segments, keys32, counts, status);
  int rc = exclusive_scan(counts, num_segments, rowptr, scratch, st);
  if (rc) return rc;
  if (n > 0) {
    csr_fill<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(keys, n, num_segments, rowptr, cursor, perm_unsorted);
    csr_rank_sort<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(keys, n, num_segments, rowptr, perm_unsorted, perm);
  }
  NT_LAUNCH_CHECK("nt_build_csr", n > 0 ? 3 : 0);
  return NT_OK;
}
