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
//             ONE pass over g = g_{h0} [E, d] (the unfused path reads it twice and makes a [V, d] intermediate with K5), as a
//             skinny tensor-core GEMM against the on-the-fly count matrix (see embed_bwd_mma_kernel below). Deterministic:
//             every CTA owns a fixed range of rows, the CTAs' tables are added in ascending order, no atomics.
//             (A first version did one shared-memory read-modify-write per (edge, slot) on per-thread columns of private tables:
//             661 us at BASELINE configs[1] - the dependent LDS/FADD/STS chains of nine warps per SM cannot hide their own
//             latency. Measured, replaced.)
#include <stdlib.h>

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

// ---- backward on the (legacy, warp-level) tensor-core path -------------------------------------------------------------------
// The table gradient is a skinny GEMM:  gT[t, c] = sum_r Cnt[t, r] * g[r, c],  Cnt[t, r] = how many slots of row r hold type t
// (a small-integer matrix, exact in TF32). One CTA per SM walks a contiguous range of 64-row blocks:
//   * two BUILDER warps turn the ids of block b + 1 into Cnt (one thread per row: it zeroes its own column of the [types][64]
//     count tile and bumps one cell per slot - no two threads share a cell, so no atomics);
//   * twelve MMA warps own 32 adjacent columns each (four interleaved 8-column tiles): per 8-row k-step they load the g fragment
//     straight from global memory (two 16-byte loads per lane, two k-steps ahead in registers), split it into TF32 hi + lo (two
//     mma.sync.m16n8k8 per tile: the product is exact to 2^-22) and accumulate all `MT` 16-type tiles in registers;
//   * one barrier per block hands the count tile over. Per-CTA tables, then a fixed-order sum (same kernel as above).
// mma.sync is the warp-level tensor-core path (SASS HMMA): right for a 58 x 300 output whose M is far below a tcgen05 tile; the
// kernel is bound by streaming g once.
constexpr int EFM_MMA_WARPS = 12, EFM_BUILD_WARPS = 2, EFM_THREADS = (EFM_MMA_WARPS + EFM_BUILD_WARPS) * 32;
constexpr int EFM_ROWS = 64, EFM_LD = 68;  // rows per block; padded row length of the count tile (conflict-free fragment loads)
constexpr int EFM_NT = 4;                   // 8-column MMA tiles per warp = 32 adjacent columns -> 384 columns per pass

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Column mapping. A warp owns 32 ADJACENT columns c0 .. c0 + 31 and runs four 8-column MMA tiles over them, interleaved: column
// n of tile i is physical column c0 + 4 n + i. The B fragment of a lane (n = lane / 4) for all four tiles is then ONE aligned
// 16-byte load per row (columns c0 + 4 n .. + 3) instead of four scalar loads, and the eight lanes of a row read 128 contiguous bytes.
template <int MT>
__global__ void __launch_bounds__(EFM_THREADS, 1)
embed_bwd_mma_kernel(const float* __restrict__ g, const int64_t* __restrict__ node_types, int bv, const int64_t* __restrict__ edge_types, int be,
                     const int32_t* __restrict__ src, int64_t n_rows, int64_t V, int Tv, int Te, int d, int col0, int cols, int64_t blocks_per_cta,
                     int flush_blocks, float* __restrict__ partial) {
  extern __shared__ __align__(16) float cnt_smem[];  // [2][MT * 16][EFM_LD]
  constexpr int TYPES = MT * 16, NT = EFM_NT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const int64_t nblocks = (n_rows + EFM_ROWS - 1) / EFM_ROWS;
  const int64_t b0 = (int64_t)blockIdx.x * blocks_per_cta;
  int64_t b1 = b0 + blocks_per_cta;
  if (b1 > nblocks) b1 = nblocks;
  const bool builder = warp >= EFM_MMA_WARPS;

  auto build = [&](int64_t blk, int buf) {  // builder threads only: thread i owns row i of the block (= column i of the count tile)
    const int i = tid - EFM_MMA_WARPS * 32;
    float* col = cnt_smem + (size_t)buf * TYPES * EFM_LD + i;
#pragma unroll 4
    for (int t = 0; t < TYPES; ++t) col[t * EFM_LD] = 0.f;
    const int64_t r = blk * EFM_ROWS + i;
    if (r < n_rows) {
      int64_t s = src ? (int64_t)__ldg(src + r) : r;
      if (s < 0 || s >= V) s = 0;
      for (int j = 0; j < bv; ++j) {
        int64_t k = __ldg(node_types + s * bv + j);
        if (k < 0 || k >= Tv) k = 0;  // reported by the forward pass (status flag); stay in bounds here
        col[(int)k * EFM_LD] += 1.f;
      }
      for (int j = 0; j < be; ++j) {
        int64_t k = __ldg(edge_types + r * be + j);
        if (k < 0 || k >= Te) k = 0;
        col[(Tv + (int)k) * EFM_LD] += 1.f;
      }
    }
  };

  float acc[NT][MT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[i][m][x] = 0.f;

  const int c_lane = warp * 32 + 4 * grp;            // first of this lane's four columns (window-relative); cols % 4 == 0
  const bool col_ok = c_lane < cols;
  const bool warp_on = warp * 32 < cols;             // warp-uniform: any column of this warp inside the window
  struct Frag { float4 lo, hi; };                    // rows tig and tig + 4 of a k-step, columns c_lane .. + 3
  auto load_b = [&](int64_t kstep_global) {          // k-step index counted from this CTA's first row
    Frag f;
    f.lo = f.hi = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t r = b0 * EFM_ROWS + kstep_global * 8 + tig;
    if (col_ok && r < b1 * EFM_ROWS) {
      if (r < n_rows) f.lo = ldg4_stream(g + r * d + col0 + c_lane);
      if (r + 4 < n_rows) f.hi = ldg4_stream(g + (r + 4) * d + col0 + c_lane);
    }
    return f;
  };

  auto flush = [&](int64_t plane) {
    // accumulator fragment of tile i: c0/c1 -> (type 16 m + grp, tile columns 2 tig, 2 tig + 1), c2/c3 -> type + 8;
    // tile column n is physical column warp * 32 + 4 n + i
    const int T = Tv + Te;
    float* dst = partial + (size_t)plane * T * cols;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int ca = warp * 32 + 4 * (2 * tig) + i, cb = ca + 4;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int t0 = 16 * m + grp;
        if (t0 < T) {
          if (ca < cols) dst[(size_t)t0 * cols + ca] = acc[i][m][0];
          if (cb < cols) dst[(size_t)t0 * cols + cb] = acc[i][m][1];
        }
        if (t0 + 8 < T) {
          if (ca < cols) dst[(size_t)(t0 + 8) * cols + ca] = acc[i][m][2];
          if (cb < cols) dst[(size_t)(t0 + 8) * cols + cb] = acc[i][m][3];
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[i][m][x] = 0.f;
      }
    }
  };

  if (builder && b0 < b1) build(b0, 0);
  __syncthreads();
  Frag f1, f2;  // the fragments of the next two k-steps, in flight
  if (!builder) { f1 = load_b(0); f2 = load_b(1); }
  // The tensor core TRUNCATES when it adds into its fp32 accumulator, so the error of a long accumulation chain grows with its
  // length: every `flush_blocks` blocks the accumulators are written out as one more partial table (70 KB each at d = 300 - cheap)
  // and restarted from zero; the partial tables are then added in fp32 (round to nearest) in fixed order.
  for (int64_t blk = b0; blk < b0 + blocks_per_cta; ++blk) {
    const bool last_of_group = ((blk - b0 + 1) % flush_blocks == 0) || blk + 1 == b0 + blocks_per_cta;
    if (blk >= b1) {  // a CTA with fewer blocks than the others still writes its (zero) tables: the sum reads every plane
      if (!builder && last_of_group) flush(((blk - b0) / flush_blocks) * gridDim.x + blockIdx.x);
      continue;       // uniform over the CTA: no barrier is skipped by part of it
    }
    const int buf = (int)((blk - b0) & 1);
    if (builder) {
      if (blk + 1 < b1) build(blk + 1, buf ^ 1);
    } else if (warp_on) {
      const float* cnt = cnt_smem + (size_t)buf * TYPES * EFM_LD;
#pragma unroll 1
      for (int ks = 0; ks < EFM_ROWS / 8; ++ks) {
        const Frag cur = f1;
        f1 = f2;
        f2 = load_b((blk - b0) * (EFM_ROWS / 8) + ks + 2);
        uint32_t a[MT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float* p = cnt + (size_t)(16 * m + grp) * EFM_LD + ks * 8 + tig;
          a[m][0] = __float_as_uint(p[0]);
          a[m][1] = __float_as_uint(p[8 * EFM_LD]);
          a[m][2] = __float_as_uint(p[4]);
          a[m][3] = __float_as_uint(p[8 * EFM_LD + 4]);
        }
        const float b_lo[NT] = {cur.lo.x, cur.lo.y, cur.lo.z, cur.lo.w}, b_hi[NT] = {cur.hi.x, cur.hi.y, cur.hi.z, cur.hi.w};
        uint32_t h0[NT], h1[NT], l0[NT], l1[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0[i]) : "f"(b_lo[i]));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1[i]) : "f"(b_hi[i]));
          l0[i] = __float_as_uint(b_lo[i] - __uint_as_float(h0[i]));
          l1[i] = __float_as_uint(b_hi[i] - __uint_as_float(h1[i]));
        }
        // all sixteen hi products first, then the sixteen lo products: consecutive MMAs never depend on each other
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_tf32_16x8x8(acc[i][m], a[m], h0[i], h1[i]);
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_tf32_16x8x8(acc[i][m], a[m], l0[i], l1[i]);
      }
    }
    if (!builder && last_of_group) flush(((blk - b0) / flush_blocks) * gridDim.x + blockIdx.x);
    __syncthreads();  // count tile of block blk + 1 complete; the one of block blk free for block blk + 2
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

// ---- backward on the tcgen05 weight-gradient kernel (wgrad_pair.cu) ------------------------------------------------------------
// gT = cnt^T . g is the SAME contraction over edges as the weight gradient gW^T = m^T . g, with the count matrix in the place of the
// messages: cnt [rows, ld] (ld = 64 or 128 >= Tv + Te, fp32, small integers) is written once by the kernel below (52 MB at BASELINE
// configs[1], ~5 % of what the step's other kernels move) and the CTA-pair kernel streams it beside g - HBM-bound, where the
// mma.sync kernel above is bound by the legacy tensor path's issue rate (202 us for a 246 MB read).
__global__ void __launch_bounds__(256) embed_count_kernel(const int64_t* __restrict__ node_types, int bv, const int64_t* __restrict__ edge_types, int be,
                                                          const int32_t* __restrict__ src, int64_t n_rows, int64_t V, int Tv, int Te, int ld_shift,
                                                          float* __restrict__ cnt) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r = t >> (ld_shift - 2);                 // ld / 4 threads per row, four types each
  if (r >= n_rows) return;
  const int q = (int)(t & ((1 << (ld_shift - 2)) - 1));
  int64_t s = src ? (int64_t)__ldg(src + r) : r;
  if (s < 0 || s >= V) s = 0;                            // reported by the forward pass (status flag); stay in bounds here
  // branch-free bumps (an array indexed by k & 3 would live in local memory: the first version of this kernel took 44 us)
  const int t0 = 4 * q;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
  for (int j = 0; j < bv; ++j) {
    int64_t k = __ldg(node_types + s * bv + j);
    if (k < 0 || k >= Tv) k = 0;
    const int kk = (int)k - t0;
    c0 += kk == 0 ? 1.f : 0.f;
    c1 += kk == 1 ? 1.f : 0.f;
    c2 += kk == 2 ? 1.f : 0.f;
    c3 += kk == 3 ? 1.f : 0.f;
  }
  for (int j = 0; j < be; ++j) {
    int64_t k = __ldg(edge_types + r * be + j);
    if (k < 0 || k >= Te) k = 0;
    const int kk = (int)k + Tv - t0;
    c0 += kk == 0 ? 1.f : 0.f;
    c1 += kk == 1 ? 1.f : 0.f;
    c2 += kk == 2 ? 1.f : 0.f;
    c3 += kk == 3 ? 1.f : 0.f;
  }
  stg4(cnt + (r << ld_shift) + 4 * q, make_float4(c0, c1, c2, c3));
}

size_t pair_count_wgrad_workspace_bytes(int64_t E, int64_t d, int64_t ld);
int pair_count_wgrad(const float* g, const float* cnt, int64_t E, int64_t d, int64_t ld, int64_t Tv, int64_t Te, float* g_tab_v, float* g_tab_e,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);

static bool embbwd_use_tc() {  // NOTORCH_B200_EMBBWD=mma keeps the warp-level kernel (A/B timing); read per call
  const char* e = getenv("NOTORCH_B200_EMBBWD");
  return !(e && e[0] == 'm');
}
static int64_t count_ld(int64_t T) { return T <= 64 ? 64 : T <= 128 ? 128 : 0; }
static size_t count_bytes(int64_t n_rows, int64_t ld) { return ((size_t)n_rows * ld * sizeof(float) + 1023) / 1024 * 1024; }
// bytes of the tcgen05 path (count matrix + split-K planes); 0 when it does not apply
static size_t embbwd_tc_workspace_bytes(int64_t n_rows, int64_t T, int64_t d) {
  const int64_t ld = count_ld(T);
  if (!ld || d % 4 != 0 || n_rows <= 0 || n_rows >= INT32_MAX / 2) return 0;
  const size_t planes = pair_count_wgrad_workspace_bytes(n_rows, d, ld);
  return planes ? count_bytes(n_rows, ld) + planes : 0;
}

// ---- launch geometry of the backward (shared by the workspace query and the launchers) ----------------------------------------
struct EfmPlan {
  int mt;      // 16-type tiles (1, 2, 4 or 8)
  int cols;    // columns per pass
  int grid;    // CTAs
  int64_t blocks_per_cta;
  int flush_blocks;  // blocks per accumulation chain (64 rows each)
  int planes;        // partial tables = grid * ceil(blocks_per_cta / flush_blocks)
};

static bool efm_plan(int64_t n_rows, int64_t T, int64_t d, EfmPlan* p) {
  if (d % 4 != 0 || T <= 0 || T > 64 || n_rows <= 0) return false;  // four 16-type tiles of accumulators per warp at most
  p->mt = T <= 16 ? 1 : T <= 32 ? 2 : 4;
  const int pass_cols = EFM_MMA_WARPS * EFM_NT * 8;
  p->cols = (int)(d < pass_cols ? d : pass_cols);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int64_t nblocks = cdiv(n_rows, EFM_ROWS);
  p->grid = (int)(nblocks < sms ? nblocks : sms);
  p->blocks_per_cta = cdiv(nblocks, p->grid);
  p->grid = (int)cdiv(nblocks, p->blocks_per_cta);
  p->flush_blocks = 6;  // 6 blocks x 8 k-steps x (hi, lo) = 96 truncating accumulations per chain
  p->planes = p->grid * (int)cdiv(p->blocks_per_cta, p->flush_blocks);
  return true;
}

template <int MT>
static cudaError_t efm_launch(const EfmPlan& plan, const float* g, const int64_t* node_types, int bv, const int64_t* edge_types, int be, const int32_t* src,
                              int64_t n_rows, int64_t V, int Tv, int Te, int d, int col0, int w, float* partial, cudaStream_t st) {
  const size_t smem = (size_t)2 * MT * 16 * EFM_LD * sizeof(float);
  static PerDeviceOnce once;
  cudaError_t e = once.run([] { return cudaFuncSetAttribute(embed_bwd_mma_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * EFM_LD * 4); });
  if (e != cudaSuccess) return e;
  embed_bwd_mma_kernel<MT><<<plan.grid, EFM_THREADS, smem, st>>>(g, node_types, bv, edge_types, be, src, n_rows, V, Tv, Te, d, col0, w,
                                                                 plan.blocks_per_cta, plan.flush_blocks, partial);
  return cudaSuccess;
}

// shared by nt_embed_edge_init_backward (two id sources, node ids gathered through src) and nt_embedding_bag_backward (one source)
int embed_bwd_mma(const float* g, const int64_t* node_types, int64_t bv, const int64_t* edge_types, int64_t be, const int32_t* src, int64_t n_rows, int64_t V,
                  int64_t Tv, int64_t Te, int64_t d, float* g_tab_v, float* g_tab_e, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t tc_bytes = embbwd_tc_workspace_bytes(n_rows, Tv + Te, d);
  if (tc_bytes && workspace_bytes >= tc_bytes && embbwd_use_tc() && (g_tab_e || Te == 0) && aligned16(g) && aligned16(workspace)) {
    const int64_t ld = count_ld(Tv + Te);
    float* cnt = static_cast<float*>(workspace);
    const int ld_shift = ld == 64 ? 6 : 7;
    embed_count_kernel<<<(unsigned)cdiv(n_rows * (ld / 4), 256), 256, 0, st>>>(node_types, (int)bv, edge_types, (int)be, src, n_rows, V, (int)Tv, (int)Te,
                                                                               ld_shift, cnt);
    const size_t off = count_bytes(n_rows, ld);
    const int rc = pair_count_wgrad(g, cnt, n_rows, d, ld, Tv, Te, g_tab_v, g_tab_e ? g_tab_e : g_tab_v, static_cast<uint8_t*>(workspace) + off,
                                    workspace_bytes - off, st);
    if (rc != NT_ERR_UNSUPPORTED) return rc;
  }
  EfmPlan plan;
  if (!efm_plan(n_rows, Tv + Te, d, &plan)) return NT_ERR_UNSUPPORTED;
  if (workspace_bytes < (size_t)plan.planes * (size_t)(Tv + Te) * plan.cols * sizeof(float)) {
    set_error("embedding backward: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  const int T = (int)(Tv + Te);
  int launches = 0;
  for (int64_t col0 = 0; col0 < d; col0 += plan.cols) {
    const int w = (int)(col0 + plan.cols <= d ? plan.cols : d - col0);
    cudaError_t e;
    float* part = static_cast<float*>(workspace);
    switch (plan.mt) {
      case 1: e = efm_launch<1>(plan, g, node_types, (int)bv, edge_types, (int)be, src, n_rows, V, (int)Tv, (int)Te, (int)d, (int)col0, w, part, st); break;
      case 2: e = efm_launch<2>(plan, g, node_types, (int)bv, edge_types, (int)be, src, n_rows, V, (int)Tv, (int)Te, (int)d, (int)col0, w, part, st); break;
      default: e = efm_launch<4>(plan, g, node_types, (int)bv, edge_types, (int)be, src, n_rows, V, (int)Tv, (int)Te, (int)d, (int)col0, w, part, st); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(embed_bwd_mma_kernel)");
    embed_edge_init_bwd_reduce<<<(unsigned)cdiv((int64_t)T * (w / 4), 256), 256, 0, st>>>(part, plan.planes, (int)Tv, (int)Te, (int)d, (int)col0, w, g_tab_v,
                                                                                         g_tab_e);
    launches += 2;
  }
  NT_LAUNCH_CHECK("embed_bwd_mma", launches);
  return NT_OK;
}

size_t embed_bwd_mma_workspace_bytes(int64_t n_rows, int64_t T, int64_t d) {
  EfmPlan plan;
  const size_t mma = efm_plan(n_rows, T, d, &plan) ? (size_t)plan.planes * (size_t)T * plan.cols * sizeof(float) + 256 : 0;
  const size_t tc = embbwd_tc_workspace_bytes(n_rows, T, d);
  return mma > tc ? mma : tc;
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
  if (E <= 0 || d <= 0) return 0;
  return embed_bwd_mma_workspace_bytes(E, num_node_types + num_edge_types, d);
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
  if (!aligned16(g_table_v) || !aligned16(g_table_e) || nt_embed_edge_init_backward_workspace_bytes(E, num_node_types, num_edge_types, d) == 0) {
    set_error("nt_embed_edge_init_backward: needs d %% 4 == 0, 16-byte aligned tables and at most 64 types in total");
    return NT_ERR_UNSUPPORTED;
  }
  if (!workspace || !aligned16(workspace) || workspace_bytes < nt_embed_edge_init_backward_workspace_bytes(E, num_node_types, num_edge_types, d)) {
    set_error("nt_embed_edge_init_backward: workspace too small (nt_embed_edge_init_backward_workspace_bytes)");
    return NT_ERR_WORKSPACE;
  }
  return embed_bwd_mma(static_cast<const float*>(g), node_types, bag_v, edge_types, bag_e, src, E, V, num_node_types, num_edge_types, d,
                       static_cast<float*>(g_table_v), static_cast<float*>(g_table_e), workspace, workspace_bytes, st);
}
