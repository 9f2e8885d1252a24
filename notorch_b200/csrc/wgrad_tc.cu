// K4b — weight gradient of one message-passing depth on tcgen05 tensor cores (3xTF32).
//
//   gW[o, i] = sum_e g_u[e, o] * m[e, i]      g_u = dropout-mask(g),  m[e,:] = n[src[e],:] - act(h[rev[e],:])
//   gb[o]    = sum_e g_u[e, o]                 (an extra all-ones column of m when d % 32 != 0)
//
// The reduction runs over EDGES, which is the row index of both operands in HBM ([E, d] row-major):
// both operands are "MN-major" for the MMA (contiguous along M / N, strided along K). tcgen05 takes
// TF32 operands in either majorness, so the producers copy rows as they lie in memory into
// 128-byte-swizzled (32-byte base) MN-major tiles (32 features per 128-byte row, 4 edges per swizzle atom) —
// no transposition. One CTA owns a 128-row block of output features (o), all columns of one N tile,
// and a contiguous range of edges; partial results of the edge ranges go to a workspace and are added in
// a fixed order afterwards (deterministic split-K, no atomics).
//
// Warp roles: 0-3 epilogue (TMEM -> partial sums, once at the end), 4 MMA issuer, 5-12 producers.
#include <mutex>

#include "tc_common.cuh"

namespace nt {
namespace wg {

using namespace nt::tc;

constexpr int BLOCK_E = 32;                 // edges per K-block (4 MMA k-steps of 8)
constexpr int STAGES = 2;
constexpr int MAX_N = 320;                  // widest N tile (multiple of 32)
constexpr int CHUNK_BYTES = BLOCK_E * 128;  // one 32-feature chunk of a K-block: 32 k-rows x 128 B
constexpr int A_PART_BYTES = (TILE_M / 32) * CHUNK_BYTES;  // 16 KiB
constexpr int B_PART_BYTES = (MAX_N / 32) * CHUNK_BYTES;   // 40 KiB
constexpr int STAGE_BYTES = 2 * A_PART_BYTES + 2 * B_PART_BYTES;  // hi + lo of both: 112 KiB
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, FIRST_P_WARP = 5, NUM_P_WARPS = 8;
constexpr int NUM_P_THREADS = NUM_P_WARPS * 32;
constexpr int THREADS = (FIRST_P_WARP + NUM_P_WARPS) * 32;  // 416
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

struct Geometry {
  int d, m_blocks, n_tiles, n_tile, n_a, n_b, ones_col, splits;
  int64_t kb_total, kb_per_split;
  int ld_partial;  // row length of the partial buffers = n_tiles * n_tile
};

static Geometry make_geometry(int64_t E, int d, int sms) {
  Geometry g;
  g.d = d;
  g.m_blocks = (d + TILE_M - 1) / TILE_M;
  g.ones_col = (d % 32 != 0) ? 1 : 0;  // a spare padded column exists: carry the bias gradient in it
  int n_pad = (d + 31) / 32 * 32;
  if (n_pad <= MAX_N) { g.n_tile = n_pad; g.n_tiles = 1; }
  else { g.n_tile = 256; g.n_tiles = (n_pad + 255) / 256; }
  if (g.n_tile <= 256) { g.n_a = g.n_tile; g.n_b = 0; }
  else { g.n_a = 160; g.n_b = g.n_tile - 160; }
  g.ld_partial = g.n_tiles * g.n_tile;
  g.kb_total = (E + BLOCK_E - 1) / BLOCK_E;
  int64_t units = (int64_t)g.m_blocks * g.n_tiles;
  int64_t s = sms / units;
  if (s < 1) s = 1;
  if (s > g.kb_total) s = g.kb_total > 0 ? g.kb_total : 1;
  g.splits = (int)s;
  g.kb_per_split = (g.kb_total + s - 1) / s;
  return g;
}

struct Params {
  const float* g;
  const float* h;
  const float* n;
  const int32_t* src;
  const int32_t* rev;
  float* partial;  // [splits][m_blocks*128][ld_partial]
  int64_t E;
  Geometry geo;
  int act;
  float act_param;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;
};

__device__ __forceinline__ uint32_t mn_swz(uint32_t r, uint32_t c16) { return mn_swz32(r, c16); }

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 m) {
  float4 hi = make_float4(tf32_rna(m.x), tf32_rna(m.y), tf32_rna(m.z), tf32_rna(m.w));
  float4 lo = make_float4(tf32_rna(m.x - hi.x), tf32_rna(m.y - hi.y), tf32_rna(m.z - hi.z), tf32_rna(m.w - hi.w));
  *reinterpret_cast<float4*>(hi_base + off) = hi;
  *reinterpret_cast<float4*>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t bar_full = sbase + OFF_BAR;          // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * STAGES;   // [STAGES]
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;
  const uint32_t tmem_slot = bar_tmem_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * STAGES + 1));

  const Geometry& geo = p.geo;
  const int d = geo.d;
  const int units = geo.m_blocks * geo.n_tiles;
  const int unit = blockIdx.x % units;   // CTAs of one edge range are adjacent: they share the gathered rows in L2
  const int split = blockIdx.x / units;
  const int mb = unit % geo.m_blocks, nt = unit / geo.m_blocks;
  const int o0 = mb * TILE_M, i0 = nt * geo.n_tile;
  const int64_t kb_lo = (int64_t)split * geo.kb_per_split;
  int64_t kb_hi = kb_lo + geo.kb_per_split;
  if (kb_hi > geo.kb_total) kb_hi = geo.kb_total;
  const int64_t nkb = kb_hi > kb_lo ? kb_hi - kb_lo : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, NUM_P_WARPS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < NUM_EPI_WARPS) {
    // ===================================== EPILOGUE (once) =====================================
    const int o = o0 + warp * 32 + lane;
    float* dst = p.partial + ((int64_t)split * geo.m_blocks * TILE_M + o) * geo.ld_partial + i0;
    if (nkb > 0) {
      mbar_wait(bar_tmem_full, 0);
      tc_fence_after();
      for (int cc = 0; cc < geo.n_tile / 16; ++cc) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + cc * 16 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    } else {
      for (int c = 0; c < geo.n_tile; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA ISSUER =====================================
    const uint32_t idesc_a = make_idesc_tf32(geo.n_a, true);
    const uint32_t idesc_b = make_idesc_tf32(geo.n_b > 0 ? geo.n_b : 32, true);
    int s = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st0 = sbase + s * STAGE_BYTES;
        const uint32_t a_hi = mnmajor_desc_lo(st0, CHUNK_BYTES), a_lo = mnmajor_desc_lo(st0 + A_PART_BYTES, CHUNK_BYTES);
        const uint32_t b_hi = mnmajor_desc_lo(st0 + 2 * A_PART_BYTES, CHUNK_BYTES), b_lo = mnmajor_desc_lo(st0 + 2 * A_PART_BYTES + B_PART_BYTES, CHUNK_BYTES);
        const uint32_t d1 = tmem_base + (uint32_t)geo.n_a;
        const uint32_t boff = (uint32_t)(geo.n_a / 32) * (CHUNK_BYTES >> 4);  // second N half: n_a / 32 chunks further
#pragma unroll
        for (int j = 0; j < BLOCK_E / 8; ++j) {
          const uint32_t k16 = j * (1024u >> 4);  // next 8 edges = the next two 4-row swizzle atoms of every chunk
          const uint32_t acc = (kb | j) != 0 ? 1u : 0u;
          if (p.products == 3) {
            umma_tf32_lo(tmem_base, a_lo + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc_a, acc);
            umma_tf32_lo(tmem_base, a_hi + k16, b_lo + k16, MNMAJOR_SW128B32_DESC_HI, idesc_a, 1u);
            umma_tf32_lo(tmem_base, a_hi + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc_a, 1u);
            if (geo.n_b > 0) {
              umma_tf32_lo(d1, a_lo + k16, b_hi + boff + k16, MNMAJOR_SW128B32_DESC_HI, idesc_b, acc);
              umma_tf32_lo(d1, a_hi + k16, b_lo + boff + k16, MNMAJOR_SW128B32_DESC_HI, idesc_b, 1u);
              umma_tf32_lo(d1, a_hi + k16, b_hi + boff + k16, MNMAJOR_SW128B32_DESC_HI, idesc_b, 1u);
            }
          } else {
            umma_tf32_lo(tmem_base, a_hi + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc_a, acc);
            if (geo.n_b > 0) umma_tf32_lo(d1, a_hi + k16, b_hi + boff + k16, MNMAJOR_SW128B32_DESC_HI, idesc_b, acc);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (kb == nkb - 1) umma_commit(bar_tmem_full);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================================== PRODUCERS =====================================
    const int pt = threadIdx.x - FIRST_P_WARP * 32;  // 0..255
    const int nch = geo.n_tile / 4;                   // 16-byte chunks per B row (<= 80)
    const int b_tasks = BLOCK_E * nch;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      const int64_t e0 = (kb_lo + kb) * BLOCK_E;
      uint8_t* st = smem + s * STAGE_BYTES;
      uint8_t *a_hi = st, *a_lo = st + A_PART_BYTES, *b_hi = st + 2 * A_PART_BYTES, *b_lo = b_hi + B_PART_BYTES;

      // ---- A: g_u[e0..e0+32, o0..o0+128): 32 rows x 32 chunks, 4 per thread (loads before the stage wait)
      float4 va[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = pt + NUM_P_THREADS * j, r = q >> 5, c = q & 31;
        const int64_t e = e0 + r;
        const int o = o0 + 4 * c;
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < p.E && o < d) {
          va[j] = ldg4(p.g + e * d + o);
          if (p.drop_p > 0.f) {
            float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)o, p.drop_thr, p.inv_keep);
            va[j] = make_float4(va[j].x * sc.x, va[j].y * sc.y, va[j].z * sc.z, va[j].w * sc.w);
          }
        }
      }
      // ---- B, first batch of gathers also issued before the wait
      constexpr int B_BATCH = 5;
      float4 vn[B_BATCH], vh[B_BATCH];
      auto load_b = [&](int j0) {
#pragma unroll
        for (int jj = 0; jj < B_BATCH; ++jj) {
          const int q = pt + NUM_P_THREADS * (j0 + jj);
          vn[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          vh[jj] = vn[jj];
          if (q < b_tasks) {
            const int r = q / nch, c = q - r * nch;
            const int64_t e = e0 + r;
            const int i = i0 + 4 * c;
            if (e < p.E) {
              if (i < d) {
                vn[jj] = ldg4(p.n + (int64_t)__ldg(p.src + e) * d + i);
                vh[jj] = act_fwd4(ldg4_stream(p.h + (int64_t)__ldg(p.rev + e) * d + i), p.act, p.act_param);
              } else if (geo.ones_col && i == d) {
                vn[jj].x = 1.f;  // column d of m is all ones: D[:, d] = sum_e g_u[e, :] = bias gradient
              }
            }
          }
        }
      };
      auto store_b = [&](int j0) {
#pragma unroll
        for (int jj = 0; jj < B_BATCH; ++jj) {
          const int q = pt + NUM_P_THREADS * (j0 + jj);
          if (q < b_tasks) {
            const int r = q / nch, c = q - r * nch;
            const uint32_t off = (uint32_t)(c >> 3) * CHUNK_BYTES + mn_swz((uint32_t)r, (uint32_t)(c & 7));
            split_store(b_hi, b_lo, off, make_float4(vn[jj].x - vh[jj].x, vn[jj].y - vh[jj].y, vn[jj].z - vh[jj].z, vn[jj].w - vh[jj].w));
          }
        }
      };
      load_b(0);
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = pt + NUM_P_THREADS * j, r = q >> 5, c = q & 31;
        split_store(a_hi, a_lo, (uint32_t)(c >> 3) * CHUNK_BYTES + mn_swz((uint32_t)r, (uint32_t)(c & 7)), va[j]);
      }
      store_b(0);
      for (int j0 = B_BATCH; j0 * NUM_P_THREADS < b_tasks; j0 += B_BATCH) {
        load_b(j0);
        store_b(j0);
      }
      fence_proxy_async();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(bar_full + 8 * s);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// gW[o,i] = sum_s partial[s][o][i] in ascending s (deterministic); gb[o] = the ones column when present
__global__ void __launch_bounds__(256) wgrad_tc_reduce(const float* __restrict__ partial, Geometry geo, float* __restrict__ gW, float* __restrict__ gb) {
  const int d = geo.d;
  const int cols = d + (geo.ones_col ? 1 : 0);
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (int64_t)d * cols) return;
  const int o = (int)(t / cols), i = (int)(t - (int64_t)o * cols);
  const int64_t plane = (int64_t)geo.m_blocks * TILE_M * geo.ld_partial;
  const float* src = partial + (int64_t)o * geo.ld_partial + i;
  float s = 0.f;
  for (int z = 0; z < geo.splits; ++z) s += __ldg(src + z * plane);
  if (i < d) gW[(int64_t)o * d + i] = s;
  else if (gb) gb[o] = s;
}

// column sums of g_u (bias gradient) when no spare ones column exists (d % 32 == 0): two fixed-order stages
constexpr int CS_ROWS = 512;
__global__ void __launch_bounds__(256) colsum_partial(const float* __restrict__ g, int64_t E, int d, float drop_p, float inv_keep, uint32_t thr,
                                                       uint64_t seed, uint64_t offset, float* __restrict__ part) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= d) return;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = r0 + CS_ROWS < E ? r0 + CS_ROWS : E;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    float v = __ldg(g + r * d + c);
    if (drop_p > 0.f) v *= dropout_scale1(seed, offset, (uint64_t)r * (uint64_t)d + (uint64_t)c, thr, inv_keep);
    s += v;
  }
  part[(int64_t)blockIdx.y * d + c] = s;
}
__global__ void __launch_bounds__(256) colsum_final(const float* __restrict__ part, int64_t nblk, int d, float* __restrict__ gb) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  for (int64_t b = 0; b < nblk; ++b) s += __ldg(part + b * d + c);
  gb[c] = s;
}

}  // namespace wg

// gb[o] = sum_e g_u[e, o] — stand-alone, for hidden sizes where no spare ones row/column exists. Uses `workspace` as scratch
// AFTER the caller's kernels (same stream), so it may alias the partial buffers that have already been reduced.
int tc_bias_grad(const float* g, int64_t E, int64_t d, float drop_p, uint64_t seed, uint64_t offset, float* gb, void* workspace, size_t workspace_bytes,
                 cudaStream_t st) {
  const int64_t nblk = cdiv(E, wg::CS_ROWS);
  if ((size_t)nblk * d * sizeof(float) > workspace_bytes) {
    set_error("tc_bias_grad: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  float* part = static_cast<float*>(workspace);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  double t = (double)drop_p * 4294967296.0;
  const uint32_t thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  dim3 grid_cs((unsigned)cdiv(d, 256), (unsigned)nblk);
  wg::colsum_partial<<<grid_cs, 256, 0, st>>>(g, E, (int)d, drop_p, inv_keep, thr, seed, offset, part);
  wg::colsum_final<<<(unsigned)cdiv(d, 256), 256, 0, st>>>(part, nblk, (int)d, gb);
  NT_LAUNCH_CHECK("tc_bias_grad", 2);
  return NT_OK;
}

size_t tc_wgrad_workspace_bytes(int64_t E, int64_t d) {
  if (d % 4 != 0 || E <= 0) return 0;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wg::Geometry geo = wg::make_geometry(E, (int)d, sms);
  size_t partial = (size_t)geo.splits * geo.m_blocks * tc::TILE_M * geo.ld_partial * sizeof(float);
  size_t cs = geo.ones_col ? 0 : (size_t)cdiv(E, wg::CS_ROWS) * d * sizeof(float);
  return partial + cs + 1024;
}

int tc_layer_wgrad(const float* g, const float* h, const float* n, const int32_t* src, const int32_t* rev, int64_t E, int64_t d, int act, float act_param,
                   float drop_p, uint64_t seed, uint64_t offset, float* gW, float* gb, void* workspace, size_t workspace_bytes, int products,
                   cudaStream_t st) {
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wg::Params p{};
  p.geo = wg::make_geometry(E, (int)d, sms);
  if (workspace_bytes < tc_wgrad_workspace_bytes(E, d)) {
    set_error("tc_layer_wgrad: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  p.g = g; p.h = h; p.n = n; p.src = src; p.rev = rev; p.partial = static_cast<float*>(workspace); p.E = E;
  p.act = act; p.act_param = act_param; p.products = products;
  p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  double t = (double)drop_p * 4294967296.0;
  p.drop_thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  p.seed = seed; p.offset = offset;

  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run([] { return cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tc_kernel)");
  const int grid = p.geo.splits * p.geo.m_blocks * p.geo.n_tiles;
  wg::wgrad_tc_kernel<<<grid, wg::THREADS, wg::SMEM_BYTES, st>>>(p);
  const int64_t total = d * (d + (p.geo.ones_col ? 1 : 0));
  wg::wgrad_tc_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p.geo, gW, p.geo.ones_col ? gb : nullptr);
  int launched = 2;
  if (gb && !p.geo.ones_col) {
    const size_t partial_bytes = (size_t)p.geo.splits * p.geo.m_blocks * tc::TILE_M * p.geo.ld_partial * sizeof(float);
    float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + ((partial_bytes + 255) / 256) * 256);
    const int64_t nblk = cdiv(E, wg::CS_ROWS);
    dim3 grid_cs((unsigned)cdiv(d, 256), (unsigned)nblk);
    wg::colsum_partial<<<grid_cs, 256, 0, st>>>(g, E, (int)d, p.drop_p, p.inv_keep, p.drop_thr, seed, offset, part);
    wg::colsum_final<<<(unsigned)cdiv(d, 256), 256, 0, st>>>(part, nblk, (int)d, gb);
    launched += 2;
  }
  NT_LAUNCH_CHECK("tc_layer_wgrad", launched);
  return NT_OK;
}

}  // namespace nt
