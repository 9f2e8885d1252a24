// K4b — weight gradient on tcgen05 with BOTH operands streamed by TMA (the fast path).
//
//   gW^T[i, o] = sum_e m[e, i] * g_u[e, o]        (D = A^T-free: A = m, B = g_u, both MN-major TF32)
//
// m [E, d] is the message tensor K2 forward already produced (n[src] - act(h[rev])) and wrote out as a
// side output, g [E, d] the incoming gradient: two dense row-major tensors, so each 16-edge K-block is a
// handful of 2-D TMA boxes (32 features x 16 edges, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B — the one
// shared-memory layout tcgen05 accepts for MN-major TF32). Out-of-range rows / columns are zero-filled by
// the copy engine, so there is no tail code.
//
// Pipeline per CTA (4 stages): TMA warp -> raw fp32 tiles in smem -> 8 transform warps split every value
// into TF32 hi / lo parts in place (same offsets: no layout math, no bank conflicts) -> MMA warp issues
// lo.hi + hi.lo + hi.hi -> after the CTA's last K-block the 4 epilogue warps drain TMEM to a partial
// buffer. Edge ranges are reduced afterwards in a fixed order (deterministic split-K). The bias gradient
// rides along as an all-ones feature row of m when d % 128 != 0.
#include <cuda.h>

#include <mutex>

#include "tc_common.cuh"

namespace nt {
namespace wg2 {

using namespace nt::tc;

constexpr int BLOCK_E = 16;                  // edges per K-block (2 MMA k-steps of 8)
constexpr int STAGES = 4;
constexpr int MAX_N = 320;
constexpr int CHUNK_BYTES = BLOCK_E * 128;   // 32 features x 16 edges
constexpr int A_CHUNKS = TILE_M / 32;        // 4
constexpr int B_CHUNKS_MAX = MAX_N / 32;     // 10
constexpr int PART_BYTES = (A_CHUNKS + B_CHUNKS_MAX) * CHUNK_BYTES;  // 28 KiB (hi or lo)
constexpr int STAGE_BYTES = 2 * PART_BYTES;                          // 56 KiB
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, TMA_WARP = 5, FIRST_X_WARP = 6, NUM_X_WARPS = 8;
constexpr int NUM_X_THREADS = NUM_X_WARPS * 32;
constexpr int THREADS = (FIRST_X_WARP + NUM_X_WARPS) * 32;  // 448
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int PREFETCH_BLOCKS = 12;          // L2 prefetch distance of the TMA warp, in K-blocks
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

struct Geometry {
  int d, m_blocks, n_tiles, n_tile, n_a, n_b, ones_row, splits, ld_partial;
  int64_t kb_total, kb_per_split;
};

static Geometry make_geometry(int64_t E, int d, int sms) {
  Geometry g;
  g.d = d;
  g.m_blocks = (d + TILE_M - 1) / TILE_M;
  g.ones_row = (d % TILE_M != 0) ? 1 : 0;  // a spare padded feature row of m exists: make it all ones -> D[d, :] = bias gradient
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
  float* partial;  // [splits][m_blocks*128 (i)][ld_partial (o)]
  int64_t E;
  Geometry geo;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_m, const __grid_constant__ CUtensorMap tmap_g,
                                                                const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t bar_raw = sbase + OFF_BAR;           // [STAGES] TMA bytes landed
  const uint32_t bar_ready = bar_raw + 8 * STAGES;    // [STAGES] hi/lo split done
  const uint32_t bar_empty = bar_ready + 8 * STAGES;  // [STAGES] MMAs retired
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;
  const uint32_t tmem_slot = bar_tmem_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (3 * STAGES + 1));

  const Geometry& geo = p.geo;
  const int d = geo.d;
  const int units = geo.m_blocks * geo.n_tiles;
  const int unit = blockIdx.x % units;  // CTAs of one edge range are adjacent: the g tiles they share hit in L2
  const int split = blockIdx.x / units;
  const int mb = unit % geo.m_blocks, nt = unit / geo.m_blocks;
  const int i0 = mb * TILE_M, o0 = nt * geo.n_tile;
  const int b_chunks = geo.n_tile / 32;
  const int64_t kb_lo = (int64_t)split * geo.kb_per_split;
  int64_t kb_hi = kb_lo + geo.kb_per_split;
  if (kb_hi > geo.kb_total) kb_hi = geo.kb_total;
  const int64_t nkb = kb_hi > kb_lo ? kb_hi - kb_lo : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_raw + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, NUM_X_WARPS);
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
    const int i = i0 + warp * 32 + lane;
    float* dst = p.partial + ((int64_t)split * geo.m_blocks * TILE_M + i) * geo.ld_partial + o0;
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
      mbar_wait(bar_ready + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st0 = sbase + s * STAGE_BYTES;
        const uint32_t a_hi = mnmajor_desc_lo(st0, CHUNK_BYTES), b_hi = mnmajor_desc_lo(st0 + A_CHUNKS * CHUNK_BYTES, CHUNK_BYTES);
        const uint32_t a_lo = mnmajor_desc_lo(st0 + PART_BYTES, CHUNK_BYTES), b_lo = mnmajor_desc_lo(st0 + PART_BYTES + A_CHUNKS * CHUNK_BYTES, CHUNK_BYTES);
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
  } else if (warp == TMA_WARP) {
    // ===================================== TMA PRODUCER =====================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(A_CHUNKS + b_chunks) * CHUNK_BYTES;
      for (int64_t kb = 0; kb < nkb; ++kb) {
        const int e0 = (int)((kb_lo + kb) * BLOCK_E);
        if (kb + PREFETCH_BLOCKS < nkb) {  // warm L2 well ahead of the smem ring
          const int ep = e0 + PREFETCH_BLOCKS * BLOCK_E;
          for (int c = 0; c < A_CHUNKS; ++c)
            if (i0 + 32 * c < d) tma_prefetch_2d(&tmap_m, i0 + 32 * c, ep);
          for (int c = 0; c < b_chunks; ++c)
            if (o0 + 32 * c < d) tma_prefetch_2d(&tmap_g, o0 + 32 * c, ep);
        }
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t st = sbase + s * STAGE_BYTES;
        mbar_arrive_expect_tx(bar_raw + 8 * s, bytes);
        for (int c = 0; c < A_CHUNKS; ++c) tma_load_2d(st + c * CHUNK_BYTES, &tmap_m, i0 + 32 * c, e0, bar_raw + 8 * s);
        for (int c = 0; c < b_chunks; ++c) tma_load_2d(st + (A_CHUNKS + c) * CHUNK_BYTES, &tmap_g, o0 + 32 * c, e0, bar_raw + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================================== TRANSFORM (hi / lo split in place) =====================================
    const int pt = threadIdx.x - FIRST_X_WARP * 32;  // 0..255
    const int units16 = (A_CHUNKS + b_chunks) * (CHUNK_BYTES / 16);
    // the all-ones feature row of m (bias gradient) lives in this CTA's A tile iff i0 <= d < i0 + 128
    const bool ones_here = geo.ones_row && d >= i0 && d < i0 + TILE_M;
    const int ones_chunk = ones_here ? (d - i0) / 32 : -1, ones_c16 = ones_here ? ((d - i0) % 32) / 4 : -1;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      const int64_t e0 = (kb_lo + kb) * BLOCK_E;
      uint8_t* hi = smem + s * STAGE_BYTES;
      uint8_t* lo = hi + PART_BYTES;
      mbar_wait(bar_raw + 8 * s, ph);
      for (int u = pt; u < units16; u += NUM_X_THREADS) {
        float4 v = *reinterpret_cast<const float4*>(hi + u * 16);
        const int chunk = u >> 7, r = (u & 127) >> 3, pos = u & 7;
        if (chunk == ones_chunk || (p.drop_p > 0.f && chunk >= A_CHUNKS)) {
          const int c16 = ((((pos >> 1) ^ (r & 3)) << 1) | (pos & 1));  // undo the 32-byte-unit swizzle
          if (chunk == ones_chunk) {
            if (c16 == ones_c16 && e0 + r < p.E) v.x = 1.f;
          } else {
            const int o = o0 + (chunk - A_CHUNKS) * 32 + c16 * 4;
            if (e0 + r < p.E && o < d) {
              float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)(e0 + r) * (uint64_t)d + (uint64_t)o, p.drop_thr, p.inv_keep);
              v = make_float4(v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w);
            }
          }
        }
        const float4 h4 = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        const float4 l4 = make_float4(tf32_rna(v.x - h4.x), tf32_rna(v.y - h4.y), tf32_rna(v.z - h4.z), tf32_rna(v.w - h4.w));
        *reinterpret_cast<float4*>(hi + u * 16) = h4;
        *reinterpret_cast<float4*>(lo + u * 16) = l4;
      }
      fence_proxy_async();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(bar_ready + 8 * s);
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

// gW[o,i] = sum_z partial[z][i][o] in ascending z; gb[o] = the all-ones row i == d when present.
__global__ void __launch_bounds__(256) wgrad_tma_reduce(const float* __restrict__ partial, Geometry geo, float* __restrict__ gW, float* __restrict__ gb) {
  const int d = geo.d;
  const int rows = d + (geo.ones_row ? 1 : 0);
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (int64_t)rows * d) return;
  const int i = (int)(t / d), o = (int)(t - (int64_t)i * d);  // o fastest: coalesced reads of the partial planes
  const int64_t plane = (int64_t)geo.m_blocks * TILE_M * geo.ld_partial;
  const float* src = partial + (int64_t)i * geo.ld_partial + o;
  float s = 0.f;
  for (int z = 0; z < geo.splits; ++z) s += __ldg(src + z * plane);
  if (i < d) gW[(int64_t)o * d + i] = s;
  else if (gb) gb[o] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int d) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return NT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)BLOCK_E};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return NT_ERR_CUDA;
  }
  return NT_OK;
}

}  // namespace wg2

size_t tma_wgrad_workspace_bytes(int64_t E, int64_t d) {
  if (d % 4 != 0 || E <= 0) return 0;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wg2::Geometry geo = wg2::make_geometry(E, (int)d, sms);
  return (size_t)geo.splits * geo.m_blocks * tc::TILE_M * geo.ld_partial * sizeof(float) + 1024;
}

// returns NT_ERR_UNSUPPORTED when the bias gradient cannot ride along (d % 128 == 0) and gb is requested
int tma_layer_wgrad(const float* g, const float* m, int64_t E, int64_t d, float drop_p, uint64_t seed, uint64_t offset, float* gW, float* gb,
                    void* workspace, size_t workspace_bytes, int products, cudaStream_t st) {
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wg2::Params p{};
  p.geo = wg2::make_geometry(E, (int)d, sms);
  if (gb && !p.geo.ones_row) return NT_ERR_UNSUPPORTED;
  if (workspace_bytes < tma_wgrad_workspace_bytes(E, d)) {
    set_error("tma_layer_wgrad: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  CUtensorMap map_m, map_g;
  int rc = wg2::make_map(&map_m, m, E, (int)d);
  if (rc) return rc;
  rc = wg2::make_map(&map_g, g, E, (int)d);
  if (rc) return rc;
  p.partial = static_cast<float*>(workspace);
  p.E = E;
  p.products = products;
  p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  double t = (double)drop_p * 4294967296.0;
  p.drop_thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  p.seed = seed;
  p.offset = offset;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(wg2::wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg2::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tma_kernel)");
  const int grid = p.geo.splits * p.geo.m_blocks * p.geo.n_tiles;
  wg2::wgrad_tma_kernel<<<grid, wg2::THREADS, wg2::SMEM_BYTES, st>>>(map_m, map_g, p);
  const int64_t total = (d + (p.geo.ones_row ? 1 : 0)) * d;
  wg2::wgrad_tma_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p.geo, gW, gb);
  NT_LAUNCH_CHECK("tma_layer_wgrad", 2);
  return NT_OK;
}

}  // namespace nt
