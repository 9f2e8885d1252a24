// K4b — weight gradient on tcgen05 with BOTH operands streamed as dense tiles (the fast path).
//
//   gW^T[i, o] = sum_e m[e, i] * g_u[e, o]        (D = A^T-free: A = m, B = g_u, both MN-major TF32)
//
// m [E, d] is the message tensor K2 forward already produced (n[src] - act(h[rev])) and wrote out as a
// side output, g [E, d] the incoming gradient: two dense row-major tensors, so each 16-edge K-block is
// 14 chunks of 32 features x 16 edges in the SWIZZLE_128B_BASE32B layout (the one shared-memory layout
// tcgen05 accepts for MN-major TF32).
//
// Pipeline per CTA (4 stages): the eight producer warps copy "their" 16-byte units of the next K-blocks
// straight from global memory into that layout with cp.async (LDGSTS, zero-filled outside [E, d]) and,
// once a K-block has landed (cp.async.wait_group: a thread only ever touches its own units, so no barrier
// is needed), split every value into TF32 hi / lo parts in place -> MMA warp issues lo.hi + hi.lo + hi.hi
// -> after the CTA's last K-block the 4 epilogue warps drain TMEM to a partial buffer. Edge ranges are
// reduced afterwards in a fixed order (deterministic split-K). The bias gradient rides along as an
// all-ones feature row of m when d % 128 != 0.
//
// The first version streamed the tiles with cp.async.bulk.tensor (TMA): the copy engine sustains only
// ~6.5 clk per 128-byte box row (scripts/probes/probe_tma_bw.cu: 19.6 B/clk/SM), i.e. ~1450 clk per
// K-block of 224 rows against ~960 clk of MMA work, so the kernel was copy-engine bound (312 us).
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
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, FIRST_X_WARP = 6, NUM_X_WARPS = 8;  // warp 5 is idle (it used to drive the copy engine)
constexpr int NUM_X_THREADS = NUM_X_WARPS * 32;
constexpr int THREADS = (FIRST_X_WARP + NUM_X_WARPS) * 32;  // 448
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
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
  const float* m;  // [E, d]
  const float* g;  // [E, d]
  float* partial;  // [splits][m_blocks*128 (i)][ld_partial (o)]
  int64_t E;
  Geometry geo;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;
};

template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tma_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t bar_ready = sbase + OFF_BAR;         // [STAGES] hi/lo split done
  const uint32_t bar_empty = bar_ready + 8 * STAGES;  // [STAGES] MMAs retired
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;
  const uint32_t tmem_slot = bar_tmem_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * STAGES + 1));

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
      mbar_wait_relaxed(bar_tmem_full, 0);
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
  } else if (warp >= FIRST_X_WARP) {
    // ===================================== PRODUCER: cp.async tiles, then hi / lo split in place =====================================
    // Thread pt owns the 16-byte units u = pt + 256 k of every stage: chunk u >> 7 (32 features), edge row (u & 127) >> 3,
    // physical slot u & 7 (the 32-byte-unit XOR swizzle undone to find the feature). It copies those units itself and later
    // rewrites the same bytes, so cp.async.wait_group is the only synchronisation the raw data needs.
    const int pt = threadIdx.x - FIRST_X_WARP * 32;  // 0..255
    constexpr int MAX_UNITS = (A_CHUNKS + B_CHUNKS_MAX) * (CHUNK_BYTES / 16) / NUM_X_THREADS;  // 7
    const int units16 = (A_CHUNKS + b_chunks) * (CHUNK_BYTES / 16);
    const int r = (pt & 127) >> 3, pos = pt & 7;                       // identical for all of this thread's units (256 = 2 chunks apart)
    const int c16 = ((((pos >> 1) ^ (r & 3)) << 1) | (pos & 1));       // logical 16-byte chunk inside the 128-byte feature row
    const int E_i = (int)p.E;
    // the all-ones feature row of m (bias gradient) lives in this CTA's A tile iff i0 <= d < i0 + 128
    const bool ones_here = geo.ones_row && d >= i0 && d < i0 + TILE_M;
    const int ones_chunk = ones_here ? (d - i0) / 32 : -1, ones_c16 = ones_here ? ((d - i0) % 32) / 4 : -1;

    auto feature_of = [&](int chunk) { return chunk < A_CHUNKS ? i0 + 32 * chunk + 4 * c16 : o0 + 32 * (chunk - A_CHUNKS) + 4 * c16; };
    auto issue = [&](int64_t kb, int s) {
      const int e = (int)((kb_lo + kb) * BLOCK_E) + r;
      const uint32_t dst = sbase + s * STAGE_BYTES + pt * 16;
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k) {
        const int u = pt + k * NUM_X_THREADS;
        if (u < units16) {
          const int chunk = u >> 7;
          const int f = feature_of(chunk);
          const bool ok = e < E_i && f < d;
          const float* src = (chunk < A_CHUNKS ? p.m : p.g) + (ok ? (int64_t)e * d + f : 0);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + k * (NUM_X_THREADS * 16)), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
      }
    };
    int ls = 0, cs = 0;
    uint32_t lph = 0, cph = 0;
    int64_t lkb = 0;
    auto load_step = [&]() {
      if (lkb < nkb) {
        mbar_wait(bar_empty + 8 * ls, lph ^ 1);
        issue(lkb, ls);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      ++lkb;
      if (++ls == STAGES) { ls = 0; lph ^= 1; }
    };
#pragma unroll 1
    for (int j = 0; j < STAGES - 1; ++j) load_step();
#pragma unroll 1
    for (int64_t kb = 0; kb < nkb; ++kb) {
      const int e = (int)((kb_lo + kb) * BLOCK_E) + r;
      uint8_t* hi = smem + cs * STAGE_BYTES;
      uint8_t* lo = hi + PART_BYTES;
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
      float4 v[MAX_UNITS];
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k) {  // all reads first: the in-place stores below must not serialise the units
        const int u = pt + k * NUM_X_THREADS;
        if (u < units16) v[k] = *reinterpret_cast<const float4*>(hi + u * 16);
      }
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k) {
        const int u = pt + k * NUM_X_THREADS;
        if (u < units16) {
          const int chunk = u >> 7;
          if (chunk == ones_chunk) {
            if (c16 == ones_c16 && e < E_i) v[k].x = 1.f;
          } else if (DROP && chunk >= A_CHUNKS) {
            const int o = feature_of(chunk);
            if (e < E_i && o < d) {
              float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)o, p.drop_thr, p.inv_keep);
              v[k] = make_float4(v[k].x * sc.x, v[k].y * sc.y, v[k].z * sc.z, v[k].w * sc.w);
            }
          }
          const float4 h4 = make_float4(tf32_rna(v[k].x), tf32_rna(v[k].y), tf32_rna(v[k].z), tf32_rna(v[k].w));
          const float4 l4 = make_float4(tf32_rna(v[k].x - h4.x), tf32_rna(v[k].y - h4.y), tf32_rna(v[k].z - h4.z), tf32_rna(v[k].w - h4.w));
          *reinterpret_cast<float4*>(hi + u * 16) = h4;
          *reinterpret_cast<float4*>(lo + u * 16) = l4;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(bar_ready + 8 * cs);
      if (++cs == STAGES) { cs = 0; cph ^= 1; }
      load_step();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    (void)cph;
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
  p.m = m;
  p.g = g;
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
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wg2::wgrad_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg2::SMEM_BYTES);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(wg2::wgrad_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg2::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_tma_kernel)");
  const int grid = p.geo.splits * p.geo.m_blocks * p.geo.n_tiles;
  if (drop_p > 0.f) wg2::wgrad_tma_kernel<true><<<grid, wg2::THREADS, wg2::SMEM_BYTES, st>>>(p);
  else wg2::wgrad_tma_kernel<false><<<grid, wg2::THREADS, wg2::SMEM_BYTES, st>>>(p);
  const int64_t total = (d + (p.geo.ones_row ? 1 : 0)) * d;
  wg2::wgrad_tma_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p.geo, gW, gb);
  NT_LAUNCH_CHECK("tma_layer_wgrad", 2);
  return NT_OK;
}

}  // namespace nt
