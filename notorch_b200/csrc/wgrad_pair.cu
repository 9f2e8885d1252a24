// K4b on a CTA pair (tcgen05 cta_group::2): the weight gradient with M = 256 feature rows per MMA.
//
//   gW^T[i, o] = sum_e m[e, i] * g_u[e, o]        (A = m, B = g_u, both MN-major TF32; see wgrad_tma.cu)
//
// Why a pair: the single-CTA kernel (wgrad_tma.cu) is bound by the per-K-block producer chain (wait for the stage, issue the
// copies, wait for them, hi / lo split, proxy fence, arrive: ~2400 clk) against ~960 clk of MMA work per 16-edge K-block
// (ncu: tensor pipe 43 % active at d = 300, 32 % at d = 1024). A pair shares the B operand: each CTA stages its own 128
// feature rows of m plus HALF of the g columns, i.e. 9 instead of 14 chunks per 16 edges at d = 300 (8 instead of 12 at
// d = 1024), which lets a K-block cover 32 edges in the same shared memory: twice the MMA work (M = 256) per handshake.
//
// Roles per CTA (448 threads): warps 0-3 epilogue (drain the accumulator of every segment, see "accumulator windows" below),
// warp 4 MMA issuer (leader CTA only; both CTAs allocate tensor memory), warp 5 idle, warps 6-13 producers (cp.async copies of "their" 16-byte units, STAGES - 1
// K-blocks ahead, then the TF32 hi / lo split in place - a thread only touches its own units, so the raw data needs no
// barrier). Barriers: ready[s] lives in the leader (16 producer warps arrive, the peer's remotely), empty[s] and tmem_full
// are signalled in both CTAs by tcgen05.commit ... multicast::cluster.
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace nt {
namespace wgp {

using namespace nt::tc;

constexpr int BLOCK_E = 32;                  // edges per K-block (4 MMA k-steps of 8)
constexpr int STAGES = 3;
constexpr int MAX_N = 320;
constexpr int CHUNK_BYTES = BLOCK_E * 128;   // 32 features x 32 edges
constexpr int A_CHUNKS = TILE_M / 32;        // 4: this CTA's 128 feature rows of m
constexpr int B_CHUNKS_MAX = MAX_N / 64;     // 5: this CTA's half of the g columns
constexpr int PART_BYTES = (A_CHUNKS + B_CHUNKS_MAX) * CHUNK_BYTES;  // 36 KiB (hi or lo)
constexpr int STAGE_BYTES = 2 * PART_BYTES;                          // 72 KiB
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, FIRST_X_WARP = 6, NUM_X_WARPS = 8;
constexpr int NUM_X_THREADS = NUM_X_WARPS * 32;
constexpr int THREADS = (FIRST_X_WARP + NUM_X_WARPS) * 32;  // 448
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

struct Geometry {
  // d = width of B (g: the output columns o), da = width of A (m: the feature rows i). The weight gradient has da == d; the
  // embedding-table gradient (embed_count_wgrad below) multiplies a [E, 64 | 128] count matrix with g.
  int d, da, m_units, n_tiles, n_tile, n_a, n_b, ones_row, ld_partial;
  // The last unit of feature rows runs at HALF height (M = 128: 64 rows per CTA) when at most 128 rows of it are real
  // (d = 300: rows 256..300 incl. the all-ones bias row): half the MMA work and half the m columns to stage. Such units get
  // fewer edge splits than the full-height ones, in proportion to their cost per edge.
  int half_last, full_units, half_units, splits, splits_last, planes;
  int64_t kb_total, kb_per_split, kb_per_split_last;
  // Accumulation chains are cut every `seg_kb` K-blocks (32 edges each): the tensor core TRUNCATES when it adds into its fp32
  // accumulator, so the error of a chain grows in proportion to its length (measured at BASELINE configs[1]: 4.6e-5 of the largest
  // weight-gradient entry for chains of 5.5 k edges). At the end of a segment the epilogue adds the accumulator into the split's
  // partial plane in fp32 (round to nearest) and the next segment restarts from zero.
  int seg_kb;
};

static int segment_kblocks() {
  static const int v = [] {
    const char* e = getenv("NOTORCH_B200_WGRAD_SEG");
    const int x = e ? atoi(e) : 16;
    return x > 0 ? x : (1 << 30);
  }();
  return v;
}

// cost of a half-height unit per edge relative to a full one = the ratio of the chunks each CTA stages per K-block (2 + B : 4 + B);
// measured 0.78-0.8 at d = 300 (B = 5 chunks: 7 / 9) - the kernel is bound by its producers, not by the MMAs (1 : 2)
static float half_unit_cost(int b_chunks) {
  static const float w = [] {
    const char* e = getenv("NOTORCH_B200_WGRAD_HALF_COST");
    const float v = e ? (float)atof(e) : 0.f;
    return v > 0.05f && v <= 1.f ? v : 0.f;
  }();
  return w > 0.f ? w : (2.f + (float)b_chunks) / (4.f + (float)b_chunks);
}

static Geometry make_geometry(int64_t E, int d, int sms, int da = -1, bool want_ones = true) {
  Geometry g;
  if (da < 0) da = d;
  g.d = d;
  g.da = da;
  g.m_units = (da + 2 * TILE_M - 1) / (2 * TILE_M);     // pair units of 256 feature rows
  g.ones_row = (want_ones && da % (2 * TILE_M) != 0) ? 1 : 0;  // a spare padded feature row of m exists: all ones -> D[da, :] = bias gradient
  int n_pad = (d + 63) / 64 * 64;                        // each CTA holds half of every MMA's N, in 32-feature chunks
  if (n_pad <= MAX_N) { g.n_tile = n_pad; g.n_tiles = 1; }
  else { g.n_tile = 256; g.n_tiles = (n_pad + 255) / 256; }
  if (g.n_tile <= 256) { g.n_a = g.n_tile; g.n_b = 0; }
  else { g.n_a = 128; g.n_b = g.n_tile - 128; }
  g.ld_partial = g.n_tiles * g.n_tile;
  g.kb_total = (E + BLOCK_E - 1) / BLOCK_E;
  static const bool allow_half = [] { const char* e = getenv("NOTORCH_B200_WGRAD_HALF"); return !(e && e[0] == '0'); }();
  const int last_rows = da + g.ones_row - (g.m_units - 1) * 2 * TILE_M;
  g.half_last = (allow_half && last_rows <= TILE_M) ? 1 : 0;
  g.full_units = (g.m_units - g.half_last) * g.n_tiles;
  g.half_units = g.half_last * g.n_tiles;
  const int clusters = sms / 2 > 0 ? sms / 2 : 1;
  int64_t sf = 0, sl = 0;
  if (g.full_units > 0) {
    sf = (int64_t)((float)clusters / ((float)g.full_units + half_unit_cost(g.n_tile / 64) * (float)g.half_units));
    if (sf < 1) sf = 1;
    if (sf > g.kb_total) sf = g.kb_total > 0 ? g.kb_total : 1;
  }
  if (g.half_units > 0) {
    sl = (clusters - g.full_units * sf) / g.half_units;
    if (sl < 1) sl = 1;
    if (sl > g.kb_total) sl = g.kb_total > 0 ? g.kb_total : 1;
  }
  g.splits = (int)sf;
  g.splits_last = (int)sl;
  g.planes = (int)(sf > sl ? sf : sl);
  g.kb_per_split = sf > 0 ? (g.kb_total + sf - 1) / sf : 0;
  g.kb_per_split_last = sl > 0 ? (g.kb_total + sl - 1) / sl : 0;
  g.seg_kb = segment_kblocks();
  return g;
}

struct Params {
  const float* m;  // [E, da]
  const float* g;  // [E, d]
  float* partial;  // [splits][m_units * 256 (i)][ld_partial (o)]
  int64_t E;
  Geometry geo;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;
  int prefetch;  // K-blocks of L2 look-ahead (0 = off)
  int ablate;  // debug (NOTORCH_B200_WGRAD_ABLATE): 1 = no MMAs, 2 = no hi / lo split, 4 = no loads; results are then meaningless
};

// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both MN-major, M = 256 or 128 over the CTA pair
__device__ __forceinline__ uint32_t make_idesc_pair_mn(int n, int m) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (3u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// plane[row, cols] += accumulator[row, cols] for one thread's row: `nch` chunks of 16 columns; the plane stores a row's 4-column
// groups `cg_stride` floats apart (layout [column / 4][row][4]). The plane lives in L2 (~700 clk away): four statically named
// register sets hold the values of chunks cc .. cc + 3 and each is refilled with chunk cc + 4 as soon as it has been consumed (no
// indexing, no moves - a move out of a register with a load in flight would wait for it). Without this every 16-column chunk pays a
// full round trip and a drain costs more than the segment it closes (measured: +124 us per launch at BASELINE configs[1]).
// Out of line on purpose: inlined into the warp-specialised kernel the register allocator spilled the in-flight loads to local
// memory, which serialises them again.
// L2 residency hints: the partial planes (19-28 MB in all) are re-read and re-written every segment while 0.7 GB of operands
// stream past them; ncu showed the planes falling out of L2 between drains (DRAM traffic 721 -> 903 MB with the drains).
// evict_last on the planes (and evict_first on the streamed operands, see the producers) keeps them resident.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_plane(const float* ptr, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_plane(float* ptr, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

static __device__ __noinline__ void drain_add_rows(uint32_t taddr, int nch, float* p0, uint32_t cg_stride) {
  const uint32_t chunk_stride = 4u * cg_stride;
  const uint64_t keep = l2_policy_evict_last();
  float4 r0[4], r1[4], r2[4], r3[4];
  auto fetch = [&](int cc, float4 (&o)[4]) {
    const float* pc = p0 + (uint32_t)cc * chunk_stride;
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = ld_plane(pc + (uint32_t)q * cg_stride, keep);
  };
  auto consume = [&](int cc, const float4 (&o)[4]) {
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)(cc * 16), v);
    tmem_ld_wait();
    float* pc = p0 + (uint32_t)cc * chunk_stride;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_plane(pc + (uint32_t)q * cg_stride,
               make_float4(o[q].x + __uint_as_float(v[4 * q]), o[q].y + __uint_as_float(v[4 * q + 1]), o[q].z + __uint_as_float(v[4 * q + 2]),
                           o[q].w + __uint_as_float(v[4 * q + 3])),
               keep);
  };
  if (nch > 0) fetch(0, r0);
  if (nch > 1) fetch(1, r1);
  if (nch > 2) fetch(2, r2);
  if (nch > 3) fetch(3, r3);
#pragma unroll 1
  for (int cc0 = 0; cc0 < nch; cc0 += 4) {
    consume(cc0, r0);
    if (cc0 + 4 < nch) fetch(cc0 + 4, r0);
    if (cc0 + 1 < nch) {
      consume(cc0 + 1, r1);
      if (cc0 + 5 < nch) fetch(cc0 + 5, r1);
    }
    if (cc0 + 2 < nch) {
      consume(cc0 + 2, r2);
      if (cc0 + 6 < nch) fetch(cc0 + 6, r2);
    }
    if (cc0 + 3 < nch) {
      consume(cc0 + 3, r3);
      if (cc0 + 7 < nch) fetch(cc0 + 7, r3);
    }
  }
}

template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1) wgrad_pair_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t bar_ready = sbase + OFF_BAR;         // [STAGES] (leader's copy is the live one)
  const uint32_t bar_empty = bar_ready + 8 * STAGES;  // [STAGES]
  // Accumulator windows. A segment's accumulator is drained while the NEXT segment's MMAs already run into another window of
  // tensor memory (segment parity w): with one window the issuer stood still for the whole drain (~8.8 k clk of every ~31 k clk
  // segment at d = 300: +51 us per launch).
  //   * at most 256 columns per segment (N tile <= 256, or a half-height unit): windows [0, 256) and [256, 512), no wait at all
  //     beyond "the drain of segment s - 2 is over";
  //   * N = 128 + 192 (d = 300): 640 columns do not exist, so the 192-column accumulator of MMA "b" alternates between [0, 192)
  //     and [320, 512) and the 128-column accumulator of MMA "a" stays at [192, 320). The last K-block of a segment issues its
  //     "a" MMAs first and commits them on their own barrier, the first K-block of the next segment issues its "b" MMAs first:
  //     the epilogue drains "a" under those ~2.4 k clk of "b" work and only the rest of that drain is exposed.
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;  // [2], per window
  const uint32_t bar_tmem_empty = bar_tmem_full + 16;     // [2]; leader's copy is the live one: all 8 epilogue warps of the pair arrive
  const uint32_t bar_a_full = bar_tmem_empty + 16;        // the single-buffered "a" accumulator: every segment
  const uint32_t bar_a_free = bar_a_full + 8;             // leader's copy is the live one
  const uint32_t tmem_slot = bar_a_free + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * STAGES + 6));

  const Geometry& geo = p.geo;
  const int d = geo.d, da = geo.da;
  const int pair_id = blockIdx.x >> 1;
  // full-height units first (pairs of one edge range are adjacent: the g tiles they share hit in L2), then the half-height ones
  const int n_full = geo.full_units * geo.splits;
  const bool half = pair_id >= n_full;
  int mu, nt, split;
  if (!half) {
    const int unit = pair_id % geo.full_units, mf = geo.m_units - geo.half_last;
    split = pair_id / geo.full_units;
    mu = unit % mf;
    nt = unit / mf;
  } else {
    const int q = pair_id - n_full;
    split = q / geo.half_units;
    mu = geo.m_units - 1;
    nt = q % geo.half_units;
  }
  const int rows_cta = half ? TILE_M / 2 : TILE_M;          // feature rows of m this CTA stages and accumulates
  const int a_chunks = rows_cta / 32;
  const int i0 = mu * 2 * TILE_M + (int)rank * rows_cta;
  const int o0 = nt * geo.n_tile;
  const int ha = geo.n_a / 2, hb = geo.n_b / 2;             // this CTA's share of the two MMAs' N (multiples of 32)
  const int b_chunks = (ha + hb) / 32;
  const int64_t per_split = half ? geo.kb_per_split_last : geo.kb_per_split;
  const int64_t kb_lo = (int64_t)split * per_split;
  int64_t kb_hi = kb_lo + per_split;
  if (kb_hi > geo.kb_total) kb_hi = geo.kb_total;
  const int64_t nkb = kb_hi > kb_lo ? kb_hi - kb_lo : 0;
  const bool split_ab = !half && geo.n_b > 0;  // "a" single-buffered at [192, 320), "b" alternating (see the barriers below)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_ready + 8 * s, 2 * NUM_X_WARPS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(bar_tmem_full + 8 * w, 1);
      mbar_init(bar_tmem_empty + 8 * w, 2 * NUM_EPI_WARPS);
    }
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_free, 2 * NUM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < NUM_EPI_WARPS) {
    // ===================================== EPILOGUE (once per segment) =====================================
    // Accumulator layout in tensor memory. M = 256: lane = feature row (128 per CTA), column = output column. M = 128 (half-height
    // unit, 64 rows per CTA): lanes 0-63 hold the rows for the first half of each MMA's N, lanes 64-127 the same rows for the
    // second half, so an MMA of width N occupies N / 2 columns.
    // Partial plane layout: [output column / 4][feature row][4] - a warp's 32 rows of one 4-column group are 512 contiguous bytes,
    // so the per-segment read-modify-write below is fully coalesced (a row-major plane would make every lane touch its own line).
    const int row = half ? (warp & 1) * 32 + lane : warp * 32 + lane;
    const int64_t plane_rows = (int64_t)geo.m_units * 2 * TILE_M;
    float* plane = p.partial + (int64_t)split * plane_rows * geo.ld_partial;
    float* prow_base = plane + (((int64_t)o0 >> 2) * plane_rows + (i0 + row)) * 4;       // this thread's row in the unit's first column group
    const uint32_t cg_stride = (uint32_t)plane_rows * 4u;                                 // floats between column groups (< 2^24)
    auto at = [&](int col) { return prow_base + (uint32_t)(col >> 2) * cg_stride; };      // col % 4 == 0
    const int seg = geo.seg_kb;
    const int64_t nseg = nkb > 0 ? (nkb + seg - 1) / seg : 0;
    const uint32_t empty_leader = map_to_cta(bar_tmem_empty, 0);
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint64_t keep = l2_policy_evict_last();
    auto drain_store = [&](uint32_t tcol0, int ncols, int out_col0) {  // first segment: the plane takes the accumulator as is
      for (int cc = 0; cc < ncols / 16; ++cc) {
        uint32_t v[16];
        tmem_ld16(lane_base + tcol0 + (uint32_t)(cc * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_plane(at(out_col0 + cc * 16 + q * 4),
                   make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), keep);
      }
    };
    // later segments: plane += accumulator (drain_add_rows above)
    auto drain_add = [&](uint32_t tcol0, int ncols, int out_col0) { drain_add_rows(lane_base + tcol0, ncols / 16, at(out_col0), cg_stride); };
    auto drain = [&](uint32_t tcol0, int ncols, int out_col0, bool accumulate) {
      if (accumulate) drain_add(tcol0, ncols, out_col0);
      else drain_store(tcol0, ncols, out_col0);
    };
    const uint32_t a_free_leader = map_to_cta(bar_a_free, 0);
    auto arrive_leader = [&](uint32_t local_bar, uint32_t leader_bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(local_bar);
        else mbar_arrive_cluster(leader_bar);
      }
    };
    for (int64_t sg = 0; sg < nseg; ++sg) {
      const uint32_t w = (uint32_t)(sg & 1), wph = (uint32_t)((sg >> 1) & 1);
      if (split_ab) {
        mbar_wait_relaxed(bar_a_full, w);
        tc_fence_after();
        drain(192u, geo.n_a, 0, sg > 0);
        if (sg + 1 < nseg) arrive_leader(bar_a_free, a_free_leader);  // the next segment's "a" MMAs overwrite [192, 320)
        mbar_wait_relaxed(bar_tmem_full + 8 * w, wph);
        tc_fence_after();
        drain(w ? 320u : 0u, geo.n_b, geo.n_a, sg > 0);
      } else {
        mbar_wait_relaxed(bar_tmem_full + 8 * w, wph);
        tc_fence_after();
        const uint32_t wb = w * 256u;
        if (!half) {
          drain(wb, geo.n_tile, 0, sg > 0);
        } else {
          const int hi_half = warp >> 1;
          drain(wb, geo.n_a / 2, hi_half * (geo.n_a / 2), sg > 0);
          if (geo.n_b > 0) drain(wb + (uint32_t)geo.n_a, geo.n_b / 2, geo.n_a + hi_half * (geo.n_b / 2), sg > 0);
        }
      }
      if (sg + 2 < nseg) arrive_leader(bar_tmem_empty + 8 * w, empty_leader + 8 * w);  // segment sg + 2 reuses this window
    }
    if (nkb == 0 && (!half || warp < 2)) {
      for (int c = 0; c < geo.n_tile; c += 4) *reinterpret_cast<float4*>(at(c)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA ISSUER (leader CTA only) =====================================
    if (leader) {
      const int mma_m = half ? TILE_M : 2 * TILE_M;
      const uint32_t idesc_a = make_idesc_pair_mn(geo.n_a, mma_m);
      const uint32_t idesc_b = make_idesc_pair_mn(geo.n_b > 0 ? geo.n_b : 64, mma_m);
      int s = 0;
      uint32_t ph = 0;
      const int seg = geo.seg_kb;
      int kb_in_seg = 0;
      uint32_t seg_idx = 0;
      const uint32_t boff = (uint32_t)(ha / 32) * (CHUNK_BYTES >> 4);  // this CTA's columns of the second MMA follow its ha / 32 chunks of the first
      const bool three = p.products == 3, two = p.products == 2, no_mma = (p.ablate & 1) != 0;
      // the MMAs of K-block stage `s` into accumulator `dacc` ("a": B columns from chunk 0, "b": from chunk ha / 32); one elected lane
      auto issue_part = [&](int s_, uint32_t dacc, uint32_t bsel, uint32_t idesc, int kbs) {
        const uint32_t st0 = sbase + s_ * STAGE_BYTES;
        const uint32_t a_hi = mnmajor_desc_lo(st0, CHUNK_BYTES), b_hi = mnmajor_desc_lo(st0 + A_CHUNKS * CHUNK_BYTES, CHUNK_BYTES) + bsel;
        const uint32_t a_lo = mnmajor_desc_lo(st0 + PART_BYTES, CHUNK_BYTES), b_lo = mnmajor_desc_lo(st0 + PART_BYTES + A_CHUNKS * CHUNK_BYTES, CHUNK_BYTES) + bsel;
#pragma unroll
        for (int j = 0; j < BLOCK_E / 8; ++j) {
          const uint32_t k16 = j * (1024u >> 4);  // next 8 edges = the next two 4-row swizzle atoms of every chunk
          const uint32_t acc = (kbs | j) != 0 ? 1u : 0u;
          if (no_mma) {
          } else if (three) {
            umma2_tf32_lo(dacc, a_lo + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc, acc);
            umma2_tf32_lo(dacc, a_hi + k16, b_lo + k16, MNMAJOR_SW128B32_DESC_HI, idesc, 1u);
            umma2_tf32_lo(dacc, a_hi + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc, 1u);
          } else if (two) {  // A is exact in TF32 (a matrix of small integers): its lo part is zero and is neither computed nor multiplied
            umma2_tf32_lo(dacc, a_hi + k16, b_lo + k16, MNMAJOR_SW128B32_DESC_HI, idesc, acc);
            umma2_tf32_lo(dacc, a_hi + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc, 1u);
          } else {
            umma2_tf32_lo(dacc, a_hi + k16, b_hi + k16, MNMAJOR_SW128B32_DESC_HI, idesc, acc);
          }
        }
      };
#pragma unroll 1
      for (int64_t kb = 0; kb < nkb; ++kb) {
        const uint32_t w = seg_idx & 1u;
        // window w was last used by segment seg_idx - 2: both CTAs' epilogues must have drained it
        if (kb_in_seg == 0 && seg_idx >= 2) mbar_wait(bar_tmem_empty + 8 * w, ((seg_idx >> 1) - 1u) & 1u);
        mbar_wait(bar_ready + 8 * s, ph);
        tc_fence_after();
        const bool seg_end = kb_in_seg + 1 == seg || kb == nkb - 1;
        const uint32_t d0 = tmem_base + (split_ab ? 192u : w * 256u);                                   // MMA "a"
        const uint32_t d1 = tmem_base + (split_ab ? (w ? 320u : 0u) : w * 256u + (uint32_t)geo.n_a);   // MMA "b"
        if (split_ab && kb_in_seg == 0 && seg_idx > 0) {
          // first K-block of a segment: "b" first (its window is free), then wait until the epilogue has drained "a"
          if (elect_one()) issue_part(s, d1, boff, idesc_b, 0);
          __syncwarp();
          mbar_wait(bar_a_free, (seg_idx - 1u) & 1u);
          tc_fence_after();
          if (elect_one()) {
            issue_part(s, d0, 0u, idesc_a, 0);
            if (seg_end) umma2_commit_both(bar_a_full);
            umma2_commit_both(bar_empty + 8 * s);
            if (seg_end) umma2_commit_both(bar_tmem_full + 8 * w);
          }
        } else if (elect_one()) {
          // "a" before "b"; in the last K-block of a segment "a" is committed on its own barrier so that its drain starts under
          // the "b" MMAs
          issue_part(s, d0, 0u, idesc_a, kb_in_seg);
          if (split_ab && seg_end) umma2_commit_both(bar_a_full);
          if (geo.n_b > 0) issue_part(s, d1, boff, idesc_b, kb_in_seg);
          umma2_commit_both(bar_empty + 8 * s);
          if (seg_end) umma2_commit_both(bar_tmem_full + 8 * w);
        }
        __syncwarp();
        if (seg_end) { kb_in_seg = 0; ++seg_idx; } else ++kb_in_seg;
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= FIRST_X_WARP) {
    // ===================================== PRODUCER: cp.async tiles, then hi / lo split in place =====================================
    // Thread pt owns the 16-byte units u = pt + 256 k of every stage: chunk u >> 8 (32 features x 32 edges = 256 units), edge
    // row (u & 255) >> 3, physical slot u & 7 (the 32-byte-unit XOR swizzle undone to find the feature).
    const int pt = threadIdx.x - FIRST_X_WARP * 32;  // 0..255: one unit per chunk
    constexpr int MAX_UNITS = A_CHUNKS + B_CHUNKS_MAX;  // 9
    const uint64_t stream_pol = l2_policy_evict_first();  // m and g are read once per unit: do not let them push the partial planes out of L2
    const int n_chunks = A_CHUNKS + b_chunks;
    const int first_split = p.products == 2 ? A_CHUNKS : 0;  // exact A operand: nothing to split, the raw tile is the hi part
    const int r = pt >> 3, pos = pt & 7;
    const int c16 = ((((pos >> 1) ^ (r & 3)) << 1) | (pos & 1));  // logical 16-byte chunk inside the 128-byte feature row
    const int E_i = (int)p.E;
    const uint32_t ready_leader = map_to_cta(bar_ready, 0);
    // the all-ones feature row of m (bias gradient) lives in this CTA's A tile iff i0 <= d < i0 + 128
    const bool ones_here = geo.ones_row && da >= i0 && da < i0 + rows_cta;
    const int ones_chunk = ones_here ? (da - i0) / 32 : -1, ones_c16 = ones_here ? ((da - i0) % 32) / 4 : -1;

    auto feature_of = [&](int chunk) {  // first feature of this thread's unit in chunk `chunk`
      if (chunk < A_CHUNKS) return i0 + 32 * chunk + 4 * c16;
      const int f = 32 * (chunk - A_CHUNKS);  // offset inside this CTA's share of the g columns
      return f < ha ? o0 + (int)rank * ha + f + 4 * c16 : o0 + geo.n_a + (int)rank * hb + (f - ha) + 4 * c16;
    };
    auto issue = [&](int64_t kb, int s) {
      const int e = (int)((kb_lo + kb) * BLOCK_E) + r;
      const uint32_t dst = sbase + s * STAGE_BYTES + pt * 16;
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k) {
        if (k < n_chunks && (k < a_chunks || k >= A_CHUNKS)) {
          const int f = feature_of(k);
          const int width = k < A_CHUNKS ? da : d;
          const bool ok = e < E_i && f < width;
          const float* src = (k < A_CHUNKS ? p.m : p.g) + (ok ? (int64_t)e * width + f : 0);
          asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst + k * CHUNK_BYTES), "l"(src), "r"(ok ? 16 : 0), "l"(stream_pol)
                       : "memory");
        }
      }
    };
    int ls = 0, cs = 0;
    uint32_t lph = 0;
    int64_t lkb = 0;
    // L2 look-ahead (NOTORCH_B200_WGRAD_PF=n, default off). A stage's loads are issued one K-block period before they are needed, so
    // a period cannot be shorter than the memory latency of a K-block's rows; the 32 edge rows of a K-block are ONE contiguous span
    // of each operand, so one thread per CTA can ask the copy engine to bring the span of K-block lkb + n into L2 ahead of time
    // (the leader the A operand's, its peer the B operand's). Measured at BASELINE configs[1]: 284 us without, 306 / 308 / 314 / 316 us
    // with n = 2 / 4 / 6 / 10 (d = 2048: 1134 -> 1330 us) - the prefetches compete with the demand loads for the same DRAM queues
    // and the kernel is not latency-bound in that sense. Kept as a switch for the record, not used.
    const int pf = p.prefetch;
    auto prefetch_span = [&](int64_t kb) {
      const int64_t e0 = (kb_lo + kb) * BLOCK_E;
      if (kb >= nkb || e0 >= p.E) return;
      const int64_t rows = p.E - e0 < BLOCK_E ? p.E - e0 : BLOCK_E;
      const int width = leader ? da : d;
      const float* base = (leader ? p.m : p.g) + e0 * width;
      l2_prefetch_bulk(base, (uint32_t)(rows * width * 4));
    };
    if (pf > 0 && pt == 0)
      for (int j = STAGES - 1; j < pf; ++j) prefetch_span(j);
    auto load_step = [&]() {
      if (lkb < nkb) {
        if (pf > 0 && pt == 0) prefetch_span(lkb + pf);
        mbar_wait(bar_empty + 8 * ls, lph ^ 1);
        if (!(p.ablate & 4)) issue(lkb, ls);
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
      uint8_t* hi = smem + cs * STAGE_BYTES + pt * 16;
      uint8_t* lo = hi + PART_BYTES;
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
      float4 v[MAX_UNITS];
      if (!(p.ablate & 2)) {
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k)  // all reads first: the in-place stores below must not serialise the units
        if (k >= first_split && k < n_chunks && (k < a_chunks || k >= A_CHUNKS)) v[k] = *reinterpret_cast<const float4*>(hi + k * CHUNK_BYTES);
#pragma unroll
      for (int k = 0; k < MAX_UNITS; ++k) {
        if (k >= first_split && k < n_chunks && (k < a_chunks || k >= A_CHUNKS)) {
          if (k == ones_chunk) {
            if (c16 == ones_c16 && e < E_i) v[k].x = 1.f;
          } else if (DROP && k >= A_CHUNKS) {
            const int o = feature_of(k);
            if (e < E_i && o < d) {
              float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)o, p.drop_thr, p.inv_keep);
              v[k] = make_float4(v[k].x * sc.x, v[k].y * sc.y, v[k].z * sc.z, v[k].w * sc.w);
            }
          }
          // hi / lo split without conversions (tc_common.cuh, SPLIT_NOCVT): the fp32 data stays in place as the hi operand and only
          // lo = v - trunc(v) is computed and stored. Units that were changed in registers (dropout scale, the all-ones row) are
          // written back.
          const bool rewritten = (k == ones_chunk) || (DROP && k >= A_CHUNKS);
          float4 h4, l4;
          tf32_split4<SPLIT_NOCVT>(v[k], h4, l4);
          if (rewritten) *reinterpret_cast<float4*>(hi + k * CHUNK_BYTES) = h4;
          *reinterpret_cast<float4*>(lo + k * CHUNK_BYTES) = l4;
        }
      }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_ready + 8 * cs);
        else mbar_arrive_cluster(ready_leader + 8 * cs);
      }
      if (++cs == STAGES) cs = 0;
      load_step();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may retire while its partner can still read its shared memory or signal its barriers
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// sum of one float4 position over `nz` planes in ascending plane order; eight loads are in flight at a time (the planes are tens of MB
// apart: one load per addition made these reductions 15-25 us of pure latency)
__device__ __forceinline__ float4 sum_planes(const float4* __restrict__ src, int64_t stride, int nz) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int z = 0;
  for (; z + 8 <= nz; z += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (z + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) s = make_float4(s.x + v[u].x, s.y + v[u].y, s.z + v[u].z, s.w + v[u].w);
  }
  for (; z < nz; ++z) {
    const float4 v = __ldg(src + z * stride);
    s = make_float4(s.x + v.x, s.y + v.y, s.z + v.z, s.w + v.w);
  }
  return s;
}

// gW[o,i] = sum_z partial[z](i, o) in ascending z; gb[o] = the all-ones row i == d when present. One thread per (4 output columns,
// feature row): a coalesced float4 read of every plane (layout [o / 4][i][4]) and four coalesced stores along i.
__global__ void __launch_bounds__(256) wgrad_pair_reduce(const float* __restrict__ partial, Geometry geo, float* __restrict__ gW, float* __restrict__ gb) {
  const int d = geo.d;
  const int rows = d + (geo.ones_row ? 1 : 0);
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int o4 = (int)(t / rows), i = (int)(t - (int64_t)o4 * rows);  // i fastest
  if (4 * o4 >= d) return;
  const int64_t plane_rows = (int64_t)geo.m_units * 2 * TILE_M;
  const int64_t plane = plane_rows * geo.ld_partial;
  const float4* src = reinterpret_cast<const float4*>(partial) + ((int64_t)o4 * plane_rows + i);
  const int nz = (geo.half_last && i >= (geo.m_units - 1) * 2 * TILE_M) ? geo.splits_last : geo.splits;
  const float4 s = sum_planes(src, plane / 4, nz);
  const float vals[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int o = 4 * o4 + q;
    if (o >= d) break;
    if (i < d) gW[(int64_t)o * d + i] = vals[q];
    else if (gb) gb[o] = vals[q];
  }
}

// Embedding-table gradient out of the same planes: gT[i, o] = sum_z partial[z](i, o), rows i < Tv into g_tab_v [Tv, d], the next Te
// rows into g_tab_e [Te, d] (row-major, no transpose). One thread per (4 output columns, table row).
__global__ void __launch_bounds__(256) wgrad_pair_reduce_tables(const float* __restrict__ partial, Geometry geo, int Tv, int Te, float* __restrict__ g_tab_v,
                                                                float* __restrict__ g_tab_e) {
  const int d = geo.d, rows = Tv + Te;
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int o4 = (int)(t / rows), i = (int)(t - (int64_t)o4 * rows);  // i fastest: coalesced plane reads
  if (4 * o4 >= d) return;
  const int64_t plane_rows = (int64_t)geo.m_units * 2 * TILE_M;
  const int64_t plane = plane_rows * geo.ld_partial;
  const float4* src = reinterpret_cast<const float4*>(partial) + ((int64_t)o4 * plane_rows + i);
  const int nz = (geo.half_last && i >= (geo.m_units - 1) * 2 * TILE_M) ? geo.splits_last : geo.splits;
  const float4 s = sum_planes(src, plane / 4, nz);
  float* out = i < Tv ? g_tab_v + (int64_t)i * d : g_tab_e + (int64_t)(i - Tv) * d;
  *reinterpret_cast<float4*>(out + 4 * o4) = s;  // d % 4 == 0
}

}  // namespace wgp

// host-only view of the work decomposition (no CUDA call): lets the CPU test suite check its invariants for any (E, d, #SMs)
void pair_wgrad_geometry(int64_t E, int64_t d, int sms, int64_t out[12]) {
  const wgp::Geometry g = wgp::make_geometry(E, (int)d, sms);
  const int64_t v[12] = {g.m_units, g.n_tiles, g.n_tile, g.n_a, g.n_b, g.half_last, g.full_units, g.half_units, g.splits, g.splits_last,
                         g.kb_per_split, g.kb_per_split_last};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
}

size_t pair_wgrad_workspace_bytes(int64_t E, int64_t d) {
  if (d % 4 != 0 || E <= 0) return 0;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wgp::Geometry geo = wgp::make_geometry(E, (int)d, sms);
  return (size_t)geo.planes * geo.m_units * 2 * tc::TILE_M * geo.ld_partial * sizeof(float) + 1024;
}

static int launch_pair_kernel(wgp::Params& p, cudaStream_t st) {
  static const int ablate = [] { const char* e = getenv("NOTORCH_B200_WGRAD_ABLATE"); return e ? atoi(e) : 0; }();
  p.ablate = ablate;
  const char* pfe = getenv("NOTORCH_B200_WGRAD_PF");  // read per call (A/B timing)
  p.prefetch = pfe ? atoi(pfe) : 0;
  if (p.prefetch < 0 || p.prefetch > 64) p.prefetch = 0;
  static PerDeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t e = cudaFuncSetAttribute(wgp::wgrad_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgp::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgp::wgrad_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgp::SMEM_BYTES);
    return e;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(wgrad_pair_kernel)");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * (p.geo.full_units * p.geo.splits + p.geo.half_units * p.geo.splits_last)));
  cfg.blockDim = dim3(wgp::THREADS);
  cfg.dynamicSmemBytes = wgp::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = p.drop_p > 0.f ? cudaLaunchKernelEx(&cfg, wgp::wgrad_pair_kernel<true>, p) : cudaLaunchKernelEx(&cfg, wgp::wgrad_pair_kernel<false>, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(wgrad_pair_kernel)");
  return NT_OK;
}

// returns NT_ERR_UNSUPPORTED when the bias gradient cannot ride along (d % 256 == 0) and gb is requested
int pair_layer_wgrad(const float* g, const float* m, int64_t E, int64_t d, float drop_p, uint64_t seed, uint64_t offset, float* gW, float* gb,
                     void* workspace, size_t workspace_bytes, int products, cudaStream_t st) {
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  wgp::Params p{};
  p.geo = wgp::make_geometry(E, (int)d, sms);
  if (products != 3) p.geo.seg_kb = 1 << 30;  // single-pass TF32 (tf32 / bf16 modes): the operand rounding (1e-3) dwarfs the chain error - no drains
  if (gb && !p.geo.ones_row) return NT_ERR_UNSUPPORTED;
  if (workspace_bytes < pair_wgrad_workspace_bytes(E, d)) {
    set_error("pair_layer_wgrad: workspace too small");
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
  const int rc = launch_pair_kernel(p, st);
  if (rc) return rc;
  const int64_t total = (d + (p.geo.ones_row ? 1 : 0)) * ((d + 3) / 4);  // one thread per (4 output columns, feature row)
  wgp::wgrad_pair_reduce<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p.geo, gW, gb);
  NT_LAUNCH_CHECK("pair_layer_wgrad", 2);
  return NT_OK;
}

// Embedding-table gradient on the same kernel: gT[t, :] = sum_e cnt[e, t] * g[e, :] with cnt [E, ld] (ld = 64 or 128 >= Tv + Te) the
// matrix of slot counts (embed_fused.cu builds it) in the place of the messages. Counts are small integers, exact in TF32: two
// products (cnt * g_lo + cnt * g_hi). ld <= 128 makes every unit a half-height one, so every CTA pair takes one range of edges and
// g is read once. (Measured the other way round as well - g as the A side, so that its features fill M = 256: 244 us against 177,
// because the kernel's time is set by its K-blocks per pair, 156 against 87, not by its MMAs.)
static wgp::Geometry count_geometry(int64_t E, int64_t d, int64_t ld) {
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  return wgp::make_geometry(E, (int)d, sms, (int)ld, false);
}

size_t pair_count_wgrad_workspace_bytes(int64_t E, int64_t d, int64_t ld) {
  if (d % 4 != 0 || E <= 0 || (ld != 64 && ld != 128)) return 0;
  const wgp::Geometry geo = count_geometry(E, d, ld);
  if (!geo.half_last || geo.m_units != 1) return 0;  // NOTORCH_B200_WGRAD_HALF=0
  return (size_t)geo.planes * geo.m_units * 2 * tc::TILE_M * geo.ld_partial * sizeof(float) + 1024;
}

int pair_count_wgrad(const float* g, const float* cnt, int64_t E, int64_t d, int64_t ld, int64_t Tv, int64_t Te, float* g_tab_v, float* g_tab_e,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (Tv + Te > ld || (ld != 64 && ld != 128) || d % 4 != 0) return NT_ERR_UNSUPPORTED;
  wgp::Params p{};
  p.geo = count_geometry(E, d, ld);
  if (!p.geo.half_last || p.geo.m_units != 1) return NT_ERR_UNSUPPORTED;
  if (workspace_bytes < pair_count_wgrad_workspace_bytes(E, d, ld)) {
    set_error("pair_count_wgrad: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  p.m = cnt;
  p.g = g;
  p.partial = static_cast<float*>(workspace);
  p.E = E;
  p.products = 2;
  p.drop_p = 0.f;
  p.inv_keep = 1.f;
  const int rc = launch_pair_kernel(p, st);
  if (rc) return rc;
  const int64_t total = (Tv + Te) * (d / 4);
  wgp::wgrad_pair_reduce_tables<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(p.partial, p.geo, (int)Tv, (int)Te, g_tab_v, g_tab_e);
  NT_LAUNCH_CHECK("pair_count_wgrad", 2);
  return NT_OK;
}

}  // namespace nt
