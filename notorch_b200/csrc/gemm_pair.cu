// K2 forward, K4a dgrad and the dense Linear of the atom variant, second generation: CTA-pair tcgen05 (cta_group::2), A operand
// gathered with cp.async (LDGSTS), weight tiles on the copy engine (cp.async.bulk).
//
//   D[256 edges, N] = A[256, K] . B[N, K]^T     fp32 in, fp32 out, 3xTF32 error-compensated (see gemm_tc.cu)
//
// One cluster of two CTAs (one TPC) owns a 256-edge tile: each CTA stages ITS 128 rows of A and HALF of the weight
// tile (N/2 rows of W), the leader CTA's elected thread issues tcgen05.mma.cta_group::2 with M = 256, and each CTA's
// tensor memory receives the accumulator rows of its own 128 edges. Per CTA that halves the W bytes pulled from L2
// (the first-generation kernel re-streamed 778 KB of W per 128-edge tile, which alone saturates the L2->SM path at
// the tensor-core rate) and frees enough shared memory for a third pipeline stage.
//
// The A operand is produced by eight warps in two steps:
//   1. each thread copies "its" 16-byte units of the next tiles straight from global memory into the 128-byte-swizzled
//      K-major layout with cp.async (LDGSTS): rows n[src[e]] and h[rev[e]] for K2, rows of g for K4a. No registers are
//      tied up, so STAGES - 1 K-blocks (64 KB per SM) stay in flight. (tile::gather4 on the copy engine was measured
//      first: it sustains only ~17 B/clk/SM for 128-byte rows - scripts/probes/probe_tma_bw.cu - about half of what this
//      kernel needs; explicit L2 prefetches of the next tile's rows were measured too and bought nothing.)
//   2. the same thread rewrites the units IN PLACE once they have landed (cp.async.wait_group — a thread's own copies,
//      no barrier): m = n - act(h) (K2) or dropout(g) (K4a), split into TF32 hi (over the n tile) and lo (over the h
//      tile); K2 also streams m to global memory for the weight gradient. In bf16 mode the tile is rounded to bf16 instead
//      (64-byte rows, one kind::f16 MMA pass).
// W arrives as two bulk copies (hi, lo) of the pre-split, pre-swizzled per-CTA image written by pair_weight_prepare.
//
// Synchronisation (mbarriers; L = lives in the leader CTA and is also arrived on remotely by the peer):
//   w_full[s]     local   copy engine -> A-producer warp 0      (tx bytes of the W copies)
//   ready[s]      L       16 A-producer warps -> MMA issuer     (A converted in both CTAs, W landed in both CTAs)
//   empty[s]      local   tcgen05.commit multicast to both CTAs -> W producer and A producers of each CTA
//   tmem_full     local   tcgen05.commit multicast -> epilogue warps of each CTA
//   tmem_empty    L       8 epilogue warps -> MMA issuer
#include <cuda_bf16.h>
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace nt {
namespace pair {

using namespace nt::tc;

constexpr int BLOCK_K = 32;
constexpr int STAGES = 3;
constexpr int MAX_N_TILE = 304;
constexpr int A_PART_BYTES = TILE_M * 128;                  // 16 KiB: one raw / hi / lo tile of 128 rows x 32 fp32
constexpr int W_PART_BYTES = (MAX_N_TILE / 2) * 128;        // 19 KiB: this CTA's half of the weight tile (hi or lo)
constexpr int STAGE_BYTES = 2 * A_PART_BYTES + 2 * W_PART_BYTES;  // 70 KiB
constexpr int EPI_COLS = 32;
constexpr int EPI_WARP_BYTES = 32 * EPI_COLS * 4;  // 4 KiB staging tile per epilogue warp
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, TMA_WARP = 5, FIRST_X_WARP = 6, NUM_X_WARPS = 8;
constexpr int NUM_X_THREADS = NUM_X_WARPS * 32;
constexpr int THREADS = (FIRST_X_WARP + NUM_X_WARPS) * 32;  // 448

constexpr int OFF_A0 = 0, OFF_A1 = A_PART_BYTES, OFF_WHI = 2 * A_PART_BYTES, OFF_WLO = OFF_WHI + W_PART_BYTES;  // inside a stage
constexpr int OFF_EPI = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_EPI + NUM_EPI_WARPS * EPI_WARP_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
static_assert(STAGE_BYTES % 1024 == 0 && W_PART_BYTES % 1024 == 0, "swizzle-128B tiles need 1 KiB alignment");

struct Geometry {
  int d, n_tile, n_tiles, k_blocks, n_a, n_b, rows_per_cta;
  // Long reductions (d >= 640): TWO tensor-memory windows that the epilogue adds in fp32 - the main product hi x hi goes to the
  // first, the two small cross products (lo x hi, hi x lo) to the second. The tensor core truncates when it adds into its
  // accumulator, so the residue of an accumulation chain grows with the number of MMAs added into it; every MMA into a window
  // truncates the WHOLE running sum, however small its own contribution, so keeping the cross terms out of the main window cuts
  // its chain to a third (d = 1024: 128 instead of 384 truncations; the cross window holds values 2^-11 times smaller, its
  // truncation error is negligible). Measured: a 5-layer d = 1024 block was 1.8e-5 off the fp64 oracle with one accumulator. The
  // first version split even / odd K-blocks instead (chains of one half). Costs the epilogue / MMA overlap.
  int split_acc;
  size_t part_bytes;  // bytes of one (hi or lo) image
};

// smallest number of 32-wide K-blocks for which the main / cross accumulators are used (NOTORCH_B200_SPLIT_MIN_KB, default 20 = d >= 640;
// a large value switches it off for A/B timing: d = 1024 then runs 35 % faster in K2 and a 5-depth block is 1.8e-5 instead of 9.8e-6 off)
static int split_min_kblocks() {
  static const int v = [] {
    const char* e = getenv("NOTORCH_B200_SPLIT_MIN_KB");
    const int x = e ? atoi(e) : 20;
    return x > 0 ? x : 20;
  }();
  return v;
}

__host__ __device__ inline Geometry make_geometry(int d) {
  Geometry g;
  g.d = d;
  int d16 = (d + 15) / 16 * 16;
  g.n_tile = d16 <= MAX_N_TILE ? d16 : 256;
  g.n_tiles = (d + g.n_tile - 1) / g.n_tile;
  g.k_blocks = (d + BLOCK_K - 1) / BLOCK_K;
  if (g.n_tile <= 256) { g.n_a = g.n_tile; g.n_b = 0; }
  else { g.n_a = ((g.n_tile / 2) + 15) / 16 * 16; g.n_b = g.n_tile - g.n_a; }
  g.rows_per_cta = g.n_tile / 2;  // n_a / 2 rows of the first MMA followed by n_b / 2 rows of the second
  g.split_acc = 0;  // decided on the host (launch<>): needs getenv
  g.part_bytes = (size_t)g.n_tiles * g.k_blocks * g.n_tile * 128;
  return g;
}

struct Params {
  const float* a0;  // K2: n [V,d]      K4a: g [E,d]
  const float* a1;  // K2: h [E,d]      K4a: unused
  const int32_t* src;
  const int32_t* rev;
  const uint8_t* wimg;  // hi image followed by lo image (pair layout)
  const float* bias;
  const float* resid;   // K2 residual input h (nullable)
  float* out;
  float* m_out;         // K2: optional copy of the message tensor m [E,d]
  int64_t E;
  Geometry geo;
  int act;
  float act_param;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;
  unsigned long long* trace;
  int ablate;  // timing experiments only (NOTORCH_B200_ABLATE): 1 epilogue without global traffic, 2 no cp.async, 4 no W copies, 8 no MMAs, 16 no transform, 32 no m_out
};

// kind::tf32 instruction descriptor for the pair: D = F32, A = B = TF32, both K-major, M = 256 (128 rows per CTA)
__host__ __device__ __forceinline__ uint32_t make_idesc_pair(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// The kernel is specialised at compile time on (dropout on/off, ReLU vs. the generic activation switch): with the six-way
// switch (expf / erff / tanhf per element) and Philox inlined into every unrolled unit the hot loops were ~75 KB of SASS and
// ncu showed the instruction cache as a bottleneck (gcc__cache_requests_type_instruction at 73 % of peak, 12 % of warp
// samples stalled on no_instructions).
// ---- bf16 operand mode (BASELINE configs[4]: "bf16 W_h with fp32 accumulation") -----------------------------------------
// Operands are rounded to bf16 and multiplied by ONE tcgen05.mma.kind::f16 pass (fp32 accumulation in TMEM); a K-block of
// 32 values is a 64-byte row, i.e. the 64-byte-swizzled K-major layout (Swizzle<2,4,3>: 16-byte chunk ^= (row >> 1) & 3).
__host__ __device__ __forceinline__ uint32_t make_idesc_pair_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
constexpr uint32_t KMAJOR_SW64_DESC_HI = (512u >> 4) | (1u << 14) | (4u << 29);  // SBO = 512 B, version 1, SWIZZLE_64B
// byte offset of the four bf16 values of (row r, 4-value chunk c in [0, 8)) inside a 64-byte-swizzled K-major tile
__host__ __device__ __forceinline__ uint32_t swz64_bf16(uint32_t r, uint32_t c) {
  return (r >> 3) * 512u + (r & 7u) * 64u + ((((c >> 1) ^ ((r >> 1) & 3u))) << 4) + (c & 1u) * 8u;
}
__device__ __forceinline__ void umma2_bf16_lo(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// Debug trace (CTA 0 only): regions 0 = epilogue thread 0, 1 = MMA issuer, 2 = transform thread 0, 3 = copy-engine warp lane 0.
template <bool TRACE>
__device__ __forceinline__ void trace_event_t(const Params& p, int region, uint32_t& cursor, int ev, int64_t tile, int aux = 0) {
  if (TRACE && p.trace != nullptr && blockIdx.x == 0 && cursor < 16000u) {
    p.trace[1 + region * 16000 + cursor] = ((unsigned long long)ev << 56) | ((unsigned long long)(tile & 0xFFFF) << 40) |
                                             ((unsigned long long)(aux & 0xFF) << 32) | (unsigned long long)(clock64() & 0xFFFFFFFFull);
    ++cursor;
  }
}

template <int MODE, bool DROP, bool RELU, bool TRACE, bool BF16, bool SPLIT>  // SPLIT = main / cross-product accumulators (d >= 640, 3xTF32); BF16 = bf16 operands, one kind::f16 pass; MODE 0 = K2 forward, 1 = K4a dgrad, 2 = dense forward (atom message passing); TRACE = the role-timeline build (scripts/trace_pair.py)
__global__ void __launch_bounds__(THREADS, 1)
layer_gemm_pair(const Params p) {
  auto trace_event = [](const Params& pp, int region, uint32_t& cursor, int ev, int64_t tile, int aux = 0) {
    trace_event_t<TRACE>(pp, region, cursor, ev, tile, aux);
  };
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t bar_w = sbase + OFF_BAR;              // [STAGES]
  const uint32_t bar_ready = bar_w + 8 * STAGES;       // [STAGES] (leader's copy is the live one)
  const uint32_t bar_empty = bar_ready + 8 * STAGES;   // [STAGES]
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;
  const uint32_t bar_tmem_empty = bar_tmem_full + 8;   // (leader's copy is the live one)
  const uint32_t tmem_slot = bar_tmem_empty + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (3 * STAGES + 2));

  const Geometry& geo = p.geo;
  const int d = geo.d;
  // tile bookkeeping is 32-bit and re-derived inside every role (blockIdx / gridDim / kernel parameters are free to read):
  // values computed here would stay live across the whole role dispatch and were spilled to local memory
#define NT_PAIR_TILE_VARS \
  const int pair_tiles = (int)((p.E + 2 * TILE_M - 1) / (2 * TILE_M)), first_tile = (int)(blockIdx.x >> 1), tile_stride = (int)(gridDim.x >> 1)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_w + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, 2 * NUM_X_WARPS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem_full, 1);
    mbar_init(bar_tmem_empty, 2 * NUM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {  // the same warp of both CTAs allocates the pair's tensor memory (all 512 columns)
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  uint32_t tcur = 0;

  if (warp < NUM_EPI_WARPS) {
    // ===================================== EPILOGUE =====================================
    // 32 accumulator columns per step: two tcgen05.ld of 16 columns -> XOR-swizzled 4 KiB staging tile (thread = row) ->
    // 128-byte row segments (8 lanes per row, 4 rows per instruction): + bias, dropout, + residual, coalesced stores.
    // Residual rows and the bias piece are fetched TWO steps ahead into registers. (The first version moved 16 columns per
    // step in 64-byte segments: 19 dependent chains per tile and twice the L1TEX requests; the role trace showed the
    // epilogue busy 55 k of every 60 k clk.)
    NT_PAIR_TILE_VARS;
    uint8_t* stage = smem + OFF_EPI + warp * EPI_WARP_BYTES;
    const uint32_t tmem_empty_leader = map_to_cta(bar_tmem_empty, 0);
    uint32_t tphase = 0;
    const int rem = geo.n_tile % EPI_COLS;                              // 0 or 16: one narrow step when n_tile is not a multiple of 32
    const int chunks = geo.n_tile / EPI_COLS + (rem ? 1 : 0);
    const int sub = lane & 7, rsub = lane >> 3;
    const bool has_resid = MODE != 1 && p.resid != nullptr;
    // overlap of the two TMEM windows (with two accumulators per pass everything is "shared": hand back after the last step)
    const int shared_chunks = SPLIT ? chunks : (2 * geo.n_tile > 512 ? (2 * geo.n_tile - 512) / EPI_COLS : 0);
    int tw = 0;
    for (int tile = first_tile; tile < pair_tiles; tile += tile_stride) {
      const int64_t row0 = (int64_t)tile * (2 * TILE_M) + rank * TILE_M + warp * 32;
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
        // TMEM windows alternate between columns [0, n_tile) and [512 - n_tile, 512); where they overlap (n_tile > 256) this
        // pass drains the overlap first and then hands tensor memory back, so the next tile's MMAs run under the rest of
        // the epilogue. Window 0 shares its LAST columns, window 1 its first: window 0 therefore puts the narrow step first,
        // so that the overlap is a whole number of 32-column steps in both.
        const int col_base = (tw && !SPLIT) ? 512 - geo.n_tile : 0;
        const bool narrow_first = tw == 0 && rem != 0 && shared_chunks > 0 && !SPLIT;
        const int first = (tw == 0 && shared_chunks > 0 && !SPLIT) ? chunks - shared_chunks : 0;
        auto chunk_at = [&](int k) { int c = k + first; return c >= chunks ? c - chunks : c; };
        auto chunk_c0 = [&](int cc) { return narrow_first ? (cc == 0 ? 0 : rem + EPI_COLS * (cc - 1)) : EPI_COLS * cc; };
        auto chunk_w = [&](int cc) { return rem == 0 ? EPI_COLS : (narrow_first ? (cc == 0 ? rem : EPI_COLS) : (cc == chunks - 1 ? rem : EPI_COLS)); };
        auto prefetch = [&](int k, float4 (&dst)[8], float4& bias4) {  // residual rows + bias piece of the k-th step of this pass
          bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) dst[it] = bias4;
          if (k >= chunks) return;
          const int cc = chunk_at(k);
          const int col = nt * geo.n_tile + chunk_c0(cc) + sub * 4;
          if (sub * 4 >= chunk_w(cc) || col >= d) return;
          if (MODE != 1 && p.bias != nullptr) bias4 = ldg4(p.bias + col);
          if (!has_resid || (p.ablate & 1)) return;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int64_t e = row0 + it * 4 + rsub;
            if (e < p.E) dst[it] = ldg4_stream(p.resid + e * d + col);
          }
        };
        float4 rA[8], rB[8], bA, bB;
        prefetch(0, rA, bA);
        prefetch(1, rB, bB);
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 1, tile);
        mbar_wait_relaxed(bar_tmem_full, tphase);
        tc_fence_after();
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 2, tile);
        if (shared_chunks == 0 && lane == 0) mbar_arrive_cluster(tmem_empty_leader);
        auto do_chunk = [&](int k, float4 (&cur)[8], float4& bias4) {
          const int cc = chunk_at(k);
          const int c0 = chunk_c0(cc), w = chunk_w(cc);
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(col_base + c0);
          uint32_t v[16];
          // thread = accumulator row: 16 columns -> pieces `half * 4 .. + 3` of its 128-byte staging row (XOR-swizzled)
          auto stage16 = [&](uint32_t ta, int half) {
            tmem_ld16(ta, v);
            if (SPLIT) {  // main + cross-product accumulators: both loads in flight together, added in registers (fp32, round to nearest)
              uint32_t x[16];
              tmem_ld16(ta + 256u, x);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(x[q]));
            } else {
              tmem_ld_wait();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((q + 4 * half) ^ (lane & 7)) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          };
          stage16(taddr, 0);
          if (w > 16) stage16(taddr + 16, 1);
          if (shared_chunks > 0 && k == shared_chunks - 1) {  // the overlap has left tensor memory: the next pass may start its MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
            if (threadIdx.x == 0) trace_event(p, 0, tcur, 3, tile);
          }
          __syncwarp();
          const int col = nt * geo.n_tile + c0 + sub * 4;
          if (sub * 4 < w && col < d) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + rsub;
              const int64_t e = row0 + r;
              if (e < p.E) {
                float4 acc = *reinterpret_cast<const float4*>(stage + r * 128 + ((sub ^ (r & 7)) << 4));
                if (MODE != 1) {
                  acc = make_float4(acc.x + bias4.x, acc.y + bias4.y, acc.z + bias4.z, acc.w + bias4.w);
                  if (DROP) {
                    float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)col, p.drop_thr, p.inv_keep);
                    acc = make_float4(acc.x * sc.x, acc.y * sc.y, acc.z * sc.z, acc.w * sc.w);
                  }
                  if (has_resid) acc = make_float4(cur[it].x + acc.x, cur[it].y + acc.y, cur[it].z + acc.z, cur[it].w + acc.w);
                }
                if (!(p.ablate & 1)) stg4(p.out + e * d + col, acc);
              }
            }
          }
          __syncwarp();
          prefetch(k + 2, cur, bias4);
        };
        for (int k = 0; k < chunks; k += 2) {
          do_chunk(k, rA, bA);
          if (k + 1 < chunks) do_chunk(k + 1, rB, bB);
        }
        tc_fence_before();
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 4, tile);
        tphase ^= 1;
        tw ^= 1;
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA ISSUER (leader CTA only) =====================================
    if (leader) {
      NT_PAIR_TILE_VARS;
      const uint32_t idesc_a = BF16 ? make_idesc_pair_bf16(geo.n_a) : make_idesc_pair(geo.n_a);
      const uint32_t idesc_b = BF16 ? make_idesc_pair_bf16(geo.n_b > 0 ? geo.n_b : 16) : make_idesc_pair(geo.n_b > 0 ? geo.n_b : 16);
      int s = 0, tw = 0;
      uint32_t ph = 0, tphase = 0;
      for (int tile = first_tile; tile < pair_tiles; tile += tile_stride) {
        for (int nt = 0; nt < geo.n_tiles; ++nt) {
          const uint32_t win_base = (tw && !SPLIT) ? (uint32_t)(512 - geo.n_tile) : 0u;
          if (lane == 0) trace_event(p, 1, tcur, 10, tile);
          mbar_wait(bar_tmem_empty, tphase ^ 1);
          tc_fence_after();
          if (lane == 0) trace_event(p, 1, tcur, 11, tile);
#pragma unroll 1
          for (int kb = 0; kb < geo.k_blocks; ++kb) {
            mbar_wait(bar_ready + 8 * s, ph);
            tc_fence_after();
            if (lane == 0) trace_event(p, 1, tcur, 12, tile, kb);
            if (elect_one()) {
              const int rem = d - kb * BLOCK_K;
              const uint32_t st0 = sbase + s * STAGE_BYTES;
              const uint32_t col_base = SPLIT ? 0u : win_base;
              const int kfirst = kb;                                   // 0 on an accumulator's first K-block
              const uint32_t d0 = tmem_base + col_base, d1 = d0 + (uint32_t)geo.n_a;
              const uint32_t x0 = SPLIT ? d0 + 256u : d0, x1 = SPLIT ? d1 + 256u : d1;  // where the cross products accumulate
              if (BF16) {
                const int ksteps = (p.ablate & 8) ? 0 : (rem >= BLOCK_K ? 2 : (rem + 15) / 16);  // K = 16 bf16 per MMA
                const uint32_t a_bf = kmajor_desc_lo(st0 + OFF_A0), w_bf = kmajor_desc_lo(st0 + OFF_WHI);
                const uint32_t woff = (uint32_t)(geo.n_a / 2) * (64u >> 4);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  if (j < ksteps) {
                    const uint32_t k16 = j * 2;  // 16 bf16 = 32 bytes inside the 64-byte swizzle row
                    const uint32_t acc = (kfirst | j) != 0 ? 1u : 0u;
                    umma2_bf16_lo(d0, a_bf + k16, w_bf + k16, KMAJOR_SW64_DESC_HI, idesc_a, acc);
                    if (geo.n_b > 0) umma2_bf16_lo(d1, a_bf + k16, w_bf + woff + k16, KMAJOR_SW64_DESC_HI, idesc_b, acc);
                  }
                }
              } else {
              const int ksteps = (p.ablate & 8) ? 0 : (rem >= BLOCK_K ? BLOCK_K / 8 : (rem + 7) / 8);
              const uint32_t a_hi = kmajor_desc_lo(st0 + OFF_A0), a_lo = kmajor_desc_lo(st0 + OFF_A1);
              const uint32_t w_hi = kmajor_desc_lo(st0 + OFF_WHI), w_lo = kmajor_desc_lo(st0 + OFF_WLO);
              const uint32_t woff = (uint32_t)(geo.n_a / 2) * (128u >> 4);  // this CTA's rows of the second MMA follow its n_a / 2 rows of the first
#pragma unroll
              for (int j = 0; j < BLOCK_K / 8; ++j) {
                if (j < ksteps) {
                  const uint32_t k16 = j * 2;
                  const uint32_t acc = (kfirst | j) != 0 ? 1u : 0u;
                  if (p.products == 3) {
                    umma2_tf32_lo(x0, a_lo + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, acc);
                    umma2_tf32_lo(x0, a_hi + k16, w_lo + k16, KMAJOR_SW128_DESC_HI, idesc_a, 1u);
                    umma2_tf32_lo(d0, a_hi + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, SPLIT ? acc : 1u);
                    if (geo.n_b > 0) {
                      umma2_tf32_lo(x1, a_lo + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, acc);
                      umma2_tf32_lo(x1, a_hi + k16, w_lo + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, 1u);
                      umma2_tf32_lo(d1, a_hi + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, SPLIT ? acc : 1u);
                    }
                  } else {
                    umma2_tf32_lo(d0, a_hi + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, acc);
                    if (geo.n_b > 0) umma2_tf32_lo(d1, a_hi + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, acc);
                  }
                }
              }
              }
              umma2_commit_both(bar_empty + 8 * s);
              if (kb == geo.k_blocks - 1) umma2_commit_both(bar_tmem_full);
            }
            __syncwarp();
            if (lane == 0) trace_event(p, 1, tcur, 14, tile, kb);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          tphase ^= 1;
          tw ^= 1;
        }
      }
    }
  } else if (warp == TMA_WARP) {
    // ===================================== W PRODUCER (bulk copies of this CTA's half of the weight tile) =====================================
    NT_PAIR_TILE_VARS;
    int s = 0;
    uint32_t ph = 0;
    const uint32_t w_bytes = (uint32_t)geo.rows_per_cta * (BF16 ? 64u : 128u);
    const bool need_lo = !BF16 && p.products == 3;
    for (int tile = first_tile; tile < pair_tiles; tile += tile_stride) {
      __syncwarp();
#pragma unroll 1
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
#pragma unroll 1
        for (int kb = 0; kb < geo.k_blocks; ++kb) {
          mbar_wait_relaxed(bar_empty + 8 * s, ph ^ 1);
          if (elect_one()) {
            const uint32_t st0 = sbase + s * STAGE_BYTES;
            trace_event(p, 3, tcur, 30, tile, kb);
            const size_t off = (((size_t)nt * geo.k_blocks + kb) * 2 + rank) * w_bytes;
            if (p.ablate & 4) {
              mbar_arrive(bar_w + 8 * s);
            } else {
              mbar_arrive_expect_tx(bar_w + 8 * s, need_lo ? 2 * w_bytes : w_bytes);
              bulk_copy_g2s(st0 + OFF_WHI, p.wimg + off, w_bytes, bar_w + 8 * s);
              if (need_lo) bulk_copy_g2s(st0 + OFF_WLO, p.wimg + geo.part_bytes + off, w_bytes, bar_w + 8 * s);  // (bf16 image: one part only)
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================================== A PRODUCER: cp.async gather, then TF32 hi / lo split in place =====================================
    // Thread pt owns the 16-byte units u = pt + 256 i (i < 4) of each 16 KiB tile: row (pt >> 3) + 32 i, physical chunk pt & 7.
    // It copies exactly those units from global memory itself (cp.async, no registers, STAGES - 1 K-blocks ahead) and later
    // rewrites them, so the raw data needs no barrier at all: cp.async.wait_group orders a thread's own copies.
    const int pt = threadIdx.x - FIRST_X_WARP * 32;  // 0..255
    const int r0 = pt >> 3, c = (pt & 7) ^ (r0 & 7);  // tile row (mod 32) and LOGICAL 16-byte chunk inside the 128-byte K-block row
    const uint32_t ready_leader = map_to_cta(bar_ready, 0);
    const float* __restrict__ a0g = p.a0;
    const float* __restrict__ a1g = p.a1;

    // 32-bit bookkeeping throughout (E < 2^31): with 64-bit tile counters the two cursors were spilled to local memory and
    // ncu showed the reloads as ~15 % of all stall samples
    NT_PAIR_TILE_VARS;
    const int n_tiles_i = pair_tiles, first_i = first_tile, stride_i = tile_stride;
    const int E_i = (int)p.E;
    struct Cursor { int tile, nt, kb, s; uint32_t ph; };
    auto advance = [&](Cursor& q) {
      if (++q.kb == geo.k_blocks) { q.kb = 0; if (++q.nt == geo.n_tiles) { q.nt = 0; q.tile += stride_i; } }
      if (++q.s == STAGES) { q.s = 0; q.ph ^= 1; }
    };
    // row indices (-1 = row past E) of this thread's four rows, for the tile the load cursor is in and for the one after it
    int ra[4], rb[4], na[4], nb[4];
    auto fetch_rows = [&](int t, int (&xa)[4], int (&xb)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = t * (2 * TILE_M) + (int)rank * TILE_M + r0 + 32 * i;
        const bool ok = t < n_tiles_i && e < E_i;
        if (MODE == 0) {
          xa[i] = ok ? __ldg(p.src + e) : -1;
          xb[i] = ok ? __ldg(p.rev + e) : -1;
        } else {
          xa[i] = ok ? e : -1;
          xb[i] = -1;
        }
      }
    };
    auto issue_loads = [&](const Cursor& q) {
      const int k0 = q.kb * BLOCK_K + c * 4;
      const bool kvalid = k0 < d;
      const uint32_t dst0 = sbase + q.s * STAGE_BYTES + OFF_A0 + pt * 16, dst1 = dst0 + (OFF_A1 - OFF_A0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = kvalid && ra[i] >= 0;
        const float* src = ok ? a0g + (int64_t)ra[i] * d + k0 : a0g;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + i * (NUM_X_THREADS * 16)), "l"(src), "r"(ok ? 16 : 0) : "memory");
        if (MODE == 0) {
          const bool okb = kvalid && rb[i] >= 0;
          const float* srcb = okb ? a1g + (int64_t)rb[i] * d + k0 : a1g;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst1 + i * (NUM_X_THREADS * 16)), "l"(srcb), "r"(okb ? 16 : 0) : "memory");
        }
      }
    };
    auto load_step = [&](Cursor& q) {  // refill the stage the load cursor points at (once its previous MMAs have retired), then advance
      if (q.tile < n_tiles_i) {
        mbar_wait(bar_empty + 8 * q.s, q.ph ^ 1);
        if (!(p.ablate & 2)) issue_loads(q);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int before = q.tile;
      advance(q);
      if (q.tile != before) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ra[i] = na[i]; rb[i] = nb[i]; }
        fetch_rows(q.tile + stride_i, na, nb);
      }
    };

    Cursor ld{first_i, 0, 0, 0, 0}, cp{first_i, 0, 0, 0, 0};
    fetch_rows(first_i, ra, rb);
    fetch_rows(first_i + stride_i, na, nb);
#pragma unroll 1
    for (int j = 0; j < STAGES - 1; ++j) load_step(ld);
#pragma unroll 1
    while (cp.tile < n_tiles_i) {
      const int e0 = cp.tile * (2 * TILE_M) + (int)rank * TILE_M;
      uint8_t* a0 = smem + cp.s * STAGE_BYTES + OFF_A0;
      uint8_t* a1 = a0 + (OFF_A1 - OFF_A0);
      if (pt == 0) trace_event(p, 2, tcur, 20, cp.tile, cp.kb);
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
      if (pt == 0) trace_event(p, 2, tcur, 21, cp.tile, cp.kb);
      const int col = cp.kb * BLOCK_K + c * 4;
      if (!(p.ablate & 16)) {
        // all shared-memory reads first (the in-place stores below would otherwise serialise the four units: the
        // compiler cannot move a later load above an earlier store to the same tile)
        float4 raw0[4], raw1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = pt + i * NUM_X_THREADS;
          raw0[i] = *reinterpret_cast<const float4*>(a0 + u * 16);
          if (MODE == 0) raw1[i] = *reinterpret_cast<const float4*>(a1 + u * 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = pt + i * NUM_X_THREADS;
          const int e = e0 + r0 + 32 * i;
          float4 m = raw0[i];
          if (MODE == 0) {
            float4 a;
            if (RELU) a = make_float4(raw1[i].x < 0.f ? 0.f : raw1[i].x, raw1[i].y < 0.f ? 0.f : raw1[i].y, raw1[i].z < 0.f ? 0.f : raw1[i].z, raw1[i].w < 0.f ? 0.f : raw1[i].w);
            else a = act_fwd4(raw1[i], p.act, p.act_param);
            m = make_float4(m.x - a.x, m.y - a.y, m.z - a.z, m.w - a.w);
            if (col >= d || e >= E_i) m = make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (MODE == 1 && DROP && e < E_i && col < d) {
            const float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)col, p.drop_thr, p.inv_keep);
            m = make_float4(m.x * sc.x, m.y * sc.y, m.z * sc.z, m.w * sc.w);
          }
          if (!BF16) {
            // K2 forms m in registers and has to write it anyway: fully rounded split. Rows used as they are (dgrad without dropout,
            // dense forward) stay in place as the hi operand and only lo is stored (tc_common.cuh).
            constexpr bool IN_PLACE = MODE == 2 || (MODE == 1 && !DROP);
            float4 hi, lo;
            tf32_split4<IN_PLACE ? SPLIT_INPLACE : SPLIT_RNA>(m, hi, lo);
            if (!IN_PLACE) *reinterpret_cast<float4*>(a0 + u * 16) = hi;
            *reinterpret_cast<float4*>(a1 + u * 16) = lo;
          }
          raw0[i] = m;
        }
        if (BF16) {
          // the bf16 tile (8 KiB, 64-byte rows) overwrites the raw fp32 tile: every producer thread must have read its raw units first
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_X_THREADS) : "memory");
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 lo2 = __float22bfloat162_rn(make_float2(raw0[i].x, raw0[i].y));
            const __nv_bfloat162 hi2 = __float22bfloat162_rn(make_float2(raw0[i].z, raw0[i].w));
            uint2 packed;
            packed.x = *reinterpret_cast<const uint32_t*>(&lo2);
            packed.y = *reinterpret_cast<const uint32_t*>(&hi2);
            *reinterpret_cast<uint2*>(a0 + swz64_bf16((uint32_t)(r0 + 32 * i), (uint32_t)c)) = packed;
          }
        }
        if (MODE == 0 && p.m_out != nullptr && cp.nt == 0 && col < d && !(p.ablate & 32)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = e0 + r0 + 32 * i;
            if (e < E_i) stg4_stream(p.m_out + (int64_t)e * d + col, raw0[i]);
          }
        }
      }
      if (pt == 0) trace_event(p, 2, tcur, 23, cp.tile, cp.kb);
      fence_proxy_async();  // generic-proxy writes to this CTA's shared memory -> visible to the tensor core (async proxy)
      __syncwarp();
      if (pt == 0) trace_event(p, 2, tcur, 24, cp.tile, cp.kb);
      if (lane == 0) {
        if (warp == FIRST_X_WARP) mbar_wait(bar_w + 8 * cp.s, cp.ph);  // this CTA's half of W has landed, too
        if (leader) mbar_arrive(bar_ready + 8 * cp.s);
        else mbar_arrive_cluster(ready_leader + 8 * cp.s);
      }
      if (pt == 0) trace_event(p, 2, tcur, 22, cp.tile, cp.kb);
      advance(cp);
      load_step(ld);
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

// W [d,d] -> hi/lo TF32 parts; per (N tile, K block) first the leader's rows, then the peer's, each a 128-byte-swizzled
// K-major tile of rows_per_cta rows: n_a / 2 rows of the first MMA followed by n_b / 2 rows of the second.
__global__ void __launch_bounds__(256) pair_weight_prepare_kernel(const float* __restrict__ W, Geometry geo, int transpose, uint8_t* __restrict__ image) {
  const int64_t total = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * BLOCK_K;
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % BLOCK_K);
  int64_t q = t / BLOCK_K;
  const int j = (int)(q % geo.rows_per_cta);
  q /= geo.rows_per_cta;
  const int rank = (int)(q % 2);
  q /= 2;
  const int kb = (int)(q % geo.k_blocks);
  const int nt = (int)(q / geo.k_blocks);
  const int ha = geo.n_a / 2, hb = geo.n_b / 2;
  const int r = j < ha ? rank * ha + j : geo.n_a + rank * hb + (j - ha);
  const int n = nt * geo.n_tile + r, k = kb * BLOCK_K + kk;
  float v = 0.f;
  if (n < geo.d && k < geo.d) v = transpose ? __ldg(W + (int64_t)k * geo.d + n) : __ldg(W + (int64_t)n * geo.d + k);
  const float hi = tf32_rna(v);
  const float lo = tf32_rna(v - hi);
  const size_t off = ((((size_t)nt * geo.k_blocks + kb) * 2 + rank) * geo.rows_per_cta) * 128 + swz128((uint32_t)j, (uint32_t)(kk >> 2)) + (kk & 3) * 4;
  *reinterpret_cast<float*>(image + off) = hi;
  *reinterpret_cast<float*>(image + geo.part_bytes + off) = lo;
}

// bf16 image: same (N tile, K block, CTA rank) order, one part, 64-byte-swizzled rows of 32 bf16
__global__ void __launch_bounds__(256) pair_weight_prepare_bf16_kernel(const float* __restrict__ W, Geometry geo, int transpose, uint8_t* __restrict__ image) {
  const int64_t total = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * (BLOCK_K / 4);  // one thread per 4 values
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % (BLOCK_K / 4));
  int64_t q = t / (BLOCK_K / 4);
  const int j = (int)(q % geo.rows_per_cta);
  q /= geo.rows_per_cta;
  const int rank = (int)(q % 2);
  q /= 2;
  const int kb = (int)(q % geo.k_blocks);
  const int nt = (int)(q / geo.k_blocks);
  const int ha = geo.n_a / 2, hb = geo.n_b / 2;
  const int r = j < ha ? rank * ha + j : geo.n_a + rank * hb + (j - ha);
  const int n = nt * geo.n_tile + r;
  float v[4];
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int k = kb * BLOCK_K + 4 * c + x;
    v[x] = (n < geo.d && k < geo.d) ? (transpose ? __ldg(W + (int64_t)k * geo.d + n) : __ldg(W + (int64_t)n * geo.d + k)) : 0.f;
  }
  const __nv_bfloat162 lo2 = __float22bfloat162_rn(make_float2(v[0], v[1])), hi2 = __float22bfloat162_rn(make_float2(v[2], v[3]));
  uint2 packed;
  packed.x = *reinterpret_cast<const uint32_t*>(&lo2);
  packed.y = *reinterpret_cast<const uint32_t*>(&hi2);
  const size_t off = ((((size_t)nt * geo.k_blocks + kb) * 2 + rank) * geo.rows_per_cta) * 64 + swz64_bf16((uint32_t)j, (uint32_t)c);
  *reinterpret_cast<uint2*>(image + off) = packed;
}

template <int MODE, bool DROP, bool RELU, bool TRACE, bool BF16, bool SPLIT>
static int launch_variant(const Params& p, cudaStream_t st) {
  static PerDeviceOnce once;  // one per kernel instantiation
  const cudaError_t attr_err = once.run([] {
    return cudaFuncSetAttribute(layer_gemm_pair<MODE, DROP, RELU, TRACE, BF16, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(layer_gemm_pair)");
  const int64_t pair_tiles = (p.E + 2 * TILE_M - 1) / (2 * TILE_M);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  int64_t clusters = sms / 2;
  if (pair_tiles < clusters) clusters = pair_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, layer_gemm_pair<MODE, DROP, RELU, TRACE, BF16, SPLIT>, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(layer_gemm_pair)");
  NT_LAUNCH_CHECK("layer_gemm_pair", 1);
  return NT_OK;
}

template <int MODE, bool BF16, bool SPLIT>
static int launch_by_flags(const Params& p, cudaStream_t st) {
  const bool drop = p.drop_p > 0.f;
  const bool relu = MODE != 0 || p.act == NT_ACT_RELU;  // the dense modes have no activation prologue
  if (MODE != 0) return drop ? launch_variant<MODE, true, true, false, BF16, SPLIT>(p, st) : launch_variant<MODE, false, true, false, BF16, SPLIT>(p, st);
  if (drop) return relu ? launch_variant<MODE, true, true, false, BF16, SPLIT>(p, st) : launch_variant<MODE, true, false, false, BF16, SPLIT>(p, st);
  return relu ? launch_variant<MODE, false, true, false, BF16, SPLIT>(p, st) : launch_variant<MODE, false, false, false, BF16, SPLIT>(p, st);
}

template <int MODE>
static int launch(const Params& p, cudaStream_t st) {
  if (p.products == 0) return launch_by_flags<MODE, true, false>(p, st);  // bf16 operands (the mode's error dwarfs the accumulator's)
  if (p.products == 3 && p.geo.k_blocks >= split_min_kblocks() && p.geo.n_tile <= 256) return launch_by_flags<MODE, false, true>(p, st);  // d >= 640, 3xTF32: main / cross accumulators
  const bool drop = p.drop_p > 0.f;
  const bool relu = MODE != 0 || p.act == NT_ACT_RELU;
  if (MODE != 2 && p.trace != nullptr && !drop && relu) return launch_variant<MODE, false, true, true, false, false>(p, st);  // the role-timeline build exists for the default case only
  return launch_by_flags<MODE, false, false>(p, st);
}

unsigned long long* g_trace_buffer = nullptr;

static void fill_dropout(Params& p, float drop_p, uint64_t seed, uint64_t offset) {
  static const int ablate = getenv("NOTORCH_B200_ABLATE") ? atoi(getenv("NOTORCH_B200_ABLATE")) : 0;
  p.ablate = ablate;
  p.trace = g_trace_buffer;
  p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  double t = (double)drop_p * 4294967296.0;
  p.drop_thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  p.seed = seed;
  p.offset = offset;
}

}  // namespace pair

void pair_set_trace_buffer(void* ptr) { pair::g_trace_buffer = static_cast<unsigned long long*>(ptr); }

size_t pair_weight_image_bytes(int64_t d) { return 2 * pair::make_geometry((int)d).part_bytes; }

int pair_weight_prepare(const float* W, int64_t d, int transpose, void* image, int bf16, cudaStream_t st) {
  pair::Geometry geo = pair::make_geometry((int)d);
  if (bf16) {
    const int64_t total4 = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * (pair::BLOCK_K / 4);
    pair::pair_weight_prepare_bf16_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, st>>>(W, geo, transpose, static_cast<uint8_t*>(image));
    NT_LAUNCH_CHECK("pair_weight_prepare_bf16_kernel", 1);
    return NT_OK;
  }
  const int64_t total = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * pair::BLOCK_K;
  pair::pair_weight_prepare_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(W, geo, transpose, static_cast<uint8_t*>(image));
  NT_LAUNCH_CHECK("pair_weight_prepare_kernel", 1);
  return NT_OK;
}

int pair_layer_forward(const float* h, const float* n, const int32_t* src, const int32_t* rev, const void* wimg, const float* bias, int64_t E, int64_t V,
                       int64_t d, int act, float act_param, int residual, float drop_p, uint64_t seed, uint64_t offset, float* out, float* m_out,
                       int products, cudaStream_t st) {
  pair::Params p{};
  p.m_out = m_out;
  p.src = src; p.rev = rev; p.wimg = static_cast<const uint8_t*>(wimg); p.bias = bias;
  p.resid = residual ? h : nullptr; p.out = out; p.E = E; p.geo = pair::make_geometry((int)d);
  p.act = act; p.act_param = act_param; p.products = products;
  pair::fill_dropout(p, drop_p, seed, offset);
  (void)V;
  p.a0 = n; p.a1 = h;
  return pair::launch<0>(p, st);
}

// out[r,:] = (resid ? resid[r,:] : 0) + Dropout(x[r,:] . W^T + bias): the Linear + Dropout + residual of an atom-state update
// (dense A operand; the bias / dropout / residual epilogue of K2)
int pair_dense_forward(const float* x, const void* wimg, const float* bias, const float* resid, int64_t R, int64_t d, float drop_p, uint64_t seed,
                       uint64_t offset, float* out, int products, cudaStream_t st) {
  pair::Params p{};
  p.a0 = x; p.wimg = static_cast<const uint8_t*>(wimg); p.bias = bias; p.resid = resid; p.out = out; p.E = R;
  p.geo = pair::make_geometry((int)d);
  p.act = NT_ACT_IDENTITY; p.products = products;
  pair::fill_dropout(p, drop_p, seed, offset);
  return pair::launch<2>(p, st);
}

int pair_layer_dgrad(const float* g, const void* wimg, int64_t E, int64_t d, float drop_p, uint64_t seed, uint64_t offset, float* g_m, int products,
                     cudaStream_t st) {
  pair::Params p{};
  p.wimg = static_cast<const uint8_t*>(wimg); p.out = g_m; p.E = E; p.geo = pair::make_geometry((int)d);
  p.act = NT_ACT_IDENTITY; p.products = products;
  pair::fill_dropout(p, drop_p, seed, offset);
  p.a0 = g;
  return pair::launch<1>(p, st);
}

}  // namespace nt
