// Tensor-core (tcgen05 / TMEM / bulk-TMA) version of K2 forward and K4a dgrad for sm_100a.
//
//   D[128 edges, N] = A[128, K] . B[N, K]^T        fp32 in, fp32 out, 3xTF32 error-compensated:
//   A = A_hi + A_lo, B = B_hi + B_lo (each part exactly representable in TF32), and
//   D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi accumulated in fp32 in tensor memory.
//
// Warp-specialised persistent kernel, one CTA per SM, 14 warps:
//   warps 0-3   epilogue   tcgen05.ld accumulator -> +bias, dropout, +residual -> coalesced stores
//   warp  4     MMA issuer one elected lane issues tcgen05.mma.kind::tf32 and tcgen05.commit; owns TMEM
//   warp  5     W producer bulk-TMA (cp.async.bulk, SASS UBLKCP) copies of pre-swizzled weight tiles
//   warps 6-13  A producer gathers n[src[e]] - act(h[rev[e]]) (K2) or dropout(g[e]) (K4a) rows with
//               128-bit loads, splits hi/lo and writes the 128-byte-swizzled K-major UMMA layout
// Pipelines (mbarriers): smem stage full/empty between producers and MMA; TMEM full/empty between
// MMA and epilogue. A is produced by threads (generic proxy) so each producer thread executes
// fence.proxy.async before arriving; W lands through the async proxy with complete_tx.
//
// The A operand cannot come from TMA (it is a gather-and-subtract), which is why only W is staged
// by the copy engine; W is pre-split and pre-swizzled once per step by nt_weight_prepare so that
// each (N-tile, K-block) is one contiguous bulk copy.
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace nt {
namespace tc {

constexpr int BLOCK_K = 32;  // fp32 elements per K-block = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int MAX_N_TILE = 304;
constexpr int A_STAGE_BYTES = TILE_M * 128;          // 16 KiB per part
constexpr int W_STAGE_BYTES = MAX_N_TILE * 128;      // 38 KiB per part
constexpr int EPI_COLS = 16;                         // accumulator columns per epilogue step
constexpr int EPI_WARP_BYTES = 32 * EPI_COLS * 4;    // 2 KiB staging tile per epilogue warp (XOR-swizzled)
constexpr int PF_DIST = 3;                           // producers prefetch this many K-blocks ahead into L2
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, W_WARP = 5, FIRST_A_WARP = 6, NUM_A_WARPS = 8;
constexpr int NUM_A_THREADS = NUM_A_WARPS * 32;
constexpr int THREADS = (FIRST_A_WARP + NUM_A_WARPS) * 32;  // 448

constexpr int OFF_A_HI = 0;
constexpr int OFF_A_LO = OFF_A_HI + STAGES * A_STAGE_BYTES;
constexpr int OFF_W_HI = OFF_A_LO + STAGES * A_STAGE_BYTES;
constexpr int OFF_W_LO = OFF_W_HI + STAGES * W_STAGE_BYTES;
constexpr int OFF_EPI = OFF_W_LO + STAGES * W_STAGE_BYTES;
constexpr int OFF_BAR = OFF_EPI + NUM_EPI_WARPS * EPI_WARP_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
static_assert(OFF_W_HI % 1024 == 0 && W_STAGE_BYTES % 1024 == 0 && A_STAGE_BYTES % 1024 == 0, "swizzle-128B tiles need 1 KiB alignment");

struct Geometry {
  int d, n_tile, n_tiles, k_blocks, n_a, n_b;
  size_t part_bytes;  // bytes of one (hi or lo) image
};

__host__ __device__ inline Geometry make_geometry(int d) {
  Geometry g;
  g.d = d;
  int d16 = (d + 15) / 16 * 16;
  g.n_tile = d16 <= MAX_N_TILE ? d16 : 256;
  g.n_tiles = (d + g.n_tile - 1) / g.n_tile;
  g.k_blocks = (d + BLOCK_K - 1) / BLOCK_K;
  if (g.n_tile <= 256) { g.n_a = g.n_tile; g.n_b = 0; }
  else { g.n_a = ((g.n_tile / 2) + 15) / 16 * 16; g.n_b = g.n_tile - g.n_a; }
  g.part_bytes = (size_t)g.n_tiles * g.k_blocks * g.n_tile * 128;
  return g;
}

struct Params {
  const float* a0;  // K2: n [V,d]      K4a: g [E,d]
  const float* a1;  // K2: h [E,d]      K4a: unused
  const int32_t* src;
  const int32_t* rev;
  const uint8_t* wimg;  // hi image followed by lo image
  const float* bias;
  const float* resid;  // K2 residual input h (nullable)
  float* out;
  float* m_out;  // K2: optional copy of the message tensor m [E,d] (saved for the weight gradient)
  int64_t E;
  Geometry geo;
  int act;
  float act_param;
  float drop_p, inv_keep;
  uint32_t drop_thr;
  uint64_t seed, offset;
  int products;  // 3 = 3xTF32, 1 = single-pass TF32
  unsigned long long* trace;  // timing experiments only: CTA 0 appends (event << 56 | tile << 40 | clock) records
  int ablate;    // timing experiments only (NOTORCH_B200_ABLATE): 1 no epilogue stores, 2 no producer loads, 4 W only once, 8 no MMAs
};

// Debug trace: each tracing thread owns a region of the buffer and a private cursor (plain stores, no atomics, so the
// probe costs a few cycles). Regions: 0 = epilogue thread 0, 1 = MMA issuer, 2 = producer thread 0; 16384 records each.
__device__ __forceinline__ void trace_event(const Params& p, int region, uint32_t& cursor, int ev, int64_t tile, int aux = 0) {
  if (p.trace != nullptr && blockIdx.x == 0 && cursor < 16384u) {
    p.trace[1 + region * 16384 + cursor] = ((unsigned long long)ev << 56) | ((unsigned long long)(tile & 0xFFFF) << 40) |
                                             ((unsigned long long)(aux & 0xFF) << 32) | (unsigned long long)(clock64() & 0xFFFFFFFFull);
    ++cursor;
  }
}

template <int MODE>  // 0 = K2 forward, 1 = K4a dgrad
__global__ void __launch_bounds__(THREADS, 1) layer_gemm_tc(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) printf("notorch_b200: dynamic shared memory base is not 1 KiB aligned\n");
    __trap();
  }
  const uint32_t bar_a_full = sbase + OFF_BAR;          // [STAGES]
  const uint32_t bar_w_full = bar_a_full + 8 * STAGES;  // [STAGES]
  const uint32_t bar_empty = bar_w_full + 8 * STAGES;   // [STAGES]
  const uint32_t bar_tmem_full = bar_empty + 8 * STAGES;
  const uint32_t bar_tmem_empty = bar_tmem_full + 8;
  const uint32_t tmem_slot = bar_tmem_empty + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (3 * STAGES + 2));

  const Geometry& geo = p.geo;
  const int d = geo.d;
  const int64_t m_tiles = (p.E + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_a_full + 8 * s, NUM_A_WARPS);  // one elected arrive per producer warp
      mbar_init(bar_w_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem_full, 1);
    mbar_init(bar_tmem_empty, NUM_EPI_WARPS);  // one elected arrive per epilogue warp
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
  uint32_t tcur = 0;  // debug-trace cursor of this thread

  if (warp < NUM_EPI_WARPS) {
    // ===================================== EPILOGUE =====================================
    // Per (tile, N tile): TMEM stays occupied until the last accumulator column has been read, and the next
    // tile's MMAs wait for it, so nothing in this loop may wait on DRAM: the residual rows are prefetched into
    // L2 while the MMAs of this tile are still running and then software-pipelined EPI_PF chunks ahead.
    uint8_t* stage = smem + OFF_EPI + warp * EPI_WARP_BYTES;
    uint32_t tphase = 0;
    const int chunks = geo.n_tile / EPI_COLS;
    const int sub = lane & 3, rsub = lane >> 2;
    const bool has_resid = MODE == 0 && p.resid != nullptr;
    const int shared_chunks = 2 * geo.n_tile > 512 ? (2 * geo.n_tile - 512) / EPI_COLS : 0;
    int tw = 0;
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * TILE_M + warp * 32;
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
        auto load_resid = [&](int cc, float4 (&dst)[4]) {
          const int col = nt * geo.n_tile + cc * EPI_COLS + sub * 4;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int64_t e = row0 + it * 8 + rsub;
            dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_resid && cc < chunks && col < d && e < p.E) dst[it] = ldg4_stream(p.resid + e * d + col);
          }
        };
        // TMEM windows: successive (tile, N tile) passes alternate between columns [0, n_tile) and [512 - n_tile, 512).
        // They overlap in `shared_chunks` 16-column chunks (6 at n_tile = 304, none for n_tile <= 256); this pass drains
        // the overlap FIRST and then hands TMEM back, so the next pass's MMAs run under the rest of this epilogue.
        const int col_base = tw ? 512 - geo.n_tile : 0;
        const int first = (tw == 0 && shared_chunks > 0) ? chunks - shared_chunks : 0;  // rotation of the chunk order
        auto chunk_at = [&](int k) { int c = k + first; return c >= chunks ? c - chunks : c; };
        // residual rows are software-pipelined three chunks ahead in three statically named buffers (a rotating
        // register array would make every iteration wait for the load issued by the previous one)
        float4 rA[4], rB[4], rC[4];
        load_resid(chunk_at(0), rA);
        load_resid(chunks > 1 ? chunk_at(1) : chunks, rB);
        load_resid(chunks > 2 ? chunk_at(2) : chunks, rC);
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 1, tile);  // epilogue: waiting for the accumulator
        mbar_wait(bar_tmem_full, tphase);
        tc_fence_after();
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 2, tile);  // epilogue: accumulator ready
        if (shared_chunks == 0 && lane == 0) mbar_arrive(bar_tmem_empty);  // the other window is free as soon as this pass has begun
        auto do_chunk = [&](int k, float4 (&cur)[4]) {
          const int cc = chunk_at(k);
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(col_base + cc * EPI_COLS), v);
          tmem_ld_wait();
          if (threadIdx.x == 0) trace_event(p, 0, tcur, 5, tile, k);  // epilogue: chunk k loaded from TMEM
          if (shared_chunks > 0 && k == shared_chunks - 1) {  // overlap drained: the next pass may start its MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tmem_empty);
            if (threadIdx.x == 0) trace_event(p, 0, tcur, 3, tile);  // epilogue: overlap drained, TMEM handed back
          }
          // thread = accumulator row: 16 columns into the XOR-swizzled staging tile (conflict-free both ways)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stage + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
          if (threadIdx.x == 0) trace_event(p, 0, tcur, 6, tile, k);  // epilogue: chunk k staged
          // 4 lanes per row (64 contiguous bytes), 8 rows per step: coalesced global traffic
          const int col = nt * geo.n_tile + cc * EPI_COLS + sub * 4;
          if (col < d) {
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == 0 && p.bias) bias4 = ldg4(p.bias + col);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + rsub;
              const int64_t e = row0 + r;
              if (e < p.E) {
                float4 acc = *reinterpret_cast<const float4*>(stage + r * 64 + ((sub ^ ((r >> 1) & 3)) << 4));
                if (MODE == 0) {
                  acc = make_float4(acc.x + bias4.x, acc.y + bias4.y, acc.z + bias4.z, acc.w + bias4.w);
                  if (p.drop_p > 0.f) {
                    float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)e * (uint64_t)d + (uint64_t)col, p.drop_thr, p.inv_keep);
                    acc = make_float4(acc.x * sc.x, acc.y * sc.y, acc.z * sc.z, acc.w * sc.w);
                  }
                  if (has_resid) acc = make_float4(cur[it].x + acc.x, cur[it].y + acc.y, cur[it].z + acc.z, cur[it].w + acc.w);
                }
                if (!(p.ablate & 1)) stg4(p.out + e * d + col, acc);
              }
            }
          }
          if (threadIdx.x == 0) trace_event(p, 0, tcur, 7, tile, k);  // epilogue: chunk k stored
          __syncwarp();
          load_resid(k + 3 < chunks ? chunk_at(k + 3) : chunks, cur);  // refill this buffer for chunk k + 3
          if (threadIdx.x == 0) trace_event(p, 0, tcur, 8, tile, k);  // epilogue: refill issued
        };
        for (int k = 0; k < chunks; k += 3) {
          do_chunk(k, rA);
          if (k + 1 < chunks) do_chunk(k + 1, rB);
          if (k + 2 < chunks) do_chunk(k + 2, rC);
        }
        tc_fence_before();
        if (threadIdx.x == 0) trace_event(p, 0, tcur, 4, tile);  // epilogue: done
        tphase ^= 1;
        tw ^= 1;
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA ISSUER =====================================
    const uint32_t idesc_a = make_idesc_tf32(geo.n_a);
    const uint32_t idesc_b = make_idesc_tf32(geo.n_b > 0 ? geo.n_b : 16);
    int s = 0, tw = 0;
    uint32_t ph = 0, tphase = 0;
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
        const uint32_t col_base = tw ? (uint32_t)(512 - geo.n_tile) : 0u;  // alternate TMEM windows (see the epilogue)
        if (lane == 0) trace_event(p, 1, tcur, 10, tile);  // MMA: waiting for TMEM
        mbar_wait(bar_tmem_empty, tphase ^ 1);
        tc_fence_after();
        if (lane == 0) trace_event(p, 1, tcur, 11, tile);  // MMA: TMEM granted
        for (int kb = 0; kb < geo.k_blocks; ++kb) {
          mbar_wait(bar_a_full + 8 * s, ph);
          if (lane == 0) trace_event(p, 1, tcur, 12, tile, kb);  // MMA: A stage ready
          mbar_wait(bar_w_full + 8 * s, ph);
          tc_fence_after();
          if (lane == 0) trace_event(p, 1, tcur, 13, tile, kb);  // MMA: W stage ready
          if (elect_one()) {
            const int rem = d - kb * BLOCK_K;
            const int ksteps = (p.ablate & 8) ? 0 : (rem >= BLOCK_K ? BLOCK_K / 8 : (rem + 7) / 8);
            const uint32_t a_hi = kmajor_desc_lo(sbase + OFF_A_HI + s * A_STAGE_BYTES), a_lo = kmajor_desc_lo(sbase + OFF_A_LO + s * A_STAGE_BYTES);
            const uint32_t w_hi = kmajor_desc_lo(sbase + OFF_W_HI + s * W_STAGE_BYTES), w_lo = kmajor_desc_lo(sbase + OFF_W_LO + s * W_STAGE_BYTES);
            const uint32_t d0 = tmem_base + col_base, d1 = d0 + (uint32_t)geo.n_a;
            const uint32_t woff = (uint32_t)geo.n_a * (128u >> 4);  // second N half: n_a rows further down the W tile
#pragma unroll
            for (int j = 0; j < BLOCK_K / 8; ++j) {
              if (j < ksteps) {
                const uint32_t k16 = j * 2;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row, in 16-byte units
                const uint32_t acc = (kb | j) != 0 ? 1u : 0u;
                if (p.products == 3) {
                  umma_tf32_lo(d0, a_lo + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, acc);  // small terms first
                  umma_tf32_lo(d0, a_hi + k16, w_lo + k16, KMAJOR_SW128_DESC_HI, idesc_a, 1u);
                  umma_tf32_lo(d0, a_hi + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, 1u);
                  if (geo.n_b > 0) {
                    umma_tf32_lo(d1, a_lo + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, acc);
                    umma_tf32_lo(d1, a_hi + k16, w_lo + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, 1u);
                    umma_tf32_lo(d1, a_hi + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, 1u);
                  }
                } else {
                  umma_tf32_lo(d0, a_hi + k16, w_hi + k16, KMAJOR_SW128_DESC_HI, idesc_a, acc);
                  if (geo.n_b > 0) umma_tf32_lo(d1, a_hi + k16, w_hi + woff + k16, KMAJOR_SW128_DESC_HI, idesc_b, acc);
                }
              }
            }
            umma_commit(bar_empty + 8 * s);                               // frees this smem stage when the MMAs retire
            if (kb == geo.k_blocks - 1) umma_commit(bar_tmem_full);       // accumulator complete
          }
          __syncwarp();
          if (lane == 0) trace_event(p, 1, tcur, 14, tile, kb);  // MMA: K-block issued
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tphase ^= 1;
        tw ^= 1;
      }
    }
  } else if (warp == W_WARP) {
    // ===================================== W PRODUCER (bulk TMA) =====================================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tile_bytes = (uint32_t)geo.n_tile * 128u;
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
        for (int kb = 0; kb < geo.k_blocks; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (elect_one()) {
            const size_t off = ((size_t)nt * geo.k_blocks + kb) * tile_bytes;
            const bool need_lo = p.products == 3;
            if ((p.ablate & 4) && !(tile == blockIdx.x && kb < STAGES)) {
              mbar_arrive(bar_w_full + 8 * s);  // experiment: reuse whatever W the stage holds
            } else {
              mbar_arrive_expect_tx(bar_w_full + 8 * s, need_lo ? 2 * tile_bytes : tile_bytes);
              bulk_copy_g2s(sbase + OFF_W_HI + s * W_STAGE_BYTES, p.wimg + off, tile_bytes, bar_w_full + 8 * s);
              if (need_lo) bulk_copy_g2s(sbase + OFF_W_LO + s * W_STAGE_BYTES, p.wimg + geo.part_bytes + off, tile_bytes, bar_w_full + 8 * s);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================================== A PRODUCER =====================================
    const int pt = threadIdx.x - FIRST_A_WARP * 32;  // 0..255
    const int c = pt & 7;                            // 16-byte chunk inside the 128-byte K-block row
    const int r0 = pt >> 3;                          // rows r0, r0+32, r0+64, r0+96
    int s = 0;
    uint32_t ph = 0;
    // Row bases (element offsets) of this thread's four tile rows, for the current and the next tile. The next tile's
    // indices are fetched a whole tile early so that the dependent src/rev -> row-address chain is never exposed.
    int64_t rowA[4], rowB[4], nextA[4], nextB[4];
    bool valid[4], next_valid[4];
    auto fetch_rows = [&](int64_t t, int64_t (&ra)[4], int64_t (&rb)[4], bool (&ok)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t e = t * TILE_M + r0 + 32 * i;
        ok[i] = t < m_tiles && e < p.E;
        if (MODE == 0) {
          ra[i] = ok[i] ? (int64_t)__ldg(p.src + e) * d : 0;
          rb[i] = ok[i] ? (int64_t)__ldg(p.rev + e) * d : 0;
        } else {
          ra[i] = e * d;
          rb[i] = 0;
        }
      }
    };
    fetch_rows(blockIdx.x, nextA, nextB, next_valid);
    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { rowA[i] = nextA[i]; rowB[i] = nextB[i]; valid[i] = next_valid[i]; }
      fetch_rows(tile + gridDim.x, nextA, nextB, next_valid);
      if (MODE == 0 && p.resid != nullptr && warp == FIRST_A_WARP && elect_one()) {
        // the epilogue of THIS tile (one main loop from now) adds the residual rows h[tile rows, :]: one contiguous block
        const int64_t e0 = tile * TILE_M;
        const int64_t rows = p.E - e0 < TILE_M ? p.E - e0 : TILE_M;
        l2_prefetch_bulk(p.resid + e0 * d, (uint32_t)(rows * d * 4));
      }
      for (int nt = 0; nt < geo.n_tiles; ++nt) {
        for (int kb = 0; kb < geo.k_blocks; ++kb) {
          const int k0 = kb * BLOCK_K + c * 4;
          const bool kvalid = k0 < d;
          float4 va[4], vb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // all loads first: 8 x 16 B in flight per thread
            va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = va[i];
            if (valid[i] && kvalid && !(p.ablate & 2)) {
              if (MODE == 0) {
                va[i] = ldg4(p.a0 + rowA[i] + k0);
                vb[i] = ldg4_stream(p.a1 + rowB[i] + k0);
              } else {
                va[i] = ldg4_stream(p.a0 + rowA[i] + k0);
              }
            }
          }
          {  // L2 prefetch PF_DIST K-blocks ahead (this thread's own 16-byte chunks; running into the next tile's rows at the end)
            int kp = kb + PF_DIST;
            const bool into_next = kp >= geo.k_blocks;
            if (into_next) kp -= geo.k_blocks;
            const int kp0 = kp * BLOCK_K + c * 4;
            if (kp0 < d && (!into_next || nt == geo.n_tiles - 1)) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (into_next ? next_valid[i] : valid[i]) {
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.a0 + (into_next ? nextA[i] : rowA[i]) + kp0));
                  if (MODE == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.a1 + (into_next ? nextB[i] : rowB[i]) + kp0));
                }
              }
            }
          }
          if (pt == 0) trace_event(p, 2, tcur, 20, tile, kb);  // producer: loads issued, waiting for the stage
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (pt == 0) trace_event(p, 2, tcur, 21, tile, kb);  // producer: stage free
          uint8_t* a_hi = smem + OFF_A_HI + s * A_STAGE_BYTES;
          uint8_t* a_lo = smem + OFF_A_LO + s * A_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 m;
            if (MODE == 0) {
              float4 a = act_fwd4(vb[i], p.act, p.act_param);
              m = make_float4(va[i].x - a.x, va[i].y - a.y, va[i].z - a.z, va[i].w - a.w);
              if (p.m_out != nullptr && nt == 0 && valid[i] && kvalid) stg4(p.m_out + (tile * TILE_M + r0 + 32 * i) * d + k0, m);
            } else {
              m = va[i];
              if (p.drop_p > 0.f && valid[i] && kvalid) {
                float4 sc = dropout_scale4(p.seed, p.offset, (uint64_t)(rowA[i] + k0), p.drop_thr, p.inv_keep);
                m = make_float4(m.x * sc.x, m.y * sc.y, m.z * sc.z, m.w * sc.w);
              }
            }
            float4 hi = make_float4(tf32_rna(m.x), tf32_rna(m.y), tf32_rna(m.z), tf32_rna(m.w));
            float4 lo = make_float4(tf32_rna(m.x - hi.x), tf32_rna(m.y - hi.y), tf32_rna(m.z - hi.z), tf32_rna(m.w - hi.w));
            const uint32_t off = swz128((uint32_t)(r0 + 32 * i), (uint32_t)c);
            *reinterpret_cast<float4*>(a_hi + off) = hi;
            *reinterpret_cast<float4*>(a_lo + off) = lo;
          }
          fence_proxy_async();  // make this thread's generic-proxy writes visible to the tensor core (async proxy)
          __syncwarp();         // ... for every lane of the warp, then ONE arrive per warp (256 arrives per stage serialise)
          if (lane == 0) mbar_arrive(bar_a_full + 8 * s);
          if (pt == 0) trace_event(p, 2, tcur, 22, tile, kb);  // producer: stage written
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// W [d,d] -> hi/lo TF32 parts in the swizzled K-major tile image consumed by the bulk copies above.
__global__ void __launch_bounds__(256) weight_prepare_kernel(const float* __restrict__ W, Geometry geo, int transpose, uint8_t* __restrict__ image) {
  const int64_t total = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * BLOCK_K;
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % BLOCK_K);
  int64_t q = t / BLOCK_K;
  const int r = (int)(q % geo.n_tile);
  q /= geo.n_tile;
  const int kb = (int)(q % geo.k_blocks);
  const int nt = (int)(q / geo.k_blocks);
  const int n = nt * geo.n_tile + r, k = kb * BLOCK_K + kk;
  float v = 0.f;
  if (n < geo.d && k < geo.d) v = transpose ? __ldg(W + (int64_t)k * geo.d + n) : __ldg(W + (int64_t)n * geo.d + k);
  const float hi = tf32_rna(v);
  const float lo = tf32_rna(v - hi);
  const size_t off = ((size_t)nt * geo.k_blocks + kb) * geo.n_tile * 128 + swz128((uint32_t)r, (uint32_t)(kk >> 2)) + (kk & 3) * 4;
  *reinterpret_cast<float*>(image + off) = hi;
  *reinterpret_cast<float*>(image + geo.part_bytes + off) = lo;
}

template <int MODE>
static int launch(const Params& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(layer_gemm_tc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(layer_gemm_tc)");
  const int64_t m_tiles = (p.E + TILE_M - 1) / TILE_M;
  int grid = num_sms();
  if (grid <= 0) grid = 148;
  if (m_tiles < grid) grid = (int)m_tiles;
  layer_gemm_tc<MODE><<<grid, THREADS, SMEM_BYTES, st>>>(p);
  NT_LAUNCH_CHECK("layer_gemm_tc", 1);
  return NT_OK;
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

}  // namespace tc

void tc_set_trace_buffer(void* ptr) { tc::g_trace_buffer = static_cast<unsigned long long*>(ptr); }

size_t tc_weight_image_bytes(int64_t d) { return 2 * tc::make_geometry((int)d).part_bytes; }

int tc_weight_prepare(const float* W, int64_t d, int transpose, void* image, cudaStream_t st) {
  tc::Geometry geo = tc::make_geometry((int)d);
  const int64_t total = (int64_t)geo.n_tiles * geo.k_blocks * geo.n_tile * tc::BLOCK_K;
  tc::weight_prepare_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(W, geo, transpose, static_cast<uint8_t*>(image));
  NT_LAUNCH_CHECK("weight_prepare_kernel", 1);
  return NT_OK;
}

int tc_layer_forward(const float* h, const float* n, const int32_t* src, const int32_t* rev, const void* wimg, const float* bias, int64_t E, int64_t d,
                     int act, float act_param, int residual, float drop_p, uint64_t seed, uint64_t offset, float* out, float* m_out, int products,
                     cudaStream_t st) {
  tc::Params p{};
  p.m_out = m_out;
  p.a0 = n; p.a1 = h; p.src = src; p.rev = rev; p.wimg = static_cast<const uint8_t*>(wimg); p.bias = bias;
  p.resid = residual ? h : nullptr; p.out = out; p.E = E; p.geo = tc::make_geometry((int)d);
  p.act = act; p.act_param = act_param; p.products = products;
  tc::fill_dropout(p, drop_p, seed, offset);
  return tc::launch<0>(p, st);
}

int tc_layer_dgrad(const float* g, const void* wimg, int64_t E, int64_t d, float drop_p, uint64_t seed, uint64_t offset, float* g_m, int products,
                   cudaStream_t st) {
  tc::Params p{};
  p.a0 = g; p.wimg = static_cast<const uint8_t*>(wimg); p.out = g_m; p.E = E; p.geo = tc::make_geometry((int)d);
  p.act = NT_ACT_IDENTITY; p.products = products;
  tc::fill_dropout(p, drop_p, seed, offset);
  return tc::launch<1>(p, st);
}

}  // namespace nt
