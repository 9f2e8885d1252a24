// HBM-bound row kernels: segmented reductions (K1/K3/K5), gather-add (K0 and the backward of the
// reductions) and the backward epilogue of a message-passing depth (K6).
//
// Thread mapping: the [rows, d] output is flattened to (row, 16-byte chunk) work items so every lane
// is busy for any d (d = 300 -> 75 chunks per row) and a warp always touches one contiguous span of
// the output; gathers read whole 16-byte chunks of the source rows. Segments are accumulated
// SEQUENTIALLY in ascending item order (no cross-lane tree over the segment), which makes K1/K3
// bit-identical to the reference's CPU scatter_add_ and run-to-run deterministic
// (the stock GPU path, atomicAdd in scatter_gather_elementwise_kernel, is neither).
#include <stdlib.h>

#include "common.cuh"

namespace nt {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Compile-time activation kind of the segmented reductions: 0 = identity, 1 = ReLU, 2 = the generic six-way switch. With the
// switch (expf / erff / tanhf per element) inlined into every unrolled row the ELL kernel was 97 KB of SASS and instruction
// fetch, not HBM, set its speed; the two hot cases are now ~3 KB each.
template <int AK>
__device__ __forceinline__ float4 seg_act_fwd4(float4 v, int act, float p) {
  if (AK == 0) return v;
  if (AK == 1) return make_float4(v.x < 0.f ? 0.f : v.x, v.y < 0.f ? 0.f : v.y, v.z < 0.f ? 0.f : v.z, v.w < 0.f ? 0.f : v.w);
  return act_fwd4(v, act, p);
}
template <int AK>
__device__ __forceinline__ float seg_act_bwd(float x, int act, float p) {
  if (AK == 0) return 1.f;
  if (AK == 1) return x > 0.f ? 1.f : 0.f;
  return act_bwd(x, act, p);
}

// ------------------------------------------------------------------------------------------------
// seg_reduce, vectorised: one thread per (segment, float4 chunk)
// ------------------------------------------------------------------------------------------------
// Each thread owns SEG_ITEMS (segment, chunk) work items and walks their segments in lock step, two rows per
// item per trip: the rowptr -> perm -> row dependency chain is three DRAM latencies long and segments are short
// (in-degree ~2), so the independent chains of several items are what keeps enough bytes in flight.
constexpr int SEG_ITEMS = 2;

template <bool HAS_PERM, int AK, int U = 2>  // U rows of every item per trip
__global__ void __launch_bounds__(ROW_THREADS, 5) seg_reduce_v4(const float* __restrict__ x, int d, int chunks, const int32_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ perm, int64_t total, uint64_t magic, int act,
                                                              float act_param, int mean, float scale, const float* __restrict__ base,
                                                              const float* __restrict__ dact_of, float* __restrict__ out) {
  // dact_of != nullptr (nt_seg_reduce_ex, backward form): no activation prologue; the reduced row is multiplied by act'(dact_of[s])
  const bool pre = dact_of == nullptr;
  // blocks walk the segments from the END: the producer kernel wrote its output front to back, so the tail is what is still in
  // the 126 MB L2 when this kernel starts (reading front to back would evict it before reaching it)
  const int64_t t0 = (int64_t)(gridDim.x - 1 - blockIdx.x) * (ROW_THREADS * SEG_ITEMS) + threadIdx.x;
  int s[SEG_ITEMS], c[SEG_ITEMS], lo[SEG_ITEMS], hi[SEG_ITEMS];
  float4 acc[SEG_ITEMS];
  int maxlen = 0;
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int64_t t = t0 + (int64_t)k * ROW_THREADS;
    const bool live = t < total;
    split_item(live ? t : 0, chunks, magic, s[k], c[k]);
    c[k] *= 4;
    lo[k] = live ? __ldg(rowptr + s[k]) : 0;
    hi[k] = live ? __ldg(rowptr + s[k] + 1) : 0;
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    maxlen = max(maxlen, hi[k] - lo[k]);
  }
  for (int j = 0; j < maxlen; j += U) {
    int r[SEG_ITEMS][U];
    bool ok[SEG_ITEMS][U];
    float4 v[SEG_ITEMS][U];
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = lo[k] + j + u;
        ok[k][u] = idx < hi[k];
        r[k][u] = ok[k][u] ? (HAS_PERM ? __ldg(perm + idx) : idx) : 0;
      }
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[k][u]) v[k][u] = ldg4(x + (int64_t)r[k][u] * d + c[k]);
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u)  // ascending item order within the segment: bit-identical to a sequential scatter_add_
        if (ok[k][u]) acc[k] = add4(acc[k], pre ? seg_act_fwd4<AK>(v[k][u], act, act_param) : v[k][u]);
  }
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int64_t t = t0 + (int64_t)k * ROW_THREADS;
    if (t >= total) continue;
    float4 a = acc[k];
    if (mean) {
      const float cnt = (float)max(hi[k] - lo[k], 1);  // torch_scatter.scatter_mean: count.clamp(min=1), true division
      a = make_float4(a.x / cnt, a.y / cnt, a.z / cnt, a.w / cnt);
    }
    if (scale != 1.f) a = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
    if (dact_of) {
      const float4 hv = ldg4(dact_of + (int64_t)s[k] * d + c[k]);
      a = make_float4(a.x * seg_act_bwd<AK>(hv.x, act, act_param), a.y * seg_act_bwd<AK>(hv.y, act, act_param),
                      a.z * seg_act_bwd<AK>(hv.z, act, act_param), a.w * seg_act_bwd<AK>(hv.w, act, act_param));
    }
    if (base) a = add4(ldg4_stream(base + (int64_t)s[k] * d + c[k]), a);
    stg4(out + (int64_t)s[k] * d + c[k], a);
  }
}

// ------------------------------------------------------------------------------------------------
// seg_reduce over an ELL copy of the CSR: ell[s] = the first four item ids of segment s (ascending, -1 padded)
// ------------------------------------------------------------------------------------------------
// Molecular graphs have in/out-degree <= 4, so one 16-byte load yields every row index of the segment and the four row
// loads are issued together: the rowptr -> perm -> row chain of seg_reduce_v4 (three dependent DRAM latencies, 60 % of HBM
// peak) becomes the two-level chain of the gather kernels (82-98 %). Segments longer than four continue through rowptr /
// perm. The accumulation order (ascending item id, starting from 0) is unchanged, so results stay bit-identical.
__global__ void __launch_bounds__(ROW_THREADS) csr_to_ell_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int64_t S,
                                                                 int4* __restrict__ ell) {
  int64_t s = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (s >= S) return;
  const int lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
  int v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = lo + j < hi ? (perm ? __ldg(perm + lo + j) : lo + j) : -1;
  ell[s] = make_int4(v[0], v[1], v[2], v[3]);
}

template <int AK>
__global__ void __launch_bounds__(ROW_THREADS, 4) seg_reduce_ell_v4(const float* __restrict__ x, int d, int chunks, const int32_t* __restrict__ rowptr,
                                                                   const int32_t* __restrict__ perm, const int4* __restrict__ ell, int64_t total,
                                                                   uint64_t magic, int act, float act_param, int mean, float scale,
                                                                   const float* __restrict__ base,
                                                                   const float* __restrict__ dact_of, float* __restrict__ out) {
  const bool pre = dact_of == nullptr;
  // blocks walk the segments from the END: the producer kernel wrote its output front to back, so the tail is what is still in
  // the 126 MB L2 when this kernel starts (reading front to back would evict it before reaching it)
  const int64_t t0 = (int64_t)(gridDim.x - 1 - blockIdx.x) * (ROW_THREADS * SEG_ITEMS) + threadIdx.x;
  int s[SEG_ITEMS], c[SEG_ITEMS], lo[SEG_ITEMS], hi[SEG_ITEMS];
  int4 nb[SEG_ITEMS];
  bool live[SEG_ITEMS];
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int64_t t = t0 + (int64_t)k * ROW_THREADS;
    live[k] = t < total;
    split_item(live[k] ? t : 0, chunks, magic, s[k], c[k]);
    c[k] *= 4;
    nb[k] = live[k] ? __ldg(ell + s[k]) : make_int4(-1, -1, -1, -1);
    lo[k] = live[k] ? __ldg(rowptr + s[k]) : 0;
    hi[k] = live[k] ? __ldg(rowptr + s[k] + 1) : 0;
  }
  float4 v[SEG_ITEMS][4];
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int id[4] = {nb[k].x, nb[k].y, nb[k].z, nb[k].w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (id[u] >= 0) v[k][u] = ldg4(x + (int64_t)id[u] * d + c[k]);
  }
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    if (!live[k]) continue;
    const int id[4] = {nb[k].x, nb[k].y, nb[k].z, nb[k].w};
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)  // ascending item order within the segment: bit-identical to a sequential scatter_add_
      if (id[u] >= 0) a = add4(a, pre ? seg_act_fwd4<AK>(v[k][u], act, act_param) : v[k][u]);
    for (int j = lo[k] + 4; j < hi[k]; ++j) {  // degree > 4: the rest of the segment through the CSR
      const int rr = perm ? __ldg(perm + j) : j;
      const float4 w = ldg4(x + (int64_t)rr * d + c[k]);
      a = add4(a, pre ? seg_act_fwd4<AK>(w, act, act_param) : w);
    }
    if (mean) {
      const float cnt = (float)max(hi[k] - lo[k], 1);
      a = make_float4(a.x / cnt, a.y / cnt, a.z / cnt, a.w / cnt);
    }
    if (scale != 1.f) a = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
    if (dact_of) {
      const float4 hv = ldg4(dact_of + (int64_t)s[k] * d + c[k]);
      a = make_float4(a.x * seg_act_bwd<AK>(hv.x, act, act_param), a.y * seg_act_bwd<AK>(hv.y, act, act_param),
                      a.z * seg_act_bwd<AK>(hv.z, act, act_param), a.w * seg_act_bwd<AK>(hv.w, act, act_param));
    }
    if (base) a = add4(ldg4_stream(base + (int64_t)s[k] * d + c[k]), a);
    stg4(out + (int64_t)s[k] * d + c[k], a);
  }
}

// A tiled form (a CTA owns 32 consecutive segments, the ELL record is read once into registers, the output row is swept in passes of
// eight 16-byte chunks with eight row loads in flight per thread) was measured and NOT kept: 86.8 us against 86.9 us for the per-item
// form above, bit-identical - the kernel is not bound by its index chain. Both deliver 4.15 TB/s (0.36 GB): what a PERMUTED walk over
// 1200-byte rows gets from HBM3e (a sequential stream of the same bytes: 6.2-6.5 TB/s, e.g. nt_gather_add in its K1-backward form).

// scalar fallback for d % 4 != 0 (or unaligned bases): one thread per (segment, element)
__global__ void __launch_bounds__(ROW_THREADS) seg_reduce_s(const float* __restrict__ x, int d, const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ perm, int64_t total, int act, float act_param, int mean,
                                                             float scale, const float* __restrict__ base, const float* __restrict__ dact_of,
                                                             float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  int s = (int)(t / d);
  int c = (int)(t - (int64_t)s * d);
  int lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
  const int pre_act = dact_of ? NT_ACT_IDENTITY : act;
  float acc = 0.f;
  for (int j = lo; j < hi; ++j) {
    int r = perm ? __ldg(perm + j) : j;
    acc += act_fwd(__ldg(x + (int64_t)r * d + c), pre_act, act_param);
  }
  if (mean) acc = acc / (float)max(hi - lo, 1);
  if (scale != 1.f) acc *= scale;
  if (dact_of) acc *= act_bwd(__ldg(dact_of + t), act, act_param);
  if (base) acc = __ldg(base + t) + acc;
  out[(int64_t)s * d + c] = acc;
}

// ------------------------------------------------------------------------------------------------
// gather_add: out[i] = base[i] + scale * x[idx[i]] / max(count(idx[i]), 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) gather_add_v4(const float* __restrict__ base, const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                              const int32_t* __restrict__ mean_rowptr, int d, int chunks, int64_t total,
                                                              uint64_t magic, float scale, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  int i, c;
  split_item(t, chunks, magic, i, c);
  c *= 4;
  int r = __ldg(idx + i);
  float4 v = ldg4(x + (int64_t)r * d + c);
  if (mean_rowptr) {
    float cnt = (float)max(__ldg(mean_rowptr + r + 1) - __ldg(mean_rowptr + r), 1);
    v = make_float4(v.x / cnt, v.y / cnt, v.z / cnt, v.w / cnt);
  }
  if (scale != 1.f) v = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  if (base) v = add4(ldg4_stream(base + (int64_t)i * d + c), v);
  stg4(out + (int64_t)i * d + c, v);
}

__global__ void __launch_bounds__(ROW_THREADS) gather_add_s(const float* __restrict__ base, const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                             const int32_t* __restrict__ mean_rowptr, int d, int64_t total, float scale,
                                                             float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  int i = (int)(t / d);
  int c = (int)(t - (int64_t)i * d);
  int r = __ldg(idx + i);
  float v = __ldg(x + (int64_t)r * d + c);
  if (mean_rowptr) v = v / (float)max(__ldg(mean_rowptr + r + 1) - __ldg(mean_rowptr + r), 1);
  if (scale != 1.f) v *= scale;
  if (base) v = __ldg(base + t) + v;
  out[t] = v;
}

// ------------------------------------------------------------------------------------------------
// K6 backward epilogue
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(ROW_THREADS) layer_bwd_epilogue(const float* __restrict__ g, const float* __restrict__ h, const float* __restrict__ g_n,
                                                                   const float* __restrict__ g_m, const int32_t* __restrict__ dst,
                                                                   const int32_t* __restrict__ rev_rowptr, const int32_t* __restrict__ rev_perm,
                                                                   const int32_t* __restrict__ dst_rowptr, const int32_t* __restrict__ arg, int d,
                                                                   int chunks, int64_t total, int act, float act_param, int residual, int mean,
                                                                   float* __restrict__ g_h) {
  // arg != NULL: the forward reduction was a max / min; only the edge that supplied the extreme of (atom, channel) receives g_n
  int64_t t = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  int e = (int)(t / chunks);
  int c = (int)(t - (int64_t)e * chunks) * (VEC ? 4 : 1);
  int v = __ldg(dst + e);
  int lo = __ldg(rev_rowptr + e), hi = __ldg(rev_rowptr + e + 1);
  float inv_cnt_div = 1.f;
  if (mean) inv_cnt_div = (float)max(__ldg(dst_rowptr + v + 1) - __ldg(dst_rowptr + v), 1);
  if (VEC) {
    float4 ga = ldg4(g_n + (int64_t)v * d + c);
    if (mean) ga = make_float4(ga.x / inv_cnt_div, ga.y / inv_cnt_div, ga.z / inv_cnt_div, ga.w / inv_cnt_div);
    if (arg) {
      const int4 w = __ldg(reinterpret_cast<const int4*>(arg + (int64_t)v * d + c));
      ga = make_float4(w.x == e ? ga.x : 0.f, w.y == e ? ga.y : 0.f, w.z == e ? ga.z : 0.f, w.w == e ? ga.w : 0.f);
    }
    float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = lo; j < hi; ++j) sub = add4(sub, ldg4(g_m + (int64_t)__ldg(rev_perm + j) * d + c));
    float4 hv = ldg4_stream(h + (int64_t)e * d + c);
    float4 r = make_float4(act_bwd(hv.x, act, act_param) * (ga.x - sub.x), act_bwd(hv.y, act, act_param) * (ga.y - sub.y),
                           act_bwd(hv.z, act, act_param) * (ga.z - sub.z), act_bwd(hv.w, act, act_param) * (ga.w - sub.w));
    if (residual) r = add4(ldg4_stream(g + (int64_t)e * d + c), r);
    stg4(g_h + (int64_t)e * d + c, r);
  } else {
    float ga = __ldg(g_n + (int64_t)v * d + c);
    if (mean) ga = ga / inv_cnt_div;
    if (arg && __ldg(arg + (int64_t)v * d + c) != e) ga = 0.f;
    float sub = 0.f;
    for (int j = lo; j < hi; ++j) sub += __ldg(g_m + (int64_t)__ldg(rev_perm + j) * d + c);
    float r = act_bwd(__ldg(h + (int64_t)e * d + c), act, act_param) * (ga - sub);
    if (residual) r = __ldg(g + (int64_t)e * d + c) + r;
    g_h[(int64_t)e * d + c] = r;
  }
}

// K5 + K6 fused: instead of reading a precomputed g_n[dst[e]] (K5 = segmented sum of g_m over the outgoing edges of every atom,
// an [E,d] read and a [V,d] write per depth), every edge sums the g_m rows of the outgoing edges of ITS destination atom itself,
// through the ELL copy of the by-source CSR (same ascending order, so the value is bit-identical to K5's). The ~2.2 extra rows
// per edge are rows its neighbours in the same molecule read too: they come from L2 / L1, not from DRAM.
__device__ __forceinline__ int ldg_i32_v(const int32_t* p) {
  int r;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_i4_v(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg4_v(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// Dependency depth is what bounds this kernel (ncu: 83 % of the warp samples wait on a long scoreboard at full occupancy): the
// first version walked dst[e] -> ell[v] -> rows -> sum and only THEN rev_rowptr[e] -> rev_perm[j] -> row (a loop with an
// unknown trip count, which the compiler cannot hoist) -> g[e]: five dependent memory round trips. Here every index of level
// two (the ELL record of the destination atom AND the first inverse-rev edge) is requested together, then every row of level
// three (<= 4 out-edge rows, the first inverse-rev row, h[e], g[e]) - three round trips. Longer segments (degree > 4, more than
// one inverse-rev edge: the reference's atom-offset rev_index quirk) continue through the CSR as before. REVERSED: blocks walk
// the edges from the end, so that the tail of g_m - what K4a wrote last and what is still in L2 - is read first.
template <int AK>
__global__ void __launch_bounds__(ROW_THREADS, 5) layer_bwd_epilogue_fused(const float* __restrict__ g, const float* __restrict__ h,
                                                                         const float* __restrict__ g_m, const int32_t* __restrict__ dst,
                                                                         const int32_t* __restrict__ src_rowptr, const int32_t* __restrict__ src_perm,
                                                                         const int4* __restrict__ src_ell, const int32_t* __restrict__ rev_rowptr,
                                                                         const int32_t* __restrict__ rev_perm, const int32_t* __restrict__ dst_rowptr,
                                                                         int d, int chunks, int64_t total, uint64_t magic, int act, float act_param,
                                                                         int residual, int mean, int reversed, float* __restrict__ g_h) {
  const int64_t blk = reversed ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x;
  const int64_t t = blk * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  int e, c;
  split_item(t, chunks, magic, e, c);
  c *= 4;
  // volatile loads: the compiler otherwise sinks the independent ones (h, g, src_rowptr) below the first use of a gathered row
  // level 1
  const int v = ldg_i32_v(dst + e);
  const int lo = ldg_i32_v(rev_rowptr + e), hi = ldg_i32_v(rev_rowptr + e + 1);
  const int64_t own = (int64_t)e * d + c;
  const float4 hv = ldg4_stream(h + own);
  float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (residual) gv = ldg4_stream(g + own);
  // level 2
  const int4 nb = ldg_i4_v(src_ell + v);
  const int r0 = lo < hi ? ldg_i32_v(rev_perm + lo) : -1;
  const int slo = ldg_i32_v(src_rowptr + v), shi = ldg_i32_v(src_rowptr + v + 1);
  float cnt = 1.f;
  if (mean) cnt = (float)max(ldg_i32_v(dst_rowptr + v + 1) - ldg_i32_v(dst_rowptr + v), 1);
  // level 3
  const int id[4] = {nb.x, nb.y, nb.z, nb.w};
  float4 row[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (id[u] >= 0) row[u] = ldg4_v(g_m + (int64_t)id[u] * d + c);
  float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r0 >= 0) sub = ldg4_v(g_m + (int64_t)r0 * d + c);  // 0 + x == x: same bits as the accumulation from zero
  __syncwarp(__activemask());  // scheduling fence: every load above is issued before the first use below
  float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (id[u] >= 0) ga = add4(ga, row[u]);
  for (int j = slo + 4; j < shi; ++j) ga = add4(ga, ldg4(g_m + (int64_t)__ldg(src_perm + j) * d + c));
  if (mean) ga = make_float4(ga.x / cnt, ga.y / cnt, ga.z / cnt, ga.w / cnt);
  for (int j = lo + 1; j < hi; ++j) sub = add4(sub, ldg4(g_m + (int64_t)__ldg(rev_perm + j) * d + c));
  float4 r = make_float4(seg_act_bwd<AK>(hv.x, act, act_param) * (ga.x - sub.x), seg_act_bwd<AK>(hv.y, act, act_param) * (ga.y - sub.y),
                         seg_act_bwd<AK>(hv.z, act, act_param) * (ga.z - sub.z), seg_act_bwd<AK>(hv.w, act, act_param) * (ga.w - sub.w));
  if (residual) r = add4(gv, r);
  stg4(g_h + own, r);
}

// A tiled form of this kernel (a CTA owns 16 / 32 consecutive edges, the index data of an edge is read once into registers, the row
// is swept in passes of 16-byte chunk groups - what made the pooled epilogue of pooled_backward.cu and the tiled K1 below faster) was
// measured and NOT kept: 244 us with one pass in flight (this form: 247), 326-329 us with two (80 registers, three CTAs per SM).
// ncu of this form: 240 us, 1.41 GB of L2 -> L1 sectors + 0.25 GB of stores = 6.9 TB/s, i.e. the L2 -> SM ceiling of §5.0 (L1 hit
// rate 9.6 %: the ~2.2 out-edge rows an edge gathers are its sibling in-edges' rows too, but siblings run on other SMs and a tile
// did not bring them together in time). A per-molecule form (one CTA per molecule of a device-collated batch: its g_m rows staged once
// in shared memory - a sequential read - and the out-edge sums served from there, bit-identical) measured 408 us: copy, barrier and
// sweep of a 50-edge molecule serialise inside a CTA and three such CTAs per SM keep too few loads in flight. Not kept either.

// the first form (kept for A/B timing: NOTORCH_B200_K6_VARIANT=0)
template <int AK>
__global__ void __launch_bounds__(ROW_THREADS) layer_bwd_epilogue_fused_v0(const float* __restrict__ g, const float* __restrict__ h,
                                                                            const float* __restrict__ g_m, const int32_t* __restrict__ dst,
                                                                            const int32_t* __restrict__ src_rowptr, const int32_t* __restrict__ src_perm,
                                                                            const int4* __restrict__ src_ell, const int32_t* __restrict__ rev_rowptr,
                                                                            const int32_t* __restrict__ rev_perm, const int32_t* __restrict__ dst_rowptr,
                                                                            int d, int chunks, int64_t total, uint64_t /*magic*/, int act,
                                                                            float act_param, int residual, int mean, int reversed,
                                                                            float* __restrict__ g_h) {
  const int64_t blk = reversed ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x;
  const int64_t t = blk * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  const int e = (int)(t / chunks);
  const int c = (int)(t - (int64_t)e * chunks) * 4;
  const int v = __ldg(dst + e);
  const int lo = __ldg(rev_rowptr + e), hi = __ldg(rev_rowptr + e + 1);
  const int4 nb = __ldg(src_ell + v);
  const int slo = __ldg(src_rowptr + v), shi = __ldg(src_rowptr + v + 1);
  const int id[4] = {nb.x, nb.y, nb.z, nb.w};
  float4 row[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (id[u] >= 0) row[u] = ldg4(g_m + (int64_t)id[u] * d + c);
  const float4 hv = ldg4_stream(h + (int64_t)e * d + c);
  float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (id[u] >= 0) ga = add4(ga, row[u]);
  for (int j = slo + 4; j < shi; ++j) ga = add4(ga, ldg4(g_m + (int64_t)__ldg(src_perm + j) * d + c));
  if (mean) {
    const float cnt = (float)max(__ldg(dst_rowptr + v + 1) - __ldg(dst_rowptr + v), 1);
    ga = make_float4(ga.x / cnt, ga.y / cnt, ga.z / cnt, ga.w / cnt);
  }
  float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = lo; j < hi; ++j) sub = add4(sub, ldg4(g_m + (int64_t)__ldg(rev_perm + j) * d + c));
  float4 r = make_float4(seg_act_bwd<AK>(hv.x, act, act_param) * (ga.x - sub.x), seg_act_bwd<AK>(hv.y, act, act_param) * (ga.y - sub.y),
                         seg_act_bwd<AK>(hv.z, act, act_param) * (ga.z - sub.z), seg_act_bwd<AK>(hv.w, act, act_param) * (ga.w - sub.w));
  if (residual) r = add4(ldg4_stream(g + (int64_t)e * d + c), r);
  stg4(g_h + (int64_t)e * d + c, r);
}

__global__ void __launch_bounds__(ROW_THREADS) dropout_mask_kernel(int64_t total, float p, uint64_t seed, uint64_t offset, float* __restrict__ mask) {
  int64_t t = (int64_t)blockIdx.x * ROW_THREADS + threadIdx.x;
  if (t >= total) return;
  mask[t] = p > 0.f ? (dropout_scale1(seed, offset, (uint64_t)t, dropout_threshold(p), 1.f)) : 1.f;
}

}  // namespace nt

using namespace nt;

static bool vec_ok(int64_t d, const void* a, const void* b = nullptr, const void* c = nullptr, const void* e = nullptr, const void* f = nullptr) {
  return d % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(e) && aligned16(f);
}

static int seg_reduce_impl(const char* fn, const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, const int32_t* ell,
                           int64_t num_segments, int act, float act_param, int mean, float scale, const void* base, const void* dact_of, void* out,
                           int dtype, nt_stream_t stream) {
  NT_CHECK_ARG(dtype == NT_F32 || dtype == NT_BF16, "%s: bad dtype", fn);
  if (dtype != NT_F32) { set_error("%s: only NT_F32 is implemented", fn); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && num_segments >= 0 && num_segments < INT32_MAX, "%s: bad sizes", fn);
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "%s: bad activation", fn);
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(rowptr && out, "%s: null pointer", fn);
  cudaStream_t st = as_stream(stream);
  const float* xf = static_cast<const float*>(x);
  const float* bf = static_cast<const float*>(base);
  const float* df = static_cast<const float*>(dact_of);
  float* of = static_cast<float*>(out);
  if (vec_ok(d, x, out, base, dact_of)) {
    int chunks = (int)(d / 4);
    int64_t total = num_segments * chunks;
    unsigned grid = (unsigned)cdiv(total, ROW_THREADS * SEG_ITEMS);
    const uint64_t magic = chunk_div_magic(total, chunks);
    const int ak = act == NT_ACT_IDENTITY ? 0 : act == NT_ACT_RELU ? 1 : 2;
#define NT_SEG_LAUNCH(AK)                                                                                                                        \
  do {                                                                                                                                           \
    if (ell && aligned16(ell))                                                                                                                   \
      seg_reduce_ell_v4<AK><<<grid, ROW_THREADS, 0, st>>>(xf, (int)d, chunks, rowptr, perm, reinterpret_cast<const int4*>(ell), total, magic,    \
                                                          act, act_param, mean, scale, bf, df, of);                                             \
    else if (perm)                                                                                                                               \
      seg_reduce_v4<true, AK><<<grid, ROW_THREADS, 0, st>>>(xf, (int)d, chunks, rowptr, perm, total, magic, act, act_param, mean, scale, bf, df, of); \
    else                                                                                                                                         \
      seg_reduce_v4<false, AK><<<grid, ROW_THREADS, 0, st>>>(xf, (int)d, chunks, rowptr, perm, total, magic, act, act_param, mean, scale, bf, df, of); \
  } while (0)
    if (ak == 0) NT_SEG_LAUNCH(0);
    else if (ak == 1) NT_SEG_LAUNCH(1);
    else NT_SEG_LAUNCH(2);
#undef NT_SEG_LAUNCH
  } else {
    int64_t total = num_segments * d;
    seg_reduce_s<<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, st>>>(xf, (int)d, rowptr, perm, total, act, act_param, mean, scale, bf, df, of);
  }
  NT_LAUNCH_CHECK(fn, 1);
  return NT_OK;
}

extern "C" int nt_seg_reduce(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int act, float act_param,
                             int mean, float scale, void* out, int dtype, nt_stream_t stream) {
  return seg_reduce_impl("nt_seg_reduce", x, d, rowptr, perm, nullptr, num_segments, act, act_param, mean, scale, nullptr, nullptr, out, dtype, stream);
}

extern "C" int nt_seg_reduce_ex(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int act, float act_param,
                                int mean, float scale, const void* base, const void* dact_of, void* out, int dtype, nt_stream_t stream) {
  return seg_reduce_impl("nt_seg_reduce_ex", x, d, rowptr, perm, nullptr, num_segments, act, act_param, mean, scale, base, dact_of, out, dtype, stream);
}

extern "C" int nt_seg_reduce_ell(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, const int32_t* ell, int64_t num_segments, int act,
                                 float act_param, int mean, float scale, const void* base, const void* dact_of, void* out, int dtype,
                                 nt_stream_t stream) {
  NT_CHECK_ARG(ell, "nt_seg_reduce_ell: null ell");
  return seg_reduce_impl("nt_seg_reduce_ell", x, d, rowptr, perm, ell, num_segments, act, act_param, mean, scale, base, dact_of, out, dtype, stream);
}

extern "C" int nt_csr_to_ell(const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int32_t* ell, nt_stream_t stream) {
  NT_CHECK_ARG(num_segments >= 0 && num_segments < INT32_MAX, "nt_csr_to_ell: bad sizes");
  if (num_segments == 0) return NT_OK;
  NT_CHECK_ARG(rowptr && ell && aligned16(ell), "nt_csr_to_ell: null or unaligned pointer");
  csr_to_ell_kernel<<<(unsigned)cdiv(num_segments, ROW_THREADS), ROW_THREADS, 0, as_stream(stream)>>>(rowptr, perm, num_segments,
                                                                                                   reinterpret_cast<int4*>(ell));
  NT_LAUNCH_CHECK("nt_csr_to_ell", 1);
  return NT_OK;
}

extern "C" int nt_gather_add(const void* base, const void* x, const int32_t* idx, const int32_t* mean_rowptr, int64_t n, int64_t d, float scale,
                             void* out, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_gather_add: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && n >= 0 && n < INT32_MAX, "nt_gather_add: bad sizes");
  if (n == 0) return NT_OK;
  NT_CHECK_ARG(x && idx && out, "nt_gather_add: null pointer");
  cudaStream_t st = as_stream(stream);
  if (vec_ok(d, x, out, base)) {
    int chunks = (int)(d / 4);
    int64_t total = n * chunks;
    gather_add_v4<<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, st>>>(static_cast<const float*>(base), static_cast<const float*>(x), idx,
                                                                              mean_rowptr, (int)d, chunks, total, chunk_div_magic(total, chunks), scale,
                                                                              static_cast<float*>(out));
  } else {
    int64_t total = n * d;
    gather_add_s<<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, st>>>(static_cast<const float*>(base), static_cast<const float*>(x), idx,
                                                                             mean_rowptr, (int)d, total, scale, static_cast<float*>(out));
  }
  NT_LAUNCH_CHECK("nt_gather_add", 1);
  return NT_OK;
}

static int bwd_epilogue_impl(const char* fn, const void* g, const void* h, const void* g_n, const void* g_m, const int32_t* dst,
                             const int32_t* rev_rowptr, const int32_t* rev_perm, const int32_t* dst_rowptr, const int32_t* arg, int64_t E, int64_t d,
                             int act, float act_param, int residual, int mean, void* g_h, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("%s: only NT_F32 is implemented", fn); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && E >= 0 && E < INT32_MAX, "%s: bad sizes", fn);
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "%s: bad activation", fn);
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(h && g_n && g_m && dst && rev_rowptr && rev_perm && g_h, "%s: null pointer", fn);
  NT_CHECK_ARG(!residual || g, "%s: residual needs g", fn);
  NT_CHECK_ARG(!mean || dst_rowptr, "%s: mean needs dst_rowptr", fn);
  cudaStream_t st = as_stream(stream);
  const float *gf = static_cast<const float*>(g), *hf = static_cast<const float*>(h), *gn = static_cast<const float*>(g_n),
              *gm = static_cast<const float*>(g_m);
  float* out = static_cast<float*>(g_h);
  if (vec_ok(d, g, h, g_n, g_m, g_h) && aligned16(arg)) {
    int chunks = (int)(d / 4);
    int64_t total = E * chunks;
    layer_bwd_epilogue<true><<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, st>>>(gf, hf, gn, gm, dst, rev_rowptr, rev_perm, dst_rowptr, arg,
                                                                                         (int)d, chunks, total, act, act_param, residual, mean, out);
  } else {
    int64_t total = E * d;
    layer_bwd_epilogue<false><<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, st>>>(gf, hf, gn, gm, dst, rev_rowptr, rev_perm, dst_rowptr, arg,
                                                                                          (int)d, (int)d, total, act, act_param, residual, mean, out);
  }
  NT_LAUNCH_CHECK(fn, 1);
  return NT_OK;
}

extern "C" int nt_layer_backward_epilogue(const void* g, const void* h, const void* g_n, const void* g_m, const int32_t* dst,
                                          const int32_t* rev_rowptr, const int32_t* rev_perm, const int32_t* dst_rowptr, int64_t E, int64_t d,
                                          int act, float act_param, int residual, int mean, void* g_h, int dtype, nt_stream_t stream) {
  return bwd_epilogue_impl("nt_layer_backward_epilogue", g, h, g_n, g_m, dst, rev_rowptr, rev_perm, dst_rowptr, nullptr, E, d, act, act_param, residual,
                           mean, g_h, dtype, stream);
}

extern "C" int nt_layer_backward_epilogue_arg(const void* g, const void* h, const void* g_n, const void* g_m, const int32_t* dst, const int32_t* arg,
                                              const int32_t* rev_rowptr, const int32_t* rev_perm, int64_t E, int64_t d, int act, float act_param,
                                              int residual, void* g_h, int dtype, nt_stream_t stream) {
  NT_CHECK_ARG(arg || E == 0, "nt_layer_backward_epilogue_arg: null pointer");
  return bwd_epilogue_impl("nt_layer_backward_epilogue_arg", g, h, g_n, g_m, dst, rev_rowptr, rev_perm, nullptr, arg, E, d, act, act_param, residual, 0,
                           g_h, dtype, stream);
}

extern "C" int nt_layer_backward_epilogue_fused(const void* g, const void* h, const void* g_m, const int32_t* dst, const int32_t* src_rowptr,
                                                const int32_t* src_perm, const int32_t* src_ell, const int32_t* rev_rowptr, const int32_t* rev_perm,
                                                const int32_t* dst_rowptr, int64_t E, int64_t d, int act, float act_param, int residual, int mean,
                                                void* g_h, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_layer_backward_epilogue_fused: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && E >= 0 && E < INT32_MAX, "nt_layer_backward_epilogue_fused: bad sizes");
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "nt_layer_backward_epilogue_fused: bad activation");
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(h && g_m && dst && src_rowptr && src_perm && src_ell && rev_rowptr && rev_perm && g_h, "nt_layer_backward_epilogue_fused: null pointer");
  NT_CHECK_ARG(!residual || g, "nt_layer_backward_epilogue_fused: residual needs g");
  NT_CHECK_ARG(!mean || dst_rowptr, "nt_layer_backward_epilogue_fused: mean needs dst_rowptr");
  if (!vec_ok(d, g, h, g_m, g_h, src_ell)) {
    set_error("nt_layer_backward_epilogue_fused: needs d %% 4 == 0 and 16-byte aligned rows (use nt_seg_reduce + nt_layer_backward_epilogue)");
    return NT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  const float *gf = static_cast<const float*>(g), *hf = static_cast<const float*>(h), *gm = static_cast<const float*>(g_m);
  float* out = static_cast<float*>(g_h);
  const int chunks = (int)(d / 4);
  const int64_t total = E * chunks;
  const unsigned grid = (unsigned)cdiv(total, ROW_THREADS);
  const int4* ell4 = reinterpret_cast<const int4*>(src_ell);
  // NOTORCH_B200_K6_VARIANT (A/B timing; read per call): bit 0 = hoisted index loads (default on), bit 1 = reversed block order (default on)
  const char* ve = getenv("NOTORCH_B200_K6_VARIANT");
  const int variant = ve ? atoi(ve) : 3;
  const int reversed = (variant >> 1) & 1;
  const int ak = act == NT_ACT_IDENTITY ? 0 : act == NT_ACT_RELU ? 1 : 2;
  const uint64_t magic = chunk_div_magic(total, chunks);
#define NT_K6_LAUNCH(KERNEL, AK)                                                                                                              \
  KERNEL<AK><<<grid, ROW_THREADS, 0, st>>>(gf, hf, gm, dst, src_rowptr, src_perm, ell4, rev_rowptr, rev_perm, dst_rowptr, (int)d, chunks, total, magic, \
                                           act, act_param, residual, mean, reversed, out)
  if (variant & 1) {
    if (ak == 0) NT_K6_LAUNCH(layer_bwd_epilogue_fused, 0);
    else if (ak == 1) NT_K6_LAUNCH(layer_bwd_epilogue_fused, 1);
    else NT_K6_LAUNCH(layer_bwd_epilogue_fused, 2);
  } else {
    if (ak == 0) NT_K6_LAUNCH(layer_bwd_epilogue_fused_v0, 0);
    else if (ak == 1) NT_K6_LAUNCH(layer_bwd_epilogue_fused_v0, 1);
    else NT_K6_LAUNCH(layer_bwd_epilogue_fused_v0, 2);
  }
#undef NT_K6_LAUNCH
  NT_LAUNCH_CHECK("nt_layer_backward_epilogue_fused", 1);
  return NT_OK;
}

extern "C" int nt_dropout_mask(int64_t n_rows, int64_t d, float dropout_p, uint64_t seed, uint64_t offset, float* mask, nt_stream_t stream) {
  NT_CHECK_ARG(n_rows >= 0 && d > 0 && dropout_p >= 0.f && dropout_p < 1.f, "nt_dropout_mask: bad arguments");
  int64_t total = n_rows * d;
  if (total == 0) return NT_OK;
  NT_CHECK_ARG(mask, "nt_dropout_mask: null pointer");
  dropout_mask_kernel<<<(unsigned)cdiv(total, ROW_THREADS), ROW_THREADS, 0, as_stream(stream)>>>(total, dropout_p, seed, offset, mask);
  NT_LAUNCH_CHECK("nt_dropout_mask", 1);
  return NT_OK;
}
