// Backward of the LAST message-passing depth when the only consumer of h_L is a sum read-out over each molecule's edges
// (ChempropBlock -> agg.Sum / Mean / Norm on a device-collated batch).
//
// Reference arithmetic: notorch/nn/gnn/agg.py:27,36 (H = scatter(node_out, batch_node_index)), chemprop.py:86 (node_out =
// scatter(h_L, dst)) and chemprop.py:37-41 / residual.py:28 (the depth). The gradient that reaches h_L is then a BROADCAST,
//     g[e, :] = G[mol(e), :],        G = dLoss/dH_sum  [B, d],
// and every contraction of the depth's backward over the E edges collapses to one over the B molecules:
//     gW^T = m^T g            = M^T G,            M[b, :] = sum_{e in b} m[e, :]          (one segmented pass over m, then K4b on B rows)
//     gb   = sum_e g[e, :]    = sum_b |b| G[b, :]                                          (nt_weighted_colsum below)
//     g_m  = g W              = (G W)[mol(e), :]                                           (K4a on B rows; g_m [E, d] is never written)
//     g_n[v] = sum_{e: src[e] = v} g_m[e] = outdeg(v) (G W)[mol(v)]                        (a molecule's edges connect its own atoms)
//     g_h[e] = [G[mol e]] + act'(h[e]) * (g_n[dst e] (/ indeg) - sum_{e'': rev[e''] = e} (G W)[mol e''])      (kernel below)
// The [E, d] tensors g, g_m and g_n are never materialised.
//
// The same collapse runs FORWARD (nt_pooled_message_sum below): what the read-out needs of the last depth is
//     H_sum[b] = sum_{e in b} h_L[e] = sum_{e in b} h[e] + |b| bias + (sum_{e in b} m[e]) W^T,        h = h_{L-1}
//     sum_{e in b} m[e] = sum_{e in b} (n[src e] - a[rev e]) = sum_{e in b} (outdeg(dst e) [/ indeg(dst e)] a[e] - a[rev e]),   a = act(h)
// (sum_{e in b} n[src e] = sum_{v in b} outdeg(v) n[v] and n[v] = sum_{e: dst e = v} a[e] [/ indeg v]; every atom and edge involved
// belongs to b). One pass over h with its rev gather gives M = sum m and S = sum h + |b| bias as [B, d] matrices, the Linear runs
// on B rows (nt_dense_forward) - K1, K2 and the read-out's pass over h_L are not launched for the last depth, h_L and m_L are not
// written, and the backward reuses M. h_L itself stays available: ChempropBlock computes it (the dense depth) only if it is read.
// Exact algebra, another summation order (tested against the dense path and the fp64 oracle).
#include <stdlib.h>

#include "common.cuh"

namespace nt {

constexpr int PB_THREADS = 256;

// out[c] = sum_r w_r * x[r, c], w_r = rowptr[r + 1] - rowptr[r] (rowptr == NULL: 1). One CTA per 4-column group: thread t adds rows
// t, t + 256, ... in ascending order, then a fixed-shape tree over the 256 partial sums - deterministic, no atomics.
__global__ void __launch_bounds__(PB_THREADS) weighted_colsum_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr, int64_t rows, int d,
                                                                     float* __restrict__ out) {
  __shared__ float4 part[PB_THREADS];
  const int c = blockIdx.x * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = threadIdx.x; r < rows; r += PB_THREADS) {
    const float w = rowptr ? (float)(__ldg(rowptr + r + 1) - __ldg(rowptr + r)) : 1.f;
    const float4 v = ldg4(x + r * d + c);
    s = make_float4(s.x + w * v.x, s.y + w * v.y, s.z + w * v.z, s.w + w * v.w);
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int stride = PB_THREADS / 2; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride) {
      const float4 a = part[threadIdx.x], b = part[threadIdx.x + stride];
      part[threadIdx.x] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) stg4(out + c, part[0]);
}

__device__ __forceinline__ float pb_act_bwd(int ak, float x, int act, float p) {
  if (ak == 0) return 1.f;
  if (ak == 1) return x > 0.f ? 1.f : 0.f;
  return act_bwd(x, act, p);
}

// Per-edge record for the epilogue below, built once per launch (16 bytes per edge, ~1 % of the epilogue's own traffic):
//   x = molecule of the edge, y = bits of the scale outdeg(dst e) [/ indeg(dst e)], z = molecule of the FIRST edge whose rev points
//   here (-1: none), w = how many do. Without it the epilogue walks mol[e] -> row and rev_rowptr -> rev_perm -> mol -> row: four
//   dependent L2 round trips per thread, and at full occupancy that latency - not HBM - set its speed (174 us for 0.49 GB).
__global__ void __launch_bounds__(PB_THREADS) pooled_record_kernel(const int32_t* __restrict__ mol, const int32_t* __restrict__ dst,
                                                                   const int32_t* __restrict__ src_rowptr, const int32_t* __restrict__ rev_rowptr,
                                                                   const int32_t* __restrict__ rev_perm, const int32_t* __restrict__ dst_rowptr, int64_t E,
                                                                   int mean, int4* __restrict__ rec) {
  const int64_t e = (int64_t)blockIdx.x * PB_THREADS + threadIdx.x;
  if (e >= E) return;
  const int v = __ldg(dst + e);
  const int lo = __ldg(rev_rowptr + e), hi = __ldg(rev_rowptr + e + 1);
  float scale = (float)(__ldg(src_rowptr + v + 1) - __ldg(src_rowptr + v));  // outdeg(dst[e]) identical rows of g_m
  if (mean) scale = scale / (float)max(__ldg(dst_rowptr + v + 1) - __ldg(dst_rowptr + v), 1);
  const int first = lo < hi ? __ldg(mol + __ldg(rev_perm + lo)) : -1;
  rec[e] = make_int4(__ldg(mol + e), __float_as_int(scale), first, hi - lo);
}

// Per-item form (NOTORCH_B200_K6P_ITEMS=1; the tiled form below is the default): one (edge, 16-byte chunk) item per thread. DRAM
// traffic: h[e] in, g_h[e] out; G and GW ([B, d], a few MB) stay in L2 / L1. Measured (configs[1], 0.49 GB): 168-171 us; two items
// per thread 161 us, four (102 registers, two CTAs per SM) 243-246 us.
template <int AK, int PB_ITEMS>
__global__ void __launch_bounds__(PB_THREADS, 5) layer_bwd_epilogue_pooled(const float* __restrict__ G, const float* __restrict__ GW,
                                                                           const float* __restrict__ h, const int4* __restrict__ rec,
                                                                           const int32_t* __restrict__ mol, const int32_t* __restrict__ rev_rowptr,
                                                                           const int32_t* __restrict__ rev_perm, int d, int chunks, int64_t total,
                                                                           uint64_t magic, int act, float act_param, int residual,
                                                                           float* __restrict__ g_h) {
  const int64_t t0 = (int64_t)blockIdx.x * (PB_THREADS * PB_ITEMS) + threadIdx.x;
  int e[PB_ITEMS], c[PB_ITEMS];
  bool live[PB_ITEMS];
  float4 hv[PB_ITEMS], gw[PB_ITEMS], gv[PB_ITEMS], sub[PB_ITEMS];
  int4 r4[PB_ITEMS];
  // level 1: the edges' own rows of h (the DRAM stream) and their records
#pragma unroll
  for (int k = 0; k < PB_ITEMS; ++k) {
    const int64_t t = t0 + (int64_t)k * PB_THREADS;
    live[k] = t < total;
    split_item(live[k] ? t : 0, chunks, magic, e[k], c[k]);
    c[k] *= 4;
    if (live[k]) {
      hv[k] = ldg4_stream(h + (int64_t)e[k] * d + c[k]);
      r4[k] = __ldg(rec + e[k]);
    }
  }
  // level 2: the [B, d] rows (L2 / L1)
#pragma unroll
  for (int k = 0; k < PB_ITEMS; ++k) {
    gv[k] = sub[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live[k]) {
      gw[k] = ldg4(GW + (int64_t)r4[k].x * d + c[k]);
      if (residual) gv[k] = ldg4(G + (int64_t)r4[k].x * d + c[k]);
      if (r4[k].z >= 0) sub[k] = ldg4(GW + (int64_t)r4[k].z * d + c[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < PB_ITEMS; ++k) {
    if (!live[k]) continue;
    if (r4[k].w > 1) {  // several edges' rev point here (the reference's atom-offset rev_index quirk): the rest through the CSR
      const int lo = __ldg(rev_rowptr + e[k]);
      for (int j = lo + 1; j < lo + r4[k].w; ++j) {
        const float4 r = ldg4(GW + (int64_t)__ldg(mol + __ldg(rev_perm + j)) * d + c[k]);
        sub[k] = make_float4(sub[k].x + r.x, sub[k].y + r.y, sub[k].z + r.z, sub[k].w + r.w);
      }
    }
    const float scale = __int_as_float(r4[k].y);
    float4 r = make_float4(pb_act_bwd(AK, hv[k].x, act, act_param) * (scale * gw[k].x - sub[k].x),
                           pb_act_bwd(AK, hv[k].y, act, act_param) * (scale * gw[k].y - sub[k].y),
                           pb_act_bwd(AK, hv[k].z, act, act_param) * (scale * gw[k].z - sub[k].z),
                           pb_act_bwd(AK, hv[k].w, act, act_param) * (scale * gw[k].w - sub[k].w));
    if (residual) r = make_float4(gv[k].x + r.x, gv[k].y + r.y, gv[k].z + r.z, gv[k].w + r.w);
    stg4(g_h + (int64_t)e[k] * d + c[k], r);
  }
}

// Tiled form (the default; same structure as layer_bwd_epilogue_tiled in rowwise_kernels.cu): a CTA owns 32 consecutive edges, 8 threads
// per edge read that edge's record once, then the CTA sweeps the rows in passes of eight 16-byte chunks, two passes in flight. One
// memory round trip per pass instead of (record -> rows) per 16-byte result, and the [B, d] rows of a tile's one or two molecules
// are served by L1. Measured at configs[1]: per-item forms 161-171 us (1, 2 items per thread; runs of 4 consecutive edges with the
// rows re-used from registers: 171 - it is not the L2 -> SM bytes), 4 items per thread at 102 registers 243 us.
constexpr int PBT_EDGES = 32, PBT_LANES = 8;

template <int AK>
__global__ void __launch_bounds__(PB_THREADS, 4) layer_bwd_epilogue_pooled_tiled(const float* __restrict__ G, const float* __restrict__ GW,
                                                                                 const float* __restrict__ h, const int4* __restrict__ rec,
                                                                                 const int32_t* __restrict__ mol, const int32_t* __restrict__ rev_rowptr,
                                                                                 const int32_t* __restrict__ rev_perm, int d, int chunks, int64_t E, int act,
                                                                                 float act_param, int residual, float* __restrict__ g_h) {
  const int64_t e = (int64_t)blockIdx.x * PBT_EDGES + (threadIdx.x / PBT_LANES);
  const int cl = threadIdx.x % PBT_LANES;
  if (e >= E) return;
  const int4 r4 = __ldg(rec + e);
  const int lo = r4.w > 1 ? __ldg(rev_rowptr + e) : 0;
  const float scale = __int_as_float(r4.y);
  const int64_t own = e * d, rowb = (int64_t)r4.x * d, rows = (int64_t)(r4.z >= 0 ? r4.z : 0) * d;
  constexpr int PASSES = 2;
  for (int c0 = cl; c0 < chunks; c0 += PBT_LANES * PASSES) {
    float4 hv[PASSES], gw[PASSES], gv[PASSES], sub[PASSES];
    bool on[PASSES];
#pragma unroll
    for (int q = 0; q < PASSES; ++q) {
      const int c = (c0 + q * PBT_LANES) * 4;
      on[q] = c0 + q * PBT_LANES < chunks;
      gv[q] = sub[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on[q]) {
        hv[q] = ldg4_stream(h + own + c);
        gw[q] = ldg4(GW + rowb + c);
        if (residual) gv[q] = ldg4(G + rowb + c);
        if (r4.z >= 0) sub[q] = ldg4(GW + rows + c);
      }
    }
#pragma unroll
    for (int q = 0; q < PASSES; ++q) {
      if (!on[q]) continue;
      const int c = (c0 + q * PBT_LANES) * 4;
      float4 sb = sub[q];
      for (int j = lo + 1; j < lo + r4.w; ++j) {  // several edges' rev point here (the reference's atom-offset rev_index quirk)
        const float4 r = ldg4(GW + (int64_t)__ldg(mol + __ldg(rev_perm + j)) * d + c);
        sb = make_float4(sb.x + r.x, sb.y + r.y, sb.z + r.z, sb.w + r.w);
      }
      float4 r = make_float4(pb_act_bwd(AK, hv[q].x, act, act_param) * (scale * gw[q].x - sb.x), pb_act_bwd(AK, hv[q].y, act, act_param) * (scale * gw[q].y - sb.y),
                             pb_act_bwd(AK, hv[q].z, act, act_param) * (scale * gw[q].z - sb.z), pb_act_bwd(AK, hv[q].w, act, act_param) * (scale * gw[q].w - sb.w));
      if (residual) r = make_float4(gv[q].x + r.x, gv[q].y + r.y, gv[q].z + r.z, gv[q].w + r.w);
      stg4(g_h + own + c, r);
    }
  }
}

// ---- forward: M[b] = sum_{e in b} (w_e act(h[e]) - act(h[rev e])),  S[b] = [sum_{e in b} h[e]] + |b| bias -----------------------
// per-edge record {rev[e], bits of w_e = outdeg(dst e) [/ indeg(dst e)]}: the main loop then issues its row loads without walking
// dst -> rowptr first
__global__ void __launch_bounds__(PB_THREADS) pooled_fwd_record_kernel(const int32_t* __restrict__ rev, const int32_t* __restrict__ dst,
                                                                       const int32_t* __restrict__ src_rowptr, const int32_t* __restrict__ dst_rowptr,
                                                                       int64_t E, int mean, int2* __restrict__ rec) {
  const int64_t e = (int64_t)blockIdx.x * PB_THREADS + threadIdx.x;
  if (e >= E) return;
  const int v = __ldg(dst + e);
  float w = (float)(__ldg(src_rowptr + v + 1) - __ldg(src_rowptr + v));
  if (mean) w = w / (float)max(__ldg(dst_rowptr + v + 1) - __ldg(dst_rowptr + v), 1);
  int r = __ldg(rev + e);
  if (r < 0 || r >= E) r = (int)e;  // out-of-range indices are reported by nt_build_csr; stay in bounds here
  rec[e] = make_int2(r, __float_as_int(w));
}

template <int AK>
__device__ __forceinline__ float4 pb_act4(float4 v, int act, float p) {
  if (AK == 0) return v;
  if (AK == 1) return make_float4(v.x < 0.f ? 0.f : v.x, v.y < 0.f ? 0.f : v.y, v.z < 0.f ? 0.f : v.z, v.w < 0.f ? 0.f : v.w);
  return act_fwd4(v, act, p);
}

// one thread per (molecule, 16-byte chunk); the molecule's edges are walked in ascending order, two per trip (four row loads in
// flight per thread): sequential fp32 accumulation, deterministic
template <int AK>
__global__ void __launch_bounds__(PB_THREADS) pooled_message_sum_kernel(const float* __restrict__ h, const int2* __restrict__ rec,
                                                                        const int32_t* __restrict__ eptr, const float* __restrict__ bias, int d,
                                                                        int chunks, int64_t total, uint64_t magic, int act, float act_param,
                                                                        int residual, float* __restrict__ M, float* __restrict__ S) {
  const int64_t t = (int64_t)(gridDim.x - 1 - blockIdx.x) * PB_THREADS + threadIdx.x;  // from the end: the tail of h is what the producer left in L2
  if (t >= total) return;
  int b, c;
  split_item(t, chunks, magic, b, c);
  c *= 4;
  const int lo = __ldg(eptr + b), hi = __ldg(eptr + b + 1);
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f), s = m;
  for (int e = lo; e < hi; e += 2) {
    const bool two = e + 1 < hi;
    const int2 r0 = __ldg(rec + e);
    const int2 r1 = two ? __ldg(rec + e + 1) : make_int2(0, 0);
    const float4 h0 = ldg4_stream(h + (int64_t)e * d + c);
    const float4 g0 = ldg4(h + (int64_t)r0.x * d + c);
    float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = h1;
    if (two) {
      h1 = ldg4_stream(h + (int64_t)(e + 1) * d + c);
      g1 = ldg4(h + (int64_t)r1.x * d + c);
    }
    {
      const float w = __int_as_float(r0.y);
      const float4 a = pb_act4<AK>(h0, act, act_param), ar = pb_act4<AK>(g0, act, act_param);
      m = make_float4(m.x + (w * a.x - ar.x), m.y + (w * a.y - ar.y), m.z + (w * a.z - ar.z), m.w + (w * a.w - ar.w));
      s = make_float4(s.x + h0.x, s.y + h0.y, s.z + h0.z, s.w + h0.w);
    }
    if (two) {
      const float w = __int_as_float(r1.y);
      const float4 a = pb_act4<AK>(h1, act, act_param), ar = pb_act4<AK>(g1, act, act_param);
      m = make_float4(m.x + (w * a.x - ar.x), m.y + (w * a.y - ar.y), m.z + (w * a.z - ar.z), m.w + (w * a.w - ar.w));
      s = make_float4(s.x + h1.x, s.y + h1.y, s.z + h1.z, s.w + h1.w);
    }
  }
  if (!residual) s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) {
    const float n = (float)(hi - lo);
    const float4 bv = ldg4(bias + c);
    s = make_float4(s.x + n * bv.x, s.y + n * bv.y, s.z + n * bv.z, s.w + n * bv.w);
  }
  stg4(M + (int64_t)b * d + c, m);
  stg4(S + (int64_t)b * d + c, s);
}

}  // namespace nt

using namespace nt;

extern "C" int nt_weighted_colsum(const void* x, const int32_t* rowptr, int64_t rows, int64_t d, void* out, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_weighted_colsum: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(rows >= 0 && d > 0 && d < (1 << 20), "nt_weighted_colsum: bad sizes");
  NT_CHECK_ARG(out && (x || rows == 0), "nt_weighted_colsum: null pointer");
  if (d % 4 != 0 || !aligned16(x) || !aligned16(out)) {
    set_error("nt_weighted_colsum: needs d %% 4 == 0 and 16-byte aligned rows");
    return NT_ERR_UNSUPPORTED;
  }
  weighted_colsum_kernel<<<(unsigned)(d / 4), PB_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(x), rowptr, rows, (int)d, static_cast<float*>(out));
  NT_LAUNCH_CHECK("nt_weighted_colsum", 1);
  return NT_OK;
}

extern "C" size_t nt_layer_backward_epilogue_pooled_workspace_bytes(int64_t E) { return E > 0 ? (size_t)E * sizeof(int4) + 256 : 256; }

extern "C" int nt_layer_backward_epilogue_pooled(const void* gH, const void* gHW, const void* h, const int32_t* mol_of_edge, const int32_t* dst,
                                                 const int32_t* src_rowptr, const int32_t* rev_rowptr, const int32_t* rev_perm, const int32_t* dst_rowptr,
                                                 int64_t E, int64_t B, int64_t d, int act, float act_param, int residual, int mean, void* g_h,
                                                 void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_layer_backward_epilogue_pooled: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && E >= 0 && E < INT32_MAX && B >= 0 && B < INT32_MAX, "nt_layer_backward_epilogue_pooled: bad sizes");
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "nt_layer_backward_epilogue_pooled: bad activation");
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(gHW && h && mol_of_edge && dst && src_rowptr && rev_rowptr && rev_perm && g_h, "nt_layer_backward_epilogue_pooled: null pointer");
  NT_CHECK_ARG(!residual || gH, "nt_layer_backward_epilogue_pooled: residual needs gH");
  NT_CHECK_ARG(!mean || dst_rowptr, "nt_layer_backward_epilogue_pooled: mean needs dst_rowptr");
  if (d % 4 != 0 || !aligned16(gH) || !aligned16(gHW) || !aligned16(h) || !aligned16(g_h)) {
    set_error("nt_layer_backward_epilogue_pooled: needs d %% 4 == 0 and 16-byte aligned rows");
    return NT_ERR_UNSUPPORTED;
  }
  if (!workspace || !aligned16(workspace) || workspace_bytes < nt_layer_backward_epilogue_pooled_workspace_bytes(E)) {
    set_error("nt_layer_backward_epilogue_pooled: workspace too small (nt_layer_backward_epilogue_pooled_workspace_bytes)");
    return NT_ERR_WORKSPACE;
  }
  const int chunks = (int)(d / 4);
  const int64_t total = E * chunks;
  const uint64_t magic = chunk_div_magic(total, chunks);
  const char* ie = getenv("NOTORCH_B200_K6P_ITEMS");  // A/B timing, read per call: 0 (default) = tiled form; 1 = per-item form
  const int items = ie ? atoi(ie) : 0;
  cudaStream_t st = as_stream(stream);
  const float *Gf = static_cast<const float*>(gH), *GWf = static_cast<const float*>(gHW), *hf = static_cast<const float*>(h);
  float* out = static_cast<float*>(g_h);
  int4* rec = static_cast<int4*>(workspace);
  pooled_record_kernel<<<(unsigned)cdiv(E, PB_THREADS), PB_THREADS, 0, st>>>(mol_of_edge, dst, src_rowptr, rev_rowptr, rev_perm, dst_rowptr, E, mean, rec);
  if (items != 1) {
    const unsigned tgrid = (unsigned)cdiv(E, PBT_EDGES);
#define NT_PBT_LAUNCH(AK)                                                                                                                          \
  layer_bwd_epilogue_pooled_tiled<AK><<<tgrid, PB_THREADS, 0, st>>>(Gf, GWf, hf, rec, mol_of_edge, rev_rowptr, rev_perm, (int)d, chunks, E, act, act_param, \
                                                                    residual, out)
    if (act == NT_ACT_IDENTITY) NT_PBT_LAUNCH(0);
    else if (act == NT_ACT_RELU) NT_PBT_LAUNCH(1);
    else NT_PBT_LAUNCH(2);
#undef NT_PBT_LAUNCH
    NT_LAUNCH_CHECK("nt_layer_backward_epilogue_pooled", 2);
    return NT_OK;
  }
  const unsigned grid = (unsigned)cdiv(total, PB_THREADS);
#define NT_PB_LAUNCH(AK)                                                                                                                            \
  layer_bwd_epilogue_pooled<AK, 1><<<grid, PB_THREADS, 0, st>>>(Gf, GWf, hf, rec, mol_of_edge, rev_rowptr, rev_perm, (int)d, chunks, total, magic, act, \
                                                                act_param, residual, out)
  if (act == NT_ACT_IDENTITY) NT_PB_LAUNCH(0);
  else if (act == NT_ACT_RELU) NT_PB_LAUNCH(1);
  else NT_PB_LAUNCH(2);
#undef NT_PB_LAUNCH
  NT_LAUNCH_CHECK("nt_layer_backward_epilogue_pooled", 2);
  return NT_OK;
}

extern "C" size_t nt_pooled_message_sum_workspace_bytes(int64_t E) { return E > 0 ? (size_t)E * sizeof(int2) + 256 : 256; }

extern "C" int nt_pooled_message_sum(const void* h, const int32_t* rev, const int32_t* dst, const int32_t* src_rowptr, const int32_t* dst_rowptr,
                                     const int32_t* mol_edge_ptr, const void* bias, int64_t E, int64_t B, int64_t d, int act, float act_param,
                                     int residual, int mean, void* M, void* S, void* workspace, size_t workspace_bytes, int dtype,
                                     nt_stream_t stream) {
  if (dtype != NT_F32) { set_error("nt_pooled_message_sum: only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && E >= 0 && E < INT32_MAX && B >= 0 && B < INT32_MAX, "nt_pooled_message_sum: bad sizes");
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "nt_pooled_message_sum: bad activation");
  if (B == 0) return NT_OK;
  NT_CHECK_ARG(mol_edge_ptr && M && S && (E == 0 || (h && rev && dst && src_rowptr)), "nt_pooled_message_sum: null pointer");
  NT_CHECK_ARG(!mean || dst_rowptr, "nt_pooled_message_sum: mean needs dst_rowptr");
  if (d % 4 != 0 || !aligned16(h) || !aligned16(bias) || !aligned16(M) || !aligned16(S)) {
    set_error("nt_pooled_message_sum: needs d %% 4 == 0 and 16-byte aligned rows");
    return NT_ERR_UNSUPPORTED;
  }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 7u) != 0 || workspace_bytes < nt_pooled_message_sum_workspace_bytes(E)) {
    set_error("nt_pooled_message_sum: workspace too small (nt_pooled_message_sum_workspace_bytes)");
    return NT_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int2* rec = static_cast<int2*>(workspace);
  if (E > 0) pooled_fwd_record_kernel<<<(unsigned)cdiv(E, PB_THREADS), PB_THREADS, 0, st>>>(rev, dst, src_rowptr, dst_rowptr, E, mean, rec);
  const int chunks = (int)(d / 4);
  const int64_t total = B * chunks;
  const uint64_t magic = chunk_div_magic(total, chunks);
  const unsigned grid = (unsigned)cdiv(total, PB_THREADS);
  const float *hf = static_cast<const float*>(h), *bf = static_cast<const float*>(bias);
  float *Mf = static_cast<float*>(M), *Sf = static_cast<float*>(S);
#define NT_PM_LAUNCH(AK) \
  pooled_message_sum_kernel<AK><<<grid, PB_THREADS, 0, st>>>(hf, rec, mol_edge_ptr, bf, (int)d, chunks, total, magic, act, act_param, residual, Mf, Sf)
  if (act == NT_ACT_IDENTITY) NT_PM_LAUNCH(0);
  else if (act == NT_ACT_RELU) NT_PM_LAUNCH(1);
  else NT_PM_LAUNCH(2);
#undef NT_PM_LAUNCH
  NT_LAUNCH_CHECK("nt_pooled_message_sum", 2);
  return NT_OK;
}
