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
// The [E, d] tensors g, g_m and g_n are never materialised: the last depth's backward reads m once and h once and writes g_h.
// Exact algebra, another summation order (tested against the dense path and the fp64 oracle).
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

// one thread per (edge, 16-byte chunk). DRAM traffic: h[e] in, g_h[e] out; G and GW ([B, d], a few MB) stay in L2 / L1.
template <int AK>
__global__ void __launch_bounds__(PB_THREADS, 5) layer_bwd_epilogue_pooled(const float* __restrict__ G, const float* __restrict__ GW,
                                                                           const float* __restrict__ h, const int4* __restrict__ rec,
                                                                           const int32_t* __restrict__ mol, const int32_t* __restrict__ rev_rowptr,
                                                                           const int32_t* __restrict__ rev_perm, int d, int chunks, int64_t total,
                                                                           uint64_t magic, int act, float act_param, int residual,
                                                                           float* __restrict__ g_h) {
  const int64_t t = (int64_t)blockIdx.x * PB_THREADS + threadIdx.x;
  if (t >= total) return;
  int e, c;
  split_item(t, chunks, magic, e, c);
  c *= 4;
  // level 1: the edge's own row of h (the DRAM stream) and its record
  const float4 hv = ldg4_stream(h + (int64_t)e * d + c);
  const int4 r4 = __ldg(rec + e);
  // level 2: the [B, d] rows (L2 / L1)
  const float4 gw = ldg4(GW + (int64_t)r4.x * d + c);
  float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (residual) gv = ldg4(G + (int64_t)r4.x * d + c);
  float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r4.z >= 0) sub = ldg4(GW + (int64_t)r4.z * d + c);
  if (r4.w > 1) {  // several edges' rev point here (the reference's atom-offset rev_index quirk): the rest through the CSR
    const int lo = __ldg(rev_rowptr + e);
    for (int j = lo + 1; j < lo + r4.w; ++j) {
      const float4 r = ldg4(GW + (int64_t)__ldg(mol + __ldg(rev_perm + j)) * d + c);
      sub = make_float4(sub.x + r.x, sub.y + r.y, sub.z + r.z, sub.w + r.w);
    }
  }
  const float scale = __int_as_float(r4.y);
  float4 r = make_float4(pb_act_bwd(AK, hv.x, act, act_param) * (scale * gw.x - sub.x), pb_act_bwd(AK, hv.y, act, act_param) * (scale * gw.y - sub.y),
                         pb_act_bwd(AK, hv.z, act, act_param) * (scale * gw.z - sub.z), pb_act_bwd(AK, hv.w, act, act_param) * (scale * gw.w - sub.w));
  if (residual) r = make_float4(gv.x + r.x, gv.y + r.y, gv.z + r.z, gv.w + r.w);
  stg4(g_h + (int64_t)e * d + c, r);
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
  const unsigned grid = (unsigned)cdiv(total, PB_THREADS);
  cudaStream_t st = as_stream(stream);
  const float *Gf = static_cast<const float*>(gH), *GWf = static_cast<const float*>(gHW), *hf = static_cast<const float*>(h);
  float* out = static_cast<float*>(g_h);
  int4* rec = static_cast<int4*>(workspace);
  pooled_record_kernel<<<(unsigned)cdiv(E, PB_THREADS), PB_THREADS, 0, st>>>(mol_of_edge, dst, src_rowptr, rev_rowptr, rev_perm, dst_rowptr, E, mean, rec);
#define NT_PB_LAUNCH(AK)                                                                                                                           \
  layer_bwd_epilogue_pooled<AK><<<grid, PB_THREADS, 0, st>>>(Gf, GWf, hf, rec, mol_of_edge, rev_rowptr, rev_perm, (int)d, chunks, total, magic, act, \
                                                             act_param, residual, out)
  if (act == NT_ACT_IDENTITY) NT_PB_LAUNCH(0);
  else if (act == NT_ACT_RELU) NT_PB_LAUNCH(1);
  else NT_PB_LAUNCH(2);
#undef NT_PB_LAUNCH
  NT_LAUNCH_CHECK("nt_layer_backward_epilogue_pooled", 2);
  return NT_OK;
}
