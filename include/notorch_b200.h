/*
 * notorch_b200 — C ABI of the B200-native D-MPNN hot path (libnotorch_b200.so).
 *
 * This is the drop-in boundary: the exact entry points a binding in the reference
 * (davidegraff/notorch, pure Python on top of PyTorch) would call for its message-passing hot
 * path. The reference has no FFI of its own; each function below names the reference code it
 * replaces (paths relative to the reference tree). The Python side of this repository
 * (notorch_b200/_lib.py, ops.py) binds them with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in _host. Row-major, contiguous.
 *  - Every function is asynchronous on `stream` (a cudaStream_t), does not allocate or free device
 *    memory, does not synchronise, and keeps no pointer after it returns. Scratch space is passed
 *    in by the caller (`workspace`, sized by the matching *_workspace_bytes function).
 *  - Return value: NT_OK (0) or an nt_status error code; nt_last_error_string() describes the last
 *    error on the calling thread. No exception crosses the ABI; nothing calls exit/abort.
 *  - Thread-safe and re-entrant (PyTorch calls backward from its autograd thread).
 *  - Index tensors handed to the floating-point kernels are int32 (built once per batch by
 *    nt_graph_prepare from the reference's int64 tensors); sizes must be < 2^31.
 *  - dtype: NT_F32 activations only in this round (NT_BF16 selects the bf16 weight image in nt_weight_prepare and returns
 *    NT_ERR_UNSUPPORTED everywhere else); the bf16 OPERAND mode is a gemm_mode, see NT_GEMM_BF16.
 */
#ifndef NOTORCH_B200_H_
#define NOTORCH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nt_stream_t; /* cudaStream_t */

enum nt_status {
  NT_OK = 0,
  NT_ERR_ARG = 1,         /* null pointer, negative size, size >= 2^31, bad enum */
  NT_ERR_ALIGN = 2,       /* pointer / leading dimension not aligned as the kernel requires */
  NT_ERR_UNSUPPORTED = 3, /* dtype / activation / shape this build does not implement */
  NT_ERR_CUDA = 4,        /* a CUDA runtime call or kernel launch failed */
  NT_ERR_WORKSPACE = 5    /* workspace too small */
};

enum nt_dtype { NT_F32 = 0, NT_BF16 = 1 };

/* torch.nn activation modules accepted for ChempropLayer(act=...) (chemprop.py:17,24,37). */
enum nt_act {
  NT_ACT_IDENTITY = 0,
  NT_ACT_RELU = 1,       /* nn.ReLU (reference default) */
  NT_ACT_LEAKY_RELU = 2, /* nn.LeakyReLU, act_param = negative_slope */
  NT_ACT_ELU = 3,        /* nn.ELU, act_param = alpha */
  NT_ACT_SILU = 4,       /* nn.SiLU */
  NT_ACT_GELU = 5,       /* nn.GELU (erf form) */
  NT_ACT_TANH = 6        /* nn.Tanh */
};

/* Arithmetic path of the W contraction. */
enum nt_gemm_mode {
  NT_GEMM_TF32X3 = 0, /* tcgen05 tensor cores, 3xTF32 error-compensated split, fp32 accumulate in TMEM */
  NT_GEMM_FP32 = 1,   /* fp32 FFMA on CUDA cores (strict fp32; also used when d % 4 != 0) */
  NT_GEMM_TF32 = 2,   /* tcgen05, single-pass TF32 (10-bit mantissa; NOT within the fp32 parity bound) */
  NT_GEMM_BF16 = 3    /* BASELINE configs[4]: W_h and the message operand rounded to bf16, ONE tcgen05.mma.kind::f16 pass, fp32
                         accumulation in TMEM (K2, K4a, nt_dense_forward); the weight gradient runs as a single TF32 pass.
                         Activations stay fp32 in HBM. Weight image: nt_weight_prepare(..., dtype = NT_BF16). Stated bound:
                         rel-to-max 2e-2 on embeddings, 5e-2 on gradients (tests/test_bf16_mode.py) */
};

/* Bumped whenever an entry point's argument list changes; the ctypes binding (notorch_b200/_lib.py) refuses a library whose
 * nt_version() differs, so a stale .so can never be called with a newer signature table. */
#define NT_ABI_VERSION 201

const char* nt_last_error_string(void);
int nt_version(void); /* = NT_ABI_VERSION of the build */
/* Number of kernels this library has launched in this process so far (all threads). */
long long nt_kernel_launch_count(void);
/* 1 if the current device is compute capability 10.x (tcgen05 available), else 0; <0 on error. */
int nt_device_supported(void);

/* ------------------------------------------------------------------------------------------------
 * Batch collation  — replaces BatchedGraph.from_graphs, notorch/data/models/graph.py:186-223.
 *
 * Input: B molecules in packed form: num_atoms[B], num_edges[B] (int32) and the molecule-LOCAL
 * edge_index [2,E] / rev_index [E] (int32) concatenated in batch order (the per-molecule contract of
 * notorch/transforms/graph.py:32-43). Output: the reference's int64 tensors, bit-exact:
 * edge_index [2,E] and rev_index [E] with the cumulative ATOM count added to BOTH (graph.py:199-200;
 * rev_offset_mode = 0), batch_node_index [V], batch_edge_index [E]; plus the int32 molecule row
 * pointers mol_atom_ptr [B+1], mol_edge_ptr [B+1]. rev_offset_mode = 1 adds the cumulative EDGE
 * count to rev_index instead (structurally correct reverse edge; a labelled deviation).
 * ---------------------------------------------------------------------------------------------- */
size_t nt_collate_workspace_bytes(int64_t B);
int nt_collate(const int32_t* num_atoms, const int32_t* num_edges, int64_t B,
               const int32_t* local_edge_index, const int32_t* local_rev_index, int64_t V, int64_t E,
               int rev_offset_mode,
               int64_t* edge_index, int64_t* rev_index, int64_t* batch_node_index, int64_t* batch_edge_index,
               int32_t* mol_atom_ptr, int32_t* mol_edge_ptr,
               void* workspace, size_t workspace_bytes, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stable CSR of `n` items grouped by an int64 key in [0, num_segments): rowptr [S+1], perm [n] with
 * perm[rowptr[s]..rowptr[s+1]) = ids of the items with key s in ASCENDING id order (so a sequential
 * segmented sum reproduces the summation order of the reference's CPU scatter_add_).
 * Also writes keys32 [n] = (int32) keys (nullable). *status (device int32, caller-zeroed) gets bit 0
 * set if any key is out of range — what the reference would raise as an indexing error at
 * chemprop.py:39-40. Replaces the index handling inside torch_scatter.scatter / aten::index.
 * ---------------------------------------------------------------------------------------------- */
size_t nt_build_csr_workspace_bytes(int64_t n, int64_t num_segments);
int nt_build_csr(const int64_t* keys, int64_t n, int64_t num_segments,
                 int32_t* keys32, int32_t* rowptr, int32_t* perm, int32_t* status,
                 void* workspace, size_t workspace_bytes, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * out[s,:] = reduce_{j in [rowptr[s], rowptr[s+1])} act(x[perm[j],:])      (sequential, ascending j)
 *   mean != 0 : divide by max(count, 1)  (torch_scatter.scatter_mean: a true division)
 *   scale     : multiply the result by `scale` (1.0f = off; used by the Norm read-out extension)
 * K1 edge->atom   : chemprop.py:37+39 (act = layer act) and chemprop.py:86 (act = identity)
 * K3 read-out     : agg.py:27 (Sum) / agg.py:36 (Mean), rowptr/perm = CSR of batch_node_index
 * K5 backward     : gradient of x[src] gathers (aten::index backward), rowptr/perm = CSR of src
 * perm may be NULL (identity: segments are contiguous row ranges).
 * ---------------------------------------------------------------------------------------------- */
int nt_seg_reduce(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments,
                  int act, float act_param, int mean, float scale, void* out, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * nt_seg_reduce with an epilogue, for the atom-state message-passing variant (extension, not in the
 * reference tree):
 *   dact_of == NULL : out[s,:] = (base ? base[s,:] : 0) + scale * reduce_j act(x[perm[j],:])               (forward: n = S_e + sum act(h)[src])
 *   dact_of != NULL : out[s,:] = (base ? base[s,:] : 0) + act'(dact_of[s,:]) * scale * reduce_j x[perm[j],:]   (backward through act)
 * ---------------------------------------------------------------------------------------------- */
int nt_seg_reduce_ex(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments,
                     int act, float act_param, int mean, float scale, const void* base, const void* dact_of,
                     void* out, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ELL acceleration of the segmented reductions (same arithmetic and accumulation order as nt_seg_reduce / nt_seg_reduce_ex):
 * nt_csr_to_ell writes ell[s] = the first four item ids of segment s (ascending, -1 padded; int32 x 4, 16-byte aligned), built
 * once per batch next to the CSR; nt_seg_reduce_ell reads all row indices of a segment with one load (degree <= 4 for molecular
 * graphs; longer segments continue through rowptr / perm). Replaces the same reference lines as nt_seg_reduce.
 * ---------------------------------------------------------------------------------------------- */
int nt_csr_to_ell(const int32_t* rowptr, const int32_t* perm, int64_t num_segments, int32_t* ell, nt_stream_t stream);
int nt_seg_reduce_ell(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, const int32_t* ell,
                      int64_t num_segments, int act, float act_param, int mean, float scale, const void* base,
                      const void* dact_of, void* out, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * out[i,:] = (base ? base[i,:] : 0) + scale * x[idx[i],:] / (mean_rowptr ? max(count(idx[i]),1) : 1)
 * K0 edge_init      : chemprop.py:83  h0 = x_v[src] + x_e            (base = x_e, x = x_v, idx = src)
 * backward of K1/K3 : g[e] = gE[e] + g_node[dst[e]] (/ indeg for mean);  g_x[v] = gH[batch[v]] (/count)
 * ---------------------------------------------------------------------------------------------- */
int nt_gather_add(const void* base, const void* x, const int32_t* idx, const int32_t* mean_rowptr,
                  int64_t n, int64_t d, float scale, void* out, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weight staging for the tensor-core path: splits W [d,d] (row-major [out,in], nn.Linear layout,
 * chemprop.py:26) into TF32 hi/lo parts and writes them as the shared-memory image the fused kernels
 * stream with bulk TMA copies (K-blocks of 32, 128-byte-swizzled K-major tiles, zero padded).
 * transpose = 0: B operand W (forward);  1: B operand W^T (dgrad).  Two images of
 * nt_weight_image_bytes(d) each are written back to back into `image`.
 * ---------------------------------------------------------------------------------------------- */
size_t nt_weight_image_bytes(int64_t d);
int nt_weight_prepare(const void* W, int64_t d, int transpose, void* image, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Dense Linear + Dropout + residual on tcgen05 (same CTA-pair kernel as K2, dense A operand):
 *   out[r,:] = (resid ? resid[r,:] : 0) + Dropout_p(x[r,:] . W^T + bias)
 * The atom-state update of the atom message-passing variant (row A10); weight_image = nt_weight_prepare(W, transpose = 0).
 * ---------------------------------------------------------------------------------------------- */
int nt_dense_forward(const void* x, const void* weight_image, const void* bias, const void* resid, int64_t R, int64_t d,
                     float dropout_p, uint64_t seed, uint64_t offset, void* out, int dtype, int gemm_mode, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2 — one fused message-passing depth, forward. Replaces chemprop.py:40-41 + residual.py:28:
 *   m[e,:]  = n[src[e],:] - act(h[rev[e],:])
 *   u[e,:]  = m[e,:] . W^T + bias                       (tcgen05 / fp32 per `gemm_mode`)
 *   u       = dropout_p(u)        (Philox keyed by (seed, offset, e*d + c); p = 0 -> identity)
 *   out[e,:]= (residual ? h[e,:] : 0) + u[e,:]
 * n is K1's output (act already applied inside K1). rev is an arbitrary in-range gather index.
 * weight_image: from nt_weight_prepare(transpose=0) (ignored for NT_GEMM_FP32, may be NULL).
 * m_out (nullable, [E,d]): if given, the message tensor m is also written out (saved for K4b so that the
 * weight gradient can stream it with TMA instead of gathering it again).
 * ---------------------------------------------------------------------------------------------- */
int nt_layer_forward(const void* h, const void* n, const int32_t* src, const int32_t* rev,
                     const void* W, const void* weight_image, const void* bias,
                     int64_t E, int64_t V, int64_t d, int act, float act_param, int residual,
                     float dropout_p, uint64_t seed, uint64_t offset,
                     void* out, void* m_out, int dtype, int gemm_mode, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4a — dgrad: g_m[e,:] = (mask . g[e,:] / (1-p)) . W        (backward of aten::addmm wrt input)
 * weight_image: from nt_weight_prepare(transpose=1).
 * ---------------------------------------------------------------------------------------------- */
int nt_layer_backward_dgrad(const void* g, const void* W, const void* weight_image,
                            int64_t E, int64_t d, float dropout_p, uint64_t seed, uint64_t offset,
                            void* g_m, int dtype, int gemm_mode, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4b — wgrad: gW[o,i] = sum_e g_u[e,o] * m[e,i],  gb[o] = sum_e g_u[e,o]   (gb nullable)
 * with g_u = mask . g / (1-p). m is either the tensor K2 saved (`m` != NULL: both operands are streamed
 * by TMA) or recomputed from (n, h, src, rev) as in K2 (`m` == NULL). Deterministic: split over edge
 * ranges into `workspace`, then a fixed-order reduction (no atomics).
 * ---------------------------------------------------------------------------------------------- */
size_t nt_layer_backward_wgrad_workspace_bytes(int64_t E, int64_t d);
int nt_layer_backward_wgrad(const void* g, const void* m, const void* h, const void* n, const int32_t* src, const int32_t* rev,
                            int64_t E, int64_t V, int64_t d, int act, float act_param,
                            float dropout_p, uint64_t seed, uint64_t offset,
                            void* gW, void* gb, void* workspace, size_t workspace_bytes,
                            int dtype, int gemm_mode, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K6 — backward epilogue of one depth (backward of chemprop.py:37,39,40 and residual.py:28):
 *   g_a[e,:] = g_n[dst[e],:] / (mean ? max(indeg(dst[e]),1) : 1) - sum_{j in revinv(e)} g_m[rev_perm[j],:]
 *   g_h[e,:] = (residual ? g[e,:] : 0) + act'(h[e,:]) * g_a[e,:]
 * rev_rowptr/rev_perm: CSR of rev_index (the inverse map of the arbitrary gather `rev`).
 * dst_rowptr: CSR row pointers of dst (only read when mean != 0).
 * ---------------------------------------------------------------------------------------------- */
int nt_layer_backward_epilogue(const void* g, const void* h, const void* g_n, const void* g_m,
                               const int32_t* dst, const int32_t* rev_rowptr, const int32_t* rev_perm,
                               const int32_t* dst_rowptr, int64_t E, int64_t d,
                               int act, float act_param, int residual, int mean,
                               void* g_h, int dtype, nt_stream_t stream);

/* K6 after a max / min forward reduction (chemprop.py:39 with reduce in {"max","min"}): torch_scatter routes the gradient of an
 * arg-reduction to the argument row only, so g_a[e,c] = (arg[dst[e],c] == e ? g_n[dst[e],c] : 0) - sum_{j in revinv(e)} g_m[..];
 * arg is the [V,d] int32 output of nt_seg_extreme. */
int nt_layer_backward_epilogue_arg(const void* g, const void* h, const void* g_n, const void* g_m,
                                   const int32_t* dst, const int32_t* arg, const int32_t* rev_rowptr,
                                   const int32_t* rev_perm, int64_t E, int64_t d, int act, float act_param,
                                   int residual, void* g_h, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K5 + K6 fused (same arithmetic as nt_seg_reduce over the by-source CSR followed by nt_layer_backward_epilogue, bit-identical):
 * every edge sums the g_m rows of the outgoing edges of its destination atom itself (src_ell / src_rowptr / src_perm: the CSR of
 * edge_index[0] and its ELL copy from nt_csr_to_ell), so g_n is never written or read.
 * ---------------------------------------------------------------------------------------------- */
int nt_layer_backward_epilogue_fused(const void* g, const void* h, const void* g_m, const int32_t* dst,
                                     const int32_t* src_rowptr, const int32_t* src_perm, const int32_t* src_ell,
                                     const int32_t* rev_rowptr, const int32_t* rev_perm, const int32_t* dst_rowptr,
                                     int64_t E, int64_t d, int act, float act_param, int residual, int mean,
                                     void* g_h, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of the LAST depth under a sum read-out over each molecule's edges (notorch/nn/gnn/agg.py:27,36 behind
 * chemprop.py:86,88 on a device-collated batch): the gradient reaching h_L is the broadcast g[e] = gH[mol(e)], so
 *   gW^T = (sum_{e in b} m[e])^T gH   -> nt_seg_reduce over m, then nt_layer_backward_wgrad on B rows
 *   gb   = sum_b |b| gH[b]            -> nt_weighted_colsum (rowptr = the molecules' edge pointers; NULL = unit weights)
 *   g_m  = (gH W)[mol(e)]             -> nt_layer_backward_dgrad on B rows = gHW [B, d]; g_m [E, d] is never written
 *   g_h[e] = [gH[mol e]] + act'(h[e]) * (outdeg(dst e) gHW[mol e] (/ indeg(dst e)) - sum_{e'': rev[e''] = e} gHW[mol e''])
 * mol_of_edge [E] = batch_edge_index as int32 (molecule of every edge); src_rowptr / dst_rowptr / rev_rowptr / rev_perm as in
 * nt_layer_backward_epilogue_fused. Requires that a molecule's edges connect only its own atoms (BatchedGraph.from_packed).
 * ---------------------------------------------------------------------------------------------- */
int nt_weighted_colsum(const void* x, const int32_t* rowptr, int64_t rows, int64_t d, void* out, int dtype, nt_stream_t stream);
/* Forward of the same collapse (chemprop.py:37-41 + residual.py:28 + agg.py:27 for the LAST depth): one pass over h = h_{L-1},
 *   M[b,:] = sum_{e in b} (outdeg(dst e) [/ indeg(dst e)] act(h[e,:]) - act(h[rev e,:]))   = sum_{e in b} m[e,:]
 *   S[b,:] = (residual ? sum_{e in b} h[e,:] : 0) + |b| bias                                  (bias may be NULL)
 * then H_sum = S + M . W^T through nt_dense_forward(x = M, resid = S) on B rows. mol_edge_ptr [B + 1]: the molecules' contiguous edge
 * ranges; workspace: 8 bytes per edge. */
size_t nt_pooled_message_sum_workspace_bytes(int64_t E);
int nt_pooled_message_sum(const void* h, const int32_t* rev, const int32_t* dst, const int32_t* src_rowptr,
                          const int32_t* dst_rowptr, const int32_t* mol_edge_ptr, const void* bias, int64_t E, int64_t B,
                          int64_t d, int act, float act_param, int residual, int mean, void* M, void* S, void* workspace,
                          size_t workspace_bytes, int dtype, nt_stream_t stream);
size_t nt_layer_backward_epilogue_pooled_workspace_bytes(int64_t E); /* 16 bytes per edge: the per-edge index record */
int nt_layer_backward_epilogue_pooled(const void* gH, const void* gHW, const void* h, const int32_t* mol_of_edge,
                                      const int32_t* dst, const int32_t* src_rowptr, const int32_t* rev_rowptr,
                                      const int32_t* rev_perm, const int32_t* dst_rowptr, int64_t E, int64_t B, int64_t d,
                                      int act, float act_param, int residual, int mean, void* g_h, void* workspace,
                                      size_t workspace_bytes, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Prediction head (row N3): the nn.Linear layers of notorch/nn/mlp.py:58-62 on the [B, d] molecule vectors, strict fp32 (FFMA),
 * rectangular: W is [out_features, in_features] row-major like nn.Linear.weight.
 *   nt_linear_forward          out = x W^T + bias                      (bias may be NULL)
 *   nt_linear_backward_input   gx  = g W
 *   nt_linear_backward_weight  gW  = g^T x, gb = column sums of g      (gb may be NULL; split over rows, fixed-order reduction)
 * With rows == 0, nt_linear_backward_weight writes zeros.
 * ---------------------------------------------------------------------------------------------- */
int nt_linear_forward(const void* x, const void* W, const void* bias, int64_t rows, int64_t out_features,
                      int64_t in_features, void* out, int dtype, nt_stream_t stream);
int nt_linear_backward_input(const void* g, const void* W, int64_t rows, int64_t out_features, int64_t in_features,
                             void* gx, int dtype, nt_stream_t stream);
size_t nt_linear_backward_weight_workspace_bytes(int64_t rows, int64_t out_features, int64_t in_features);
int nt_linear_backward_weight(const void* g, const void* x, int64_t rows, int64_t out_features, int64_t in_features,
                              void* gW, void* gb, void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GraphEmbedding (row N1, the step before the block) — replaces the two nn.EmbeddingBag(mode="sum") of
 * notorch/nn/gnn/embed.py:20-24 on 2-D index input: out[i,:] = sum_{j<bag} table[idx[i,j],:]  (idx int64 [n,bag]).
 * *status bit 0 is set if an index is outside [0, num_types). Backward: g_table[t,:] = sum over all (i,j) with
 * idx[i,j] == t of g[i,:], deterministic (per-CTA partial tables + fixed-order sum, no atomics).
 * ---------------------------------------------------------------------------------------------- */
int nt_embedding_bag_sum(const void* table, int64_t num_types, const int64_t* idx, int64_t n, int64_t bag, int64_t d,
                         void* out, int32_t* status, int dtype, nt_stream_t stream);
size_t nt_embedding_bag_backward_workspace_bytes(int64_t n, int64_t num_types, int64_t d);
int nt_embedding_bag_backward(const void* g, const int64_t* idx, int64_t n, int64_t bag, int64_t num_types, int64_t d,
                              void* g_table, void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GraphEmbedding FUSED into the edge initialisation (row N1 as SURVEY.md §8f words it) — replaces
 * notorch/nn/gnn/embed.py:20-24 followed by notorch/nn/gnn/chemprop.py:83 in one kernel:
 *   h0[e,:] = (sum_j table_v[node_types[src[e], j], :]) + (sum_k table_e[edge_types[e, k], :])
 * node_types int64 [V, bag_v], edge_types int64 [E, bag_e], src int32 [E]; both tables are staged in shared memory, x_v / x_e
 * are never written. Bit-identical to nt_embedding_bag_sum (twice) + nt_gather_add. *status bit 0: a type id or src out of range.
 * Backward: ONE pass over g = dL/dh0 [E, d] yields both table gradients (per-thread column ownership over private
 * shared-memory tables, fixed-order sums: deterministic, no atomics). NT_ERR_UNSUPPORTED when d % 4 != 0 or the combined
 * vocabulary does not fit shared memory (callers then use the unfused entry points above). bag_v + bag_e <= 32.
 * ---------------------------------------------------------------------------------------------- */
int nt_embed_edge_init(const void* table_v, int64_t num_node_types, const void* table_e, int64_t num_edge_types,
                       const int64_t* node_types, int64_t bag_v, const int64_t* edge_types, int64_t bag_e, const int32_t* src,
                       int64_t E, int64_t V, int64_t d, void* h0, int32_t* status, int dtype, nt_stream_t stream);
size_t nt_embed_edge_init_backward_workspace_bytes(int64_t E, int64_t num_node_types, int64_t num_edge_types, int64_t d);
int nt_embed_edge_init_backward(const void* g, const int64_t* node_types, int64_t bag_v, const int64_t* edge_types, int64_t bag_e,
                                const int32_t* src, int64_t E, int64_t V, int64_t num_node_types, int64_t num_edge_types, int64_t d,
                                void* g_table_v, void* g_table_e, void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Remaining read-outs of notorch/nn/gnn/agg.py (row N2). Segments are given as a CSR (rowptr, perm; perm NULL =
 * contiguous rows); every loop is sequential in ascending row order (deterministic).
 *  nt_seg_max            agg.Max (agg.py:45, torch_scatter.scatter_max): out[s,c] = max, arg = first row attaining it
 *                        (-1 and value 0 for an empty segment). Backward routes g to the arg row.
 *  nt_row_dot            out[i] = scale * <x[i,:], y[row(i),:]> + bias, row(i) = y_index[i] | 0 (y_rows == 1) | i
 *  nt_seg_softmax(_backward)  torch_scatter.scatter_softmax over a [n] score vector (agg.py:60,83)
 *  nt_seg_weighted_sum   out[s,:] = scale * sum_j w[r_j] * x[r_j,:]          (agg.py:61,84: scatter_sum(alpha * x))
 *  nt_row_scale_gather   out[i,:] = scale * w[i] * y[y_index ? y_index[i] : 0, :]   (backward of the two above)
 *  nt_weighted_col_sum   out[c] = scale * sum_i w[i] * x[i,c]  (w NULL = 1), two fixed-order stages
 * ---------------------------------------------------------------------------------------------- */
int nt_seg_max(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments,
               void* out, int32_t* arg, int dtype, nt_stream_t stream);
/* nt_seg_extreme: scatter(act(x), index, reduce = is_min ? "min" : "max") with its argument (chemprop.py:39,86 for the block-level
 * arg-reductions): out[s,c] = extreme over the segment of act(x[r,c]), arg[s,c] = the FIRST row attaining it; empty: 0 / -1.
 * nt_seg_max_backward is the backward of both (g flows to the argument row). */
int nt_seg_extreme(const void* x, int64_t d, const int32_t* rowptr, const int32_t* perm, int64_t num_segments,
                   int act, float act_param, int is_min, void* out, int32_t* arg, int dtype, nt_stream_t stream);
int nt_seg_max_backward(const void* g, const int32_t* arg, const int32_t* seg_of_row, int64_t n, int64_t d,
                        void* gx, int dtype, nt_stream_t stream);
int nt_row_dot(const void* x, const void* y, const int32_t* y_index, int64_t y_rows, int64_t n, int64_t d,
               float scale, float bias, void* out, int dtype, nt_stream_t stream);
int nt_seg_softmax(const void* s, const int32_t* rowptr, const int32_t* perm, int64_t num_segments, void* alpha,
                   int dtype, nt_stream_t stream);
int nt_seg_softmax_backward(const void* alpha, const void* g_alpha, const int32_t* rowptr, const int32_t* perm,
                            int64_t num_segments, void* g_s, int dtype, nt_stream_t stream);
int nt_seg_weighted_sum(const void* x, const void* w, int64_t d, const int32_t* rowptr, const int32_t* perm,
                        int64_t num_segments, float scale, void* out, int dtype, nt_stream_t stream);
int nt_row_scale_gather(const void* y, const void* w, const int32_t* y_index, int64_t n, int64_t d, float scale,
                        void* out, int dtype, nt_stream_t stream);
size_t nt_weighted_col_sum_workspace_bytes(int64_t n, int64_t d);
int nt_weighted_col_sum(const void* x, const void* w, int64_t n, int64_t d, float scale, void* out,
                        void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream);

/* Dropout keep-mask exactly as K2/K4 compute it (1.0f keep / 0.0f drop), for tests. */
int nt_dropout_mask(int64_t n_rows, int64_t d, float dropout_p, uint64_t seed, uint64_t offset,
                    float* mask, nt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NOTORCH_B200_H_ */
