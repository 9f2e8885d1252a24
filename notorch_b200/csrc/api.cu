// extern "C" entry points that dispatch between the arithmetic paths, plus error plumbing.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <string.h>

#include "common.cuh"
#include "../../include/notorch_b200_debug.h"

namespace nt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return NT_ERR_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

// gemm_simt.cu
int simt_wgrad_splits(int64_t E, int64_t d);
int simt_layer_forward(const float*, const float*, const int32_t*, const int32_t*, const float*, const float*, int64_t, int64_t, int, float, int, float,
                       uint64_t, uint64_t, float*, cudaStream_t);
int simt_layer_dgrad(const float*, const float*, int64_t, int64_t, float, uint64_t, uint64_t, float*, cudaStream_t);
int simt_layer_wgrad(const float*, const float*, const float*, const int32_t*, const int32_t*, int64_t, int64_t, int, float, float, uint64_t, uint64_t,
                     float*, float*, float*, cudaStream_t);
size_t simt_linear_wgrad_workspace_bytes(int64_t rows, int64_t n_out, int64_t k_in);
int simt_linear_forward(const float*, const float*, const float*, int64_t, int64_t, int64_t, float*, cudaStream_t);
int simt_linear_dgrad(const float*, const float*, int64_t, int64_t, int64_t, float*, cudaStream_t);
int simt_linear_wgrad(const float*, const float*, int64_t, int64_t, int64_t, float*, float*, float*, cudaStream_t);
// gemm_pair.cu (K2 / K4a / dense forward on CTA pairs)
void pair_set_trace_buffer(void* ptr);
size_t pair_weight_image_bytes(int64_t d);
int pair_weight_prepare(const float*, int64_t, int, void*, int, cudaStream_t);
int pair_layer_forward(const float*, const float*, const int32_t*, const int32_t*, const void*, const float*, int64_t, int64_t, int64_t, int, float, int,
                       float, uint64_t, uint64_t, float*, float*, int, cudaStream_t);
int pair_layer_dgrad(const float*, const void*, int64_t, int64_t, float, uint64_t, uint64_t, float*, int, cudaStream_t);
int pair_dense_forward(const float*, const void*, const float*, const float*, int64_t, int64_t, float, uint64_t, uint64_t, float*, int, cudaStream_t);
// wgrad_pair.cu (K4b on a CTA pair)
size_t pair_wgrad_workspace_bytes(int64_t E, int64_t d);
void pair_wgrad_geometry(int64_t E, int64_t d, int sms, int64_t out[12]);
int pair_layer_wgrad(const float*, const float*, int64_t, int64_t, float, uint64_t, uint64_t, float*, float*, void*, size_t, int, cudaStream_t);
// wgrad_tc.cu
int tc_bias_grad(const float*, int64_t, int64_t, float, uint64_t, uint64_t, float*, void*, size_t, cudaStream_t);
size_t tc_wgrad_workspace_bytes(int64_t E, int64_t d);
int tc_layer_wgrad(const float*, const float*, const float*, const int32_t*, const int32_t*, int64_t, int64_t, int, float, float, uint64_t, uint64_t,
                   float*, float*, void*, size_t, int, cudaStream_t);

// MMA passes per product: 3 = 3xTF32, 1 = single TF32 pass, 0 = bf16 operands (one kind::f16 pass)
static int products_of(int gemm_mode) { return gemm_mode == NT_GEMM_TF32 ? 1 : gemm_mode == NT_GEMM_BF16 ? 0 : 3; }
// the weight gradient has no bf16 kernel: in bf16 mode it runs as a single TF32 pass (no less precise than bf16 operands)
static int wgrad_products_of(int gemm_mode) { return gemm_mode == NT_GEMM_TF32X3 ? 3 : 1; }

static bool tc_shape_ok(int64_t d, const void* a, const void* b, const void* c, const void* e) {
  return d % 4 == 0 && d >= 4 && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(e);
}

}  // namespace nt

using namespace nt;

extern "C" const char* nt_last_error_string(void) { return g_err; }
extern "C" int nt_version(void) { return NT_ABI_VERSION; }
extern "C" long long nt_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int nt_debug_wgrad_geometry(int64_t E, int64_t d, int num_sms_, int64_t* out12) {
  NT_CHECK_ARG(E >= 0 && E < INT32_MAX && d > 0 && d % 4 == 0 && d < (1 << 20) && num_sms_ > 0 && out12, "nt_debug_wgrad_geometry: bad arguments");
  pair_wgrad_geometry(E, d, num_sms_, out12);
  return NT_OK;
}

extern "C" void nt_debug_set_trace_buffer(void* device_u64_buffer) { pair_set_trace_buffer(device_u64_buffer); }

// host restatement of split_item (common.cuh): unsigned __int128 stands in for __umul64hi
extern "C" int nt_debug_split_item(int64_t total, int64_t chunks, int64_t t, int64_t* out3) {
  NT_CHECK_ARG(total > 0 && chunks > 0 && chunks < (1 << 20) && t >= 0 && t < total && out3, "nt_debug_split_item: bad arguments");
  const uint64_t magic = chunk_div_magic(total, (int)chunks);
  int64_t row, chunk;
  if (magic) {
    const uint32_t q = (uint32_t)(((unsigned __int128)(uint64_t)t * magic) >> 64);
    row = (int32_t)q;
    chunk = (int32_t)((uint32_t)t - q * (uint32_t)chunks);
  } else {
    row = t / chunks;
    chunk = t - row * chunks;
  }
  out3[0] = magic != 0;
  out3[1] = row;
  out3[2] = chunk;
  return NT_OK;
}

extern "C" int nt_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

#define NT_COMMON_LAYER_CHECKS(fn)                                                                        \
  if (dtype != NT_F32) { set_error(fn ": only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }       \
  NT_CHECK_ARG(d > 0 && d < (1 << 20) && E >= 0 && E < INT32_MAX, fn ": bad sizes");                      \
  NT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, fn ": dropout_p must be in [0, 1)");                  \
  NT_CHECK_ARG(gemm_mode == NT_GEMM_TF32X3 || gemm_mode == NT_GEMM_FP32 || gemm_mode == NT_GEMM_TF32 || gemm_mode == NT_GEMM_BF16, fn ": bad gemm_mode")

extern "C" size_t nt_weight_image_bytes(int64_t d) { return d > 0 ? pair_weight_image_bytes(d) : 0; }

extern "C" int nt_weight_prepare(const void* W, int64_t d, int transpose, void* image, int dtype, nt_stream_t stream) {
  NT_CHECK_ARG(dtype == NT_F32 || dtype == NT_BF16, "nt_weight_prepare: bad dtype");
  NT_CHECK_ARG(W && image && d > 0 && d < (1 << 20), "nt_weight_prepare: bad arguments");
  if (!aligned16(image)) { set_error("nt_weight_prepare: image must be 16-byte aligned"); return NT_ERR_ALIGN; }
  return pair_weight_prepare(static_cast<const float*>(W), d, transpose != 0, image, dtype == NT_BF16, as_stream(stream));
}

extern "C" int nt_layer_forward(const void* h, const void* n, const int32_t* src, const int32_t* rev, const void* W, const void* weight_image,
                                const void* bias, int64_t E, int64_t V, int64_t d, int act, float act_param, int residual, float dropout_p,
                                uint64_t seed, uint64_t offset, void* out, void* m_out, int dtype, int gemm_mode, nt_stream_t stream) {
  NT_COMMON_LAYER_CHECKS("nt_layer_forward");
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "nt_layer_forward: bad activation");
  NT_CHECK_ARG(V >= 0, "nt_layer_forward: bad V");
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(h && n && src && rev && W && out, "nt_layer_forward: null pointer");
  cudaStream_t st = as_stream(stream);
  if (gemm_mode != NT_GEMM_FP32 && tc_shape_ok(d, h, n, out, bias) && aligned16(m_out)) {
    NT_CHECK_ARG(weight_image, "nt_layer_forward: tensor-core path needs weight_image (nt_weight_prepare)");
    return pair_layer_forward(static_cast<const float*>(h), static_cast<const float*>(n), src, rev, weight_image, static_cast<const float*>(bias), E,
                              V, d, act, act_param, residual, dropout_p, seed, offset, static_cast<float*>(out), static_cast<float*>(m_out),
                              products_of(gemm_mode), st);
  }
  if (m_out) { set_error("nt_layer_forward: m_out is only produced by the tensor-core path"); return NT_ERR_UNSUPPORTED; }
  return simt_layer_forward(static_cast<const float*>(h), static_cast<const float*>(n), src, rev, static_cast<const float*>(W),
                            static_cast<const float*>(bias), E, d, act, act_param, residual, dropout_p, seed, offset, static_cast<float*>(out), st);
}

extern "C" int nt_dense_forward(const void* x, const void* weight_image, const void* bias, const void* resid, int64_t R, int64_t d, float dropout_p,
                                uint64_t seed, uint64_t offset, void* out, int dtype, int gemm_mode, nt_stream_t stream) {
  const int64_t E = R;
  NT_COMMON_LAYER_CHECKS("nt_dense_forward");
  if (R == 0) return NT_OK;
  NT_CHECK_ARG(x && out && weight_image, "nt_dense_forward: null pointer");
  if (gemm_mode == NT_GEMM_FP32 || !tc_shape_ok(d, x, out, bias, resid)) {
    set_error("nt_dense_forward: needs the tensor-core path (gemm_mode tf32x3 / tf32, d %% 4 == 0, 16-byte aligned rows)");
    return NT_ERR_UNSUPPORTED;
  }
  return pair_dense_forward(static_cast<const float*>(x), weight_image, static_cast<const float*>(bias), static_cast<const float*>(resid), R, d,
                            dropout_p, seed, offset, static_cast<float*>(out), products_of(gemm_mode), as_stream(stream));
}

// ---- prediction head: rectangular fp32 Linear on [rows, in_features] molecule vectors (notorch/nn/mlp.py:58-62) ----
#define NT_LINEAR_CHECKS(fn)                                                                                              \
  if (dtype != NT_F32) { set_error(fn ": only NT_F32 is implemented"); return NT_ERR_UNSUPPORTED; }                        \
  NT_CHECK_ARG(rows >= 0 && rows < INT32_MAX && out_features > 0 && in_features > 0 && out_features < (1 << 20) && in_features < (1 << 20), \
               fn ": bad sizes")

extern "C" int nt_linear_forward(const void* x, const void* W, const void* bias, int64_t rows, int64_t out_features, int64_t in_features, void* out,
                                 int dtype, nt_stream_t stream) {
  NT_LINEAR_CHECKS("nt_linear_forward");
  if (rows == 0) return NT_OK;
  NT_CHECK_ARG(x && W && out, "nt_linear_forward: null pointer");
  return simt_linear_forward(static_cast<const float*>(x), static_cast<const float*>(W), static_cast<const float*>(bias), rows, out_features,
                             in_features, static_cast<float*>(out), as_stream(stream));
}

extern "C" int nt_linear_backward_input(const void* g, const void* W, int64_t rows, int64_t out_features, int64_t in_features, void* gx, int dtype,
                                        nt_stream_t stream) {
  NT_LINEAR_CHECKS("nt_linear_backward_input");
  if (rows == 0) return NT_OK;
  NT_CHECK_ARG(g && W && gx, "nt_linear_backward_input: null pointer");
  return simt_linear_dgrad(static_cast<const float*>(g), static_cast<const float*>(W), rows, out_features, in_features, static_cast<float*>(gx),
                           as_stream(stream));
}

extern "C" size_t nt_linear_backward_weight_workspace_bytes(int64_t rows, int64_t out_features, int64_t in_features) {
  if (rows < 0 || out_features <= 0 || in_features <= 0) return 0;
  return simt_linear_wgrad_workspace_bytes(rows, out_features, in_features);
}

extern "C" int nt_linear_backward_weight(const void* g, const void* x, int64_t rows, int64_t out_features, int64_t in_features, void* gW, void* gb,
                                         void* workspace, size_t workspace_bytes, int dtype, nt_stream_t stream) {
  NT_LINEAR_CHECKS("nt_linear_backward_weight");
  NT_CHECK_ARG(gW, "nt_linear_backward_weight: null pointer");
  NT_CHECK_ARG(rows == 0 || (g && x), "nt_linear_backward_weight: null pointer");
  if (workspace_bytes < simt_linear_wgrad_workspace_bytes(rows, out_features, in_features) || !workspace) {
    set_error("nt_linear_backward_weight: workspace too small (nt_linear_backward_weight_workspace_bytes)");
    return NT_ERR_WORKSPACE;
  }
  return simt_linear_wgrad(static_cast<const float*>(g), static_cast<const float*>(x), rows, out_features, in_features, static_cast<float*>(gW),
                           static_cast<float*>(gb), static_cast<float*>(workspace), as_stream(stream));
}

extern "C" int nt_layer_backward_dgrad(const void* g, const void* W, const void* weight_image, int64_t E, int64_t d, float dropout_p, uint64_t seed,
                                       uint64_t offset, void* g_m, int dtype, int gemm_mode, nt_stream_t stream) {
  NT_COMMON_LAYER_CHECKS("nt_layer_backward_dgrad");
  if (E == 0) return NT_OK;
  NT_CHECK_ARG(g && W && g_m, "nt_layer_backward_dgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  if (gemm_mode != NT_GEMM_FP32 && tc_shape_ok(d, g, g_m, nullptr, nullptr)) {
    NT_CHECK_ARG(weight_image, "nt_layer_backward_dgrad: tensor-core path needs weight_image (nt_weight_prepare, transpose=1)");
    return pair_layer_dgrad(static_cast<const float*>(g), weight_image, E, d, dropout_p, seed, offset, static_cast<float*>(g_m),
                            products_of(gemm_mode), st);
  }
  return simt_layer_dgrad(static_cast<const float*>(g), static_cast<const float*>(W), E, d, dropout_p, seed, offset, static_cast<float*>(g_m), st);
}

extern "C" size_t nt_layer_backward_wgrad_workspace_bytes(int64_t E, int64_t d) {
  if (d <= 0) return 0;
  size_t simt = (size_t)simt_wgrad_splits(E, d) * (size_t)d * (size_t)(d + 1) * sizeof(float);
  size_t tcb = tc_wgrad_workspace_bytes(E, d);
  size_t pairb = pair_wgrad_workspace_bytes(E, d);
  if (pairb > tcb) tcb = pairb;
  return (simt > tcb ? simt : tcb) + 256;
}

extern "C" int nt_layer_backward_wgrad(const void* g, const void* m, const void* h, const void* n, const int32_t* src, const int32_t* rev, int64_t E, int64_t V,
                                       int64_t d, int act, float act_param, float dropout_p, uint64_t seed, uint64_t offset, void* gW, void* gb,
                                       void* workspace, size_t workspace_bytes, int dtype, int gemm_mode, nt_stream_t stream) {
  NT_COMMON_LAYER_CHECKS("nt_layer_backward_wgrad");
  NT_CHECK_ARG(act >= NT_ACT_IDENTITY && act <= NT_ACT_TANH, "nt_layer_backward_wgrad: bad activation");
  NT_CHECK_ARG(V >= 0 && gW, "nt_layer_backward_wgrad: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (E == 0) {
    NT_CUDA(cudaMemsetAsync(gW, 0, (size_t)d * d * sizeof(float), st));
    if (gb) NT_CUDA(cudaMemsetAsync(gb, 0, (size_t)d * sizeof(float), st));
    return NT_OK;
  }
  NT_CHECK_ARG(g && (m || (h && n && src && rev)), "nt_layer_backward_wgrad: null pointer");
  if (!workspace || workspace_bytes < nt_layer_backward_wgrad_workspace_bytes(E, d)) {
    set_error("nt_layer_backward_wgrad: workspace too small");
    return NT_ERR_WORKSPACE;
  }
  if (m) {  // both operands dense (m saved by K2): the CTA-pair kernel
    if (gemm_mode == NT_GEMM_FP32 || !tc_shape_ok(d, g, m, workspace, nullptr)) {
      set_error("nt_layer_backward_wgrad: a saved m needs the tensor-core path (d %% 4 == 0, 16-byte aligned)");
      return NT_ERR_UNSUPPORTED;
    }
    auto wgrad = pair_layer_wgrad;
    int rc = wgrad(static_cast<const float*>(g), static_cast<const float*>(m), E, d, dropout_p, seed, offset, static_cast<float*>(gW),
                   static_cast<float*>(gb), workspace, workspace_bytes, wgrad_products_of(gemm_mode), st);
    if (rc != NT_ERR_UNSUPPORTED) return rc;
    // no spare padded feature row for the bias gradient (d % 128 == 0 resp. d % 256 == 0): weight gradient here, column sums of g_u below
    rc = wgrad(static_cast<const float*>(g), static_cast<const float*>(m), E, d, dropout_p, seed, offset, static_cast<float*>(gW), nullptr,
               workspace, workspace_bytes, wgrad_products_of(gemm_mode), st);
    if (rc) return rc;
    return tc_bias_grad(static_cast<const float*>(g), E, d, dropout_p, seed, offset, static_cast<float*>(gb), workspace, workspace_bytes, st);
  }
  NT_CHECK_ARG(h && n && src && rev, "nt_layer_backward_wgrad: null pointer");
  if (gemm_mode != NT_GEMM_FP32 && tc_shape_ok(d, g, h, n, workspace)) {
    int rc = tc_layer_wgrad(static_cast<const float*>(g), static_cast<const float*>(h), static_cast<const float*>(n), src, rev, E, d, act, act_param,
                            dropout_p, seed, offset, static_cast<float*>(gW), static_cast<float*>(gb), workspace, workspace_bytes,
                            wgrad_products_of(gemm_mode), st);
    if (rc != NT_ERR_UNSUPPORTED) return rc;
  }
  return simt_layer_wgrad(static_cast<const float*>(g), static_cast<const float*>(h), static_cast<const float*>(n), src, rev, E, d, act, act_param,
                          dropout_p, seed, offset, static_cast<float*>(gW), static_cast<float*>(gb), static_cast<float*>(workspace), st);
}
