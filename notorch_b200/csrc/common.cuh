// Shared device/host helpers for libnotorch_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/notorch_b200.h"

namespace nt {

// ---- error plumbing (thread-local message; no exceptions across the ABI) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define NT_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      nt::set_error(__VA_ARGS__);               \
      return NT_ERR_ARG;                        \
    }                                           \
  } while (0)

#define NT_CUDA(call)                                          \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return nt::cuda_fail(e__, #call);  \
  } while (0)

// checks the launch(es) just issued and adds `n` to the process-wide kernel-launch counter
#define NT_LAUNCH_CHECK(name, n)                                    \
  do {                                                              \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return nt::cuda_fail(e__, name);        \
    nt::count_launches(n);                                          \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline cudaStream_t as_stream(nt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();
void count_launches(int n);

// Function attributes (the > 48 KiB dynamic shared-memory opt-in) are PER DEVICE: a process-wide once-flag would leave every
// GPU but the first without it. One bit per device ordinal; devices >= 64 simply repeat the (cheap) call.
struct PerDeviceOnce {
  std::atomic<uint64_t> done{0};
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
    e = f();
    if (e == cudaSuccess && dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return e;
  }
};

// ---- activations (closed set compiled into every kernel; chemprop.py:17,24,37) -------------
__device__ __forceinline__ float act_fwd(float x, int act, float p) {
  switch (act) {
    case NT_ACT_RELU: return x < 0.f ? 0.f : x;  // NaN propagates like torch.relu
    case NT_ACT_LEAKY_RELU: return x > 0.f ? x : x * p;
    case NT_ACT_ELU: return x > 0.f ? x : p * expm1f(x);
    case NT_ACT_SILU: return x / (1.f + expf(-x));
    case NT_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case NT_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ float act_bwd(float x, int act, float p) {  // d act / d x at x
  switch (act) {
    case NT_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case NT_ACT_LEAKY_RELU: return x > 0.f ? 1.f : p;
    case NT_ACT_ELU: return x > 0.f ? 1.f : p * expf(x);
    case NT_ACT_SILU: {
      float s = 1.f / (1.f + expf(-x));
      return s * (1.f + x * (1.f - s));
    }
    case NT_ACT_GELU: {
      float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
      float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case NT_ACT_TANH: {
      float t = tanhf(x);
      return 1.f - t * t;
    }
    default: return 1.f;
  }
}

__device__ __forceinline__ float4 act_fwd4(float4 v, int act, float p) {
  return make_float4(act_fwd(v.x, act, p), act_fwd(v.y, act, p), act_fwd(v.z, act, p), act_fwd(v.w, act, p));
}

// ---- Philox4x32-10 (counter-based; mask is a pure function of (seed, offset, element id)) ---
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep-probability threshold: keep iff u32 >= thr, thr = p * 2^32
__device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

// Scale factors for elements [idx, idx+4) where idx % 4 == 0: 0 (dropped) or 1/(1-p).
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thr, float inv_keep) {
  uint4 r = philox4x32(seed, offset, idx >> 2);
  return make_float4(r.x >= thr ? inv_keep : 0.f, r.y >= thr ? inv_keep : 0.f,
                     r.z >= thr ? inv_keep : 0.f, r.w >= thr ? inv_keep : 0.f);
}

__device__ __forceinline__ float dropout_scale1(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thr, float inv_keep) {
  uint4 r = philox4x32(seed, offset, idx >> 2);
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return v >= thr ? inv_keep : 0.f;
}

// ---- streaming 128-bit global access ---------------------------------------------------------

// (row, 16-byte chunk) of a flattened work item t = row * chunks + chunk without the emulated 64-bit division (~70 instructions, twice
// per thread in the segmented reductions: ncu showed these HBM kernels half issue-bound). magic = ceil(2^64 / chunks) makes
// umul64hi(t, magic) == t / chunks exactly for every t < 2^32 (error term t * r / (chunks * 2^64) < 2^-32 < 1 / chunks); the host passes
// magic = 0 (general path) when the item count does not fit in 32 bits or chunks == 1.
__host__ inline uint64_t chunk_div_magic(int64_t total, int chunks) {
  return (chunks > 1 && total < (int64_t(1) << 32)) ? (~uint64_t(0)) / (uint64_t)chunks + 1 : 0;
}
__device__ __forceinline__ void split_item(int64_t t, int chunks, uint64_t magic, int& row, int& chunk) {
  if (magic) {
    const uint32_t q = (uint32_t)__umul64hi((uint64_t)t, magic);
    row = (int)q;
    chunk = (int)((uint32_t)t - q * (uint32_t)chunks);
  } else {
    row = (int)(t / chunks);
    chunk = (int)(t - (int64_t)row * chunks);
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stg4_stream(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace nt
