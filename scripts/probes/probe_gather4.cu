// Probe: semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (tensor-map box shape, swizzle, OOB fill).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o notorch_b200/_build/probe_gather4 scripts/probes/probe_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, const int* idx, int col0, float* out, int nrows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar), s = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -777.f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nrows * 128));
  __syncthreads();
  if (threadIdx.x < nrows / 4) {
    const int l = threadIdx.x;
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(s + 512 * l), "l"(&map), "r"(col0), "r"(idx[4 * l]), "r"(idx[4 * l + 1]), "r"(idx[4 * l + 2]), "r"(idx[4 * l + 3]), "r"(b) : "memory");
  }
  // bounded wait
  uint32_t ok = 0;
  for (long long it = 0; it < 20000000 && !ok; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
  if (threadIdx.x == 0 && !ok) printf("TIMEOUT waiting for gather4 bytes\n");
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
  const int R = 1000, d = 300, nrows = 128;
  std::vector<float> h((size_t)R * d);
  for (int r = 0; r < R; ++r) for (int c = 0; c < d; ++c) h[(size_t)r * d + c] = r * 1000.f + c;
  std::vector<int> idx(nrows);
  for (int i = 0; i < nrows; ++i) idx[i] = (i * 37 + 11) % R;
  idx[5] = R + 3;  // out-of-range row: expect zero fill (or whatever the hardware does)
  float *dh, *dout; int* didx;
  cudaMalloc(&dh, h.size() * 4); cudaMalloc(&dout, nrows * 128); cudaMalloc(&didx, nrows * 4);
  cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(didx, idx.data(), nrows * 4, cudaMemcpyHostToDevice);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int boxrows : {1, 4}) {
    for (int col0 : {32, 288}) {
      CUtensorMap map;
      cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)d * 4};
      cuuint32_t box[2] = {32u, (cuuint32_t)boxrows}; cuuint32_t estr[2] = {1u, 1u};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("box rows %d col0 %d: encode -> %d\n", boxrows, col0, (int)r);
      if (r != CUDA_SUCCESS) continue;
      cudaMemset(dout, 0, nrows * 128);
      probe<<<1, 128, 16384>>>(map, didx, col0, dout, nrows);
      cudaError_t e = cudaDeviceSynchronize();
      printf("  kernel -> %s\n", cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      std::vector<float> o(nrows * 32);
      cudaMemcpy(o.data(), dout, nrows * 128, cudaMemcpyDeviceToHost);
      // expected: row r at byte r*128 within 1 KiB atoms, 16-byte chunk c stored at chunk (c ^ (r & 7))
      int bad = 0, untouched = 0;
      for (int r2 = 0; r2 < nrows; ++r2) for (int c = 0; c < 32; ++c) {
        const int chunk = c / 4, phys = (r2 / 8) * 256 + (r2 % 8) * 32 + ((chunk ^ (r2 % 8)) * 4) + c % 4;
        const int col = col0 + c;
        float want = (idx[r2] < R && col < d) ? idx[r2] * 1000.f + col : 0.f;
        if (o[phys] == -777.f) ++untouched;
        if (o[phys] != want) { if (bad < 6) printf("  mismatch row %d col %d: got %.1f want %.1f\n", r2, c, o[phys], want); ++bad; }
      }
      printf("  => %d mismatches, %d untouched of %d\n", bad, untouched, nrows * 32);
    }
  }
  return 0;
}
