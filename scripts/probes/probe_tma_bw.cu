// Probe: per-SM copy-engine throughput on B200 for the transfer shapes the GEMM kernels use, all 148 SMs running.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o notorch_b200/_build/probe_tma_bw scripts/probes/probe_tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok && clock64() - t0 < 2000000000LL) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
constexpr int STAGE = 38912;  // bytes reserved per stage
// method 0: one bulk copy of `bytes` from wsrc (+ iteration offset, wrapping in wbytes)         [W stream, first-generation]
// method 1: 32 lanes x bulk copies of bytes/32
// method 2: 2-D box {32, 128} fp32 (16 KB) of the [rows, d] tensor at distinct rows per CTA       [dgrad A]
// method 3: 32 x gather4 (16 KB) random rows                                                      [K2 A], x2 when bytes == 32768
// method 4: 2-D box {32, rows_box} of the W image viewed as [n, 32] fp32
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap map, const uint8_t* wsrc, int wbytes, const int* idx, int nidx, int method,
                                               int bytes, int stages, int iters, long long rows, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t s0 = smem_u32(smem), b0 = smem_u32(bars);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    // producer: keeps `stages` transfers in flight; consumer (same warp) waits in order
    for (int it = 0; it < iters + stages; ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(b0 + 8 * s, ((it - stages) / stages) & 1);
      if (it < iters) {
        const uint32_t dst = s0 + s * STAGE, bar = b0 + 8 * s;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        __syncwarp();
        const long long base_row = ((long long)blockIdx.x * iters + it) * 128 % (rows - 128);
        if (method == 0) {
          if (lane == 0) {
            const uint8_t* src = wsrc + ((long long)it * bytes) % (wbytes - bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
          }
        } else if (method == 1) {
          const int part = bytes / 32;
          const uint8_t* src = wsrc + ((long long)it * bytes) % (wbytes - bytes) + lane * part;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + lane * part), "l"(src), "r"(part), "r"(bar) : "memory");
        } else if (method == 2) {
          if (lane == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(&map), "r"((it % 9) * 32), "r"((int)base_row), "r"(bar) : "memory");
        } else if (method == 3) {
          const int* ip = idx + ((blockIdx.x * 131 + it * 128 + lane * 4) % (nidx - 4));
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       ::"r"(dst + 512 * lane), "l"(&map), "r"((it % 9) * 32), "r"(ip[0]), "r"(ip[1]), "r"(ip[2]), "r"(ip[3]), "r"(bar) : "memory");
          if (bytes == 32768) {
            const int* iq = idx + ((blockIdx.x * 977 + it * 128 + lane * 4 + 64) % (nidx - 4));
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(dst + 16384 + 512 * lane), "l"(&map), "r"((it % 9) * 32), "r"(iq[0]), "r"(iq[1]), "r"(iq[2]), "r"(iq[3]), "r"(bar) : "memory");
          }
        } else if (method == 4) {
          if (lane == 0) {
            const int nrows = bytes / 128;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(&map), "r"(0), "r"((it * nrows) % (int)(rows - nrows)), "r"(bar) : "memory");
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
// NW warps share the 64 gather4 instructions of one 32 KB stage (method 0) or issue one elected bulk copy each (method 1)
__global__ void __launch_bounds__(256, 1) probe_mw(const __grid_constant__ CUtensorMap map, const uint8_t* wsrc, int wbytes, const int* idx, int nidx, int method,
                                                   int stages, int iters, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t s0 = smem_u32(smem), b0 = smem_u32(bars);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  long long t0 = clock64();
  const int per_warp = 64 / nw;
  for (int it = 0; it < iters + stages; ++it) {
    const int s = it % stages;
    if (it >= stages) mbar_wait(b0 + 8 * s, ((it - stages) / stages) & 1);
    if (it < iters) {
      const uint32_t dst = s0 + s * STAGE, bar = b0 + 8 * s;
      if (method == 0) {
        if (warp == 0 && elect_one()) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768) : "memory");
        for (int g = warp * per_warp + lane; g < (warp + 1) * per_warp; g += 32) {
          const int* ip = idx + ((blockIdx.x * 131 + it * 256 + g * 4) % (nidx - 4));
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                       ::"r"(dst + 512 * g), "l"(&map), "r"((it % 9) * 32), "r"(ip[0]), "r"(ip[1]), "r"(ip[2]), "r"(ip[3]), "r"(bar) : "memory");
        }
      } else {
        if (warp == 0 && elect_one()) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(19456) : "memory");
          const uint8_t* src = wsrc + (it & 31) * 19456;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(19456), "r"(bar) : "memory");
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const long long E = 205166; const int d = 300;
  float* act; cudaMalloc(&act, E * d * 4); cudaMemset(act, 0, E * d * 4);
  const int wbytes = 778240 * 2; uint8_t* w; cudaMalloc(&w, wbytes); cudaMemset(w, 0, wbytes);
  const int nidx = 1 << 20; std::vector<int> hidx(nidx);
  uint32_t x = 12345; for (int i = 0; i < nidx; ++i) { x = x * 1664525u + 1013904223u; hidx[i] = (x >> 8) % E; }
  // K2-like locality: indices mostly near i/2.18 (edges of one molecule touch nearby atoms)
  std::vector<int> hloc(nidx); for (int i = 0; i < nidx; ++i) hloc[i] = (int)(((long long)(i / 2) + (hidx[i] % 30)) % E);
  int *didx, *dloc; cudaMalloc(&didx, nidx * 4); cudaMalloc(&dloc, nidx * 4);
  cudaMemcpy(didx, hidx.data(), nidx * 4, cudaMemcpyHostToDevice); cudaMemcpy(dloc, hloc.data(), nidx * 4, cudaMemcpyHostToDevice);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  auto mk = [&](void* base, cuuint64_t cols, cuuint64_t rows, cuuint32_t bc, cuuint32_t br, CUtensorMapSwizzle sw) {
    CUtensorMap m; cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols * 4}; cuuint32_t box[2] = {bc, br}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode failed %d\n", (int)r);
    return m;
  };
  CUtensorMap map_tile = mk(act, d, E, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), map_g4 = mk(act, d, E, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap map_w152 = mk(w, 32, wbytes / 128, 32, 152, CU_TENSOR_MAP_SWIZZLE_NONE), map_w76 = mk(w, 32, wbytes / 128, 32, 76, CU_TENSOR_MAP_SWIZZLE_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * STAGE);
  struct Case { const char* name; int method, bytes, stages; const CUtensorMap* map; const int* idx; long long rows; };
  Case cases[] = {
    {"bulk 1 x 19456 B st1", 0, 19456, 1, &map_tile, didx, E},
    {"bulk 1 x 19456 B st2", 0, 19456, 2, &map_tile, didx, E},
    {"bulk 1 x 19456 B st3", 0, 19456, 3, &map_tile, didx, E},
    {"bulk 1 x 19456 B st5", 0, 19456, 5, &map_tile, didx, E},
    {"bulk 1 x 38912 B st1", 0, 38912, 1, &map_tile, didx, E},
    {"bulk 1 x 38912 B st5", 0, 38912, 5, &map_tile, didx, E},
    {"bulk 1 x 4864 B st1 ", 0, 4864, 1, &map_tile, didx, E},
    {"bulk 1 x 4864 B st5 ", 0, 4864, 5, &map_tile, didx, E},
    {"bulk 32 x 608 B st1 ", 1, 19456, 1, &map_tile, didx, E},
    {"bulk 32 x 608 B st5 ", 1, 19456, 5, &map_tile, didx, E},
    {"box 32x128 [E,300] st1", 2, 16384, 1, &map_tile, didx, E},
    {"box 32x128 [E,300] st2", 2, 16384, 2, &map_tile, didx, E},
    {"box 32x128 [E,300] st5", 2, 16384, 5, &map_tile, didx, E},
    {"gather4 x32 random st1", 3, 16384, 1, &map_g4, didx, E},
    {"gather4 x32 random st5", 3, 16384, 5, &map_g4, didx, E},
    {"gather4 x64 local st1 ", 3, 32768, 1, &map_g4, dloc, E},
    {"gather4 x64 local st5 ", 3, 32768, 5, &map_g4, dloc, E},
  };
  const int iters = 400;
  for (auto& c : cases) {
    if (c.method >= 0) continue;
    for (int grid : {1, 148}) {
      probe<<<grid, 64, c.stages * STAGE>>>(*c.map, w, wbytes, c.idx, nidx, c.method, c.bytes, c.stages, iters, c.rows, cyc);
      cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      std::vector<unsigned long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      printf("%s grid %3d: %7.1f B/clk/SM  (%6.0f clk per transfer)\n", c.name, grid, (double)c.bytes * iters / mx, mx / iters);
    }
  }
  cudaFuncSetAttribute(probe_mw, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * STAGE);
  for (int method : {0, 1}) for (int nw : {1, 2, 4, 8}) for (int stages : {1, 3, 5}) {
    if (method == 1 && nw > 1) continue;
    probe_mw<<<148, 32 * nw, stages * STAGE>>>(map_g4, w, wbytes, dloc, nidx, method, stages, iters, cyc);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe_mw: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const int bytes = method == 0 ? 32768 : 19456;
    printf("%s warps %d stages %d: %7.1f B/clk/SM (%6.0f clk per stage)\n", method == 0 ? "gather4 x64 local (32 KB)" : "elected bulk 19456 B     ", nw, stages, (double)bytes * iters / mx, mx / iters);
  }
  return 0;
}
