// Microbenchmark (dev tool): per-SM global store / load rate of the GEMM epilogue's access patterns.
//   A: row-per-thread (thread r owns 128 B of row r: 4 x 256-bit accesses, every instruction touches 32 lines)
//   B: line-coalesced   (each instruction covers 4 rows x 128 B: lane l -> row l/8, 16-byte chunk l%8)
//   C: B preceded by the shared-memory transpose that turns a row-per-thread register tile into pattern B
// 8 warps per CTA, 1 CTA per SM, rows 2 KB apart (fp32 N = 512), working set = rows x 2 KB.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
               "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)) : "memory");
}
template <int PAT, bool LOAD>
__global__ void __launch_bounds__(256, 1) k(float* buf, int64_t rows, int ld, int iters, float* sink) {
  __shared__ float4 sm[8][32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc = make_float4(0, 0, 0, 0);
  float4 v0 = make_float4(lane, 1, 2, 3), v1 = make_float4(4, 5, 6, 7);
  for (int it = 0; it < iters; ++it) {
    // sub-tile = 32 rows x 32 fp32 columns; walk them like the epilogue: (row block, column block)
    const int64_t sub = ((int64_t)it * gridDim.x + blockIdx.x) * 8 + warp;
    const int64_t nsub_c = ld / 32;
    const int64_t rb = (sub / nsub_c) % (rows / 32), cb = sub % nsub_c;
    float* base = buf + rb * 32 * ld + cb * 32;
    if (PAT == 0) {
      float* p = base + (int64_t)lane * ld;
      if (LOAD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float4 t = *reinterpret_cast<const float4*>(p + 4 * j); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) st256(p + 8 * j, v0, v1);
      }
    } else {
      if (PAT == 2 && !LOAD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[warp][lane * 8 + (j ^ (lane & 7))] = v0;
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + (lane >> 3), c = lane & 7;
        float* p = base + (int64_t)r * ld + c * 4;
        if (LOAD) {
          float4 t = *reinterpret_cast<const float4*>(p);
          if (PAT == 2) sm[warp][r * 8 + (c ^ (r & 7))] = t; else { acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        } else {
          float4 t = PAT == 2 ? sm[warp][r * 8 + (c ^ (r & 7))] : v0;
          *reinterpret_cast<float4*>(p) = t;
        }
      }
      if (PAT == 2 && LOAD) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) { float4 t = sm[warp][lane * 8 + (j ^ (lane & 7))]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
      }
      __syncwarp();
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}
template <int PAT, bool LOAD> void run(const char* name, float* buf, int64_t rows, int ld, float* sink) {
  const int iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PAT, LOAD><<<148, 256>>>(buf, rows, ld, 8, sink);
  cudaEventRecord(e0);
  k<PAT, LOAD><<<148, 256>>>(buf, rows, ld, iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 148.0 * 8 * iters * 4096;
  printf("%-34s rows=%7lld  %8.1f us  %7.1f GB/s  %6.1f B/clk/SM (1.9 GHz)  err=%s\n", name, (long long)rows, ms * 1e3, bytes / ms / 1e6, bytes / 148 / (ms * 1e-3 * 1.9e9), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* sink; cudaMalloc(&sink, 4);
  for (int64_t rows : {int64_t(8192), int64_t(262144)}) {     // 16 MB (L2 resident) and 512 MB (HBM)
    const int ld = 512;
    float* buf; cudaMalloc(&buf, rows * ld * 4); cudaMemset(buf, 0, rows * ld * 4);
    run<0, false>("store A row-per-thread", buf, rows, ld, sink);
    run<1, false>("store B line-coalesced", buf, rows, ld, sink);
    run<2, false>("store C smem transpose + B", buf, rows, ld, sink);
    run<0, true>("load  A row-per-thread", buf, rows, ld, sink);
    run<1, true>("load  B line-coalesced", buf, rows, ld, sink);
    run<2, true>("load  C B + smem transpose", buf, rows, ld, sink);
    cudaFree(buf);
  }
  return 0;
}
