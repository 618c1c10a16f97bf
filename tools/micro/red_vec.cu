// Microbenchmark: L2 reduction throughput for scalar / v2 / v4 float atomics on random addresses
// of an L2-resident grid (decides whether a 2x2-block deposit layout pays off).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>
__global__ void k(float* grid, unsigned cells, int per_thread) {
  unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < per_thread; ++i) {
    unsigned h = hash(id * 977u + i);
    unsigned c = (h % (cells / 4)) * 4;  // 16-byte aligned group of 4 cells
    if (MODE == 1) {
      atomicAdd(grid + c, 1.f); atomicAdd(grid + c + 1, 1.f);
      atomicAdd(grid + c + 2, 1.f); atomicAdd(grid + c + 3, 1.f);
    } else if (MODE == 2) {
      atomicAdd(reinterpret_cast<float2*>(grid + c), make_float2(1.f, 1.f));
      atomicAdd(reinterpret_cast<float2*>(grid + c + 2), make_float2(1.f, 1.f));
    } else {
      atomicAdd(reinterpret_cast<float4*>(grid + c), make_float4(1.f, 1.f, 1.f, 1.f));
    }
  }
}

template <int MODE>
float run(float* grid, unsigned cells, int blocks, int per_thread) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256>>>(grid, cells, per_thread);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(grid, cells, per_thread);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

int main() {
  const unsigned cells = 1u << 20;  // 4 MB of floats
  float* grid; cudaMalloc(&grid, cells * sizeof(float)); cudaMemset(grid, 0, cells * sizeof(float));
  const int blocks = 148 * 16, per_thread = 16;
  const double groups = double(blocks) * 256 * per_thread;
  float t1 = run<1>(grid, cells, blocks, per_thread);
  float t2 = run<2>(grid, cells, blocks, per_thread);
  float t4 = run<4>(grid, cells, blocks, per_thread);
  printf("groups of 4 cells per launch: %.0f\n", groups);
  printf("4 x scalar: %.3f ms  %.1f G groups/s\n", t1, groups / t1 / 1e6);
  printf("2 x v2    : %.3f ms  %.1f G groups/s\n", t2, groups / t2 / 1e6);
  printf("1 x v4    : %.3f ms  %.1f G groups/s\n", t4, groups / t4 / 1e6);
  return 0;
}
