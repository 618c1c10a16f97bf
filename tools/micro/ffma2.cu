// Microbenchmark: FP32 multiply-add issue rate on sm_100a, scalar FFMA against the packed
// FFMA2 (fma.rn.f32x2) form, alone and interleaved 1:1 with an ALU-pipe op (FMNMX).  Decides
// whether the compute-bound kernels (observables epilogue, non-linear runs) should pair particles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/ffma2 ffma2.cu && bin/ffma2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float b, float c, int iters) {
  float2 a[CHAINS];
  float lim[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    lim[i] = 1e30f - i;
  }
  const float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (MODE == 0 || MODE == 2) {
        a[i].x = fmaf(a[i].x, b2.x, c2.x);
        a[i].y = fmaf(a[i].y, b2.y, c2.y);
      } else {
        a[i] = __ffma2_rn(a[i], b2, c2);
      }
      if (MODE == 2) {
        lim[i] = fminf(lim[i], a[i].x);  // 2 ALU ops per 2 FFMA
        lim[i] = fmaxf(lim[i], a[i].y);
      } else if (MODE == 3) {
        lim[i] = fminf(lim[i], a[i].x);  // 2 ALU ops per FFMA2
        lim[i] = fmaxf(lim[i], a[i].y);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y + lim[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int sms, double clock_ghz) {
  const int iters = 4096, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 0.999f, 0.001f, iters);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(out, 0.999f, 0.001f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double fma = double(blocks) * 256 * iters * CHAINS * 2;
  printf("%-28s %8.3f ms  %7.2f T multiply-adds/s  %6.1f per SM per clock (at %.2f GHz)\n", name, ms,
         fma / ms / 1e9, fma / (ms * 1e-3) / sms / (clock_ghz * 1e9), clock_ghz);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float* out;
  cudaMalloc(&out, sizeof(float) * prop.multiProcessorCount * 8 * 256);
  run<0>("FFMA", out, prop.multiProcessorCount, ghz);
  run<1>("FFMA2", out, prop.multiProcessorCount, ghz);
  run<2>("FFMA + FMNMX (1:1)", out, prop.multiProcessorCount, ghz);
  run<3>("FFMA2 + 2 FMNMX", out, prop.multiProcessorCount, ghz);
  return 0;
}
