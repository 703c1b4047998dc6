// Microbenchmark: FP32 FMA throughput with scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float x, float y) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ unsigned long long f2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float f1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE> __global__ void kern(float* out, int iters, float s) {
  float a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = f1(a[i], s, 0.5f);
  } else {
    unsigned long long p[4], ss = pk(s, s), h = pk(0.5f, 0.5f);
    for (int i = 0; i < 4; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = f2(p[i], ss, h);
    for (int i = 0; i < 4; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); a[2 * i] = x; a[2 * i + 1] = y; }
  }
  float t = 0; for (int i = 0; i < 8; ++i) t += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) kern<0><<<148 * 8, 256>>>(out, iters, 0.999f); else kern<1><<<148 * 8, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 8 * 256 * iters * 32.0;
    printf("mode %s: %.3f ms  %.2f TFLOP/s (fp32, 2 flop/fma)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9);
  }
  return 0;
}
