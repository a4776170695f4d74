// FFMA vs FFMA2 (fma.rn.f32x2, sm_100a) issue throughput: is the packed form one issue slot for two FMAs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_bench tools/ffma2_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float &x, float &y, float a0, float a1) {
  uint64_t v, h, a;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(h) : "f"(0.5f));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(h), "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
template <int PACKED>
__global__ void k(float *o, const float *in, int n) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x + i) & 1023];
  float a0 = in[5], a1 = in[6];
  for (int t = 0; t < n; ++t) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (PACKED) fma2(v[i], v[i + 1], a0, a1);
      else { v[i] = __fmaf_rn(v[i], 0.5f, a0); v[i + 1] = __fmaf_rn(v[i + 1], 0.5f, a1); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *in, *o;
  cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&o, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n = 20000;
  for (int warps = 1; warps <= 8; warps *= 2)
    for (int packed = 0; packed < 2; ++packed) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (packed) k<1><<<148, 128 * warps>>>(o, in, n); else k<0><<<148, 128 * warps>>>(o, in, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fmas = 148.0 * 128 * warps * 32.0 * n;
      printf("warps/scheduler=%d %s: %.3f ms  %.2f TFMA/s (lane-FMAs)\n", warps, packed ? "FFMA2" : "FFMA ", ms, fmas / ms / 1e9);
    }
  return 0;
}
