// Throughput of the special-function (XU) pipe and of packed fp32 math on one B200, per SM and clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int OP> __device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 5) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 6) { y = fmaf(x, 1.0001f, 0.5f); }
  else if (OP == 7) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x), "f"(x + 1.f)); y = __uint_as_float(u << 16); }
  else { float2 a = make_float2(x, x + 1.f); a = __ffma2_rn(a, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); y = a.x + a.y; }
  return y;
}

template <int OP> __global__ void k(float* out, int iters) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i) + 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, float* out, int sms, float mhz) {
  const int iters = 4096, threads = 1024, blocks = sms * 2;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(out, 16);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%-14s %8.3f ms  %7.2f lane-ops / clk / SM (at %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / sms / (mhz * 1e6), mhz);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float mhz = khz / 1000.f;
  float* out;
  cudaMalloc(&out, sizeof(float) * 1024 * p.multiProcessorCount * 2);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("ex2.approx", out, p.multiProcessorCount, mhz);
  run<1>("rcp.approx", out, p.multiProcessorCount, mhz);
  run<2>("tanh.approx", out, p.multiProcessorCount, mhz);
  run<3>("rsqrt.approx", out, p.multiProcessorCount, mhz);
  run<4>("lg2.approx", out, p.multiProcessorCount, mhz);
  run<5>("sin.approx", out, p.multiProcessorCount, mhz);
  run<6>("ffma", out, p.multiProcessorCount, mhz);
  run<7>("cvt.bf16x2", out, p.multiProcessorCount, mhz);
  run<8>("ffma2 (+fadd)", out, p.multiProcessorCount, mhz);
  return 0;
}
