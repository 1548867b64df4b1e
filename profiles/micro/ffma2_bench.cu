// Packed fp32x2 FMA (FFMA2, sm_100) against scalar FFMA: issue slots vs FMA-pipe lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
// Variants (same number of fp32 FMAs per thread in A and B):
//   A  16 independent scalar FFMA chains
//   B   8 independent FFMA2 chains (16 fp32 FMAs)
//   C  A + 8 LOP3 per 16 FMAs: 24 issue slots, fma pipe 16 cycles, alu pipe 16 cycles
//   D  B + 8 LOP3 per 16 FMAs: 16 issue slots if an FFMA2 takes one
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float m, float c, unsigned salt) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack(acc[2 * i], acc[2 * i + 1]);
  const u64 mm = pack(m, m), cc = pack(c, c);
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], m, c);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = (h[i] & salt) ^ (unsigned)it;  // one LOP3 each (alu pipe, 8 of them = 16 pipe cycles)
    }
  }
  float s = 0.f;
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a, b;
      unpack(p[i], a, b);
      s += a + b;
    }
  }
  unsigned hs = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) hs += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)hs;
}

// dependent chains: one warp, 1 or 2 independent chains, cycles per dependent instruction (latency)
template <int MODE>
__global__ void latency(float* out, long long* cycles, int iters, float m, float c) {
  float a = threadIdx.x * 0.001f, b = a + 1.0f;
  u64 p = pack(a, b);
  const u64 mm = pack(m, m), cc = pack(c, c);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a = fmaf(a, m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) p = fma2(p, mm, cc);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // two scalar chains = the work of one packed chain
        a = fmaf(a, m, c);
        b = fmaf(b, m, c);
      }
    }
  }
  long long t1 = clock64();
  float x, y;
  unpack(p, x, y);
  out[threadIdx.x] = a + b + x + y;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
void run_latency(const char* name, float* out) {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  latency<MODE><<<1, 32>>>(out, cyc, iters, 0.999f, 0.001f);
  latency<MODE><<<1, 32>>>(out, cyc, iters, 0.999f, 0.001f);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %6.2f cycles per step of the chain\n", name, (double)h / (16.0 * iters));
  cudaFree(cyc);
}

template <int MODE>
void run(const char* name, float* out, int blocks) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bench<MODE><<<blocks, 256>>>(out, 100, 0.999f, 0.001f, 0x9e3779b9u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f, 0x9e3779b9u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)blocks * 256 * 16.0 * iters;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-40s %8.3f ms  %7.2f TFLOP/s fp32 (FMA = 2 flop)  ~%.1f cycles per warp-iteration at %d MHz\n", name, ms,
         2.0 * fmas / ms / 1e9, ms * 1e-3 * clk * 1e3 / ((double)blocks / 148 * 8 / 4 * iters), clk / 1000);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8;
  float* out;
  cudaMalloc(&out, (size_t)blocks * 256 * 4);
  run<0>("A 16 x FFMA", out, blocks);
  run<1>("B  8 x FFMA2", out, blocks);
  run<2>("C 16 x FFMA  + 8 LOP3", out, blocks);
  run<3>("D  8 x FFMA2 + 8 LOP3", out, blocks);
  run_latency<0>("latency: FFMA chain", out);
  run_latency<1>("latency: FFMA2 chain", out);
  run_latency<2>("latency: two interleaved FFMA chains", out);
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
