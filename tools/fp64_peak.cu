// fp64_peak.cu -- measures the FP64 roofline denominators on this GPU:
//   dfma : dependent-chain-free DFMA stream (CUDA-core FP64 pipe)
//   dmma : mma.sync.m8n8k4.f64 stream (FP64 tensor path)
//   mix  : both in the same warps
// Prints one JSON line.  Timing: CUDA events, best of 5 after warm-up.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_mix(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; acc[i] = i * 0.5; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      dmma(c0[i], c1[i], a, b);
      acc[i] = fma(acc[i], a, b);
      acc[i] = fma(acc[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float best_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  const int threads = 256, blocksPerSM = 4, iters = 20000;
  constexpr int ILP = 8;
  int blocks = sms * blocksPerSM;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  double nThreads = (double)blocks * threads;
  float tf = best_ms([&] { k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  float tm = best_ms([&] { k_dmma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  float tx = best_ms([&] { k_mix<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  double fmaF = nThreads * ILP * iters;                      // DFMA count
  double fmaM = (nThreads / 32) * ILP * (double)iters * 256;  // 8x8x4 FMA per warp mma
  double fmaX = fmaM + 2 * fmaF;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, "
         "\"mix_tflops\": %.3f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, \"mix_ms\": %.3f}\n",
         p.name, sms, 2 * fmaF / tf * 1e-9, 2 * fmaM / tm * 1e-9, 2 * fmaX / tx * 1e-9, tf, tm,
         tx);
  cudaFree(out);
  return 0;
}
