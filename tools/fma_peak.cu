// FP64 / FP32 FMA throughput of the device ("measure it with an FMA
// micro-kernel and record it beside the HBM figure").  8 independent chains per thread.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fma_peak tools/fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <class S>
__global__ void fma_kernel(S* out, int iters, S a, S b) {
  S x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (S)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  S s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == (S)12345.678) out[0] = s;   // keep the chains alive
}

template <class S>
double run(const char* name, int warps_per_sm) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  S* out;
  cudaMalloc(&out, sizeof(S));
  const int threads = 32 * warps_per_sm, blocks = sms, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fma_kernel<S><<<blocks, threads>>>(out, iters, (S)0.999, (S)0.001);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  fma_kernel<S><<<blocks, threads>>>(out, iters, (S)0.999, (S)0.001);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
  const double tf = flops / (ms * 1e-3) / 1e12;
  printf("{\"dtype\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", name,
         warps_per_sm, ms, tf);
  cudaFree(out);
  return tf;
}

int main() {
  for (int w : {4, 8, 16, 32}) run<double>("f64", w);
  for (int w : {4, 8, 16, 32}) run<float>("f32", w);
  return 0;
}
