// fp64_latency.cu -- dependent-issue latency and per-warp issue interval of DFMA / DADD /
// MUFU.RCP64H-based division on the device, measured with clock64() from one warp per SM
// sub-partition.  Explains the per-warp cycles-per-instruction the one-thread-per-problem
// sweeps run at.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ddiv_chain(double* out, long long* cyc, int iters, double a) {
  double x = threadIdx.x + 1.5;
  long long t0 = clock64();
  for (int k = 0; k < iters; ++k) x = a / x + 1.0;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
static void run(const char* name, int warps) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 32 * 64 * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 1024);
  const int iters = 4096;
  dfma_chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  dfma_chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"test\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"cycles_per_dfma_per_warp\": %.2f}\n",
         name, ILP, warps, (double)h / (iters * ILP));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1>("dfma_dependent", 1);
  run<2>("dfma", 1);
  run<4>("dfma", 1);
  run<8>("dfma", 1);
  run<16>("dfma", 1);
  run<8>("dfma", 4);    // one warp per sub-partition
  run<8>("dfma", 8);    // two warps per sub-partition (the occupancy of the FP64 sweeps)
  run<8>("dfma", 16);
  run<1>("dfma_dependent", 8);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 32 * 64);
  cudaMalloc(&cyc, sizeof(long long) * 8);
  ddiv_chain<<<1, 32>>>(out, cyc, 1024, 3.0);
  ddiv_chain<<<1, 32>>>(out, cyc, 1024, 3.0);
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"test\": \"ddiv_dependent(+dadd)\", \"cycles_per_op\": %.1f}\n", (double)h / 1024);
  return 0;
}
