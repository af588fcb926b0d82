// cost_glue.cu -- the step either side of the solve in the imitation-learning loop:
// il_env.py:159-162 tiles the diagonal cost (q, p) into the dense API tensors
// C[T,B,n,n], c[T,B,n] with `.repeat`, and autograd later sums the dense gradients
// dC, dc back to (dq, dp).  Both are pure HBM streams; here each is one pass at
// copy bandwidth (16-byte stores / one read of dC,dc), deterministic summation order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dilqr.h"

namespace {

constexpr int kRedBlocks = 148 * 4;   // partial-sum rows of the gradient reduction
constexpr int kRedThreads = 256;

// C[r][i][j] = (i == j) q[i], r over T*B rows; one 16-byte store per thread-iteration.
template <class S>
__global__ void tile_C_kernel(const S* __restrict__ q, S* __restrict__ C, int n, size_t total) {
  constexpr int V = 16 / sizeof(S);
  extern __shared__ unsigned char smem_raw[];
  S* row = reinterpret_cast<S*>(smem_raw);       // one n*n row image, repeated access
  const int nn = n * n;
  for (int e = threadIdx.x; e < nn; e += blockDim.x) row[e] = (e / n == e % n) ? q[e / n] : S(0);
  __syncthreads();
  const size_t nvec = total / V;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const size_t base = v * V;
    int e = (int)(base % (size_t)nn);
    S vals[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      vals[k] = row[e];
      e = (e + 1 == nn) ? 0 : e + 1;
    }
    if constexpr (sizeof(S) == 8)
      reinterpret_cast<double2*>(C)[v] = make_double2(vals[0], vals[1]);
    else
      reinterpret_cast<float4*>(C)[v] = make_float4(vals[0], vals[1], vals[2], vals[3]);
  }
  // tail (total not a multiple of V)
  if (blockIdx.x == 0 && threadIdx.x < (int)(total - nvec * V)) {
    const size_t i = nvec * V + threadIdx.x;
    C[i] = row[i % (size_t)nn];
  }
}

template <class S>
__global__ void tile_c_kernel(const S* __restrict__ p, S* __restrict__ c, int n, size_t total) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    c[i] = p[i % (size_t)n];
}

// partial[blk][0..n) = sum over this block's rows of diag(dC[r]); [n..2n) = sum of dc[r].
// A thread walks whole rows, so every sector of dC is fetched exactly once.
template <class S, int MAXN>
__global__ void reduce_kernel(const S* __restrict__ dC, const S* __restrict__ dc, int n,
                              size_t rows, double* __restrict__ partial) {
  double aq[MAXN], ap[MAXN];
#pragma unroll
  for (int i = 0; i < MAXN; ++i) aq[i] = ap[i] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int nn = n * n;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const S* m = dC + r * nn;
    const S* v = dc + r * n;
#pragma unroll
    for (int i = 0; i < MAXN; ++i)
      if (i < n) {
        aq[i] += (double)m[i * n + i];
        ap[i] += (double)v[i];
      }
  }
  __shared__ double red[kRedThreads / 32][2 * MAXN];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MAXN; ++i) {
    double a = aq[i], b = ap[i];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      red[w][i] = a;
      red[w][MAXN + i] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * n) {
    const int i = threadIdx.x % n, which = threadIdx.x / n;
    double s = 0.0;
    for (int k = 0; k < kRedThreads / 32; ++k) s += red[k][which * MAXN + i];
    partial[(size_t)blockIdx.x * 2 * n + threadIdx.x] = s;
  }
}

template <class S>
__global__ void reduce_final_kernel(const double* __restrict__ partial, int n, int blocks,
                                    S* __restrict__ dq, S* __restrict__ dp) {
  const int i = threadIdx.x;
  if (i >= 2 * n) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += partial[(size_t)b * 2 * n + i];
  if (i < n) dq[i] = (S)s;
  else dp[i - n] = (S)s;
}

template <class S>
int tile(int n, int T, int B, const void* q, const void* p, void* C, void* c, cudaStream_t st) {
  const size_t rows = (size_t)T * B;
  if (C) {
    const size_t total = rows * n * n;
    const int blocks = 148 * 8;
    tile_C_kernel<S><<<blocks, 256, (size_t)n * n * sizeof(S), st>>>(
        static_cast<const S*>(q), static_cast<S*>(C), n, total);
  }
  if (c)
    tile_c_kernel<S><<<148 * 4, 256, 0, st>>>(static_cast<const S*>(p), static_cast<S*>(c), n,
                                              rows * n);
  return cudaGetLastError() == cudaSuccess ? DILQR_OK : DILQR_ECUDA;
}

template <class S>
int reduce(int n, int T, int B, const void* dC, const void* dc, void* dq, void* dp, void* partial,
           size_t partial_bytes, cudaStream_t st) {
  if (partial_bytes < (size_t)kRedBlocks * 2 * n * sizeof(double)) return DILQR_EWORKSPACE;
  const size_t rows = (size_t)T * B;
  double* part = static_cast<double*>(partial);
  const S* a = static_cast<const S*>(dC);
  const S* b = static_cast<const S*>(dc);
  if (n <= 8) reduce_kernel<S, 8><<<kRedBlocks, kRedThreads, 0, st>>>(a, b, n, rows, part);
  else if (n <= 20) reduce_kernel<S, 20><<<kRedBlocks, kRedThreads, 0, st>>>(a, b, n, rows, part);
  else return DILQR_EUNSUPPORTED;
  reduce_final_kernel<S><<<1, 64, 0, st>>>(part, n, kRedBlocks, static_cast<S*>(dq),
                                           static_cast<S*>(dp));
  return cudaGetLastError() == cudaSuccess ? DILQR_OK : DILQR_ECUDA;
}

}  // namespace

extern "C" {

int dilqr_tile_cost(int dtype, int n, int T, int n_batch, const void* q, const void* p, void* C,
                    void* c, void* stream) {
  if (n <= 0 || T <= 0 || n_batch <= 0 || !q || !p) return DILQR_EINVAL;
  if ((reinterpret_cast<uintptr_t>(C) & 15) != 0) return DILQR_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DILQR_F32) return tile<float>(n, T, n_batch, q, p, C, c, st);
  if (dtype == DILQR_F64) return tile<double>(n, T, n_batch, q, p, C, c, st);
  return DILQR_EINVAL;
}

size_t dilqr_tile_cost_grad_workspace_bytes(int n) {
  return (size_t)kRedBlocks * 2 * (size_t)n * sizeof(double);
}

int dilqr_tile_cost_grad(int dtype, int n, int T, int n_batch, const void* dC, const void* dc,
                         void* dq, void* dp, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (n <= 0 || T <= 0 || n_batch <= 0 || !dC || !dc || !dq || !dp || !workspace)
    return DILQR_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DILQR_F32)
    return reduce<float>(n, T, n_batch, dC, dc, dq, dp, workspace, workspace_bytes, st);
  if (dtype == DILQR_F64)
    return reduce<double>(n, T, n_batch, dC, dc, dq, dp, workspace, workspace_bytes, st);
  return DILQR_EINVAL;
}

}  // extern "C"
