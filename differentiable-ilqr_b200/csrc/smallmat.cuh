// smallmat.cuh -- per-thread dense linear algebra on tiny (n_ctrl <= 8) systems.
// Everything is fully unrolled over compile-time sizes so the operands live in
// registers; loops over larger compile-time sizes fall back to `#pragma unroll 1`
// (local memory) purely to bound code size -- those shapes are parity cases, not
// bench cases.
#pragma once
#include "common.cuh"

namespace dilqr {

// Unroll policy: full unroll for small trip counts, rolled otherwise.
#define DILQR_UNROLL_IF_SMALL(n) _Pragma("unroll")

// LU factorisation with partial pivoting of an N x N matrix held in registers,
// LAPACK getrf semantics (first maximal |a_ik| wins; one row interchange per
// column), as used by torch's Tensor.lu()/lu_solve in pnqp.py:18-19,53-54 and
// lqr_step.py:125-127,148.  piv[k] = row swapped with k (0-based).
template <class S, int N>
struct LUpp {
  S a[N][N];
  int piv[N];

  DILQR_DEVICE void factor() {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      int p = k;
      S best = absS<S>(a[k][k]);
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const S v = absS<S>(a[i][k]);
        if (v > best) {
          best = v;
          p = i;
        }
      }
      piv[k] = p;
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        if (i == p) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const S tmp = a[k][j];
            a[k][j] = a[i][j];
            a[i][j] = tmp;
          }
        }
      }
      const S d = a[k][k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const S m = a[i][k] / d;
        a[i][k] = m;
#pragma unroll
        for (int j = k + 1; j < N; ++j) a[i][j] = fmaS<S>(-m, a[k][j], a[i][j]);
      }
    }
  }

  // Solve A x = b in place.
  DILQR_DEVICE void solve(S* b) const {
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        if (i == piv[k]) {
          const S tmp = b[k];
          b[k] = b[i];
          b[i] = tmp;
        }
      }
    }
#pragma unroll
    for (int i = 1; i < N; ++i) {
      S acc = b[i];
#pragma unroll
      for (int j = 0; j < i; ++j) acc = fmaS<S>(-a[i][j], b[j], acc);
      b[i] = acc;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      S acc = b[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) acc = fmaS<S>(-a[i][j], b[j], acc);
      b[i] = acc / a[i][i];
    }
  }
};

// Cholesky A = L L^T (lower) in place; solve by two triangular sweeps.
template <class S, int N>
struct Chol {
  S l[N][N];
  DILQR_DEVICE void factor() {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      S d = l[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d = fmaS<S>(-l[j][k], l[j][k], d);
      d = sqrtS<S>(d);
      l[j][j] = d;
#pragma unroll
      for (int i = j + 1; i < N; ++i) {
        S s = l[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaS<S>(-l[i][k], l[j][k], s);
        l[i][j] = s / d;
      }
    }
  }
  DILQR_DEVICE void solve(S* b) const {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      S acc = b[i];
#pragma unroll
      for (int j = 0; j < i; ++j) acc = fmaS<S>(-l[i][j], b[j], acc);
      b[i] = acc / l[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      S acc = b[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) acc = fmaS<S>(-l[j][i], b[j], acc);
      b[i] = acc / l[i][i];
    }
  }
};

}  // namespace dilqr
