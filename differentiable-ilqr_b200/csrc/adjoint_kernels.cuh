// adjoint_kernels.cuh -- the adjoint (KKT) LQR solves of the DiLQR backward for
// env_dx dynamics, factored once and replayed per Richardson pass.
//
// The reference performs every adjoint solve as a full
//   mpc_backup.MPC(lqr_iter=1, u_zero_I=active set)(0, QuadCost(C,-r), LinDx(F,None))
// (lqr_step_explicit.py:276-303): Riccati sweep + rollout + line search.  For a
// fixed problem (C, F = D(tau*), active set) only the affine terms depend on r:
//   q_t = -r_t + F_t' v_{t+1},  k_t = -Hm_t^{-1} (q_u masked),
//   v_t = q_x + (Q_xu + K_t' Q_uu) k_t + K_t' q_u          (lqr_step.py:156-158)
// so the quadratic part (K_t, Hm_t^{-1}, Q_xu + K'Q_uu, Q_uu) is computed ONCE
// (adjoint_factor_kernel) and each Richardson pass only runs the cheap affine
// backward sweep + the linear rollout dx_{t+1} = F_t [dx_t; K_t dx_t + k_t], fused
// with the Richardson update w_t = g_t - Lam_t dtau_t (adjoint_pass_kernel).
// F_t is re-derived from (x*_t, u*_t) in registers instead of being read from HBM.
//
// Line search of the reference's adjoint MPC: the step alpha = 1 is the exact
// minimiser of the (masked) QP, whose optimal value sum_t (q_u'k + 1/2 k'Q_uu k) is
// <= 0 whenever Q_uu > 0 on the free set, so the reference accepts alpha = 1.  Each
// pass evaluates that value and counts the problems where it is positive (or NaN)
// into resid[2]; the host falls back to the generic (line-searching) kernels if the
// count is non-zero.
#pragma once
#include "common.cuh"
#include "dynamics.cuh"
#include "smallmat.cuh"
#include "../../include/dilqr.h"

namespace dilqr {

DILQR_DEVICE void bulk_s2g(void* dst_global, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
DILQR_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DILQR_DEVICE void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DILQR_DEVICE void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DILQR_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class S>
struct AdjParams {
  int T, B, Bp;
  int bounds_kind;   // 0: no active set; 1: scalar bounds -> I = |u - bound| <= 1e-8
  int gain_solve;
  int final_pass;
  S lo, hi;
  const S* C;
  const S* x;
  const S* u;
  const S* g;
  const S* Lam;
  S* w;
  S* fac;     // [T][NFAC][Bp]
  S* kvec;    // [T][NC][Bp]
  S* dtau;    // [T][N][Bp]
  S* dC;
  S* dc;
  S* df;
  S* dx_out;
  S* du_out;
  unsigned long long* resid;   // [0] max|dw| [1] max|w| (double bits), [2] #rejected problems
  DynParams<S> dyn;
};

template <class S, int DYN>
struct Adj {
  using D = Dyn<S, DYN>;
  static constexpr int NS = D::NS, NC = D::NC, N = D::N;
  // per (t, problem) factor record: K[NC][NS], G[NC][NC] (inverse of the masked
  // Q_uu), M[NS][NC] = Q_xu + K'Q_uu, Quu[NC][NC]
  static constexpr int OFF_K = 0, OFF_G = NC * NS, OFF_M = OFF_G + NC * NC,
                       OFF_Q = OFF_M + NS * NC, NFAC = OFF_Q + NC * NC;

  DILQR_DEVICE static bool active(const AdjParams<S>& p, S uv) {
    if (p.bounds_kind == 0) return false;
    return (absS<S>(uv - p.lo) <= S(1e-8)) || (absS<S>(uv - p.hi) <= S(1e-8));
  }

  // F_t = D(x_t, u_t); sin/cos of the new angle are components of x_{t+1}
  DILQR_DEVICE static void jac_at(const AdjParams<S>& p, const S* tau, const S* xnext,
                                  S (*F)[N]) {
    S sp, cp;
    if (D::trig_reusable(&tau[NS])) {
      sp = xnext[DYN == DYN_PENDULUM ? 1 : 3];
      cp = xnext[DYN == DYN_PENDULUM ? 0 : 2];
    } else {
      D::trig(p.dyn, tau, &tau[NS], &sp, &cp);
    }
    D::jac(p.dyn, tau, &tau[NS], sp, cp, F);
  }

  DILQR_DEVICE static void load_tau(const AdjParams<S>& p, int t, int b, S* tau) {
    const size_t tb = (size_t)t * p.B + b;
#pragma unroll
    for (int i = 0; i < NS; ++i) tau[i] = __ldg(p.x + tb * NS + i);
#pragma unroll
    for (int a = 0; a < NC; ++a) tau[NS + a] = __ldg(p.u + tb * NC + a);
  }
};

// ---------------------------------------------------------------------------
// Factor: masked Riccati sweep at tau* (lqr_step.py:99-127 / lqr_step_backup.py:
// 196-228), storing what the affine passes need.
// ---------------------------------------------------------------------------
template <class S, int DYN>
__global__ void __launch_bounds__(128) adjoint_factor_kernel(const __grid_constant__ AdjParams<S> p) {
  using A = Adj<S, DYN>;
  constexpr int NS = A::NS, NC = A::NC, N = A::N, NFAC = A::NFAC;
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int b0 = (blockIdx.x * wpb + warp) * kWarp;
  if (b0 >= p.B) return;
  const int nvalid = min(kWarp, p.B - b0);
  const bool act = lane < nvalid;
  const int b = act ? b0 + lane : b0;
  const uint32_t elems[1] = {N * N};
  const size_t per_warp = WarpStager<S>::bytes_per_warp(1, elems) + kStages * sizeof(uint64_t);
  char* wbase = smem + warp * per_warp;
  WarpStager<S> st;
  st.init(wbase + kStages * sizeof(uint64_t), reinterpret_cast<uint64_t*>(wbase), lane, nvalid, 1,
          elems);
  const int T = p.T;
  auto issue = [&](int stage, int t) {
    const S* src[1] = {p.C + ((size_t)t * p.B + b0) * (N * N)};
    st.issue(stage, src, 1);
  };
  S V[NS][NS], xnext[NS];
  issue(0, T - 1);
  for (int t = T - 1; t >= 0; --t) {
    const int sg = (T - 1 - t) & 1;
    if (t > 0) issue(sg ^ 1, t - 1);
    S tau[N];
    A::load_tau(p, t, b, tau);
    st.wait(sg);
    const S* Cs = st.lane_ptr(sg, 0);
    S Q[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) Q[i][j] = Cs[i * N + j];
    if (t < T - 1) {
      S Fm[NS][N];
      A::jac_at(p, tau, xnext, Fm);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        S Mr[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          S acc = S(0);
#pragma unroll
          for (int l = 0; l < NS; ++l) acc = fmaS<S>(Fm[l][i], V[l][k], acc);
          Mr[k] = acc;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          S acc = S(0);
#pragma unroll
          for (int k = 0; k < NS; ++k) acc = fmaS<S>(Mr[k], Fm[k][j], acc);
          Q[i][j] = Q[i][j] + acc;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) xnext[i] = tau[i];
    bool mask[NC];
#pragma unroll
    for (int a = 0; a < NC; ++a) mask[a] = A::active(p, tau[NS + a]);
    S K[NC][NS], G[NC][NC];
    if (NC == 1) {
      if (p.bounds_kind == 0) {                 // lqr_step.py:84-86
        const S r = S(1) / Q[NS][NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) K[0][j] = -(r * Q[NS][j]);
        G[0][0] = r;
      } else {                                  // lqr_step.py:101-123
        const S r = S(1) / (mask[0] ? S(1e-8) : Q[NS][NS]);
#pragma unroll
        for (int j = 0; j < NS; ++j) K[0][j] = -(r * (mask[0] ? S(0) : Q[NS][j]));
        G[0][0] = S(1) / Q[NS][NS];             // k uses the UNMASKED Q_uu (lqr_step.py:123)
      }
    } else if (p.bounds_kind == 0 && p.gain_solve == 1) {   // lqr_step_backup.py:202-205
      Chol<S, NC> ch;
#pragma unroll
      for (int a = 0; a < NC; ++a)
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2)
          ch.l[a][c2] = Q[NS + a][NS + c2] + (a == c2 ? S(1e-6) : S(0));
      ch.factor();
#pragma unroll
      for (int j = 0; j < NS + NC; ++j) {
        S rhs[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) rhs[a] = (j < NS) ? Q[NS + a][j] : (a == j - NS ? S(1) : S(0));
        ch.solve(rhs);
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          if (j < NS) K[a][j] = -rhs[a];
          else G[a][j - NS < NC ? j - NS : 0] = rhs[a];
        }
      }
    } else {                                    // lqr_step.py:88-94 / 101-127
      LUpp<S, NC> lu;
#pragma unroll
      for (int a = 0; a < NC; ++a)
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2) {
          S h = (mask[a] || mask[c2]) ? S(0) : Q[NS + a][NS + c2];
          if (a == c2 && mask[a]) h = h + S(1e-8);
          lu.a[a][c2] = h;
        }
      lu.factor();
#pragma unroll
      for (int j = 0; j < NS + NC; ++j) {
        S rhs[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a)
          rhs[a] = (j < NS) ? (mask[a] ? S(0) : Q[NS + a][j]) : (a == j - NS ? S(1) : S(0));
        lu.solve(rhs);
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          if (j < NS) K[a][j] = -rhs[a];
          else G[a][j - NS < NC ? j - NS : 0] = rhs[a];
        }
      }
    }
    S KQ[NS][NC];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int a = 0; a < NC; ++a) {
        S acc = S(0);
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2) acc = fmaS<S>(K[c2][i], Q[NS + c2][NS + a], acc);
        KQ[i][a] = acc;
      }
    if (act) {
      S* f = p.fac + (size_t)t * NFAC * p.Bp + b;
#pragma unroll
      for (int a = 0; a < NC; ++a)
#pragma unroll
        for (int j = 0; j < NS; ++j) f[(size_t)(A::OFF_K + a * NS + j) * p.Bp] = K[a][j];
#pragma unroll
      for (int a = 0; a < NC; ++a)
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2) {
          f[(size_t)(A::OFF_G + a * NC + c2) * p.Bp] = G[a][c2];
          f[(size_t)(A::OFF_Q + a * NC + c2) * p.Bp] = Q[NS + a][NS + c2];
        }
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int a = 0; a < NC; ++a)
          f[(size_t)(A::OFF_M + i * NC + a) * p.Bp] = Q[i][NS + a] + KQ[i][a];
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        S t1 = S(0), t2 = S(0), t3 = S(0);
#pragma unroll
        for (int a = 0; a < NC; ++a) {
          t1 = fmaS<S>(Q[i][NS + a], K[a][j], t1);
          t2 = fmaS<S>(K[a][i], Q[NS + a][j], t2);
          t3 = fmaS<S>(KQ[i][a], K[a][j], t3);
        }
        V[i][j] = ((Q[i][j] + t1) + t2) + t3;
      }
  }
}

// ---------------------------------------------------------------------------
// One adjoint solve with r = w (affine sweep + linear rollout).
//   final_pass == 0: fused Richardson update  w_t <- g_t - Lam_t dtau_t.
//   final_pass == 1: keeps dtau (SoA workspace + optional AoS dx_out/du_out), then a
//                    costate sweep (lqr_step.py:371-385) writes dC, dc, df.
// ---------------------------------------------------------------------------
template <class S, int DYN, bool FINAL>
__global__ void __launch_bounds__(128) adjoint_pass_kernel(const __grid_constant__ AdjParams<S> p) {
  using A = Adj<S, DYN>;
  constexpr int NS = A::NS, NC = A::NC, N = A::N, NFAC = A::NFAC;
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int b0 = (blockIdx.x * wpb + warp) * kWarp;
  if (b0 >= p.B) return;
  const int nvalid = min(kWarp, p.B - b0);
  const bool act = lane < nvalid;
  const int b = act ? b0 + lane : b0;
  const int T = p.T;
  // stage: seg0 = n*n block (Lam_t or C_t), seg1 = n vector (g_t)
  const uint32_t elems[2] = {N * N, N};
  const size_t stage_bytes = WarpStager<S>::bytes_per_warp(2, elems);
  const size_t out_bytes = FINAL ? (((size_t)kWarp * (N * N + N) * sizeof(S) + 15) & ~(size_t)15) : 0;
  const size_t per_warp = stage_bytes + kStages * sizeof(uint64_t) + out_bytes;
  char* wbase = smem + warp * per_warp;
  WarpStager<S> st;
  st.init(wbase + kStages * sizeof(uint64_t), reinterpret_cast<uint64_t*>(wbase), lane, nvalid, 2,
          elems);
  S* outC = reinterpret_cast<S*>(wbase + kStages * sizeof(uint64_t) + stage_bytes);
  S* outc = outC + kWarp * N * N;

  // ---------------- affine backward sweep
  S v[NS], xnext[NS];
  S pred = S(0);
  for (int t = T - 1; t >= 0; --t) {
    S tau[N];
    A::load_tau(p, t, b, tau);
    S q[N];
    const size_t tb = (size_t)t * p.B + b;
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = -p.w[tb * N + i];
    if (t < T - 1) {
      S Fm[NS][N];
      A::jac_at(p, tau, xnext, Fm);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        S acc = S(0);
#pragma unroll
        for (int l = 0; l < NS; ++l) acc = fmaS<S>(Fm[l][i], v[l], acc);
        q[i] = q[i] + acc;
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) xnext[i] = tau[i];
    const S* f = p.fac + (size_t)t * NFAC * p.Bp + b;
    S qm[NC], k[NC];
#pragma unroll
    for (int a = 0; a < NC; ++a) qm[a] = A::active(p, tau[NS + a]) ? S(0) : q[NS + a];
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      S acc = S(0);
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) acc = fmaS<S>(f[(size_t)(A::OFF_G + a * NC + c2) * p.Bp], qm[c2], acc);
      k[a] = -acc;
    }
    // optimal value of the masked QP: sum_t q_u'k + 1/2 k'Q_uu k
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      S hk = S(0);
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) hk = fmaS<S>(f[(size_t)(A::OFF_Q + a * NC + c2) * p.Bp], k[c2], hk);
      pred = fmaS<S>(k[a], q[NS + a] + S(0.5) * hk, pred);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S t1 = S(0), t2 = S(0);
#pragma unroll
      for (int a = 0; a < NC; ++a) {
        t1 = fmaS<S>(f[(size_t)(A::OFF_M + i * NC + a) * p.Bp], k[a], t1);
        t2 = fmaS<S>(f[(size_t)(A::OFF_K + a * NS + i) * p.Bp], q[NS + a], t2);
      }
      v[i] = (q[i] + t1) + t2;
    }
    if (act) {
#pragma unroll
      for (int a = 0; a < NC; ++a) p.kvec[((size_t)t * NC + a) * p.Bp + b] = k[a];
    }
  }
  if (act && !(pred <= S(0))) atomicAdd(&p.resid[2], 1ull);

  // ---------------- linear rollout (+ Richardson update)
  S dx[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) dx[i] = S(0);
  double dmax = 0.0, wmax = 0.0;
  auto issue_fwd = [&](int stage, int t) {
    const S* src[2];
    src[0] = (t < T - 1) ? p.Lam + ((size_t)t * p.B + b0) * (N * N) : nullptr;
    src[1] = p.g + ((size_t)t * p.B + b0) * N;
    st.issue(stage, src, 2);
  };
  if (!FINAL) issue_fwd(0, 0);
  S tau[N];
  A::load_tau(p, 0, b, tau);
  for (int t = 0; t < T; ++t) {
    const int sg = t & 1;
    if (!FINAL && t + 1 < T) issue_fwd(sg ^ 1, t + 1);
    S taun[N];
    if (t + 1 < T) A::load_tau(p, t + 1, b, taun);
    const S* f = p.fac + (size_t)t * NFAC * p.Bp + b;
    S dt[N];
#pragma unroll
    for (int i = 0; i < NS; ++i) dt[i] = dx[i];
#pragma unroll
    for (int a = 0; a < NC; ++a) {
      S acc = S(0);
      if (t > 0) {
#pragma unroll
        for (int j = 0; j < NS; ++j) acc = fmaS<S>(f[(size_t)(A::OFF_K + a * NS + j) * p.Bp], dx[j], acc);
      }
      S un = acc + p.kvec[((size_t)t * NC + a) * p.Bp + b];
      if (A::active(p, tau[NS + a])) un = S(0);          // lqr_step.py:197-198
      dt[NS + a] = un;
    }
    if (FINAL) {
      if (act) {
#pragma unroll
        for (int i = 0; i < N; ++i) p.dtau[((size_t)t * N + i) * p.Bp + b] = dt[i];
        const size_t tb = (size_t)t * p.B + b;
        if (p.dx_out) {
#pragma unroll
          for (int i = 0; i < NS; ++i) p.dx_out[tb * NS + i] = dt[i];
        }
        if (p.du_out) {
#pragma unroll
          for (int a = 0; a < NC; ++a) p.du_out[tb * NC + a] = dt[NS + a];
        }
      }
    } else {
      st.wait(sg);
      const S* Ls = st.lane_ptr(sg, 0);
      const S* gs = st.lane_ptr(sg, 1);
      const size_t tb = (size_t)t * p.B + b;
#pragma unroll
      for (int k2 = 0; k2 < N; ++k2) {
        S acc = S(0);
        if (t < T - 1) {
#pragma unroll
          for (int j = 0; j < N; ++j) acc = fmaS<S>(Ls[k2 * N + j], dt[j], acc);
        }
        const S wn = gs[k2] - acc;
        if (act) {
          const S wo = p.w[tb * N + k2];
          dmax = fmax(dmax, fabs((double)wn - (double)wo));
          wmax = fmax(wmax, fabs((double)wn));
          p.w[tb * N + k2] = wn;
        }
      }
    }
    if (t < T - 1) {
      S Fm[NS][N];
      A::jac_at(p, tau, taun, Fm);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        S acc = S(0);
#pragma unroll
        for (int j = 0; j < N; ++j) acc = fmaS<S>(Fm[i][j], dt[j], acc);
        dx[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) tau[i] = taun[i];
    }
  }
  if (!FINAL) {
    for (int o = 16; o > 0; o >>= 1) {
      dmax = fmax(dmax, __shfl_xor_sync(kFull, dmax, o));
      wmax = fmax(wmax, __shfl_xor_sync(kFull, wmax, o));
    }
    if (lane == 0) {
      atomicMax(&p.resid[0], dbits(dmax));
      atomicMax(&p.resid[1], dbits(wmax));
    }
    return;
  }

  // ---------------- final pass: costate sweep + gradient assembly
  //   dlam_t = Cxx dx + Cxu du - r_x + Fx' dlam_{t+1}          (lqr_step.py:371-385)
  //   dC_t = -1/2 (dtau tau' + tau dtau'), dc_t = -dtau, df_t = -dlam_{t+1}
  auto issue_bwd = [&](int stage, int t) {
    const S* src[2] = {p.C + ((size_t)t * p.B + b0) * (N * N), nullptr};
    st.issue(stage, src, 2);
  };
  const bool bulk_out = (nvalid == kWarp) && ((((size_t)N * N * sizeof(S) * kWarp) & 15) == 0) &&
                        ((((size_t)N * sizeof(S) * kWarp) & 15) == 0);
  S dlam[NS];
  issue_bwd(0, T - 1);
  for (int t = T - 1; t >= 0; --t) {
    const int sg = (T - 1 - t) & 1;
    if (t > 0) issue_bwd(sg ^ 1, t - 1);
    S tt[N], dt[N];
    A::load_tau(p, t, b, tt);
#pragma unroll
    for (int i = 0; i < N; ++i) dt[i] = p.dtau[((size_t)t * N + i) * p.Bp + b];
    const size_t tb = (size_t)t * p.B + b;
    if (t < T - 1 && p.df && act) {
#pragma unroll
      for (int i = 0; i < NS; ++i) p.df[tb * NS + i] = -dlam[i];
    }
    // outputs through shared memory, one bulk store per warp slab
    if (bulk_out) {
      bulk_wait_read0();
      __syncwarp();
    }
    S* oC = bulk_out ? outC + lane * (N * N) : p.dC + tb * (N * N);
    S* oc = bulk_out ? outc + lane * N : p.dc + tb * N;
    if (act || bulk_out) {
      if (p.dC) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) oC[i * N + j] = S(-0.5) * (dt[i] * tt[j] + tt[i] * dt[j]);
      }
      if (p.dc) {
#pragma unroll
        for (int i = 0; i < N; ++i) oc[i] = -dt[i];
      }
    }
    if (bulk_out) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (p.dC) bulk_s2g(p.dC + ((size_t)t * p.B + b0) * (N * N), outC, kWarp * N * N * sizeof(S));
        if (p.dc) bulk_s2g(p.dc + ((size_t)t * p.B + b0) * N, outc, kWarp * N * sizeof(S));
        bulk_commit();
      }
    }
    st.wait(sg);
    const S* Cs = st.lane_ptr(sg, 0);
    S nd[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S d1 = S(0), d2 = S(0);
#pragma unroll
      for (int j = 0; j < NS; ++j) d1 = fmaS<S>(Cs[i * N + j], dt[j], d1);
#pragma unroll
      for (int a = 0; a < NC; ++a) d2 = fmaS<S>(Cs[i * N + NS + a], dt[NS + a], d2);
      nd[i] = (d1 + d2) - p.w[tb * N + i];
    }
    if (t < T - 1) {
      S xn[NS];
      {
        const size_t nb = (size_t)(t + 1) * p.B + b;
#pragma unroll
        for (int i = 0; i < NS; ++i) xn[i] = __ldg(p.x + nb * NS + i);
      }
      S Fm[NS][N];
      A::jac_at(p, tt, xn, Fm);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        S d1 = S(0);
#pragma unroll
        for (int l = 0; l < NS; ++l) d1 = fmaS<S>(Fm[l][i], dlam[l], d1);
        nd[i] = nd[i] + d1;
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) dlam[i] = nd[i];
  }
  if (bulk_out) bulk_wait0();
}

}  // namespace dilqr
