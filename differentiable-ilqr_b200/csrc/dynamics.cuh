// dynamics.cuh -- env_dx dynamics as device functions: one Euler step and the
// analytic first-order Jacobian D = d step / d (x,u) used by the iLQR
// linearisation (mpc_explicit.py:516-546).
//
//   pendulum : env_dx/pendulum.py:60-95  (step), 444-475 (get_linear_dyn)
//   cartpole : env_dx/cartpole.py:64-97  (step), 790-839 (get_linear_dyn)
//
// The step functions follow the reference's operation order (they are compiled
// with -fmad=false, i.e. one rounding per written operation, like the eager
// PyTorch ops of the reference).  The Jacobians are algebraically simplified
// closed forms of the reference's expressions; like the reference they ignore
// the input clamp inside `step` and treat cos/sin as independent coordinates.
#pragma once
#include "common.cuh"

namespace dilqr {

enum { DYN_LINDX = 0, DYN_PENDULUM = 1, DYN_CARTPOLE = 2, DYN_ROCKET = 3 };

template <class S>
struct DynParams {
  S p[8];
};

template <class S, int DYN>
struct Dyn;

// ------------------------------------------------------------------ LinDx stub
template <class S>
struct Dyn<S, DYN_LINDX> {
  static constexpr bool kEnv = false;
  // structural non-zero pattern of F (dense for user-supplied LinDx)
  __host__ __device__ static constexpr bool nz(int, int) { return true; }
};

// ------------------------------------------------------------------- pendulum
template <class S>
struct Dyn<S, DYN_PENDULUM> {
  static constexpr bool kEnv = true;
  static constexpr int NS = 3, NC = 1, N = 4;
  static constexpr bool kTrigFromNext = true;
  // structural zeros of D (pendulum.py:450-474): only D[2][0]
  __host__ __device__ static constexpr bool nz(int i, int j) { return !(i == 2 && j == 0); }

  // pendulum.py:81-91   x = (cos th, sin th, dth), params (g, m, l), dt = 0.05
  DILQR_DEVICE static void step(const DynParams<S>& P, const S* x, const S* u, S* xn) {
    const S g = P.p[0], m = P.p[1], l = P.p[2];
    const S dt = S(0.05);
    S uc = u[0];
    uc = uc < S(-2.0) ? S(-2.0) : uc;   // torch.clamp (pendulum.py:81)
    uc = uc > S(2.0) ? S(2.0) : uc;
    const S c = x[0], s = x[1], w = x[2];
    const S th = atan2S<S>(s, c);
    const S a = ((S(-3.0) * g) / (S(2.0) * l)) * (-s);
    const S b = (S(3.0) * uc) / (m * (l * l));
    const S nw = w + dt * (a + b);
    const S nth = th + nw * dt;
    S sn, cn;
    sincosS<S>(nth, &sn, &cn);
    xn[0] = cn;
    xn[1] = sn;
    xn[2] = nw;
  }

  // The trig of the Jacobian (sin/cos of the new angle) equals (xn[1], xn[0]) of
  // `step` whenever u is inside the clamp range; otherwise it is recomputed with
  // the unclamped u (pendulum.py:451).
  DILQR_DEVICE static bool trig_reusable(const S* u) { return u[0] >= S(-2.0) && u[0] <= S(2.0); }

  DILQR_DEVICE static void trig(const DynParams<S>& P, const S* x, const S* u, S* sp, S* cp) {
    const S g = P.p[0], m = P.p[1], l = P.p[2];
    const S dt = S(0.05);
    const S c = x[0], s = x[1], w = x[2];
    const S acc = (S(3.0) * g * s) / (S(2.0) * l) + (S(3.0) * u[0]) / (l * l * m);
    const S phi = dt * (dt * acc + w) + atan2S<S>(s, c);
    sincosS<S>(phi, sp, cp);
  }

  // pendulum.py:450-474.  F is row-major [NS][N], columns (cos, sin, dth, u).
  DILQR_DEVICE static void jac(const DynParams<S>& P, const S* x, const S* /*u*/, S sp, S cp,
                               S (*F)[N]) {
    const S g = P.p[0], m = P.p[1], l = P.p[2];
    const S dt = S(0.05);
    const S c = x[0], s = x[1];
    const S r2 = c * c + s * s;
    const S a = (S(3.0) * dt * dt * g) / (S(2.0) * l);
    const S b = (S(3.0) * dt * dt) / (l * l * m);
    const S cr = c / r2 + a;
    F[0][0] = s * sp / r2;
    F[0][1] = -cr * sp;
    F[0][2] = -dt * sp;
    F[0][3] = -b * sp;
    F[1][0] = -s * cp / r2;
    F[1][1] = cr * cp;
    F[1][2] = dt * cp;
    F[1][3] = b * cp;
    F[2][0] = S(0.0);
    F[2][1] = (S(3.0) * dt * g) / (S(2.0) * l);
    F[2][2] = S(1.0);
    F[2][3] = (S(3.0) * dt) / (l * l * m);
  }
};

// ------------------------------------------------------------------- cartpole
template <class S>
struct Dyn<S, DYN_CARTPOLE> {
  static constexpr bool kEnv = true;
  static constexpr int NS = 5, NC = 1, N = 6;
  static constexpr bool kTrigFromNext = true;
  // structural non-zeros of D (cartpole.py:802-838): 16 of 30 entries.  Skipping a
  // structurally zero term is exact: x*0 + acc == acc for finite x.
  __host__ __device__ static constexpr bool nz(int i, int j) {
    return (i == 0 && j <= 1) || (i == 1 && j >= 1) || ((i == 2 || i == 3) && j >= 2 && j <= 4) ||
           (i == 4 && j >= 2);
  }

  // cartpole.py:70-95  state (x, dx, cos th, sin th, dth), params (g, m_c, m_p, l)
  DILQR_DEVICE static void step(const DynParams<S>& P, const S* x, const S* u, S* xn) {
    const S g = P.p[0], mc = P.p[1], mp = P.p[2], l = P.p[3];
    const S dt = S(0.05);
    const S M = mp + mc;
    const S pml = mp * l;
    S uc = u[0];
    uc = uc < S(-100.0) ? S(-100.0) : uc;   // torch.clamp (cartpole.py:77)
    uc = uc > S(100.0) ? S(100.0) : uc;
    const S c = x[2], s = x[3], w = x[4];
    const S th = atan2S<S>(s, c);
    const S cart_in = (uc + pml * (w * w) * s) / M;
    const S th_acc = (g * s - c * cart_in) / (l * (S(4.0 / 3.0) - mp * (c * c) / M));
    const S xacc = cart_in - pml * th_acc * c / M;
    xn[0] = x[0] + dt * x[1];
    xn[1] = x[1] + dt * xacc;
    const S nth = th + dt * w;
    S sn, cn;
    sincosS<S>(nth, &sn, &cn);
    xn[2] = cn;
    xn[3] = sn;
    xn[4] = w + dt * th_acc;
  }

  DILQR_DEVICE static bool trig_reusable(const S*) { return true; }

  DILQR_DEVICE static void trig(const DynParams<S>& /*P*/, const S* x, const S* /*u*/, S* sp,
                                S* cp) {
    const S phi = S(0.05) * x[4] + atan2S<S>(x[3], x[2]);
    sincosS<S>(phi, sp, cp);
  }

  // cartpole.py:802-838, simplified: with M = m_c+m_p, A = dth^2 l m_p s + u,
  // G = g s - c A / M, den = 4/3 - m_p c^2 / M :
  //   th_acc = G/(l den),  xacc = A/M - m_p c G/(M den).
  DILQR_DEVICE static void jac(const DynParams<S>& P, const S* x, const S* u, S sp, S cp,
                               S (*F)[N]) {
    const S g = P.p[0], mc = P.p[1], mp = P.p[2], l = P.p[3];
    const S dt = S(0.05);
    const S M = mc + mp;
    const S c = x[2], s = x[3], w = x[4];
    const S iM = S(1.0) / M;
    const S w2lmp = w * w * l * mp;
    const S A = w2lmp * s + u[0];
    const S G = g * s - c * A * iM;
    const S den = S(4.0 / 3.0) - mp * c * c * iM;
    const S iden = S(1.0) / den;
    const S r2 = c * c + s * s;
    const S ir2 = S(1.0) / r2;
    const S gs = g - c * w2lmp * iM;              // dG/ds
    const S mpc = mp * c * iM * iden;             // m_p c / (M den)
    // d th_acc / d(c, s, w, u)
    const S ta_c = (S(2.0) * mpc * G * iden - A * iM * iden) / l;
    const S ta_s = gs * iden / l;
    const S ta_w = S(-2.0) * c * w * mp * s * iM * iden;
    const S ta_u = -c * iM * iden / l;
    // d xacc / d(c, s, w, u)
    const S xa_c = -mp * G * iM * iden + mpc * A * iM - S(2.0) * mpc * mpc * G;
    const S xa_s = w2lmp * iM - mpc * gs;
    const S xa_w = S(2.0) * w * l * mp * s * iM * (S(1.0) + mpc * c);
    const S xa_u = iM * (S(1.0) + mpc * c);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) F[i][j] = S(0.0);
    F[0][0] = S(1.0);
    F[0][1] = dt;
    F[1][1] = S(1.0);
    F[1][2] = dt * xa_c;
    F[1][3] = dt * xa_s;
    F[1][4] = dt * xa_w;
    F[1][5] = dt * xa_u;
    F[2][2] = s * sp * ir2;
    F[2][3] = -c * sp * ir2;
    F[2][4] = -dt * sp;
    F[3][2] = -s * cp * ir2;
    F[3][3] = c * cp * ir2;
    F[3][4] = dt * cp;
    F[4][2] = dt * ta_c;
    F[4][3] = dt * ta_s;
    F[4][4] = S(1.0) + dt * ta_w;
    F[4][5] = dt * ta_u;
  }
};

}  // namespace dilqr
