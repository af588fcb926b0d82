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

#ifndef DILQR_RCP_PARAMS
#define DILQR_RCP_PARAMS 1
#endif

namespace dilqr {

enum { DYN_LINDX = 0, DYN_PENDULUM = 1, DYN_CARTPOLE = 2, DYN_ROCKET = 3, DYN_NN = 4 };

template <class S>
struct DynParams {
  S p[8];
  const S* aux;   // DYN_NN: packed weights  W1[H][n] b1[H] W2[ns][H] b2[ns]  (device)
  int ai[4];      // DYN_NN: {H1 | H2 << 16 (H2 = 0: one hidden layer), activation (0 sigmoid, 1 relu), passthrough,
                  //          linearisation (0 analytic grad_input, 1 central differences)}
};

template <class S, int DYN>
struct Dyn;

template <class S, int DYN>
struct EnvTables;   // generated second-order tables (env_tables_gen.cuh)

// ------------------------------------------------------------------ LinDx stub
template <class S>
struct Dyn<S, DYN_LINDX> {
  static constexpr bool kEnv = false;
  // structural non-zero pattern of F (dense for user-supplied LinDx)
  __host__ __device__ static constexpr bool nz(int, int) { return true; }
};

// ------------------------------------------------------- one-hidden-layer network
// dynamics.NNDynamics (dynamics.py:15-130) with hidden_sizes=[H]:
//   z = act(W1 [x;u] + b1),  x' = W2 z + b2 (+ x if passthrough)          (forward, :57-79)
//   d x'/d[x;u] = W2 diag(act'(.)) W1 (+ [I 0])                          (grad_input, :81-130)
// sigmoid: act' = z (1 - z); relu: rows with z <= 0 dropped (:104-110).  The weights are
// shared by all problems (uniform, read-only loads); the hidden units are streamed, so
// neither the activations nor the H x n products are ever stored.
template <class S>
struct Dyn<S, DYN_NN> {
  static constexpr bool kEnv = true;
  static constexpr bool kTrigFromNext = false;
  __host__ __device__ static constexpr bool nz(int, int) { return true; }

  DILQR_DEVICE static S act(S a, int kind) {
    if (kind == 1) return a > S(0) ? a : S(0);
    return S(1) / (S(1) + expS<S>(-a));
  }

  static constexpr int kMaxH = 128;   // two hidden layers: the first layer's width (local array)

  // hidden_sizes=[H1, H2]: weights W1[H1][n] b1[H1] W2[H2][H1] b2[H2] W3[ns][H2] b3[ns]
  template <int NS, int NC>
  DILQR_DEVICE static void step2(const DynParams<S>& P, const S* x, const S* u, S* xn) {
    constexpr int N = NS + NC;
    const int H1 = P.ai[0] & 0xffff, H2 = P.ai[0] >> 16;
    const S* W1 = P.aux;
    const S* b1 = W1 + (size_t)H1 * N;
    const S* W2 = b1 + H1;
    const S* b2 = W2 + (size_t)H2 * H1;
    const S* W3 = b2 + H2;
    const S* b3 = W3 + (size_t)NS * H2;
    S z1[kMaxH];
    for (int h = 0; h < H1; ++h) {
      S a = S(0);
#pragma unroll
      for (int j = 0; j < NS; ++j) a = fmaS<S>(__ldg(W1 + h * N + j), x[j], a);
#pragma unroll
      for (int j = 0; j < NC; ++j) a = fmaS<S>(__ldg(W1 + h * N + NS + j), u[j], a);
      z1[h] = act(a + __ldg(b1 + h), P.ai[1]);
    }
    S acc[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = S(0);
    for (int k = 0; k < H2; ++k) {
      S a = S(0);
      for (int h = 0; h < H1; ++h) a = fmaS<S>(__ldg(W2 + (size_t)k * H1 + h), z1[h], a);
      const S z2 = act(a + __ldg(b2 + k), P.ai[1]);
#pragma unroll
      for (int i = 0; i < NS; ++i) acc[i] = fmaS<S>(__ldg(W3 + i * H2 + k), z2, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S v = acc[i] + __ldg(b3 + i);
      if (P.ai[2]) v = v + x[i];
      xn[i] = v;
    }
  }

  // grad_input for two hidden layers in the reference's association (dynamics.py:98-116):
  // G = W3 (diag(d2) W2)  [ns x H1], then J = G (diag(d1) W1).
  template <int NS, int N>
  DILQR_DEVICE static void jacobian2(const DynParams<S>& P, const S* tau, S (&F)[NS][N]) {
    const int H1 = P.ai[0] & 0xffff, H2 = P.ai[0] >> 16;
    const S* W1 = P.aux;
    const S* b1 = W1 + (size_t)H1 * N;
    const S* W2 = b1 + H1;
    const S* b2 = W2 + (size_t)H2 * H1;
    const S* W3 = b2 + H2;
    S z1[kMaxH];
    S G[NS][kMaxH];
    for (int h = 0; h < H1; ++h) {
      S a = S(0);
#pragma unroll
      for (int j = 0; j < N; ++j) a = fmaS<S>(__ldg(W1 + h * N + j), tau[j], a);
      z1[h] = act(a + __ldg(b1 + h), P.ai[1]);
#pragma unroll
      for (int i = 0; i < NS; ++i) G[i][h] = S(0);
    }
    for (int k = 0; k < H2; ++k) {
      S a = S(0);
      for (int h = 0; h < H1; ++h) a = fmaS<S>(__ldg(W2 + (size_t)k * H1 + h), z1[h], a);
      const S z2 = act(a + __ldg(b2 + k), P.ai[1]);
      const S d2 = (P.ai[1] == 1) ? (z2 <= S(0) ? S(0) : S(1)) : z2 * (S(1) - z2);
      for (int h = 0; h < H1; ++h) {
        const S w = __ldg(W2 + (size_t)k * H1 + h) * d2;
#pragma unroll
        for (int i = 0; i < NS; ++i) G[i][h] = fmaS<S>(__ldg(W3 + i * H2 + k), w, G[i][h]);
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) F[i][j] = S(0);
    for (int h = 0; h < H1; ++h) {
      const S d1 = (P.ai[1] == 1) ? (z1[h] <= S(0) ? S(0) : S(1)) : z1[h] * (S(1) - z1[h]);
      S w[N];
#pragma unroll
      for (int j = 0; j < N; ++j) w[j] = __ldg(W1 + h * N + j) * d1;
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) F[i][j] = fmaS<S>(G[i][h], w[j], F[i][j]);
    }
    if (P.ai[2]) {
#pragma unroll
      for (int i = 0; i < NS; ++i) F[i][i] = F[i][i] + S(1);
    }
  }

  template <int NS, int NC>
  DILQR_DEVICE static void step(const DynParams<S>& P, const S* x, const S* u, S* xn) {
    if (P.ai[0] >> 16) {
      step2<NS, NC>(P, x, u, xn);
      return;
    }
    constexpr int N = NS + NC;
    const int H = P.ai[0];
    const S* W1 = P.aux;
    const S* b1 = W1 + (size_t)H * N;
    const S* W2 = b1 + H;
    const S* b2 = W2 + (size_t)NS * H;
    S acc[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = S(0);
    for (int h = 0; h < H; ++h) {
      S a = S(0);
#pragma unroll
      for (int j = 0; j < NS; ++j) a = fmaS<S>(__ldg(W1 + h * N + j), x[j], a);
#pragma unroll
      for (int j = 0; j < NC; ++j) a = fmaS<S>(__ldg(W1 + h * N + NS + j), u[j], a);
      const S z = act(a + __ldg(b1 + h), P.ai[1]);
#pragma unroll
      for (int i = 0; i < NS; ++i) acc[i] = fmaS<S>(__ldg(W2 + i * H + h), z, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      S v = acc[i] + __ldg(b2 + i);
      if (P.ai[2]) v = v + x[i];
      xn[i] = v;
    }
  }

  // GradMethods.FINITE_DIFF (mpc.py:567-583, util.py:10-20): central differences of the
  // step with eps = 1e-4, one input coordinate at a time.
  template <int NS, int N>
  DILQR_DEVICE static void jacobian_fd(const DynParams<S>& P, const S* tau, S (&F)[NS][N]) {
    constexpr int NC = N - NS;
    const S eps = S(1e-4), two_eps = S(2. * 1e-4);
    S tp[N], fp[NS], fm[NS];
#pragma unroll
    for (int j = 0; j < N; ++j) tp[j] = tau[j];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      tp[j] = tau[j] + eps;
      step<NS, NC>(P, tp, tp + NS, fp);
      tp[j] = tau[j] - eps;
      step<NS, NC>(P, tp, tp + NS, fm);
      tp[j] = tau[j];
#pragma unroll
      for (int i = 0; i < NS; ++i) F[i][j] = (fp[i] - fm[i]) / two_eps;
    }
  }

  template <int NS, int N>
  DILQR_DEVICE static void jacobian(const DynParams<S>& P, const S* tau, const S* /*xnext*/,
                                    S (&F)[NS][N]) {
    if (P.ai[3] == 1) {
      jacobian_fd<NS, N>(P, tau, F);
      return;
    }
    if (P.ai[0] >> 16) {
      jacobian2<NS, N>(P, tau, F);
      return;
    }
    const int H = P.ai[0];
    const S* W1 = P.aux;
    const S* b1 = W1 + (size_t)H * N;
    const S* W2 = b1 + H;
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) F[i][j] = S(0);
    for (int h = 0; h < H; ++h) {
      S w[N];
      S a = S(0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        w[j] = __ldg(W1 + h * N + j);
        a = fmaS<S>(w[j], tau[j], a);
      }
      const S z = act(a + __ldg(b1 + h), P.ai[1]);
      S d;
      if (P.ai[1] == 1) d = z <= S(0) ? S(0) : S(1);
      else d = z * (S(1) - z);
#pragma unroll
      for (int j = 0; j < N; ++j) w[j] = w[j] * d;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const S w2 = __ldg(W2 + i * H + h);
#pragma unroll
        for (int j = 0; j < N; ++j) F[i][j] = fmaS<S>(w2, w[j], F[i][j]);
      }
    }
    if (P.ai[2]) {
#pragma unroll
      for (int i = 0; i < NS; ++i) F[i][i] = F[i][i] + S(1);
    }
  }
};

// One step of the dynamics `DYN` for a problem with NS states / NC controls (the env
// models have fixed sizes; the network's sizes come from the kernel's template).
template <class S, int NS, int NC, int DYN>
DILQR_DEVICE void dyn_step(const DynParams<S>& P, const S* x, const S* u, S* xn) {
  if constexpr (DYN == DYN_NN) Dyn<S, DYN_NN>::template step<NS, NC>(P, x, u, xn);
  else Dyn<S, DYN>::step(P, x, u, xn);
}

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

  // F_t = D(x_t, u_t); `xnext` = x_{t+1} of the same rollout supplies sin/cos.
  DILQR_DEVICE static void jacobian(const DynParams<S>& P, const S* tau, const S* xnext,
                                    S (*F)[N]) {
    S sp, cp;
    if (trig_reusable(&tau[NS])) {
      sp = xnext[1];
      cp = xnext[0];
    } else {
      trig(P, tau, &tau[NS], &sp, &cp);
    }
    jac(P, tau, &tau[NS], sp, cp, F);
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
#if DILQR_RCP_PARAMS
    // 1/M is a property of theta (loop invariant): the three divisions by M -- two of them on
    // the dependent chain cart_in -> th_acc -> xacc, 124 cycles each -- become multiplications
    // (<= 1 ulp each, the order of the libdevice / glibc trig differences already in the rollout)
    const S iM = S(1.0) / M;
    const S cart_in = (uc + pml * (w * w) * s) * iM;
    const S th_acc = (g * s - c * cart_in) / (l * (S(4.0 / 3.0) - mp * (c * c) * iM));
    const S xacc = cart_in - pml * th_acc * c * iM;
#else
    const S cart_in = (uc + pml * (w * w) * s) / M;
    const S th_acc = (g * s - c * cart_in) / (l * (S(4.0 / 3.0) - mp * (c * c) / M));
    const S xacc = cart_in - pml * th_acc * c / M;
#endif
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

  DILQR_DEVICE static void jacobian(const DynParams<S>& P, const S* tau, const S* xnext,
                                    S (*F)[N]) {
    jac(P, tau, &tau[NS], xnext[3], xnext[2], F);
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
#if DILQR_RCP_PARAMS
    // 1/l is a property of theta (loop invariant): three divisions per Jacobian become
    // multiplications (<= 1 ulp each; ~120 cycles of dependent latency each saved)
    const S il = S(1.0) / l;
    const S ta_c = (S(2.0) * mpc * G * iden - A * iM * iden) * il;
    const S ta_s = gs * iden * il;
    const S ta_w = S(-2.0) * c * w * mp * s * iM * iden;
    const S ta_u = -c * iM * iden * il;
#else
    const S ta_c = (S(2.0) * mpc * G * iden - A * iM * iden) / l;
    const S ta_s = gs * iden / l;
    const S ta_w = S(-2.0) * c * w * mp * s * iM * iden;
    const S ta_u = -c * iM * iden / l;
#endif
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

// --------------------------------------------------------------------- rocket
template <class S>
struct Dyn<S, DYN_ROCKET> {
  static constexpr bool kEnv = true;
  static constexpr int NS = 13, NC = 3, N = 16;
  static constexpr bool kTrigFromNext = false;
  // structural non-zeros of D (rocket.py:340-424, 69 of 208): generated mask
  __host__ __device__ static constexpr bool nz(int i, int j) {
    return EnvTables<S, DYN_ROCKET>::nz_D(i, j);
  }

  // rocket.py:82-164.  state r(3) v(3) q(4) w(3); params (Jx, Jy, Jz, mass, l);
  // dt = 0.1; returns the UN-normalised quaternion like the reference (:158-164).
  DILQR_DEVICE static void step(const DynParams<S>& P, const S* x, const S* u, S* xn) {
    const S Jx = P.p[0], Jy = P.p[1], Jz = P.p[2], mass = P.p[3], l = P.p[4];
    const S dt = S(0.1);
    S T[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      S v = u[a];
      v = v < S(-400.0) ? S(-400.0) : v;      // torch.clamp (rocket.py:111)
      v = v > S(400.0) ? S(400.0) : v;
      T[a] = v;
    }
    const S q0 = x[6], q1 = x[7], q2 = x[8], q3 = x[9];
    const S wx = x[10], wy = x[11], wz = x[12];
    S Cb[3][3];                                 // C_B_I (rocket.py:119-123)
    Cb[0][0] = S(1) - S(2) * (q2 * q2 + q3 * q3);
    Cb[0][1] = S(2) * (q1 * q2 + q0 * q3);
    Cb[0][2] = S(2) * (q1 * q3 - q0 * q2);
    Cb[1][0] = S(2) * (q1 * q2 - q0 * q3);
    Cb[1][1] = S(1) - S(2) * (q1 * q1 + q3 * q3);
    Cb[1][2] = S(2) * (q2 * q3 + q0 * q1);
    Cb[2][0] = S(2) * (q1 * q3 + q0 * q2);
    Cb[2][1] = S(2) * (q2 * q3 - q0 * q1);
    Cb[2][2] = S(1) - S(2) * (q1 * q1 + q2 * q2);
    S d[13];
    d[0] = x[3];
    d[1] = x[4];
    d[2] = x[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {               // thrust_global = C_I_B T_B (:131)
      S acc = S(0);
#pragma unroll
      for (int j = 0; j < 3; ++j) acc = fmaS<S>(Cb[j][i], T[j], acc);
      d[3 + i] = acc / mass + (i == 0 ? S(-10.0) : S(0));
    }
    // dq = 0.5 * Omega(w) q   (rocket.py:133-144)
    d[6] = S(0.5) * (((-wx) * q1 + (-wy) * q2) + (-wz) * q3);
    d[7] = S(0.5) * ((wx * q0 + wz * q2) + (-wy) * q3);
    d[8] = S(0.5) * ((wy * q0 + (-wz) * q1) + wx * q3);
    d[9] = S(0.5) * ((wz * q0 + wy * q1) + (-wx) * q2);
    // torque = r_T_B x T_B, r_T_B = (-l/2, 0, 0)   (rocket.py:147-148)
    const S a0 = -l / S(2);
    const S tq[3] = {S(0), S(0) - a0 * T[2], a0 * T[1]};
    const S Jw[3] = {Jx * wx, Jy * wy, Jz * wz};
    const S cr[3] = {wy * Jw[2] - wz * Jw[1], wz * Jw[0] - wx * Jw[2], wx * Jw[1] - wy * Jw[0]};
    d[10] = (S(1) / Jx) * (tq[0] - cr[0]);
    d[11] = (S(1) / Jy) * (tq[1] - cr[1]);
    d[12] = (S(1) / Jz) * (tq[2] - cr[2]);
#pragma unroll
    for (int i = 0; i < 13; ++i) xn[i] = x[i] + d[i] * dt;
  }

  DILQR_DEVICE static bool trig_reusable(const S*) { return true; }

  DILQR_DEVICE static void jacobian(const DynParams<S>& P, const S* tau, const S* /*xnext*/,
                                    S (*F)[N]) {
    EnvTables<S, DYN_ROCKET>::eval_D(P, tau, &tau[NS], F);
  }
};

}  // namespace dilqr
