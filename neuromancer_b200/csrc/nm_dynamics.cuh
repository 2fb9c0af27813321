// nm_dynamics.cuh -- per-trajectory dynamics in registers: right-hand sides, explicit integrators,
// linear state-space step, and their vector-Jacobian products for the adjoint kernel.
//
// Operation order follows the reference so fp32 rounding stays within 1e-5 of it:
//   RK4   integrators.py:205-211   k2 = f(x + h*k1/2.0) ... x + h*(k1/6.0 + k2/3.0 + k3/3.0 + k4/6.0)
//   RK2   integrators.py:189-193   x + h*f(x + h*k1/2.0)
//   Euler integrators.py:153-156   x + h*f(x)
//   SSM   ode.py:44-51             fx(x) + fu(u) (+= fd(d)),  nn.Linear rows as dot products
// Subgradient conventions reproduce torch autograd: clip passes gradient on the
// closed interval, the masked branch yields 0, d sqrt(h)/dh = 1/(2 sqrt(h)) (inf at 0).
#pragma once
#include "nm_device.cuh"

namespace nm {

// ------------------------------------------------------------------------------------------------
// analytic right-hand sides  dx = f(x, u)
template <class S>
struct Rhs {
  static constexpr int NX = S::NX, NU = S::NU;
  static constexpr int NUA = NU > 0 ? NU : 1;

  static __device__ __forceinline__ void eval(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                              float (&dx)[NX]) {
    if constexpr (S::RHS == NM_RHS_TWOTANK) {            // reference ode.py:227-237
      const float c1 = p.rhs_const[0], c2 = p.rhs_const[1];
      const float h1 = fminf(fmaxf(x[0], 0.f), 1.f), h2 = fminf(fmaxf(x[1], 0.f), 1.f);
      const float pump = fminf(fmaxf(u[0], 0.f), 1.f), valve = fminf(fmaxf(u[1], 0.f), 1.f);
      const float s1 = c2 * sqrtf(h1), s2 = c2 * sqrtf(h2);
      dx[0] = c1 * (1.0f - valve) * pump - s1;
      dx[1] = c1 * valve * pump + s1 - s2;
    } else if constexpr (S::RHS == NM_RHS_VDP) {         // reference ode.py:363-368
      const float mu = p.rhs_const[0];
      dx[0] = x[1];
      dx[1] = mu * (1.f - x[0] * x[0]) * x[1] - x[0] + u[0];
    } else if constexpr (S::RHS == NM_RHS_CVDP) {        // ring of NX/2 coupled VanDerPolControl
      const float mu = p.rhs_const[0], kappa = p.rhs_const[1];
      constexpr int NO = NX / 2;
#pragma unroll
      for (int j = 0; j < NO; ++j) {
        const float pos = x[2 * j], vel = x[2 * j + 1], nxt = x[2 * ((j + 1) % NO)];
        dx[2 * j] = vel;
        dx[2 * j + 1] = mu * (1.f - pos * pos) * vel - pos + u[j] + kappa * (nxt - pos);
      }
    } else if constexpr (S::RHS == NM_RHS_LORENZ_CTRL || S::RHS == NM_RHS_LORENZ) {   // ode.py:371-414
      const float rho = p.rhs_const[0], sigma = p.rhs_const[1], beta = p.rhs_const[2];
      dx[0] = sigma * (x[1] - x[0]);
      dx[1] = x[0] * (rho - x[2]) - x[1];
      dx[2] = x[0] * x[1] - beta * x[2];
      if constexpr (S::RHS == NM_RHS_LORENZ_CTRL) { dx[0] += u[0]; dx[2] -= u[1]; }
    } else if constexpr (S::RHS == NM_RHS_CSTR) {              // ode.py:449-461
      const float qV = p.rhs_const[0], Caf = p.rhs_const[1], Tf = p.rhs_const[2], k0 = p.rhs_const[3];
      const float EoverR = p.rhs_const[4], cH = p.rhs_const[5], cU = p.rhs_const[6];
      const float rA = k0 * expf(-EoverR / x[1]) * x[0];
      dx[0] = qV * (Caf - x[0]) - rA;
      dx[1] = qV * (Tf - x[1]) + cH * rA + cU * (u[0] - x[1]);
    } else if constexpr (S::RHS == NM_RHS_DUFFING) {           // ode.py:254-262 (input is time)
      const float alpha = p.rhs_const[0], beta = p.rhs_const[1], delta = p.rhs_const[2], gamma = p.rhs_const[3], omega = p.rhs_const[4];
      dx[0] = x[1];
      dx[1] = -delta * x[1] - alpha * x[0] - beta * (x[0] * x[0] * x[0]) + gamma * cosf(omega * u[0]);
    } else if constexpr (S::RHS == NM_RHS_BRUSSELATOR) {       // ode.py:276-281
      const float alpha = p.rhs_const[0], beta = p.rhs_const[1];
      const float q = x[1] * (x[0] * x[0]);
      dx[0] = alpha + q - beta * x[0] - x[0];
      dx[1] = beta * x[0] - q;
    } else if constexpr (S::RHS == NM_RHS_LOTKA_VOLTERRA) {    // ode.py:344-349
      const float alpha = p.rhs_const[0];
      dx[0] = alpha * x[0] - 0.4f * x[0] * x[1];
      dx[1] = 0.1f * x[0] * x[1] - 0.4f * x[1];
    } else if constexpr (S::RHS == NM_RHS_LINEAR) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        float s = 0.f, su = 0.f;
#pragma unroll
        for (int j = 0; j < NX; ++j) s = fmaf(p.A[i * NX + j], x[j], s);
#pragma unroll
        for (int j = 0; j < NU; ++j) su = fmaf(p.Bm[i * NU + j], u[j], su);
        dx[i] = s + su;
      }
    }
  }

  // gx += Jx^T w ; gu += Ju^T w
  static __device__ __forceinline__ void vjp(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                             const float (&w)[NX], float (&gx)[NX], float (&gu)[NUA]) {
    if constexpr (S::RHS == NM_RHS_TWOTANK) {
      const float c1 = p.rhs_const[0], c2 = p.rhs_const[1];
      const float h1 = fminf(fmaxf(x[0], 0.f), 1.f), h2 = fminf(fmaxf(x[1], 0.f), 1.f);
      const float pump = fminf(fmaxf(u[0], 0.f), 1.f), valve = fminf(fmaxf(u[1], 0.f), 1.f);
      // d/dh of c2*sqrt(h): autograd computes grad / (2*sqrt(h)); keep the division so that h == 0
      // gives +-inf (or NaN for a zero cotangent) exactly like the reference
      const float g_s1 = (w[1] - w[0]) * c2;       // cotangent of sqrt(h1)
      const float g_s2 = -w[1] * c2;               // cotangent of sqrt(h2)
      const bool in0 = (x[0] >= 0.f) && (x[0] <= 1.f), in1 = (x[1] >= 0.f) && (x[1] <= 1.f);
      if (in0) gx[0] += g_s1 / (2.f * sqrtf(h1));
      if (in1) gx[1] += g_s2 / (2.f * sqrtf(h2));
      const float g_pump = c1 * ((1.0f - valve) * w[0] + valve * w[1]);
      const float g_valve = c1 * pump * (w[1] - w[0]);
      if (u[0] >= 0.f && u[0] <= 1.f) gu[0] += g_pump;
      if (u[1] >= 0.f && u[1] <= 1.f) gu[1] += g_valve;
    } else if constexpr (S::RHS == NM_RHS_VDP) {
      const float mu = p.rhs_const[0];
      gx[0] += w[1] * (-2.f * mu * x[0] * x[1] - 1.f);
      gx[1] += w[0] + w[1] * mu * (1.f - x[0] * x[0]);
      gu[0] += w[1];
    } else if constexpr (S::RHS == NM_RHS_CVDP) {
      const float mu = p.rhs_const[0], kappa = p.rhs_const[1];
      constexpr int NO = NX / 2;
#pragma unroll
      for (int j = 0; j < NO; ++j) {
        const float pos = x[2 * j], vel = x[2 * j + 1];
        const float wa = w[2 * j + 1];
        gx[2 * j] += wa * (-2.f * mu * pos * vel - 1.f - kappa);
        gx[2 * j + 1] += w[2 * j] + wa * mu * (1.f - pos * pos);
        gx[2 * ((j + 1) % NO)] += wa * kappa;
        gu[j] += wa;
      }
    } else if constexpr (S::RHS == NM_RHS_LORENZ_CTRL || S::RHS == NM_RHS_LORENZ) {
      const float rho = p.rhs_const[0], sigma = p.rhs_const[1], beta = p.rhs_const[2];
      gx[0] += -sigma * w[0] + (rho - x[2]) * w[1] + x[1] * w[2];
      gx[1] += sigma * w[0] - w[1] + x[0] * w[2];
      gx[2] += -x[0] * w[1] - beta * w[2];
      if constexpr (S::RHS == NM_RHS_LORENZ_CTRL) { gu[0] += w[0]; gu[1] -= w[2]; }
    } else if constexpr (S::RHS == NM_RHS_CSTR) {
      const float qV = p.rhs_const[0], k0 = p.rhs_const[3], EoverR = p.rhs_const[4], cH = p.rhs_const[5], cU = p.rhs_const[6];
      const float e = k0 * expf(-EoverR / x[1]);
      const float g_rA = -w[0] + cH * w[1];                 // cotangent of rA = e * Ca
      gx[0] += -qV * w[0] + g_rA * e;
      gx[1] += (-qV - cU) * w[1] + g_rA * x[0] * e * (EoverR / (x[1] * x[1]));
      gu[0] += cU * w[1];
    } else if constexpr (S::RHS == NM_RHS_DUFFING) {
      const float alpha = p.rhs_const[0], beta = p.rhs_const[1], delta = p.rhs_const[2], gamma = p.rhs_const[3], omega = p.rhs_const[4];
      gx[0] += w[1] * (-alpha - 3.f * beta * x[0] * x[0]);
      gx[1] += w[0] - delta * w[1];
      gu[0] += -w[1] * gamma * omega * sinf(omega * u[0]);
    } else if constexpr (S::RHS == NM_RHS_BRUSSELATOR) {
      const float beta = p.rhs_const[1];
      const float dq0 = 2.f * x[1] * x[0], dq1 = x[0] * x[0];   // d(x2*x1^2)/dx1, /dx2
      gx[0] += w[0] * (dq0 - beta - 1.f) + w[1] * (beta - dq0);
      gx[1] += (w[0] - w[1]) * dq1;
    } else if constexpr (S::RHS == NM_RHS_LOTKA_VOLTERRA) {
      const float alpha = p.rhs_const[0];
      gx[0] += w[0] * (alpha - 0.4f * x[1]) + w[1] * 0.1f * x[1];
      gx[1] += -w[0] * 0.4f * x[0] + w[1] * (0.1f * x[0] - 0.4f);
    } else if constexpr (S::RHS == NM_RHS_LINEAR) {
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) s = fmaf(p.A[i * NX + j], w[i], s);
        gx[j] += s;
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) s = fmaf(p.Bm[i * NU + j], w[i], s);
        gu[j] += s;
      }
    }
  }
};

// gth[c] += sum_i w_i d f_i / d rhs_const[c]: gradient w.r.t. the white-box constants (system identification: reference
// ode.py declares them nn.Parameters, several with requires_grad=True).  Same conventions as eval / vjp above.
constexpr int NM_MAX_CONST = 8;
template <class S> constexpr int rhs_num_const() {
  return S::RHS == NM_RHS_TWOTANK ? 2 : S::RHS == NM_RHS_VDP ? 1 : S::RHS == NM_RHS_CVDP ? 2 : (S::RHS == NM_RHS_LORENZ_CTRL || S::RHS == NM_RHS_LORENZ) ? 3 :
         S::RHS == NM_RHS_CSTR ? 7 : S::RHS == NM_RHS_DUFFING ? 5 : S::RHS == NM_RHS_BRUSSELATOR ? 2 : S::RHS == NM_RHS_LOTKA_VOLTERRA ? 1 :
         S::RHS == NM_RHS_BRUSS_HYBRID ? 2 : S::RHS == NM_RHS_LV_HYBRID ? 4 : 0;
}
// hybrid (grey-box) right-hand sides (reference ode.py:284-331): white-box equations around one unknown term m = MLP(x)
template <int RHS> constexpr bool rhs_hybrid() { return RHS == NM_RHS_BRUSS_HYBRID || RHS == NM_RHS_LV_HYBRID; }
template <class S>
__device__ __forceinline__ void hybrid_eval(const NmPlan& p, const float (&x)[2], float m, float (&dx)[2]) {
  if constexpr (S::RHS == NM_RHS_BRUSS_HYBRID) {
    const float alpha = p.rhs_const[0], beta = p.rhs_const[1];
    dx[0] = ((alpha + m) - beta * x[0]) - x[0];
    dx[1] = beta * x[0] - m;
  } else {
    const float alpha = p.rhs_const[0], beta = p.rhs_const[1], delta = p.rhs_const[2], gamma = p.rhs_const[3];
    dx[0] = alpha * x[0] - beta * m;
    dx[1] = delta * m - gamma * x[1];
  }
}
// w = cotangent of dx: wm = cotangent of m, gdir = the equations' direct contribution to the cotangent of x,
// gth += d / d rhs_const (THETA)
template <class S, bool THETA>
__device__ __forceinline__ void hybrid_vjp(const NmPlan& p, const float (&x)[2], float m, const float (&w)[2], float& wm, float (&gdir)[2],
                                           float (&gth)[NM_MAX_CONST]) {
  if constexpr (S::RHS == NM_RHS_BRUSS_HYBRID) {
    const float beta = p.rhs_const[1];
    wm = w[0] - w[1];
    gdir[0] = beta * (w[1] - w[0]) - w[0];
    gdir[1] = 0.f;
    if constexpr (THETA) { gth[0] += w[0]; gth[1] += (w[1] - w[0]) * x[0]; }
  } else {
    const float alpha = p.rhs_const[0], beta = p.rhs_const[1], delta = p.rhs_const[2], gamma = p.rhs_const[3];
    wm = delta * w[1] - beta * w[0];
    gdir[0] = alpha * w[0];
    gdir[1] = -gamma * w[1];
    if constexpr (THETA) { gth[0] += w[0] * x[0]; gth[1] -= w[0] * m; gth[2] += w[1] * m; gth[3] -= w[1] * x[1]; }
  }
}
template <class S>
__device__ __forceinline__ void rhs_vjp_theta(const NmPlan& p, const float (&x)[S::NX], const float (&u)[S::NU > 0 ? S::NU : 1],
                                              const float (&w)[S::NX], float (&gth)[NM_MAX_CONST]) {
  if constexpr (S::RHS == NM_RHS_TWOTANK) {
    const float h1 = fminf(fmaxf(x[0], 0.f), 1.f), h2 = fminf(fmaxf(x[1], 0.f), 1.f);
    const float pump = fminf(fmaxf(u[0], 0.f), 1.f), valve = fminf(fmaxf(u[1], 0.f), 1.f);
    gth[0] += w[0] * ((1.0f - valve) * pump) + w[1] * (valve * pump);
    gth[1] += (w[1] - w[0]) * sqrtf(h1) - w[1] * sqrtf(h2);
  } else if constexpr (S::RHS == NM_RHS_VDP) {
    gth[0] += w[1] * ((1.f - x[0] * x[0]) * x[1]);
  } else if constexpr (S::RHS == NM_RHS_CVDP) {
    constexpr int NO = S::NX / 2;
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      const float pos = x[2 * j], vel = x[2 * j + 1], nxt = x[2 * ((j + 1) % NO)];
      gth[0] += w[2 * j + 1] * ((1.f - pos * pos) * vel);
      gth[1] += w[2 * j + 1] * (nxt - pos);
    }
  } else if constexpr (S::RHS == NM_RHS_LORENZ_CTRL || S::RHS == NM_RHS_LORENZ) {
    gth[0] += w[1] * x[0];
    gth[1] += w[0] * (x[1] - x[0]);
    gth[2] -= w[2] * x[2];
  } else if constexpr (S::RHS == NM_RHS_CSTR) {              // constants: qV, Caf, Tf, k0, E/R, cH, cU
    const float qV = p.rhs_const[0], Caf = p.rhs_const[1], Tf = p.rhs_const[2], k0 = p.rhs_const[3], EoverR = p.rhs_const[4], cH = p.rhs_const[5];
    const float ex = expf(-EoverR / x[1]);
    const float rA = k0 * ex * x[0];
    const float g_rA = -w[0] + cH * w[1];
    gth[0] += w[0] * (Caf - x[0]) + w[1] * (Tf - x[1]);
    gth[1] += w[0] * qV;
    gth[2] += w[1] * qV;
    gth[3] += g_rA * (ex * x[0]);
    gth[4] -= g_rA * (rA / x[1]);
    gth[5] += w[1] * rA;
    gth[6] += w[1] * (u[0] - x[1]);
  } else if constexpr (S::RHS == NM_RHS_DUFFING) {
    const float gamma = p.rhs_const[3], omega = p.rhs_const[4];
    gth[0] -= w[1] * x[0];
    gth[1] -= w[1] * (x[0] * x[0] * x[0]);
    gth[2] -= w[1] * x[1];
    gth[3] += w[1] * cosf(omega * u[0]);
    gth[4] -= w[1] * gamma * u[0] * sinf(omega * u[0]);
  } else if constexpr (S::RHS == NM_RHS_BRUSSELATOR) {
    gth[0] += w[0];
    gth[1] += (w[1] - w[0]) * x[0];
  } else if constexpr (S::RHS == NM_RHS_LOTKA_VOLTERRA) {
    gth[0] += w[0] * x[0];
  }
}

// ------------------------------------------------------------------------------------------------
// explicit Runge-Kutta schemes as Butcher tableaux:  xs_s = x + h * sum_{j<s} a(s,j) k_j ,  k_s = f(xs_s, u),
// x+ = x + h * sum_s b(s) k_s.  Euler / RK2 / RK4 evaluate their stages in the reference's exact operation
// order (below); the other schemes use the generic tableau arithmetic (same values to ~1 ulp per stage).
constexpr int NM_MAX_STAGES = 7;
constexpr double kQ21 = 4.58257569495584000658804719372800848898445657676797;   // sqrt(21), Luther
template <int INTEG> struct Tableau;
template <> struct Tableau<NM_INT_EULER> {
  static constexpr int NS = 1;
  static constexpr double a(int, int) { return 0.0; }
  static constexpr double b(int) { return 1.0; }
};
template <> struct Tableau<NM_INT_RK2> {
  static constexpr int NS = 2;
  static constexpr double a(int s, int j) { return (s == 1 && j == 0) ? 0.5 : 0.0; }
  static constexpr double b(int s) { return s == 1 ? 1.0 : 0.0; }
};
template <> struct Tableau<NM_INT_RK4> {
  static constexpr int NS = 4;
  static constexpr double a(int s, int j) { return (j == s - 1) ? (s == 3 ? 1.0 : 0.5) : 0.0; }
  static constexpr double b(int s) { return (s == 0 || s == 3) ? 1.0 / 6.0 : 1.0 / 3.0; }
};
template <> struct Tableau<NM_INT_EULER_TRAP> {              // pred = x + h k1 ; corr = x + 0.5 h (k1 + f(pred))
  static constexpr int NS = 2;
  static constexpr double a(int s, int j) { return (s == 1 && j == 0) ? 1.0 : 0.0; }
  static constexpr double b(int) { return 0.5; }
};
template <> struct Tableau<NM_INT_RK4_TRAP> {                // RK4 predictor, trapezoidal corrector
  static constexpr int NS = 5;
  static constexpr double a(int s, int j) {
    if (s < 4) return Tableau<NM_INT_RK4>::a(s, j);
    return Tableau<NM_INT_RK4>::b(j);
  }
  static constexpr double b(int s) { return (s == 0 || s == 4) ? 0.5 : 0.0; }
};
template <> struct Tableau<NM_INT_LUTHER> {                  // integrators.py:247-264
  static constexpr int NS = 7;
  static constexpr double a(int s, int j) {
    constexpr double q = kQ21;
    const double A[7][6] = {
        {0, 0, 0, 0, 0, 0},
        {1, 0, 0, 0, 0, 0},
        {3.0 / 8, 1.0 / 8, 0, 0, 0, 0},
        {8.0 / 27, 2.0 / 27, 8.0 / 27, 0, 0, 0},
        {(-21 + 9 * q) / 392, (-56 + 8 * q) / 392, (336 - 48 * q) / 392, (-63 + 3 * q) / 392, 0, 0},
        {(-1155 - 255 * q) / 1960, (-280 - 40 * q) / 1960, -320 * q / 1960, (63 + 363 * q) / 1960, (2352 + 392 * q) / 1960, 0},
        {(330 + 105 * q) / 180, 120.0 / 180, (-200 + 280 * q) / 180, (126 - 189 * q) / 180, (-686 - 126 * q) / 180, (490 - 70 * q) / 180}};
    return j < s ? A[s][j] : 0.0;
  }
  static constexpr double b(int s) {
    const double B[7] = {1.0 / 20, 0, 16.0 / 45, 0, 49.0 / 180, 49.0 / 180, 1.0 / 20};
    return B[s];
  }
};
template <> struct Tableau<NM_INT_RKF> {                     // integrators.py:287-301 (5th-order output)
  static constexpr int NS = 6;
  static constexpr double a(int s, int j) {
    const double A[6][5] = {
        {0, 0, 0, 0, 0},
        {1.0 / 4, 0, 0, 0, 0},
        {3.0 / 32, 9.0 / 32, 0, 0, 0},
        {1932.0 / 2197, -7200.0 / 2197, 7296.0 / 2197, 0, 0},
        {439.0 / 216, -8.0, 3680.0 / 513, -845.0 / 4104, 0},
        {-8.0 / 27, 2.0, -3544.0 / 2565, 1859.0 / 4104, -11.0 / 40}};
    return j < s ? A[s][j] : 0.0;
  }
  static constexpr double b(int s) {
    const double B[6] = {16.0 / 135, 0, 6656.0 / 12825, 28561.0 / 56430, -9.0 / 50, 2.0 / 55};
    return B[s];
  }
};

// ------------------------------------------------------------------------------------------------
// Second-order schemes (reference integrators.py:340-404): the block is ddx = f(x[, u]) on HALF = NX / 2 coordinates,
// the state is X = [x, dx].  They are not Runge-Kutta tableaus on X (position and velocity advance with different
// coefficients, LeapFrog even with h^2), so they get their own stage functions; k[j][0 .. HALF) holds a_j = f(q_j).
//   LeapFrog:  q0 = x;  q1 = x + dx h + 0.5 a0 h^2;  X+ = [q1, dx + 0.5 (a0 + a1) h]
//   Yoshida4:  q1 = x + c1 dx h;  v1 = dx + d1 a1 h;  q2 = q1 + c2 v1 h;  v2 = v1 + d2 a2 h;  q3 = q2 + c3 v2 h;
//              v3 = v2 + d3 a3 h;  X+ = [q3 + c4 v3 h, v3]          (operation order of the reference expressions)
template <int INTEG> constexpr bool second_order() { return INTEG == NM_INT_LEAPFROG || INTEG == NM_INT_YOSHIDA4; }
template <> struct Tableau<NM_INT_LEAPFROG> {
  static constexpr int NS = 2;
  static constexpr double a(int, int) { return 0.0; }
  static constexpr double b(int) { return 0.0; }
};
template <> struct Tableau<NM_INT_YOSHIDA4> {
  static constexpr int NS = 3;
  static constexpr double a(int, int) { return 0.0; }
  static constexpr double b(int) { return 0.0; }
};
struct Yoshida {                                             // integrators.py:381-390, evaluated in double like the python floats
  static constexpr double cbrt2 = 1.2599210498948731648;
  static constexpr double w0 = -cbrt2 / (2.0 - cbrt2), w1 = 1.0 / (2.0 - cbrt2);
  static constexpr float c(int i) { return static_cast<float>((i == 0 || i == 3) ? w1 / 2.0 : (w0 + w1) / 2.0); }   // c1..c4
  static constexpr float d(int i) { return static_cast<float>(i == 1 ? w0 : w1); }                                   // d1..d3
};
// position q_St the right-hand side is evaluated at in stage St, from X and a_0 .. a_{St-1}
template <int INTEG, int St, int NX>
__device__ __forceinline__ void so_position(float h, const float (&X)[NX], const float (&k)[NM_MAX_STAGES][NX], float (&q)[NX / 2]) {
  constexpr int HALF = NX / 2;
  if constexpr (INTEG == NM_INT_LEAPFROG) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      if constexpr (St == 0) q[i] = X[i];
      else q[i] = (X[i] + X[HALF + i] * h) + (0.5f * k[0][i]) * (h * h);
    }
  } else {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      float qq = X[i] + (Yoshida::c(0) * X[HALF + i]) * h, v = X[HALF + i];
      static_for<0, St>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        v = v + (Yoshida::d(J) * k[J][i]) * h;
        qq = qq + (Yoshida::c(J + 1) * v) * h;
      });
      q[i] = qq;
    }
  }
}
template <int INTEG, int NX>
__device__ __forceinline__ void so_combine(float h, const float (&X)[NX], const float (&k)[NM_MAX_STAGES][NX], float (&Xn)[NX]) {
  constexpr int HALF = NX / 2;
  if constexpr (INTEG == NM_INT_LEAPFROG) {
    float q1[HALF];
    so_position<INTEG, 1, NX>(h, X, k, q1);
#pragma unroll
    for (int i = 0; i < HALF; ++i) { Xn[i] = q1[i]; Xn[HALF + i] = X[HALF + i] + (0.5f * (k[0][i] + k[1][i])) * h; }
  } else {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      float qq = X[i] + (Yoshida::c(0) * X[HALF + i]) * h, v = X[HALF + i];
      static_for<0, 3>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        v = v + (Yoshida::d(J) * k[J][i]) * h;
        qq = qq + (Yoshida::c(J + 1) * v) * h;
      });
      Xn[i] = qq; Xn[HALF + i] = v;
    }
  }
}
// adjoint: running cotangents gq (of the current position) and gv (of the current velocity), stages in reverse;
// ga is LeapFrog's pending cotangent of a0 (it enters both the position and the velocity update)
template <int INTEG, int NX>
__device__ __forceinline__ void so_adj_init(float h, const float (&lam)[NX], float (&gq)[NX / 2], float (&gv)[NX / 2], float (&ga)[NX / 2]) {
  constexpr int HALF = NX / 2;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    gq[i] = lam[i]; ga[i] = 0.f;
    if constexpr (INTEG == NM_INT_LEAPFROG) gv[i] = lam[HALF + i];
    else gv[i] = lam[HALF + i] + (Yoshida::c(3) * h) * lam[i];
  }
}
template <int INTEG, int St, int NX>                         // cotangent w of a_St
__device__ __forceinline__ void so_adj_cot(float h, const float (&gv)[NX / 2], const float (&ga)[NX / 2], float (&w)[NX / 2]) {
#pragma unroll
  for (int i = 0; i < NX / 2; ++i) {
    if constexpr (INTEG == NM_INT_LEAPFROG) w[i] = St == 1 ? 0.5f * h * gv[i] : ga[i];
    else w[i] = (Yoshida::d(St) * h) * gv[i];
  }
}
template <int INTEG, int St, int NX>                         // gqs = (d f / d q_St)^T w arrived: step back over stage St
__device__ __forceinline__ void so_adj_push(float h, const float (&gqs)[NX / 2], float (&gq)[NX / 2], float (&gv)[NX / 2], float (&ga)[NX / 2]) {
#pragma unroll
  for (int i = 0; i < NX / 2; ++i) {
    gq[i] += gqs[i];
    if constexpr (INTEG == NM_INT_LEAPFROG) {
      if constexpr (St == 1) { ga[i] = 0.5f * h * gv[i] + 0.5f * (h * h) * gq[i]; gv[i] += h * gq[i]; }
    } else gv[i] += (Yoshida::c(St) * h) * gq[i];
  }
}

// stage input xs_S from x and the earlier stage derivatives k[0..S-1]
template <int INTEG, int S_, int NX>
__device__ __forceinline__ void stage_input(float h, const float (&x)[NX], const float (&k)[NM_MAX_STAGES][NX], float (&xs)[NX]) {
  if constexpr (S_ == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i];
  } else if constexpr (INTEG == NM_INT_RK2 || INTEG == NM_INT_RK4) {
    // reference order:  x + h*k/2.0   (RK4 last stage:  x + h*k)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if constexpr (INTEG == NM_INT_RK4 && S_ == 3) xs[i] = x[i] + h * k[S_ - 1][i];
      else xs[i] = x[i] + h * k[S_ - 1][i] / 2.0f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      float acc = 0.f;
      static_for<0, S_>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        constexpr float A = static_cast<float>(Tableau<INTEG>::a(S_, J));
        if constexpr (A != 0.f) acc = fmaf(A, k[J][i], acc);
      });
      xs[i] = fmaf(h, acc, x[i]);
    }
  }
}
// x / d for a compile-time constant d, correctly rounded like the IEEE division it replaces (Markstein: with r = RN(1/d)
// and q0 = RN(x r) within one ulp of x/d, the residual rem = x - d q0 is exact in an FMA and RN(q0 + rem r) is the correctly
// rounded quotient; checked against x / d on every significand of a binade for d = 3 and 6 in tests/test_plan_host.py).
// Three instructions instead of a reciprocal + Newton steps + a slow-path call: the RK4 combine below has 4 NX of them per
// step, which made it the longest serial stretch of a rollout step.  (Not preserved: x = +-inf gives NaN; subnormal
// quotients may differ in the last bit.)
__device__ __forceinline__ float div_const(float x, float d) {
  const float r = 1.0f / d;
  const float q0 = x * r;
  return fmaf(fmaf(-d, q0, x), r, q0);
}
// x+ from the stage derivatives
template <int INTEG, int NX>
__device__ __forceinline__ void combine_stages(float h, const float (&x)[NX], const float (&k)[NM_MAX_STAGES][NX], float (&xn)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if constexpr (INTEG == NM_INT_EULER) xn[i] = x[i] + h * k[0][i];
    else if constexpr (INTEG == NM_INT_RK2) xn[i] = x[i] + h * k[1][i];
    else if constexpr (INTEG == NM_INT_RK4) xn[i] = x[i] + h * (div_const(k[0][i], 6.0f) + div_const(k[1][i], 3.0f) + div_const(k[2][i], 3.0f) + div_const(k[3][i], 6.0f));
    else {
      float acc = 0.f;
      static_for<0, Tableau<INTEG>::NS>([&](auto sc) {
        constexpr int Sx = decltype(sc)::value;
        constexpr float Bv = static_cast<float>(Tableau<INTEG>::b(Sx));
        if constexpr (Bv != 0.f) acc = fmaf(Bv, k[Sx][i], acc);
      });
      xn[i] = fmaf(h, acc, x[i]);
    }
  }
}
// adjoint bookkeeping of stage S:  w = h*b(S)*lam + gk[S]  is the cotangent of k_S; after the RHS vjp gave
// gxs (cotangent of xs_S) it is pushed to x and to the earlier stages:  gk[j] += h*a(S,j)*gxs
template <int INTEG, int S_, int NX>
__device__ __forceinline__ void stage_cotangent(float h, const float (&lam)[NX], const float (&gk)[NM_MAX_STAGES][NX], float (&w)[NX]) {
  constexpr float Bv = static_cast<float>(Tableau<INTEG>::b(S_));
#pragma unroll
  for (int i = 0; i < NX; ++i) w[i] = fmaf(h * Bv, lam[i], gk[S_][i]);
}
template <int INTEG, int S_, int NX>
__device__ __forceinline__ void stage_push(float h, const float (&gxs)[NX], float (&gx)[NX], float (&gk)[NM_MAX_STAGES][NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) gx[i] += gxs[i];
  static_for<0, S_>([&](auto jc) {
    constexpr int J = decltype(jc)::value;
    constexpr float A = static_cast<float>(Tableau<INTEG>::a(S_, J));
    if constexpr (A != 0.f) {
#pragma unroll
      for (int i = 0; i < NX; ++i) gk[J][i] = fmaf(h * A, gxs[i], gk[J][i]);
    }
  });
}

// ------------------------------------------------------------------------------------------------
// one step of the (non-MLP) dynamics and its adjoint
template <class S>
struct Dyn {
  static constexpr int NX = S::NX, NU = S::NU, ND = S::ND;
  static constexpr int NUA = NU > 0 ? NU : 1, NDA = ND > 0 ? ND : 1;

  static __device__ __forceinline__ void step(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                              const float (&d)[NDA], float (&xn)[NX]) {
    if constexpr (S::DYN == NM_DYN_SSM) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        float sx = 0.f, su = 0.f;
#pragma unroll
        for (int j = 0; j < NX; ++j) sx = fmaf(x[j], p.A[i * NX + j], sx);
#pragma unroll
        for (int j = 0; j < NU; ++j) su = fmaf(u[j], p.Bm[i * NU + j], su);
        float v = sx + su;
        if constexpr (ND > 0) {
          float sd = 0.f;
#pragma unroll
          for (int j = 0; j < ND; ++j) sd = fmaf(d[j], p.E[i * ND + j], sd);
          v += sd;
        }
        xn[i] = v;
      }
    } else {
      constexpr int NS = Tableau<S::INTEG>::NS;
      float k[NM_MAX_STAGES][NX];
      static_for<0, NS>([&](auto sc) {
        constexpr int St = decltype(sc)::value;
        float xs[NX];
        stage_input<S::INTEG, St, NX>(p.h, x, k, xs);
        Rhs<S>::eval(p, xs, u, k[St]);
      });
      combine_stages<S::INTEG, NX>(p.h, x, k, xn);
    }
  }

  // given lam = dL/dx+ :  gx = (dx+/dx)^T lam,  gu = (dx+/du)^T lam   (gx, gu are overwritten)
  static __device__ __forceinline__ void step_vjp(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                                  const float (&lam)[NX], float (&gx)[NX], float (&gu)[NUA]) {
    float unused[NM_MAX_CONST];
    step_vjp_impl<false>(p, x, u, lam, gx, gu, unused);
  }
  // the same, also accumulating the gradient w.r.t. the white-box constants into gth
  static __device__ __forceinline__ void step_vjp_theta(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                                        const float (&lam)[NX], float (&gx)[NX], float (&gu)[NUA], float (&gth)[NM_MAX_CONST]) {
    step_vjp_impl<true>(p, x, u, lam, gx, gu, gth);
  }
  template <bool THETA>
  static __device__ __forceinline__ void step_vjp_impl(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                                       const float (&lam)[NX], float (&gx)[NX], float (&gu)[NUA], float (&gth)[NM_MAX_CONST]) {
#pragma unroll
    for (int j = 0; j < NUA; ++j) gu[j] = 0.f;
    if constexpr (S::DYN == NM_DYN_SSM) {
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) s = fmaf(p.A[i * NX + j], lam[i], s);
        gx[j] = s;
      }
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) s = fmaf(p.Bm[i * NU + j], lam[i], s);
        gu[j] = s;
      }
    } else {
      constexpr int NS = Tableau<S::INTEG>::NS;
      const float h = p.h;
      float k[NM_MAX_STAGES][NX], xs[NM_MAX_STAGES][NX], gk[NM_MAX_STAGES][NX];
      static_for<0, NS>([&](auto sc) {                   // recompute the stage inputs
        constexpr int St = decltype(sc)::value;
        stage_input<S::INTEG, St, NX>(h, x, k, xs[St]);
        if constexpr (St + 1 < NS) Rhs<S>::eval(p, xs[St], u, k[St]);
#pragma unroll
        for (int i = 0; i < NX; ++i) gk[St][i] = 0.f;
      });
#pragma unroll
      for (int i = 0; i < NX; ++i) gx[i] = lam[i];
      static_for_rev<0, NS>([&](auto sc) {
        constexpr int St = decltype(sc)::value;
        float w[NX], gxs[NX];
        stage_cotangent<S::INTEG, St, NX>(h, lam, gk, w);
#pragma unroll
        for (int i = 0; i < NX; ++i) gxs[i] = 0.f;
        Rhs<S>::vjp(p, xs[St], u, w, gxs, gu);
        if constexpr (THETA) rhs_vjp_theta<S>(p, xs[St], u, w, gth);
        stage_push<S::INTEG, St, NX>(h, gxs, gx, gk);
      });
    }
  }
};

}  // namespace nm
