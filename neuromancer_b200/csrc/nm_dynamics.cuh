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

// ------------------------------------------------------------------------------------------------
// explicit one-step schemes whose stage s reads only stage s-1:  xs = x + a_s*h*k_{s-1}
template <int INTEG> struct Tableau;
template <> struct Tableau<NM_INT_EULER> { static constexpr int NS = 1; };
template <> struct Tableau<NM_INT_RK2> { static constexpr int NS = 2; };
template <> struct Tableau<NM_INT_RK4> { static constexpr int NS = 4; };

// stage input, in the reference's operation order
template <int INTEG, int NX>
__device__ __forceinline__ void stage_input(int s, float h, const float (&x)[NX], const float (&kprev)[NX],
                                            float (&xs)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if (s == 0) xs[i] = x[i];
    else if (INTEG == NM_INT_RK4 && s == 3) xs[i] = x[i] + h * kprev[i];
    else xs[i] = x[i] + h * kprev[i] / 2.0f;
  }
}
// d xs / d k_{s-1}
template <int INTEG>
__device__ __forceinline__ float stage_coupling(int s, float h) {
  if (s == 0) return 0.f;
  if (INTEG == NM_INT_RK4 && s == 3) return h;
  return h * 0.5f;
}
// d x+ / d k_s
template <int INTEG>
__device__ __forceinline__ float stage_weight(int s, float h) {
  if (INTEG == NM_INT_EULER) return h;
  if (INTEG == NM_INT_RK2) return s == 1 ? h : 0.f;
  return (s == 0 || s == 3) ? h / 6.0f : h / 3.0f;
}
// x+ from the stage derivatives, reference operation order
template <int INTEG, int NX>
__device__ __forceinline__ void combine_stages(float h, const float (&x)[NX], const float (&k)[4][NX],
                                               float (&xn)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    if (INTEG == NM_INT_EULER) xn[i] = x[i] + h * k[0][i];
    else if (INTEG == NM_INT_RK2) xn[i] = x[i] + h * k[1][i];
    else xn[i] = x[i] + h * (k[0][i] / 6.0f + k[1][i] / 3.0f + k[2][i] / 3.0f + k[3][i] / 6.0f);
  }
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
      float k[4][NX];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float xs[NX];
        stage_input<S::INTEG, NX>(s, p.h, x, k[s > 0 ? s - 1 : 0], xs);
        Rhs<S>::eval(p, xs, u, k[s]);
      }
      combine_stages<S::INTEG, NX>(p.h, x, k, xn);
    }
  }

  // given lam = dL/dx+ :  gx = (dx+/dx)^T lam,  gu = (dx+/du)^T lam   (gx, gu are overwritten)
  static __device__ __forceinline__ void step_vjp(const NmPlan& p, const float (&x)[NX], const float (&u)[NUA],
                                                  const float (&lam)[NX], float (&gx)[NX], float (&gu)[NUA]) {
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
      float k[4][NX], xs[4][NX];
#pragma unroll
      for (int s = 0; s < NS; ++s) {                     // recompute the stage inputs
        stage_input<S::INTEG, NX>(s, h, x, k[s > 0 ? s - 1 : 0], xs[s]);
        if (s + 1 < NS) Rhs<S>::eval(p, xs[s], u, k[s]);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) gx[i] = lam[i];
      float carry[NX];                                    // cotangent reaching k_s from stage s+1's input
#pragma unroll
      for (int i = 0; i < NX; ++i) carry[i] = 0.f;
#pragma unroll
      for (int s = NS - 1; s >= 0; --s) {
        float w[NX], gxs[NX];
        const float bw = stage_weight<S::INTEG>(s, h);
#pragma unroll
        for (int i = 0; i < NX; ++i) { w[i] = fmaf(bw, lam[i], carry[i]); gxs[i] = 0.f; }
        Rhs<S>::vjp(p, xs[s], u, w, gxs, gu);
        const float cp = stage_coupling<S::INTEG>(s, h);
#pragma unroll
        for (int i = 0; i < NX; ++i) { gx[i] += gxs[i]; carry[i] = cp * gxs[i]; }
      }
    }
  }
};

}  // namespace nm
