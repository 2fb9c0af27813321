// nm_terms.cuh -- PenaltyLoss term scoring and its gradient w.r.t. trajectory entries.
//
// Follows the reference's comparator arithmetic (constraint.py:85-98 LT, :122-134 GT, :157-173 Eq):
//   LT: value = l - r ; GT: value = r - l ; penalty = relu(value) (squared for norm 2), loss = mean
//   Eq: value = l - r ; penalty = |value| (norm 1, F.l1_loss) or value^2 (norm 2, F.mse_loss)
// with the two sides broadcast to a common [t_len, d_len] element shape per trajectory
// (e.g. the terminal constraint  x[:, [-1], :] > r - 0.01  broadcasts x_N against every r_t).
#pragma once
#include "nm_device.cuh"

namespace nm {

struct VarTable {                       // trajectory variables a term atom can load from
  const float* base[NM_VAR_EXO0 + NM_MAX_EXO];
  int T[NM_VAR_EXO0 + NM_MAX_EXO];      // time length
  int D[NM_VAR_EXO0 + NM_MAX_EXO];      // last dim
};

__device__ __forceinline__ float atom_value(const NmAtom& a, const VarTable& vt, int64_t b, int te, int de) {
  if (a.var < 0) return a.offset;
  const int t = a.t_start + (a.t_len == 1 ? 0 : te);
  const int d = a.d_start + (a.d_len == 1 ? 0 : de);
  // plain (coherent) load: the forward kernel reads back entries this same thread stored earlier
  const float v = vt.base[a.var][(b * vt.T[a.var] + t) * vt.D[a.var] + d];
  return fmaf(a.scale, v, a.offset);
}

__device__ __forceinline__ float side_value(const NmSide& s, const VarTable& vt, int64_t b, int te, int de) {
  const float a = atom_value(s.a, vt, b, te, de);
  if (s.op == NM_OP_NONE) return a;
  const float c = atom_value(s.b, vt, b, te, de);
  return s.op == NM_OP_ADD ? a + c : (s.op == NM_OP_SUB ? a - c : a * c);
}

__device__ __forceinline__ float term_value(const NmTerm& tm, const VarTable& vt, int64_t b, int te, int de) {
  const float l = side_value(tm.lhs, vt, b, te, de);
  const float r = side_value(tm.rhs, vt, b, te, de);
  return tm.cmp == NM_CMP_GT ? r - l : l - r;
}

__device__ __forceinline__ float term_penalty(const NmTerm& tm, float v) {
  if (tm.cmp == NM_CMP_EQ) return tm.norm == 1 ? fabsf(v) : v * v;
  const float p = v > 0.f ? v : 0.f;
  return tm.norm == 1 ? p : p * p;
}

// d penalty / d value with autograd's conventions: relu'(0) = 0, sign(0) = 0
__device__ __forceinline__ float term_dpenalty(const NmTerm& tm, float v) {
  if (tm.cmp == NM_CMP_EQ) return tm.norm == 1 ? (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) : 2.f * v;
  return tm.norm == 1 ? (v > 0.f ? 1.f : 0.f) : (v > 0.f ? 2.f * v : 0.f);
}

// sum over this trajectory's elements of term k
__device__ __forceinline__ float term_sum_traj(const NmTerm& tm, const VarTable& vt, int64_t b) {
  float s = 0.f;
  for (int te = 0; te < tm.t_len; ++te)
    for (int de = 0; de < tm.d_len; ++de) s += term_penalty(tm, term_value(tm, vt, b, te, de));
  return s;
}

// Gradient of  sum_k coef[k] * sum_elements penalty_k  w.r.t. entry (var, t_idx, d) of trajectory b,
// accumulated into g[d] for d < dim.  coef[k] already holds upstream * weight / element count.
template <int MAXD>
__device__ __forceinline__ void term_grad_accum(const NmPlan& p, const VarTable& vt, const float* __restrict__ coef,
                                                int64_t b, int var, int t_idx, int dim, float (&g)[MAXD]) {
  for (int k = 0; k < p.n_terms; ++k) {
    const NmTerm& tm = p.terms[k];
    const float ck = coef[k];
    if (ck == 0.f) continue;
#pragma unroll
    for (int which = 0; which < 4; ++which) {
      const NmSide& side = which < 2 ? tm.lhs : tm.rhs;
      const bool second = which & 1;
      if (second && side.op == NM_OP_NONE) continue;
      const NmAtom& at = second ? side.b : side.a;
      if (at.var != var) continue;
      // time elements this atom entry participates in
      int te_lo, te_hi;
      if (at.t_len == 1) {
        if (t_idx != at.t_start) continue;
        te_lo = 0; te_hi = tm.t_len;
      } else {
        const int te = t_idx - at.t_start;
        if (te < 0 || te >= at.t_len) continue;
        te_lo = te; te_hi = te + 1;
      }
      // sign of d value / d side
      const float side_sign = (which < 2) ? (tm.cmp == NM_CMP_GT ? -1.f : 1.f) : (tm.cmp == NM_CMP_GT ? 1.f : -1.f);
#pragma unroll
      for (int d = 0; d < MAXD; ++d) {
        if (d >= dim) break;
        int de_lo, de_hi;
        if (at.d_len == 1) {
          if (d != at.d_start) continue;
          de_lo = 0; de_hi = tm.d_len;
        } else {
          const int de = d - at.d_start;
          if (de < 0 || de >= at.d_len) continue;
          de_lo = de; de_hi = de + 1;
        }
        float acc = 0.f;
        for (int te = te_lo; te < te_hi; ++te)
          for (int de = de_lo; de < de_hi; ++de) {
            const float v = term_value(tm, vt, b, te, de);
            float dv = term_dpenalty(tm, v) * side_sign;
            // d side / d atom
            if (side.op == NM_OP_SUB && second) dv = -dv;
            else if (side.op == NM_OP_MUL) dv *= atom_value(second ? side.a : side.b, vt, b, te, de);
            acc += dv;
          }
        g[d] = fmaf(ck * at.scale, acc, g[d]);
      }
    }
  }
}

}  // namespace nm
