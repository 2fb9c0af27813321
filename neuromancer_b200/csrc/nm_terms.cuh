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

// The same table by value, for the compile-time-structured term code: with constant indices the struct is scalarised
// and every entry becomes a kernel-parameter (constant bank) operand or an immediate -- no shared-memory pointer
// loads in front of the trajectory loads (those chains were 11 % of the adjoint's stall samples on c2).
template <class Args>
__device__ __forceinline__ VarTable make_vartable_regs(const NmPlan& p, const Args& a, int nx, int nu, int ny) {
  VarTable vt;
  vt.base[NM_VAR_X] = a.x_traj; vt.T[NM_VAR_X] = p.nsteps + 1; vt.D[NM_VAR_X] = nx;
  vt.base[NM_VAR_U] = a.u_traj; vt.T[NM_VAR_U] = p.nsteps; vt.D[NM_VAR_U] = nu;
  vt.base[NM_VAR_Y] = a.y_traj; vt.T[NM_VAR_Y] = p.nsteps; vt.D[NM_VAR_Y] = ny;
#pragma unroll
  for (int e = 0; e < NM_MAX_EXO; ++e) {
    vt.base[NM_VAR_EXO0 + e] = a.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
  }
  return vt;
}

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

// BarrierLoss entries with value v < 0 (reference loss.py:196-214, 240-246): barrier(v), nan / +inf -> 0, clamped to
// [0, ub]; `df` = its derivative with autograd's conventions (the masked assignments and the clamp outside [0, ub] pass none)
// (out of line on purpose: six transcendental variants inlined at every runtime-descriptor call site grew the FFMA adjoint
// ten-fold; only BarrierLoss plans ever take the call)
static __device__ __noinline__ float barrier_eval_ool(float a, float shift, float ub, int kind, float v, float& df) {
  float f;
  switch (kind) {
    case NM_BARRIER_LOG10:    f = -log10f(-v); df = -0.4342944819032518f / v; break;
    case NM_BARRIER_LOG:      f = -logf(-v); df = -1.f / v; break;
    case NM_BARRIER_INVERSE:  f = 1.f / (-v); df = 1.f / (v * v); break;
    case NM_BARRIER_SOFTEXP:  { const float e = expf(a * v); f = (e - 1.f) / a + a; df = e; } break;
    case NM_BARRIER_SOFTLOG:  { const float q = 1.f + a * (-v - a); f = -logf(q) / a; df = 1.f / q; } break;
    default:                  f = expf(v + shift); df = f; break;
  }
  if (!(f == f) || f == INFINITY) { f = 0.f; df = 0.f; }
  if (f < 0.f) { f = 0.f; df = 0.f; }
  else if (f > ub) { f = ub; df = 0.f; }
  return f;
}
__device__ __forceinline__ float barrier_eval(const NmPlan& p, int kind, float v, float& df) {
  return barrier_eval_ool(p.barrier_alpha, p.barrier_shift, p.barrier_ub, kind, v, df);
}

__device__ __forceinline__ float term_penalty(const NmPlan& pl, const NmTerm& tm, float v) {
  const int norm = NM_TERM_NORM(tm.norm), bk = NM_TERM_BARRIER(tm.norm);
  if (bk != NM_BARRIER_NONE && !(v >= 0.f)) { float df; return barrier_eval(pl, bk, v, df); }
  if (tm.cmp == NM_CMP_EQ) return norm == 1 ? fabsf(v) : v * v;
  const float p = v > 0.f ? v : 0.f;
  return norm == 1 ? p : p * p;
}

// d penalty / d value with autograd's conventions: relu'(0) = 0, sign(0) = 0
__device__ __forceinline__ float term_dpenalty(const NmPlan& pl, const NmTerm& tm, float v) {
  const int norm = NM_TERM_NORM(tm.norm), bk = NM_TERM_BARRIER(tm.norm);
  if (bk != NM_BARRIER_NONE && !(v >= 0.f)) { float df; barrier_eval(pl, bk, v, df); return df; }
  if (tm.cmp == NM_CMP_EQ) return norm == 1 ? (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) : 2.f * v;
  return norm == 1 ? (v > 0.f ? 1.f : 0.f) : (v > 0.f ? 2.f * v : 0.f);
}

__device__ __forceinline__ bool atom_is_tbcast(const NmAtom& at, const NmTerm& tm) { return at.t_len == 1 && tm.t_len > 1; }

// sum over this trajectory's elements of term k
__device__ __forceinline__ float term_sum_traj(const NmPlan& p, const NmTerm& tm, const VarTable& vt, int64_t b) {
  float s = 0.f;
  for (int te = 0; te < tm.t_len; ++te)
    for (int de = 0; de < tm.d_len; ++de) s += term_penalty(p, tm, term_value(tm, vt, b, te, de));
  return s;
}

// Gradient of  sum_k coef[k] * sum_elements penalty_k  w.r.t. entry (var, t_idx, d) of trajectory b,
// accumulated into g[d] for d < dim.  coef[k] already holds upstream * weight / element count.
template <int MAXD, int MODE = 0>
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
      if (MODE == 1 && atom_is_tbcast(at, tm)) continue;
      if (MODE == 2 && !atom_is_tbcast(at, tm)) continue;
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
            float dv = term_dpenalty(p, tm, v) * side_sign;
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

// Time-broadcast atoms only (MODE 2 of term_grad_accum), evaluated by a whole WARP for one trajectory entry row: the
// range of time elements the anchored entry participates in is strided over the lanes; the caller sums g over the lanes.
template <int MAXD>
__device__ __forceinline__ void term_grad_accum_tb_lanes(const NmPlan& p, const VarTable& vt, const float* __restrict__ coef,
                                                         int64_t b, int var, int t_idx, int dim, int lane, float (&g)[MAXD]) {
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
      if (at.var != var || !atom_is_tbcast(at, tm) || t_idx != at.t_start) continue;
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
        for (int te = lane; te < tm.t_len; te += 32)
          for (int de = de_lo; de < de_hi; ++de) {
            const float v = term_value(tm, vt, b, te, de);
            float dv = term_dpenalty(p, tm, v) * side_sign;
            if (side.op == NM_OP_SUB && second) dv = -dv;
            else if (side.op == NM_OP_MUL) dv *= atom_value(second ? side.a : side.b, vt, b, te, de);
            acc += dv;
          }
        g[d] = fmaf(ck * at.scale, acc, g[d]);
      }
    }
  }
}

// ---- pre-resolved atoms: hoist the (b, te) addressing out of the per-dimension loop -------------
struct AtomRef { const float* p; int dstep; float scale, offset; };
__device__ __forceinline__ AtomRef atom_ref(const NmAtom& a, const VarTable& vt, int64_t b, int te) {
  AtomRef r;
  r.scale = a.scale; r.offset = a.offset; r.dstep = 0; r.p = nullptr;
  if (a.var >= 0) {
    const int t = a.t_start + (a.t_len == 1 ? 0 : te);
    r.p = vt.base[a.var] + (b * vt.T[a.var] + t) * vt.D[a.var] + a.d_start;
    r.dstep = a.d_len == 1 ? 0 : 1;
  }
  return r;
}
__device__ __forceinline__ float ref_value(const AtomRef& r, int de) {
  return r.p == nullptr ? r.offset : fmaf(r.scale, r.p[de * r.dstep], r.offset);
}
struct TermRef { AtomRef la, lb, ra, rb; };
__device__ __forceinline__ TermRef term_ref(const NmTerm& tm, const VarTable& vt, int64_t b, int te) {
  TermRef r;
  r.la = atom_ref(tm.lhs.a, vt, b, te);
  r.ra = atom_ref(tm.rhs.a, vt, b, te);
  if (tm.lhs.op != NM_OP_NONE) r.lb = atom_ref(tm.lhs.b, vt, b, te); else { r.lb = r.la; }
  if (tm.rhs.op != NM_OP_NONE) r.rb = atom_ref(tm.rhs.b, vt, b, te); else { r.rb = r.ra; }
  return r;
}
__device__ __forceinline__ float combine(int op, float a, float c) {
  return op == NM_OP_NONE ? a : (op == NM_OP_ADD ? a + c : (op == NM_OP_SUB ? a - c : a * c));
}
__device__ __forceinline__ float term_value_ref(const NmTerm& tm, const TermRef& r, int de) {
  const float l = combine(tm.lhs.op, ref_value(r.la, de), tm.lhs.op != NM_OP_NONE ? ref_value(r.lb, de) : 0.f);
  const float q = combine(tm.rhs.op, ref_value(r.ra, de), tm.rhs.op != NM_OP_NONE ? ref_value(r.rb, de) : 0.f);
  return tm.cmp == NM_CMP_GT ? q - l : l - q;
}

// d(sum_k coef_k * sum penalty_k) / d V[b, t_idx, d]  for ONE trajectory entry
// An atom is "time-broadcast" when it is a single time entry compared against a whole time range
// (x[:, [-1], :] against every r_t).  Its gradient is a sum over the range; evaluated per entry in a
// thread-per-entry kernel it diverges badly, so the entry kernel skips it (MODE 1) and the adjoint
// sweep -- where all lanes sit at the same t -- adds it (MODE 2).  MODE 0 handles every atom.

template <int MODE>
__device__ __forceinline__ float term_grad_entry(const NmPlan& p, const VarTable& vt, const float* __restrict__ coef,
                                                 int64_t b, int var, int t_idx, int d) {
  float g = 0.f;
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
      if (MODE == 1 && atom_is_tbcast(at, tm)) continue;
      if (MODE == 2 && !atom_is_tbcast(at, tm)) continue;
      int te_lo, te_hi, de_lo, de_hi;
      if (at.t_len == 1) { if (t_idx != at.t_start) continue; te_lo = 0; te_hi = tm.t_len; }
      else { const int te = t_idx - at.t_start; if (te < 0 || te >= at.t_len) continue; te_lo = te; te_hi = te + 1; }
      if (at.d_len == 1) { if (d != at.d_start) continue; de_lo = 0; de_hi = tm.d_len; }
      else { const int de = d - at.d_start; if (de < 0 || de >= at.d_len) continue; de_lo = de; de_hi = de + 1; }
      const float side_sign = (which < 2) ? (tm.cmp == NM_CMP_GT ? -1.f : 1.f) : (tm.cmp == NM_CMP_GT ? 1.f : -1.f);
      float acc = 0.f;
      for (int te = te_lo; te < te_hi; ++te) {
        const TermRef r = term_ref(tm, vt, b, te);
        for (int de = de_lo; de < de_hi; ++de) {
          float dv = term_dpenalty(p, tm, term_value_ref(tm, r, de)) * side_sign;
          if (side.op == NM_OP_SUB && second) dv = -dv;
          else if (side.op == NM_OP_MUL) dv *= ref_value(second ? (which < 2 ? r.la : r.ra) : (which < 2 ? r.lb : r.rb), de);
          acc += dv;
        }
      }
      g = fmaf(ck * at.scale, acc, g);
    }
  }
  return g;
}


// ================================================================================================
// Compile-time-structured terms.  The shape-class header (build.py) records, per term, the comparator,
// norm, side ops and per atom the variable it reads and whether it is broadcast along time / dim;
// slice starts, lengths, scale, offset and weight stay runtime (NmPlan).  With the structure known the
// evaluation compiles to straight-line code (no descriptor decoding), so scoring and the term
// gradients are fused into the rollout / adjoint sweeps at a few dozen instructions per step.
template <class S, int K, int W>
struct AtomS {                                   // W: 0 lhs.a, 1 lhs.b, 2 rhs.a, 3 rhs.b
  static constexpr int VAR = S::a_var(K, W);     // -2 unused, -1 constant, >= 0 NM_VAR_*
  static constexpr bool TB = S::a_tb(K, W) != 0; // single time entry broadcast over the term's range
  static constexpr bool DB = S::a_db(K, W) != 0;
};
template <int W>
__device__ __forceinline__ const NmAtom& pick_atom(const NmTerm& tm) {
  if constexpr (W == 0) return tm.lhs.a; else if constexpr (W == 1) return tm.lhs.b;
  else if constexpr (W == 2) return tm.rhs.a; else return tm.rhs.b;
}
template <class S, int K, int W>
__device__ __forceinline__ float atom_value_s(const NmTerm& tm, const VarTable& vt, int64_t b, int te, int de) {
  using A = AtomS<S, K, W>;
  const NmAtom& a = pick_atom<W>(tm);
  if constexpr (A::VAR < 0) return a.offset;
  else {
    const int t = a.t_start + (A::TB ? 0 : te);
    const int d = a.d_start + (A::DB ? 0 : de);
    const float v = vt.base[A::VAR][(b * vt.T[A::VAR] + t) * vt.D[A::VAR] + d];
    return fmaf(a.scale, v, a.offset);
  }
}
template <class S, int K, int SIDE>
__device__ __forceinline__ float side_value_s(const NmTerm& tm, const VarTable& vt, int64_t b, int te, int de) {
  constexpr int OP = S::t_op(K, SIDE);
  const float a = atom_value_s<S, K, 2 * SIDE>(tm, vt, b, te, de);
  if constexpr (OP == NM_OP_NONE) return a;
  else {
    const float c = atom_value_s<S, K, 2 * SIDE + 1>(tm, vt, b, te, de);
    if constexpr (OP == NM_OP_ADD) return a + c; else if constexpr (OP == NM_OP_SUB) return a - c; else return a * c;
  }
}
template <class S, int K>
__device__ __forceinline__ float term_value_s(const NmTerm& tm, const VarTable& vt, int64_t b, int te, int de) {
  const float l = side_value_s<S, K, 0>(tm, vt, b, te, de);
  const float r = side_value_s<S, K, 1>(tm, vt, b, te, de);
  if constexpr (S::t_cmp(K) == NM_CMP_GT) return r - l; else return l - r;
}
template <class S, int K>
__device__ __forceinline__ float term_penalty_s(const NmPlan& pl, float v) {
  constexpr int NORM = NM_TERM_NORM(S::t_norm(K)), BK = NM_TERM_BARRIER(S::t_norm(K));
  if constexpr (BK != NM_BARRIER_NONE) { if (!(v >= 0.f)) { float df; return barrier_eval(pl, BK, v, df); } }
  if constexpr (S::t_cmp(K) == NM_CMP_EQ) { if constexpr (NORM == 1) return fabsf(v); else return v * v; }
  else { const float p = v > 0.f ? v : 0.f; if constexpr (NORM == 1) return p; else return p * p; }
}
template <class S, int K>
__device__ __forceinline__ float term_dpenalty_s(const NmPlan& pl, float v) {
  constexpr int NORM = NM_TERM_NORM(S::t_norm(K)), BK = NM_TERM_BARRIER(S::t_norm(K));
  if constexpr (BK != NM_BARRIER_NONE) { if (!(v >= 0.f)) { float df; barrier_eval(pl, BK, v, df); return df; } }
  if constexpr (S::t_cmp(K) == NM_CMP_EQ) {
    if constexpr (NORM == 1) return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); else return 2.f * v;
  } else {
    if constexpr (NORM == 1) return v > 0.f ? 1.f : 0.f; else return v > 0.f ? 2.f * v : 0.f;
  }
}

// add this trajectory's penalty sums to tsum[k]
template <class S>
__device__ __forceinline__ void score_traj_s(const NmPlan& p, const VarTable& vt, int64_t b,
                                             float (&tsum)[S::NTERMS > 0 ? S::NTERMS : 1]) {
  static_for<0, S::NTERMS>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    const NmTerm& tm = p.terms[K];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int tl = tm.t_len, dl = tm.d_len;
    int te = 0;
    for (; te + 3 < tl; te += 4) {                      // four independent chains: the (L2-latency) loads overlap
      for (int de = 0; de < dl; ++de) {
        s0 += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te, de));
        s1 += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te + 1, de));
        s2 += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te + 2, de));
        s3 += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te + 3, de));
      }
    }
    for (; te < tl; ++te)
      for (int de = 0; de < dl; ++de) s0 += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te, de));
    s0 += s2; s1 += s3;
    tsum[K] += s0 + s1;
  });
}

// d(sum_k coef_k * penalty sums)/d VAR[b, t_idx, 0..DIM)  accumulated into g
// MODE 0: every atom; 1: skip time-broadcast atoms; 2: time-broadcast atoms only (see term_grad_entry)
// (te0, te_stride): the time range of a time-broadcast atom may be strided over the lanes of a warp (the caller sums)
template <class S, int VAR, int DIM, int DIMA, int MODE = 0>
__device__ __forceinline__ void term_grad_s(const NmPlan& p, const VarTable& vt, const float* __restrict__ coef,
                                            int64_t b, int t_idx, float (&g)[DIMA], int te0 = 0, int te_stride = 1) {
  static_for<0, S::NTERMS>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    static_for<0, 4>([&](auto wc) {
      constexpr int W = decltype(wc)::value;
      using A = AtomS<S, K, W>;
      if constexpr (A::VAR == VAR && !(MODE == 1 && A::TB) && !(MODE == 2 && !A::TB)) {
        const NmTerm& tm = p.terms[K];
        const NmAtom& at = pick_atom<W>(tm);
        constexpr float side_sign = (W < 2) ? (S::t_cmp(K) == NM_CMP_GT ? -1.f : 1.f) : (S::t_cmp(K) == NM_CMP_GT ? 1.f : -1.f);
        constexpr int OP = S::t_op(K, W / 2);
        const float ck = coef[K] * at.scale * side_sign * ((OP == NM_OP_SUB && (W & 1)) ? -1.f : 1.f);
        if constexpr (!A::TB && !A::DB) {
          // one element per (t_idx, d): branch-free -- the element index is clamped into the term's range and the
          // contribution selected afterwards, so the loads of all terms are independent and issue back to back
          const int te = t_idx - at.t_start;
          const bool act_t = te >= 0 && te < at.t_len;
          const int te_c = act_t ? te : 0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            const int de = d - at.d_start;
            const bool in_d = de >= 0 && de < at.d_len;
            const int de_c = in_d ? de : 0;
            float dv = term_dpenalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te_c, de_c));
            if constexpr (OP == NM_OP_MUL) dv *= atom_value_s<S, K, (W ^ 1)>(tm, vt, b, te_c, de_c);
            g[d] = (act_t && in_d) ? fmaf(ck, dv, g[d]) : g[d];
          }
        } else {
          int te_lo, te_hi;
          if constexpr (A::TB) {
            if (t_idx != at.t_start) return;
            te_lo = 0; te_hi = tm.t_len;
          } else {
            const int te = t_idx - at.t_start;
            if (te < 0 || te >= at.t_len) return;
            te_lo = te; te_hi = te + 1;
          }
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            int de_lo, de_hi;
            if constexpr (A::DB) { if (d != at.d_start) continue; de_lo = 0; de_hi = tm.d_len; }
            else { const int de = d - at.d_start; if (de < 0 || de >= at.d_len) continue; de_lo = de; de_hi = de + 1; }
            float acc = 0.f;
            for (int te = te_lo + (A::TB ? te0 : 0); te < te_hi; te += (A::TB ? te_stride : 1))
              for (int de = de_lo; de < de_hi; ++de) {
                float dv = term_dpenalty_s<S, K>(p, term_value_s<S, K>(tm, vt, b, te, de));
                if constexpr (OP == NM_OP_MUL) dv *= atom_value_s<S, K, (W ^ 1)>(tm, vt, b, te, de);
                acc += dv;
              }
            g[d] = fmaf(ck, acc, g[d]);
          }
        }
      }
    });
  });
}

// out-of-line generic (runtime descriptor) fallbacks: keep them out of the sweeps' register budget
template <int MAXD, int MODE>
__device__ __noinline__ void term_grad_accum_ool(const NmPlan& p, const VarTable& vt, const float* __restrict__ coef,
                                                 int64_t b, int var, int t_idx, int dim, float* g) {
  float tmp[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) tmp[d] = 0.f;
  term_grad_accum<MAXD, MODE>(p, vt, coef, b, var, t_idx, dim, tmp);
#pragma unroll
  for (int d = 0; d < MAXD; ++d) if (d < dim) g[d] += tmp[d];
}

// coefficient of term k in the scalar being differentiated:
//   (d term_k + d objective|penalty + d loss) * weight / (global element count)
__device__ __forceinline__ float term_coef(const NmPlan& p, int k, const float* __restrict__ grad_scalars, int64_t b_total) {
  const NmTerm& tm = p.terms[k];
  float up = 1.f;
  if (grad_scalars != nullptr)
    up = grad_scalars[k] + grad_scalars[p.n_terms + (tm.is_objective ? 0 : 1)] + grad_scalars[p.n_terms + 2];
  return up * tm.weight / (static_cast<float>(b_total) * static_cast<float>(tm.t_len) * static_cast<float>(tm.d_len));
}

}  // namespace nm
