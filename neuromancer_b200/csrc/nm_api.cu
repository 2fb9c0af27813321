// nm_api.cu -- the C ABI of include/nm_rollout.h: kernel registry, plan validation / signature,
// dispatch to the per-shape-class launchers, and the two small finalise kernels.
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <cmath>
#include <cuda_runtime.h>
#include "nm_registry.h"
#include "nm_terms.cuh"
#include "nm_collective.cuh"
#include <algorithm>

namespace {

std::mutex g_mu;
NmKernelEntry* g_head = nullptr;
int g_count = 0;
thread_local char g_err[512] = "";
thread_local int g_launches = 0;

// Entry for a shape-class signature.  An entry whose compiled term structure equals `term_sig` is
// preferred (fused term evaluation, *fused = 1); otherwise any entry of the shape class serves the plan
// through the runtime-descriptor term kernels (*fused = 0).
NmKernelEntry* find_entry(const char* sig, const char* term_sig = nullptr, int* fused = nullptr) {
  std::lock_guard<std::mutex> lock(g_mu);
  NmKernelEntry* any = nullptr;
  if (fused) *fused = 0;
  for (NmKernelEntry* e = g_head; e; e = e->next) {
    if (std::strcmp(e->signature, sig) != 0) continue;
    if (term_sig && term_sig[0] && std::strcmp(e->term_signature, term_sig) == 0) { if (fused) *fused = 1; return e; }
    if (!any) any = e;
  }
  return any;
}

void append_mlp(std::string& s, const NmMlp& m) {
  char buf[64];
  s += "[";
  for (int i = 0; i <= m.n_layers; ++i) { std::snprintf(buf, sizeof(buf), i ? "-%d" : "%d", m.sizes[i]); s += buf; }
  std::snprintf(buf, sizeof(buf), "]a%d_b%d_o%d", m.act, m.has_bias ? 1 : 0, m.bounds);
  s += buf;
}

int validate(const NmPlan* p) {
  if (!p) { nm_set_error("null plan"); return NM_ERR_BAD_PLAN; }
  if (p->abi_version != NM_ABI_VERSION) { nm_set_error("plan.abi_version does not match the library"); return NM_ERR_ABI; }
  if (p->nx < 1 || p->nx > NM_MAX_DIM || p->nu < 0 || p->nu > NM_MAX_DIM || p->ny < 0 || p->ny > NM_MAX_DIM) { nm_set_error("nx/nu/ny out of range"); return NM_ERR_BAD_PLAN; }
  if (p->nsteps < 1) { nm_set_error("nsteps must be >= 1"); return NM_ERR_BAD_PLAN; }
  if (p->n_exo < 0 || p->n_exo > NM_MAX_EXO || p->n_terms < 0 || p->n_terms > NM_MAX_TERMS || p->n_segs < 0 || p->n_segs > NM_MAX_SEGS) { nm_set_error("count field out of range"); return NM_ERR_BAD_PLAN; }
  if (p->has_policy && (p->policy.n_layers < 1 || p->policy.n_layers > NM_MAX_LAYERS)) { nm_set_error("policy layer count out of range"); return NM_ERR_BAD_PLAN; }
  if (NM_RHS_HAS_MLP(p->rhs_kind) && (p->rhs_mlp.n_layers < 1 || p->rhs_mlp.n_layers > NM_MAX_LAYERS)) { nm_set_error("rhs layer count out of range"); return NM_ERR_BAD_PLAN; }
  for (int i = 0; i < p->n_segs; ++i) {
    const NmSegment& s = p->segs[i];
    if (s.kind == NM_SEG_EXO && (s.index < 0 || s.index >= p->n_exo || p->exo[s.index].width != s.width)) { nm_set_error("policy segment refers to a bad exo input"); return NM_ERR_BAD_PLAN; }
    if (s.kind == NM_SEG_PREVIEW) {
      if (s.index < 0 || s.index >= p->n_exo || s.preview_len < 0 || p->exo[s.index].width * (s.preview_len + 1) != s.width) { nm_set_error("preview segment shape mismatch"); return NM_ERR_BAD_PLAN; }
      if (p->pad_mode < NM_PAD_CONSTANT || p->pad_mode > NM_PAD_REFLECT) { nm_set_error("bad pad_mode"); return NM_ERR_BAD_PLAN; }
      if (p->pad_mode == NM_PAD_REFLECT && s.preview_len >= p->exo[s.index].steps) { nm_set_error("reflect padding needs preview_len < signal length"); return NM_ERR_BAD_PLAN; }
    }
  }
  for (int e = 0; e < p->n_exo; ++e)
    if (p->exo[e].width < 1 || p->exo[e].steps < 1) { nm_set_error("exo shape invalid"); return NM_ERR_BAD_PLAN; }
  if (p->dist_exo >= p->n_exo) { nm_set_error("dist_exo out of range"); return NM_ERR_BAD_PLAN; }
  if (p->act_exo >= p->n_exo || (p->act_exo >= 0 && (p->has_policy || p->exo[p->act_exo].width != p->nu))) { nm_set_error("act_exo invalid"); return NM_ERR_BAD_PLAN; }
  const int T[3] = {p->nsteps + 1, p->nsteps, p->nsteps};
  const int D[3] = {p->nx, p->nu, p->ny};
  for (int k = 0; k < p->n_terms; ++k) {
    const NmTerm& t = p->terms[k];
    if ((NM_TERM_NORM(t.norm) != 1 && NM_TERM_NORM(t.norm) != 2) || NM_TERM_BARRIER(t.norm) > NM_BARRIER_EXPSHIFT || (t.norm >> 8) != 0 ||
        (NM_TERM_BARRIER(t.norm) != NM_BARRIER_NONE && t.is_objective)) {
      nm_set_error("term norm must be 1 or 2 (+ a barrier kind in bits 4-7, constraints only)"); return NM_ERR_BAD_PLAN;
    }
    if (t.t_len < 1 || t.d_len < 1) { nm_set_error("term element shape invalid"); return NM_ERR_BAD_PLAN; }
    const NmAtom* atoms[4] = {&t.lhs.a, &t.lhs.b, &t.rhs.a, &t.rhs.b};
    const bool live[4] = {true, t.lhs.op != NM_OP_NONE, true, t.rhs.op != NM_OP_NONE};
    for (int i = 0; i < 4; ++i) {
      if (!live[i]) continue;
      const NmAtom& a = *atoms[i];
      if (a.var < 0) continue;
      int vt, vd;
      if (a.var < NM_VAR_EXO0) { vt = T[a.var]; vd = D[a.var]; }
      else { const int e = a.var - NM_VAR_EXO0; if (e >= p->n_exo) { nm_set_error("term atom refers to a missing exo input"); return NM_ERR_BAD_PLAN; } vt = p->exo[e].steps; vd = p->exo[e].width; }
      if (a.t_start < 0 || a.t_len < 1 || a.t_start + a.t_len > vt || a.d_start < 0 || a.d_len < 1 || a.d_start + a.d_len > vd) { nm_set_error("term atom slice out of range"); return NM_ERR_BAD_PLAN; }
      if ((a.t_len != 1 && a.t_len != t.t_len) || (a.d_len != 1 && a.d_len != t.d_len)) { nm_set_error("term atom shape does not broadcast to the term shape"); return NM_ERR_BAD_PLAN; }
    }
  }
  return NM_OK;
}

// structure of the loss terms (what build.py compiles into a shape class): comparator, norm, list, side
// ops and per atom the variable and its time / dim broadcast flags
void append_atom_sig(std::string& s, const NmAtom& a, const NmTerm& t) {
  if (a.var < 0) { s += "k"; return; }
  char buf[32];
  std::snprintf(buf, sizeof(buf), "v%d%c%c", a.var, (a.t_len == 1 && t.t_len > 1) ? 'b' : 'r', (a.d_len == 1 && t.d_len > 1) ? 'b' : 'r');
  s += buf;
}
std::string make_term_signature(const NmPlan* p) {
  std::string s;
  for (int k = 0; k < p->n_terms; ++k) {
    const NmTerm& t = p->terms[k];
    char buf[32];
    std::snprintf(buf, sizeof(buf), "%sT%d%d%c", k ? "|" : "", t.cmp, t.norm, t.is_objective ? 'o' : 'c');
    s += buf;
    const NmSide* sides[2] = {&t.lhs, &t.rhs};
    for (int i = 0; i < 2; ++i) {
      std::snprintf(buf, sizeof(buf), "%c%d", i ? 'R' : 'L', sides[i]->op);
      s += buf;
      append_atom_sig(s, sides[i]->a, t);
      if (sides[i]->op != NM_OP_NONE) append_atom_sig(s, sides[i]->b, t);
    }
  }
  return s;
}

std::string make_signature(const NmPlan* p) {
  std::string s;
  char buf[96];
  std::snprintf(buf, sizeof(buf), "v%d_nx%d_nu%d_ny%d", NM_ABI_VERSION, p->nx, p->nu, p->ny);
  s += buf;
  if (p->has_policy) {
    s += "_pol"; append_mlp(s, p->policy);
    s += "_seg";
    for (int i = 0; i < p->n_segs; ++i) {
      const NmSegment& g = p->segs[i];
      if (g.kind == NM_SEG_PREVIEW) std::snprintf(buf, sizeof(buf), "%sp%d.%d.%d", i ? "," : "", g.index, g.width, p->exo[g.index].width);
      else std::snprintf(buf, sizeof(buf), "%s%c%d.%d", i ? "," : "", g.kind == NM_SEG_STATE ? 'x' : (g.kind == NM_SEG_OUTPUT ? 'y' : 'e'), g.kind == NM_SEG_EXO ? g.index : 0, g.width);
      s += buf;
    }
  }
  std::snprintf(buf, sizeof(buf), "_out%d_dyn%d_int%d_rhs%d", p->has_output ? 1 : 0, p->dyn_kind, p->dyn_kind == NM_DYN_ODE ? p->integrator : 0, p->rhs_kind);
  s += buf;
  if (NM_RHS_HAS_MLP(p->rhs_kind)) append_mlp(s, p->rhs_mlp);
  const int nd = p->dist_exo >= 0 ? p->exo[p->dist_exo].width : 0;
  std::snprintf(buf, sizeof(buf), "_dist%d.%d_act%d", p->dist_exo, nd, p->act_exo);
  s += buf;
  return s;
}


constexpr size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
constexpr int kScoreThreads = 256;
constexpr int kScoreMaxBlocks = 148 * 8;
constexpr size_t kGbufMaxBytes = size_t(6) << 30;     // precompute term gradients only below 6 GiB

nm::VarTable make_vartable(const NmPlan* p, const float* x, const float* u, const float* y, const float* const* exo) {
  nm::VarTable vt{};
  vt.base[NM_VAR_X] = x; vt.T[NM_VAR_X] = p->nsteps + 1; vt.D[NM_VAR_X] = p->nx;
  vt.base[NM_VAR_U] = u; vt.T[NM_VAR_U] = p->nsteps; vt.D[NM_VAR_U] = p->nu;
  vt.base[NM_VAR_Y] = y; vt.T[NM_VAR_Y] = p->nsteps; vt.D[NM_VAR_Y] = p->ny;
  for (int e = 0; e < NM_MAX_EXO; ++e) {
    vt.base[NM_VAR_EXO0 + e] = (e < p->n_exo && exo) ? exo[e] : nullptr;
    vt.T[NM_VAR_EXO0 + e] = e < p->n_exo ? p->exo[e].steps : 0;
    vt.D[NM_VAR_EXO0 + e] = e < p->n_exo ? p->exo[e].width : 0;
  }
  return vt;
}

int max_term_steps(const NmPlan* p) {
  int m = 1;
  for (int k = 0; k < p->n_terms; ++k) m = p->terms[k].t_len > m ? p->terms[k].t_len : m;
  return m;
}
int score_grid(const NmPlan* p, int64_t B) {
  const int64_t n = B * max_term_steps(p);
  int64_t g = (n + kScoreThreads - 1) / kScoreThreads;
  if (g > kScoreMaxBlocks) g = kScoreMaxBlocks;
  return g < 1 ? 1 : int(g);
}
size_t gbuf_floats(const NmPlan* p, int64_t B) {
  return size_t(B) * (size_t(p->nsteps + 1) * p->nx + size_t(p->nsteps) * (p->nu + p->ny));
}
bool use_gbuf(const NmPlan* p, int64_t B) {
  return p->n_terms > 0 && gbuf_floats(p, B) * sizeof(float) <= kGbufMaxBytes;
}

// PenaltyLoss scoring: one thread per (trajectory, time) pair; terms are processed one after another
// (small code, few registers; the re-read of the pair's entries hits L1), block partial sums in double
__global__ void __launch_bounds__(kScoreThreads) nm_term_score_kernel(const __grid_constant__ NmPlan plan,
                                                                      const __grid_constant__ nm::VarTable vt,
                                                                      int64_t B, int tmax, double* __restrict__ partials) {
  __shared__ double red[kScoreThreads / 32];
  const int64_t n = B * tmax;
#pragma unroll 1
  for (int k = 0; k < plan.n_terms; ++k) {
    const NmTerm& tm = plan.terms[k];
    float ts = 0.f;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
      const int64_t b = idx / tmax;
      const int te = int(idx - b * tmax);
      if (te < tm.t_len) {
        const nm::TermRef r = nm::term_ref(tm, vt, b, te);
        for (int de = 0; de < tm.d_len; ++de) ts += nm::term_penalty(plan, tm, nm::term_value_ref(tm, r, de));
      }
    }
    double v = static_cast<double>(ts);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < kScoreThreads / 32; ++w) s += red[w];
      partials[size_t(blockIdx.x) * NM_MAX_TERMS + k] = s;
    }
    __syncthreads();
  }
}

// Constraint matrices (reference loss.py:86-106): one thread per (trajectory, column) entry, columns fastest -- stores
// are coalesced rows, loads follow the [t][d] order of the trajectories; every output is written exactly once.
struct CMatArgs {
  float* val; float* viol; float* eq_val; float* eq_viol; float* in_val; float* in_viol;
  int64_t B; int64_t W, Weq, Win;
  int n;                                            // constraint terms
  int term[NM_MAX_TERMS];                           // index into plan.terms
  int64_t col0[NM_MAX_TERMS + 1];                   // first column in C_values
  int64_t part0[NM_MAX_TERMS];                      // first column in its partition (eq or ineq)
};
__global__ void __launch_bounds__(256) nm_constraint_matrices_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ nm::VarTable vt,
                                                                     const __grid_constant__ CMatArgs a) {
  const int64_t total = a.B * a.W;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx / a.W, col = idx - b * a.W;
    int c = 0;
    while (c + 1 < a.n && col >= a.col0[c + 1]) ++c;
    NmTerm tm = plan.terms[a.term[c]];
    const int e = int(col - a.col0[c]);
    const int te = e / tm.d_len, de = e - te * tm.d_len;
    const float v = nm::term_value(tm, vt, b, te, de);
    tm.norm = NM_TERM_NORM(tm.norm);                // C_violations hold the plain penalty, BarrierLoss or not
    const float pen = nm::term_penalty(plan, tm, v);
    if (a.val) a.val[idx] = v;
    if (a.viol) a.viol[idx] = pen;
    const bool eq = tm.cmp == NM_CMP_EQ;
    const int64_t pw = eq ? a.Weq : a.Win, pidx = b * pw + a.part0[c] + e;
    float* pv = eq ? a.eq_val : a.in_val;
    float* pp = eq ? a.eq_viol : a.in_viol;
    if (pv) pv[pidx] = v;
    if (pp) pp[pidx] = pen;
  }
}

// Gradient of the scored scalar w.r.t. every trajectory entry: one thread per entry.  Output is one
// row per trajectory,  G[b] = [ x(T+1, nx) | u(T, nu) | y(T, ny) ],  which the adjoint kernel reads.
struct TermGradArgs {
  const float* grad_scalars; const float* gx_ext; const float* gu_ext; const float* gy_ext;
  float* G; int64_t B; int64_t b_total;
};
__global__ void __launch_bounds__(kScoreThreads) nm_term_grad_kernel(const __grid_constant__ NmPlan plan,
                                                                     const __grid_constant__ nm::VarTable vt,
                                                                     const __grid_constant__ TermGradArgs a) {
  __shared__ float coef[NM_MAX_TERMS];
  if (threadIdx.x < NM_MAX_TERMS)
    coef[threadIdx.x] = threadIdx.x < plan.n_terms ? nm::term_coef(plan, threadIdx.x, a.grad_scalars, a.b_total) : 0.f;
  __syncthreads();
  const int T = plan.nsteps, nx = plan.nx, nu = plan.nu, ny = plan.ny;
  const int nxr = (T + 1) * nx, nur = T * nu, row = nxr + nur + T * ny;
  const int64_t n = a.B * row;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx / row;
    int r = int(idx - b * row);
    int var, dim;
    const float* ext;
    if (r < nxr) { var = NM_VAR_X; dim = nx; ext = a.gx_ext ? a.gx_ext + b * nxr : nullptr; }
    else if (r < nxr + nur) { r -= nxr; var = NM_VAR_U; dim = nu; ext = a.gu_ext ? a.gu_ext + b * nur : nullptr; }
    else { r -= nxr + nur; var = NM_VAR_Y; dim = ny; ext = a.gy_ext ? a.gy_ext + b * int64_t(T) * ny : nullptr; }
    const int t = r / dim, d = r - t * dim;
    float g = nm::term_grad_entry<1>(plan, vt, coef, b, var, t, d);   // time-broadcast atoms: adjoint kernel
    if (ext) g += ext[r];
    a.G[idx] = g;
  }
}

__global__ void __launch_bounds__(256) nm_loss_finalize_kernel(const __grid_constant__ NmPlan plan, const double* __restrict__ partials,
                                                               int grid, int64_t B, float* __restrict__ scalars) {
  // thread (k, j): fixed-order strided sum over the scoring blocks, then a fixed-order tree -> deterministic
  __shared__ double part[NM_MAX_TERMS][16];
  __shared__ float term[NM_MAX_TERMS];
  const int k = threadIdx.x / 16, j = threadIdx.x % 16;
  double s = 0.0;
  if (k < plan.n_terms)
    for (int c = j; c < grid; c += 16) s += partials[size_t(c) * NM_MAX_TERMS + k];
  part[k][j] = s;
  __syncthreads();
  if (threadIdx.x < plan.n_terms) {
    double v = 0.0;
    for (int q = 0; q < 16; ++q) v += part[threadIdx.x][q];
    const NmTerm& t = plan.terms[threadIdx.x];
    const float mean = static_cast<float>(v / (static_cast<double>(B) * t.t_len * t.d_len));
    term[threadIdx.x] = t.weight * mean;            // Constraint.forward: weight * loss
    scalars[threadIdx.x] = term[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {                            // loss.py:61-108: python-float 0.0 start, list order
    float obj = 0.f, pen = 0.f;
    for (int i = 0; i < plan.n_terms; ++i) { if (plan.terms[i].is_objective) obj += term[i]; }
    for (int i = 0; i < plan.n_terms; ++i) { if (!plan.terms[i].is_objective) pen += term[i]; }
    scalars[plan.n_terms] = obj;
    scalars[plan.n_terms + 1] = pen;
    scalars[plan.n_terms + 2] = obj + pen;
  }
}

// grad[i] = sum over CTAs of partials[c][i]: 32 parameters x 8 CTA-lanes per block, fixed order
__global__ void __launch_bounds__(256) nm_grad_finalize_kernel(const float* __restrict__ partials, int grid, int64_t n, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int px = threadIdx.x % 32, cy = threadIdx.x / 32;
  const int64_t i = int64_t(blockIdx.x) * 32 + px;
  float s = 0.f;
  if (i < n)
    for (int c = cy; c < grid; c += 8) s += partials[size_t(c) * size_t(n) + i];
  part[cy][px] = s;
  __syncthreads();
  if (cy == 0 && i < n) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += part[q][px];
    out[i] = v;
  }
}

}  // namespace

extern "C" {

void nm_set_error(const char* msg) { std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

void nm_register_kernel(NmKernelEntry* e) {
  std::lock_guard<std::mutex> lock(g_mu);
  for (NmKernelEntry* q = g_head; q; q = q->next)
    if (std::strcmp(q->signature, e->signature) == 0 && std::strcmp(q->term_signature, e->term_signature) == 0) return;   // first registration wins
  e->next = g_head;
  g_head = e;
  ++g_count;
}

// ------------------------------------------------------------------------------------------------ fused clip + AdamW
// (declared in include/nm_rollout.h; reference trainer.py:251-252).  n is a few thousand to a few hundred thousand
// floats: the work is two tiny launches, what matters is that the training step never leaves the stream.
static constexpr int kOptBlocks = 64, kOptThreads = 256;
static __global__ void __launch_bounds__(kOptThreads) nm_adamw_sumsq_kernel(const float* __restrict__ g, const unsigned char* __restrict__ mask,
                                                                              int64_t n, double* __restrict__ partial) {
  __shared__ double red[kOptThreads / 32];
  double s = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float v = (mask == nullptr || mask[i]) ? g[i] : 0.f;
    s += double(v) * double(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
static __global__ void __launch_bounds__(kOptThreads) nm_adamw_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                               float* __restrict__ v, const unsigned char* __restrict__ mask, int64_t n,
                                                                               const double* __restrict__ partial, float lr, float b1, float b2, float eps,
                                                                               float wd, float bc1, float bc2_sqrt, float max_norm, float* __restrict__ norm_out) {
  double t = 0.0;
  for (int b = 0; b < kOptBlocks; ++b) t += partial[b];            // same order in every thread
  const float norm = float(sqrt(t));
  float coef = 1.f;
  if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); coef = coef < 1.f ? coef : 1.f; }
  if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  const float step_size = lr / bc1;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    if (mask != nullptr && !mask[i]) continue;
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;               // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;          //        exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

size_t nm_adamw_workspace_bytes(int64_t) { return 256 + sizeof(double) * kOptBlocks; }
int nm_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const unsigned char* mask, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float max_norm,
                  float* norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || workspace == nullptr) {
    nm_set_error("nm_adamw_step: null argument"); return NM_ERR_BAD_PLAN;
  }
  if (n < 0 || step < 1 || workspace_bytes < nm_adamw_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    nm_set_error("nm_adamw_step: bad size, step count or workspace"); return NM_ERR_WORKSPACE;
  }
  if (n == 0) return NM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  nm_adamw_sumsq_kernel<<<kOptBlocks, kOptThreads, 0, st>>>(grads, mask, n, partial);
  const float bc1 = 1.f - float(std::pow(double(beta1), double(step)));
  const float bc2_sqrt = float(std::sqrt(1.0 - std::pow(double(beta2), double(step))));
  nm_adamw_update_kernel<<<kOptBlocks, kOptThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, mask, n, partial, lr, beta1, beta2, eps,
                                                             weight_decay, bc1, bc2_sqrt, max_norm, norm_out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { nm_set_error(cudaGetErrorString(ce)); return NM_ERR_CUDA; }
  return NM_OK;
}

int nm_abi_version(void) { return NM_ABI_VERSION; }
const char* nm_last_error(void) { return g_err; }
int nm_last_launch_count(void) { return g_launches; }

size_t nm_grad_allreduce_buffer_floats(int64_t n) { return size_t(2 * ((n + 63) / 64 * 64) + nm::kArFlagWords); }
int nm_grad_allreduce(void* const* peer_bufs, const void* multicast_ptr, float* out, int64_t n, int rank, int world,
                      uint32_t epoch, int64_t tail_from, float tail_scale, void* stream) {
  if (!peer_bufs || !out || n < 1 || world < 1 || world > nm::kArMaxRanks || rank < 0 || rank >= world || epoch == 0) {
    nm_set_error("nm_grad_allreduce: bad arguments (1 <= world <= 16, epoch >= 1)"); return NM_ERR_BAD_PLAN;
  }
  nm::ArPeers peers{};
  for (int q = 0; q < world; ++q) {
    if (!peer_bufs[q] || (reinterpret_cast<uintptr_t>(peer_bufs[q]) & 15)) { nm_set_error("nm_grad_allreduce: null / misaligned peer buffer"); return NM_ERR_BAD_PLAN; }
    peers.buf[q] = static_cast<float*>(peer_bufs[q]);
  }
  const int64_t npad = (n + 63) / 64 * 64;                   // halves stay 256-byte aligned
  const int blocks = int(std::min<int64_t>(nm::kArBlocks, std::max<int64_t>(1, (npad / 4 + nm::kArThreads - 1) / nm::kArThreads)));
  nm::nm_grad_allreduce_kernel<<<blocks, nm::kArThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      peers, static_cast<const float*>(multicast_ptr), out, n, npad, rank, world, epoch, tail_scale, tail_from);
  g_launches = 1;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { nm_set_error(cudaGetErrorString(ce)); return NM_ERR_CUDA; }
  return NM_OK;
}

int nm_constraint_matrices(const NmPlan* plan, const float* const* exo, const float* x_traj, const float* u_traj, const float* y_traj,
                           float* c_values, float* c_violations, float* c_eq_values, float* c_eq_violations,
                           float* c_ineq_values, float* c_ineq_violations, int64_t widths[3], int64_t B, void* stream) {
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  if (!widths || B < 0) { nm_set_error("bad arguments"); return NM_ERR_BAD_PLAN; }
  CMatArgs a{};
  a.val = c_values; a.viol = c_violations; a.eq_val = c_eq_values; a.eq_viol = c_eq_violations; a.in_val = c_ineq_values; a.in_viol = c_ineq_violations;
  a.B = B;
  for (int k = 0; k < plan->n_terms; ++k) {
    const NmTerm& t = plan->terms[k];
    if (t.is_objective) continue;
    const int64_t w = int64_t(t.t_len) * t.d_len;
    a.term[a.n] = k; a.col0[a.n] = a.W;
    if (t.cmp == NM_CMP_EQ) { a.part0[a.n] = a.Weq; a.Weq += w; } else { a.part0[a.n] = a.Win; a.Win += w; }
    a.W += w; ++a.n;
  }
  a.col0[a.n] = a.W;
  widths[0] = a.W; widths[1] = a.Weq; widths[2] = a.Win;
  g_launches = 0;
  if (B == 0 || a.W == 0) return NM_OK;
  if (!x_traj || (plan->nu > 0 && !u_traj) || (plan->ny > 0 && !y_traj)) { nm_set_error("null trajectory pointer"); return NM_ERR_BAD_PLAN; }
  const nm::VarTable vt = make_vartable(plan, x_traj, u_traj, y_traj, exo);
  const int64_t total = B * a.W;
  const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(148) * 16));
  nm_constraint_matrices_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*plan, vt, a);
  g_launches = 1;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { nm_set_error(cudaGetErrorString(ce)); return NM_ERR_CUDA; }
  return NM_OK;
}

int nm_struct_sizes(int* out, int cap) {
  const int v[7] = {(int)sizeof(NmMlp), (int)sizeof(NmSegment), (int)sizeof(NmExo), (int)sizeof(NmAtom),
                    (int)sizeof(NmSide), (int)sizeof(NmTerm), (int)sizeof(NmPlan)};
  for (int i = 0; i < 7 && i < cap; ++i) out[i] = v[i];
  return 7;
}

int nm_plan_signature(const NmPlan* plan, char* buf, int cap) {
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  const std::string s = make_signature(plan);
  if (!buf || cap <= (int)s.size()) { nm_set_error("signature buffer too small"); return NM_ERR_BAD_PLAN; }
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// The (shape, term) lookup of every entry point.  A plan with BarrierLoss terms runs ONLY on a class compiled with its own
// term structure: the per-class kernels carry the barrier code in their compiled term evaluators and nowhere else (the
// runtime-descriptor branch inside them cost the tensor-core adjoints 2-4 %), so another class of the same shape cannot serve it.
bool plan_has_barrier(const NmPlan* p) {
  for (int k = 0; k < p->n_terms; ++k) if (NM_TERM_BARRIER(p->terms[k].norm) != NM_BARRIER_NONE) return true;
  return false;
}
std::string no_kernel_message(const NmPlan* plan) {
  return "no compiled specialisation for " + make_signature(plan) +
         (plan_has_barrier(plan) ? " with this BarrierLoss plan's term structure (neuromancer_b200.build.ensure_specialisation builds it)" : "");
}
NmKernelEntry* lookup(const NmPlan* plan, int* fused = nullptr) {
  int f = 0;
  NmKernelEntry* e = find_entry(make_signature(plan).c_str(), make_term_signature(plan).c_str(), &f);
  if (fused) *fused = f;
  if (e != nullptr && !f && plan_has_barrier(plan)) {
    nm_set_error("BarrierLoss plan: no specialisation compiled with this term structure (neuromancer_b200.build.ensure_specialisation builds it)");
    return nullptr;
  }
  return e;
}

int nm_has_kernel(const NmPlan* plan) {
  if (validate(plan) != NM_OK) return 0;
  return lookup(plan) != nullptr ? 1 : 0;
}

int nm_num_kernels(void) { return g_count; }

const char* nm_kernel_signature(int i) {
  std::lock_guard<std::mutex> lock(g_mu);
  int k = 0;
  for (NmKernelEntry* e = g_head; e; e = e->next, ++k)
    if (k == i) return e->signature;
  return nullptr;
}

const char* nm_kernel_info(const NmPlan* plan) {
  if (validate(plan) != NM_OK) return nullptr;
  int fused = 0;
  NmKernelEntry* e = lookup(plan, &fused);
  if (!e) return nullptr;
  static thread_local char buf[320];
  std::snprintf(buf, sizeof(buf), "%s fused_terms=%d", e->info, fused);
  return buf;
}

int nm_plan_term_signature(const NmPlan* plan, char* buf, int cap) {
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  const std::string s = make_term_signature(plan);
  if (!buf || cap <= (int)s.size()) { nm_set_error("signature buffer too small"); return NM_ERR_BAD_PLAN; }
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

int nm_load_plugin(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { nm_set_error(dlerror()); return NM_ERR_NO_KERNEL; }
  return NM_OK;                                     // the plugin's static registrar ran during dlopen
}

// stage derivatives kept per integrator step for MLP right-hand sides
static int integrator_stages(int integ) {
  switch (integ) { case NM_INT_EULER: return 1; case NM_INT_RK2: case NM_INT_EULER_TRAP: return 2; case NM_INT_RK4: return 4;
                   case NM_INT_RK4_TRAP: return 5; case NM_INT_RKF: return 6; case NM_INT_LUTHER: return 7;
                   case NM_INT_LEAPFROG: return 2; case NM_INT_YOSHIDA4: return 3; default: return 1; }
}
int nm_rollout_checkpoint_bytes(const NmPlan* plan, int64_t B, size_t* bytes) {
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  if (!bytes || B < 1) { nm_set_error("bad arguments"); return NM_ERR_BAD_PLAN; }
  *bytes = 0;
  // a shape class on the tensor-core path states what its kernels keep (policy mode: the raw policy output o_t)
  // (the same (shape, term) lookup as nm_rollout_forward / nm_rollout_backward: entries of one shape class may be
  // compiled with different tuning, and with it different checkpoint layouts)
  NmKernelEntry* e = lookup(plan);
  if (e != nullptr && e->chk_floats_per_step >= 0) {
    *bytes = size_t(B) * size_t(plan->nsteps) * size_t(e->chk_floats_per_step) * sizeof(float);
    return NM_OK;
  }
  if (plan->dyn_kind == NM_DYN_ODE && NM_RHS_HAS_MLP(plan->rhs_kind)) {
    const int ns = integrator_stages(plan->integrator);
    *bytes = size_t(B) * size_t(plan->nsteps) * size_t(ns - 1) * size_t(plan->nx) * sizeof(float);
  }
  return NM_OK;
}

int nm_rollout_workspace_bytes(const NmPlan* plan, int64_t B, size_t* fwd_bytes, size_t* bwd_bytes) {
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  if (B < 1 || !fwd_bytes || !bwd_bytes) { nm_set_error("bad arguments"); return NM_ERR_BAD_PLAN; }
  NmKernelEntry* e = lookup(plan);
  if (!e) { nm_set_error(no_kernel_message(plan).c_str()); return NM_ERR_NO_KERNEL; }
  rc = e->workspace(plan, B, fwd_bytes, bwd_bytes);
  if (rc != NM_OK) return rc;
  // API-owned tail: block partials of the scoring kernel / precomputed term gradients
  *fwd_bytes = align256(*fwd_bytes) + align256(size_t(score_grid(plan, B)) * NM_MAX_TERMS * sizeof(double));
  *bwd_bytes = align256(*bwd_bytes) + (use_gbuf(plan, B) ? align256(gbuf_floats(plan, B) * sizeof(float)) : 0);
  return NM_OK;
}

int nm_rollout_forward(const NmPlan* plan, const float* params, const float* x0, const float* const* exo,
                       float* x_traj, float* u_traj, float* y_traj, float* scalars, float* checkpoints,
                       void* workspace, size_t workspace_bytes, int64_t B, void* stream) {
  g_launches = 0;
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  if (B < 1 || !x0 || !x_traj || !scalars || !workspace || (plan->nu > 0 && !u_traj) || (plan->ny > 0 && !y_traj) ||
      (plan->n_params > 0 && !params)) { nm_set_error("null pointer argument"); return NM_ERR_BAD_PLAN; }
  int fused = 0;
  NmKernelEntry* e = lookup(plan, &fused);
  if (!e) { nm_set_error(no_kernel_message(plan).c_str()); return NM_ERR_NO_KERNEL; }
  size_t spec_f = 0, spec_b = 0;
  rc = e->workspace(plan, B, &spec_f, &spec_b);
  if (rc != NM_OK) return rc;
  const int sgrid = score_grid(plan, B);
  const size_t need = align256(spec_f) + align256(size_t(sgrid) * NM_MAX_TERMS * sizeof(double));
  if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255)) { nm_set_error("forward workspace too small or misaligned"); return NM_ERR_WORKSPACE; }
  NmFwdCall c{};
  c.params = params; c.x0 = x0; c.exo = exo; c.x_traj = x_traj; c.u_traj = u_traj; c.y_traj = y_traj; c.checkpoints = checkpoints;
  c.workspace = workspace; c.workspace_bytes = align256(spec_f); c.B = B; c.stream = static_cast<cudaStream_t>(stream);
  c.fused_terms = fused;
  rc = e->forward(plan, &c);
  if (rc != NM_OK) return rc;
  const double* partials = c.term_partials;            // written by the rollout kernel when fused
  int pgrid = c.grid;
  const bool api_scores = plan->n_terms > 0 && !(fused && c.scored);
  if (api_scores) {
    double* sp = reinterpret_cast<double*>(static_cast<char*>(workspace) + align256(spec_f));
    const nm::VarTable vt = make_vartable(plan, x_traj, u_traj, y_traj, exo);
    nm_term_score_kernel<<<sgrid, kScoreThreads, 0, c.stream>>>(*plan, vt, B, max_term_steps(plan), sp);
    partials = sp; pgrid = sgrid;
  }
  nm_loss_finalize_kernel<<<1, 256, 0, c.stream>>>(*plan, partials, pgrid, B, scalars);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { nm_set_error(cudaGetErrorString(ce)); return NM_ERR_CUDA; }
  g_launches = (api_scores ? 4 : 3) + c.extra_launches;   // weight prep + rollout [+ term scoring] + loss finalise
  return NM_OK;
}

int nm_rollout_backward(const NmPlan* plan, const float* params, const float* const* exo, const float* x_traj,
                        const float* u_traj, const float* y_traj, const float* checkpoints, const float* grad_scalars,
                        const float* grad_x_ext,
                        const float* grad_u_ext, const float* grad_y_ext, float* grad_params, float* grad_x0,
                        void* workspace, size_t workspace_bytes, int64_t B, int64_t b_total, void* stream) {
  g_launches = 0;
  int rc = validate(plan);
  if (rc != NM_OK) return rc;
  if (B < 1 || b_total < B || !x_traj || !workspace || (plan->nu > 0 && !u_traj) || (plan->ny > 0 && !y_traj) ||
      (plan->n_params > 0 && (!params || !grad_params))) { nm_set_error("null pointer / bad size argument"); return NM_ERR_BAD_PLAN; }
  int fused = 0;
  NmKernelEntry* e = lookup(plan, &fused);
  if (!e) { nm_set_error(no_kernel_message(plan).c_str()); return NM_ERR_NO_KERNEL; }
  size_t spec_f = 0, spec_b = 0;
  rc = e->workspace(plan, B, &spec_f, &spec_b);
  if (rc != NM_OK) return rc;
  const bool pre = !fused && use_gbuf(plan, B);
  const size_t need = align256(spec_b) + (pre ? align256(gbuf_floats(plan, B) * sizeof(float)) : 0);
  if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255)) { nm_set_error("backward workspace too small or misaligned"); return NM_ERR_WORKSPACE; }
  NmBwdCall c{};
  c.params = params; c.exo = exo; c.x_traj = x_traj; c.u_traj = u_traj; c.y_traj = y_traj; c.checkpoints = checkpoints;
  c.grad_scalars = grad_scalars; c.gx_ext = grad_x_ext; c.gu_ext = grad_u_ext; c.gy_ext = grad_y_ext;
  c.grad_x0 = grad_x0; c.workspace = workspace; c.workspace_bytes = align256(spec_b); c.B = B; c.b_total = b_total;
  c.stream = static_cast<cudaStream_t>(stream);
  c.fused_terms = fused;
  int extra = 0;
  if (pre) {
    float* G = reinterpret_cast<float*>(static_cast<char*>(workspace) + align256(spec_b));
    TermGradArgs ta{};
    ta.grad_scalars = grad_scalars; ta.gx_ext = grad_x_ext; ta.gu_ext = grad_u_ext; ta.gy_ext = grad_y_ext;
    ta.G = G; ta.B = B; ta.b_total = b_total;
    const nm::VarTable vt = make_vartable(plan, x_traj, u_traj, y_traj, exo);
    int64_t g = (int64_t(gbuf_floats(plan, B)) + kScoreThreads - 1) / kScoreThreads;
    if (g > 148 * 16) g = 148 * 16;
    nm_term_grad_kernel<<<int(g), kScoreThreads, 0, c.stream>>>(*plan, vt, ta);
    c.G_x = G; c.G_u = nullptr; c.G_y = nullptr;     // row layout, see nm_term_grad_kernel
    extra = 1;
  }
  rc = e->backward(plan, &c);
  if (rc != NM_OK) return rc;
  extra += c.extra_launches;
  g_launches = 2 + extra;
  if (plan->n_params > 0) {
    const int blocks = int((plan->n_params + 31) / 32);
    nm_grad_finalize_kernel<<<blocks, 256, 0, c.stream>>>(c.grad_partials, c.grid, plan->n_params, grad_params);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) { nm_set_error(cudaGetErrorString(ce)); return NM_ERR_CUDA; }
    g_launches = 3 + extra;                         // weight prep [+ term gradients] + adjoint + gradient finalise
  }
  return NM_OK;
}

}  // extern "C"
