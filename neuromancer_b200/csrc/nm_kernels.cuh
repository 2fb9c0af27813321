// nm_kernels.cuh -- the two fused kernels of the DPC hot path, templated on a shape class `S`
// (generated header, see neuromancer_b200/build.py):
//
//   nm_fwd_kernel<S>  persistent, batch-sliced closed-loop rollout over the whole horizon
//                     (policy MLP -> bounds -> integrator(RHS) / SSM -> trajectory stores) followed by
//                     the PenaltyLoss scoring of the tile's trajectories;
//   nm_bwd_kernel<S>  the matching adjoint / BPTT sweep: per step it recomputes the layer activations
//                     from the stored state x_t (the emitted trajectory IS the checkpoint), pulls the
//                     co-state back through integrator, RHS, bounds and MLP, and accumulates weight
//                     gradients in registers across steps and tiles.
//
// Work decomposition per CTA: a tile of TILE trajectories.  Layer contractions are CTA-cooperative
// register-tiled FFMA GEMMs on shared-memory activations [feature][traj]; everything that is
// per-trajectory (state, dynamics, bounds, loss terms, co-state) is thread-per-trajectory register
// math.  Weights are permuted once by nm_prep_kernel and staged into shared memory with one bulk TMA
// copy per CTA.
#pragma once
#include "nm_device.cuh"
#include "nm_dynamics.cuh"
#include "nm_terms.cuh"

namespace nm {

// ------------------------------------------------------------------------------------------------
// Exogenous first-layer hoist (policy MLPs whose inputs are [state / output segments | exogenous segments]):
//   z_1 = b + W[:, :KS] s_t + W[:, KS:] e_t     where e_t does not depend on the rollout,
// so  E[b, t, :] = b + W[:, KS:] e[b, t, :]  is ONE plain GEMM over all (b, t) rows before the horizon loop
// (nm_exo_gemm_fwd_kernel), the loop contracts the first layer over its KS leading inputs only and adds E as a
// per-step pre-activation, the adjoint stores dZ_1[b, t, :], and  dW[:, KS:] = sum_{b,t} dZ_1 (x) e  is one GEMM
// after it (nm_exo_gemm_bwd_kernel).  C4: 193 -> 1 contracted inputs per step.
// HoistedMlp<M, KS> is M with its first layer restricted to the KS leading inputs; KW0 keeps the fan-in of the
// parameter layout (row stride of W_0, offsets of the following layers).
template <class M, int KS>
struct HoistedMlp : M {
  static constexpr int sz(int i) { return i == 0 ? KS : M::sz(i); }
  static constexpr int KW0 = M::sz(0);
};
template <class M, class = void> struct KW0Of { static constexpr int value = M::sz(0); };
template <class M> struct KW0Of<M, std::void_t<decltype(M::KW0)>> { static constexpr int value = M::KW0; };

// Static description of one MLP's staging buffers.
//   GW0 = number of leading input features whose gradient is needed (state / output segments).
template <class M, int TILE, int THREADS, int GW0>
struct MlpCfg {
  static constexpr int NL = M::NL;
  static constexpr int LD = TILE + 4;
  static constexpr int K(int l) { return M::sz(l); }
  static constexpr int N(int l) { return M::sz(l + 1); }
  static constexpr int KW(int l) { return l == 0 ? KW0Of<M>::value : M::sz(l); }   // fan-in in the parameter layout
  static constexpr int gw(int l) { return l == 0 ? GW0 : K(l); }       // width of the data gradient of layer l
  // permuted weight block (floats): per layer  Wf[K][NPf] , bias[NPf] ; then per layer Wb[N][NPb]
  static constexpr int npf(int l) { return g_np(N(l)); }
  static constexpr int npb(int l) { return g_np(gw(l)); }
  static constexpr int wf_off(int l) { int o = 0; for (int i = 0; i < l; ++i) o += (K(i) + 1) * npf(i); return o; }
  static constexpr int bf_off(int l) { return wf_off(l) + K(l) * npf(l); }
  static constexpr int FWD_W = wf_off(NL);
  static constexpr int wb_off(int l) { int o = FWD_W; for (int i = 0; i < l; ++i) o += N(i) * npb(i); return o; }
  static constexpr int ALL_W = wb_off(NL);
  // number of parameters of layer l in the reference's flat order (W [N][K] then b [N])
  static constexpr int p_off(int l) { int o = 0; for (int i = 0; i < l; ++i) o += N(i) * KW(i) + (M::BIAS ? N(i) : 0); return o; }
  static constexpr int N_PARAMS = p_off(NL);
  static constexpr int max_width() { int m = 1; for (int l = 0; l <= NL; ++l) m = cmax(m, M::sz(l)); return m; }
  // forward-only activation buffers: layer inputs alternate between two buffers
  static constexpr int pp_rows(int parity) { int m = 0; for (int l = parity; l <= NL; l += 2) m = cmax(m, M::sz(l)); return cmax(m, 1); }
  // keep-all activation buffers (adjoint): a_l for l < NL, derivative buffers for hidden layers
  static constexpr bool DBUF = act_needs_dbuf<M::ACT>();
  static constexpr int a_off(int l) { int o = 0; for (int i = 0; i < l; ++i) o += K(i) * LD; return o; }
  static constexpr int d_off(int l) { int o = a_off(NL); for (int i = 1; i < l; ++i) o += K(i) * LD; return o; }  // l in [1, NL)
  static constexpr int KEEP_FLOATS = DBUF ? d_off(NL) : a_off(NL);
  static constexpr int delta_rows() { int m = 1; for (int l = 0; l < NL; ++l) m = cmax(m, cmax(N(l), gw(l))); return m; }
  static constexpr int dw_regs() {
    int r = 0;
    for (int l = 0; l < NL; ++l) {
      const int n = N(l), k = K(l);
      const int rj = cmin(4, n), rk = cmin(4, k);
      r += cdiv(cdiv(n, rj) * cdiv(k, rk), THREADS) * rj * rk + 1;
    }
    return r;
  }
};
template <int TILE, int THREADS> struct NoMlpCfg {
  static constexpr int NL = 0, FWD_W = 0, ALL_W = 0, N_PARAMS = 0, KEEP_FLOATS = 0;
  static constexpr int max_width() { return 1; }
  static constexpr int pp_rows(int) { return 0; }
  static constexpr int delta_rows() { return 0; }
  static constexpr int dw_regs() { return 0; }
};

// ------------------------------------------------------------------------------------------------
// Per-spec static configuration
template <class S>
struct KCfg {
  static constexpr int NX = S::NX, NU = S::NU, NY = S::NY, ND = S::ND;
  static constexpr int NUA = NU > 0 ? NU : 1, NYA = NY > 0 ? NY : 1, NDA = ND > 0 ? ND : 1;
  static constexpr bool HAS_POL = S::Pol::NL > 0;
  static constexpr bool HYB = rhs_hybrid<S::RHS>();                       // grey-box: white-box equations around m = MLP(x), 2 -> 1
  static constexpr bool HAS_RHS_MLP = S::RHS == NM_RHS_MLP || HYB;
  static constexpr int THREADS = S::THREADS, TILE = S::TILE, LD = TILE + 4;
  // gradient is needed for the policy inputs up to the end of the last state/output segment
  static constexpr int pol_gw0() {
    int off = 0, gw = 0;
    for (int i = 0; i < S::NSEG; ++i) { off += S::seg_width(i); if (S::seg_kind(i) == NM_SEG_STATE || S::seg_kind(i) == NM_SEG_OUTPUT) gw = off; }
    return gw;
  }
  // exogenous first-layer hoist: every exogenous segment follows the last state / output segment, no preview windows
  static constexpr bool hoist_ok() {
    if (!HAS_POL || S::Pol::NL < 2 || S::Pol::sz(1) % 4 != 0 || S::Pol::sz(1) > 128) return false;
    int ks = 0, seen_exo = 0;
    for (int i = 0; i < S::NSEG; ++i) {
      if (S::seg_kind(i) == NM_SEG_PREVIEW) return false;
      if (S::seg_kind(i) == NM_SEG_EXO) { seen_exo = 1; if (S::seg_width(i) % 4 != 0) return false; }
      else { if (seen_exo) return false; ks += S::seg_width(i); }
    }
    return seen_exo && ks >= 1 && ks == pol_gw0();
  }
  static constexpr bool HOIST = S::HOIST && hoist_ok();
  static constexpr int KS = pol_gw0();                                     // contracted fan-in of the first layer when hoisted
  static constexpr int KE = HAS_POL ? S::Pol::sz(0) - KS : 0;              // hoisted (exogenous) fan-in
  static constexpr int N0 = HAS_POL ? S::Pol::sz(1) : 1;
  using PolM = std::conditional_t<HOIST, HoistedMlp<typename S::Pol, pol_gw0()>, typename S::Pol>;
  using PolCfg = std::conditional_t<HAS_POL, MlpCfg<PolM, TILE, THREADS, pol_gw0()>, NoMlpCfg<TILE, THREADS>>;
  // prepared exogenous block of the hoist GEMMs: rows [KE exogenous inputs | bias][NPf] in the forward image layout
  static constexpr int EXO_NP = g_np(N0);
  static constexpr int EXO_W_FLOATS = HOIST ? round_up((KE + 1) * EXO_NP, 4) : 0;
  // second-order integrators (LeapFrog, Yoshida4): the right-hand side maps RX = NX / 2 positions [+ NU] to RX accelerations
  static constexpr bool SO = HAS_RHS_MLP && second_order<S::INTEG>();
  static constexpr int RX = SO ? NX / 2 : NX;
  static constexpr int RO = HYB ? 1 : RX;                                  // output width of the right-hand-side MLP
  static_assert(!HYB || (NX == 2 && NU == 0 && !SO), "hybrid right-hand sides are autonomous two-state systems");
  static_assert(!SO || NX % 2 == 0, "second-order integrators need an even state width");
  using RhsCfg = std::conditional_t<HAS_RHS_MLP, MlpCfg<typename S::RhsM, TILE, THREADS, RX + NU>, NoMlpCfg<TILE, THREADS>>;
  static constexpr int W_FLOATS_FWD = round_up(PolCfg::FWD_W, 4) + round_up(RhsCfg::FWD_W, 4);
  static constexpr int RHS_WF_OFF = round_up(PolCfg::FWD_W, 4);                     // forward-kernel smem offset
  static constexpr int W_FLOATS_ALL = round_up(PolCfg::ALL_W, 4) + round_up(RhsCfg::ALL_W, 4);
  static constexpr int POL_W_OFF = 0, RHS_W_OFF = round_up(PolCfg::ALL_W, 4);       // in the prepared global block
  static constexpr int EXO_W_OFF = W_FLOATS_ALL;                                     // (hoist) after the staged part
  // forward kernel shared memory (floats)
  static constexpr int FWD_PP0 = cmax(cmax(PolCfg::pp_rows(0), RhsCfg::pp_rows(0)), 4) * LD;   // >= 256 floats: reused by the loss reduction
  static constexpr int FWD_PP1 = cmax(PolCfg::pp_rows(1), RhsCfg::pp_rows(1)) * LD;
  // weights are staged in shared memory when they fit, else streamed from the prepared global block (L1/L2)
  static constexpr bool W_SMEM = S::W_SMEM;
  static constexpr int FWD_SMEM_FLOATS = (W_SMEM ? W_FLOATS_FWD : 0) + FWD_PP0 + FWD_PP1 + 64 + 128;
  // adjoint kernel shared memory (floats)
  static constexpr int DELTA_ROWS = cmax(PolCfg::delta_rows(), RhsCfg::delta_rows());
  static constexpr int BWD_SMEM_FLOATS = (W_SMEM ? W_FLOATS_ALL : 0) + PolCfg::KEEP_FLOATS + RhsCfg::KEEP_FLOATS + 2 * DELTA_ROWS * LD + 64 + 128;
  static constexpr size_t FWD_SMEM_BYTES = size_t(FWD_SMEM_FLOATS) * 4, BWD_SMEM_BYTES = size_t(BWD_SMEM_FLOATS) * 4;
  static constexpr bool DW_SPILL = (PolCfg::dw_regs() + RhsCfg::dw_regs()) > 112;   // accumulators do not fit the RF
  static constexpr int N_PARAMS = PolCfg::N_PARAMS + RhsCfg::N_PARAMS;
  static constexpr bool FITS_FWD = FWD_SMEM_BYTES <= 227 * 1024, FITS_BWD = BWD_SMEM_BYTES <= 227 * 1024;
};

struct FwdArgs {
  const float* params; const float* wprep; const float* x0;
  const float* exo[NM_MAX_EXO];
  float* x_traj; float* u_traj; float* y_traj;
  float* chk;                       // stage-derivative checkpoints [B][T][(NS-1)*NX] (MLP right-hand side), or null
  double* term_partials;            // [grid][NM_MAX_TERMS], written when fused_terms
  int fused_terms;                  // the plan's terms match the compiled term structure of S
  int64_t B; int num_tiles;
  const float* pre;                 // (hoist) E[B][T][N0] = bias + exogenous part of the first layer
};
struct BwdArgs {
  const float* params; const float* wprep;
  const float* exo[NM_MAX_EXO];
  const float* x_traj; const float* u_traj; const float* y_traj;
  const float* chk;                 // stage-derivative checkpoints written by the forward kernel, or null
  const float* grad_scalars; const float* gx_ext; const float* gu_ext; const float* gy_ext;
  // loss-term gradients precomputed by nm_term_grad_kernel (already include the *_ext gradients);
  // when null they are evaluated inline in the sweep
  const float* G_x; const float* G_u; const float* G_y;
  float* grad_partials;             // [grid][n_params]
  float* grad_x0;
  int fused_terms;                  // evaluate term gradients with the compiled term structure of S
  int64_t B; int64_t b_total; int num_tiles;
  const float* pre;                 // (hoist) E[B][T][N0]
  float* dz1;                       // (hoist) dZ_1[B][T][N0], may alias `pre` (each entry is read before it is written)
};

// ------------------------------------------------------------------------------------------------
// weight preparation: permute nn.Linear weights into the staged layouts
template <class MC>
__device__ void prep_mlp(const float* __restrict__ params, float* __restrict__ dst, bool has_bias, int gtid, int gthreads) {
  static_for<0, MC::NL>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int K = MC::K(l), KW = MC::KW(l), N = MC::N(l), NPF = MC::npf(l), GW = MC::gw(l), NPB = MC::npb(l);
    const float* __restrict__ Wg = params + MC::p_off(l);
    const float* __restrict__ bg = Wg + N * KW;
    float* __restrict__ wf = dst + MC::wf_off(l);
    for (int idx = gtid; idx < (K + 1) * NPF; idx += gthreads) {
      const int c = idx / NPF, col = idx % NPF;
      const int ng = col / g_tn(N), j = col % g_tn(N);
      const int n = ng + g_ng(N) * j;
      float v = 0.f;
      if (ng < g_ng(N) && n < N) v = c < K ? Wg[n * KW + c] : (has_bias ? bg[n] : 0.f);
      wf[idx] = v;
    }
    if constexpr (GW > 0) {
      float* __restrict__ wb = dst + MC::wb_off(l);
      for (int idx = gtid; idx < N * NPB; idx += gthreads) {
        const int n = idx / NPB, col = idx % NPB;
        const int kg = col / g_tn(GW), j = col % g_tn(GW);
        const int k = kg + g_ng(GW) * j;
        wb[idx] = (kg < g_ng(GW) && k < GW) ? Wg[n * KW + k] : 0.f;
      }
    }
  });
}

template <class S>
__global__ void nm_prep_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, float* __restrict__ wprep) {
  using KC = KCfg<S>;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  if constexpr (KC::HAS_POL)
    prep_mlp<typename KC::PolCfg>(params + plan.policy.param_offset, wprep + KC::POL_W_OFF, plan.policy.has_bias, gtid, gthreads);
  if constexpr (KC::HAS_RHS_MLP)
    prep_mlp<typename KC::RhsCfg>(params + plan.rhs_mlp.param_offset, wprep + KC::RHS_W_OFF, plan.rhs_mlp.has_bias, gtid, gthreads);
  if constexpr (KC::HOIST) {
    // exogenous rows of the first layer (+ its bias as the last row) in the forward image layout of the hoist GEMMs
    constexpr int N = KC::N0, NP = KC::EXO_NP, KW = S::Pol::sz(0), KS = KC::KS, KE = KC::KE;
    const float* __restrict__ Wg = params + plan.policy.param_offset;
    const float* __restrict__ bg = Wg + N * KW;
    float* __restrict__ we = wprep + KC::EXO_W_OFF;
    for (int idx = gtid; idx < (KE + 1) * NP; idx += gthreads) {
      const int c = idx / NP, col = idx % NP;
      const int ng = col / g_tn(N), j = col % g_tn(N);
      const int n = ng + g_ng(N) * j;
      float v = 0.f;
      if (ng < g_ng(N) && n < N) v = c < KE ? Wg[n * KW + KS + c] : (plan.policy.has_bias ? bg[n] : 0.f);
      we[idx] = v;
    }
  }
}

// hoisted first-layer pre-activation of the tile's trajectories at step t:  E[(b * T + t) * N0 + n]   (b clamped to B-1)
struct PreAct {
  const float* base; int64_t b0, B; int T, t;
  __device__ __forceinline__ const float* row(int traj, int n0) const {
    const int64_t b = b0 + traj < B ? b0 + traj : B - 1;
    return base + (b * T + t) * int64_t(n0);
  }
};

// ------------------------------------------------------------------------------------------------
// MLP forward on the tile.  Inputs a_0 must be in place and visible (caller synchronised).
//   HOIST        : layer 0 contracts over its leading inputs only; `pre` supplies bias + exogenous part (see HoistedMlp)
//   KEEP = false : layer inputs ping-pong between pp0/pp1; raw output of the last layer is left in
//                  the buffer of parity NL.
//   KEEP = true  : a_l kept at keep + a_off(l) (and act' at keep + d_off(l)); raw output -> out_raw.
// Ends with a __syncthreads().
template <class MC, class M, int TILE, int THREADS, bool KEEP, bool HOIST = false>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ W, float* __restrict__ pp0, float* __restrict__ pp1,
                                            float* __restrict__ keep, float* __restrict__ out_raw, float act_prm, int tid,
                                            const PreAct pre = PreAct{}) {
  constexpr int LD = TILE + 4;
  static_for<0, MC::NL>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int K = MC::K(l), N = MC::N(l);
    constexpr bool LAST = (l == MC::NL - 1);
    using G = GemmCfg<N, TILE, THREADS>;
    const float* in = KEEP ? keep + MC::a_off(l) : ((l & 1) ? pp1 : pp0);
    float* out = KEEP ? (LAST ? out_raw : keep + MC::a_off(l + 1)) : ((l & 1) ? pp0 : pp1);
    const float* __restrict__ bias = W + MC::bf_off(l);
    gemm_phase<K, N, TILE, THREADS, LD>(W + MC::wf_off(l), in, tid, [&](int ng, int traj0, float (&acc)[G::TMB][G::TN]) {
#pragma unroll
      for (int j = 0; j < G::TN; ++j) {
        const int n = ng + G::NG * j;
        if (n < N) {
          float v[G::TMB];
          if constexpr (HOIST && l == 0) {
#pragma unroll
            for (int i = 0; i < G::TMB; ++i) v[i] = acc[i][j] + __ldg(pre.row(traj0 + i, N) + n);
          } else {
            const float bj = bias[ng * G::TN + j];
#pragma unroll
            for (int i = 0; i < G::TMB; ++i) v[i] = acc[i][j] + bj;
          }
          if constexpr (!LAST) {
            if constexpr (KEEP && MC::DBUF) {
              float dv[G::TMB];
#pragma unroll
              for (int i = 0; i < G::TMB; ++i) act_pair_z<M::ACT>(v[i], act_prm, v[i], dv[i]);
              st_vec<G::TMB>(keep + MC::d_off(l + 1) + n * LD + traj0, dv);
            } else {
#pragma unroll
              for (int i = 0; i < G::TMB; ++i) v[i] = act_fwd<M::ACT>(v[i], act_prm);
            }
          }
          st_vec<G::TMB>(out + n * LD + traj0, v);
        }
      }
    });
    __syncthreads();
  });
}

// persistent weight-gradient accumulators of one MLP (registers)
template <class MC, int TILE, int THREADS, int L = 0, bool END = (L >= MC::NL)>
struct DwStore;
template <class MC, int TILE, int THREADS, int L>
struct DwStore<MC, TILE, THREADS, L, true> {};
template <class MC, int TILE, int THREADS, int L>
struct DwStore<MC, TILE, THREADS, L, false> {
  using D = DwCfg<MC::N(L), MC::K(L), TILE, THREADS>;
  float acc[D::SLOTS][D::RJ][D::RK];
  float db;
  DwStore<MC, TILE, THREADS, L + 1> next;
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int s = 0; s < D::SLOTS; ++s)
#pragma unroll
      for (int i = 0; i < D::RJ; ++i)
#pragma unroll
        for (int j = 0; j < D::RK; ++j) acc[s][i][j] = 0.f;
    db = 0.f;
    if constexpr (L + 1 < MC::NL) next.zero();
  }
  template <int Q> __device__ __forceinline__ auto& at() {
    if constexpr (Q == L) return *this; else return next.template at<Q>();
  }
};
struct DwNone { __device__ __forceinline__ void zero() {} };

// add the thread's tile of layer L to  dst[n*K + k].  With a trajectory split (SPLIT > 1) several thread groups
// hold partial sums of the same entries: they are added one group after the other, a barrier apart, so that the
// result does not depend on scheduling (bitwise reproducible gradients).  Called by every thread of the CTA.
template <class D, int N, int K, int KW = K>
__device__ __forceinline__ void dw_flush_tile(const float (&acc)[D::RJ][D::RK], float* __restrict__ dst, int unit) {
  const int tile = unit % D::TILES, grp = unit / D::TILES;
  const int kt = tile % D::NKT, jt = tile / D::NKT;
#pragma unroll 1
  for (int g = 0; g < D::SPLIT; ++g) {
    if (grp == g) {
#pragma unroll
      for (int i = 0; i < D::RJ; ++i)
#pragma unroll
        for (int j = 0; j < D::RK; ++j) {
          const int n = jt + D::NJ * i, k = kt + D::NKT * j;
          if (n < N && k < K) dst[n * KW + k] += acc[i][j];
        }
    }
    if constexpr (D::SPLIT > 1) __syncthreads();
  }
}

// MLP backward on the tile.  delta_{NL} (cotangent of the raw output) must be in dbuf0 and visible.
// Leaves the gradient w.r.t. the first gw(0) inputs in the delta buffer of parity NL (dbuf0 if NL even
// else dbuf1).  `gscratch` != nullptr selects spill mode: per-step tiles are added to the CTA's private
// global gradient block instead of persistent registers.  Ends with a __syncthreads().
//   HOIST: dZ_1 of the tile's trajectories is also written to `dz1` (PreAct addressing) for nm_exo_gemm_bwd_kernel.
template <class MC, class M, int TILE, int THREADS, bool SPILL, bool HOIST = false, class Store>
__device__ __forceinline__ void mlp_backward(const float* __restrict__ W, const float* __restrict__ keep,
                                             float* __restrict__ dbuf0, float* __restrict__ dbuf1, Store& store,
                                             float* __restrict__ gscratch, bool has_bias, bool trainable,
                                             float act_prm, int tid, const PreAct dz1 = PreAct{}) {
  constexpr int LD = TILE + 4;
  static_for_rev<0, MC::NL>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int K = MC::K(l), N = MC::N(l), GW = MC::gw(l);
    constexpr int step = MC::NL - 1 - l;                       // 0 for the last layer
    const float* dl = (step & 1) ? dbuf1 : dbuf0;              // delta_{l+1}  [N][LD]
    float* dout = (step & 1) ? dbuf0 : dbuf1;                  // delta_l      [GW][LD]
    const float* al = keep + MC::a_off(l);
    if constexpr (HOIST && l == 0) {
      // consecutive threads -> consecutive outputs of one trajectory: coalesced rows of N floats
      float* __restrict__ gdst = const_cast<float*>(dz1.base);
      for (int idx = tid; idx < N * TILE; idx += THREADS) {
        const int traj = idx / N, n = idx - traj * N;
        if (dz1.b0 + traj < dz1.B) gdst[((dz1.b0 + traj) * dz1.T + dz1.t) * int64_t(N) + n] = dl[n * LD + traj];
      }
    }
    if constexpr (GW > 0) {
      using G = GemmCfg<GW, TILE, THREADS>;
      gemm_phase<N, GW, TILE, THREADS, LD>(W + MC::wb_off(l), dl, tid, [&](int kg, int traj0, float (&acc)[G::TMB][G::TN]) {
#pragma unroll
        for (int j = 0; j < G::TN; ++j) {
          const int k = kg + G::NG * j;
          if (k < GW) {
            float v[G::TMB];
#pragma unroll
            for (int i = 0; i < G::TMB; ++i) v[i] = acc[i][j];
            if constexpr (l > 0) {                               // through the activation of hidden layer l
              float s[G::TMB];
              if constexpr (MC::DBUF) ld_vec<G::TMB>(s, keep + MC::d_off(l) + k * LD + traj0);
              else {
                ld_vec<G::TMB>(s, al + k * LD + traj0);
#pragma unroll
                for (int i = 0; i < G::TMB; ++i) s[i] = act_deriv_a<M::ACT>(s[i], act_prm);
              }
#pragma unroll
              for (int i = 0; i < G::TMB; ++i) v[i] *= s[i];
            }
            st_vec<G::TMB>(dout + k * LD + traj0, v);
          }
        }
      });
    }
    if (trainable) {
      using D = DwCfg<N, K, TILE, THREADS>;
      if constexpr (!SPILL) {
        auto& st = store.template at<l>();
#pragma unroll
        for (int s = 0; s < D::SLOTS; ++s) dw_tile<N, K, TILE, THREADS, LD>(st.acc[s], dl, al, s * THREADS + tid);
        if (has_bias && tid < N) {
          const float* row = dl + tid * LD;
          float s = 0.f;
#pragma unroll 4
          for (int t4 = 0; t4 < TILE; t4 += 4) {
            const float4 v = *reinterpret_cast<const float4*>(row + t4);
            s += (v.x + v.y) + (v.z + v.w);
          }
          st.db += s;
        }
      } else {
        float* gw_dst = gscratch + MC::p_off(l);
#pragma unroll 1
        for (int s = 0; s < D::SLOTS; ++s) {
          float t[D::RJ][D::RK];
#pragma unroll
          for (int i = 0; i < D::RJ; ++i)
#pragma unroll
            for (int j = 0; j < D::RK; ++j) t[i][j] = 0.f;
          dw_tile<N, K, TILE, THREADS, LD>(t, dl, al, s * THREADS + tid);
          dw_flush_tile<D, N, K, MC::KW(l)>(t, gw_dst, s * THREADS + tid);
        }
        if (has_bias && tid < N) {
          const float* row = dl + tid * LD;
          float s = 0.f;
#pragma unroll 4
          for (int t4 = 0; t4 < TILE; t4 += 4) {
            const float4 v = *reinterpret_cast<const float4*>(row + t4);
            s += (v.x + v.y) + (v.z + v.w);
          }
          gw_dst[N * MC::KW(l) + tid] += s;
        }
      }
    }
    __syncthreads();
  });
}

// write the persistent accumulators of one MLP into the CTA's gradient block (zeroed beforehand)
template <class MC, int TILE, int THREADS, class Store>
__device__ __forceinline__ void dw_writeout(Store& store, float* __restrict__ gdst, bool has_bias, int tid) {
  static_for<0, MC::NL>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int K = MC::K(l), N = MC::N(l);
    using D = DwCfg<N, K, TILE, THREADS>;
    auto& st = store.template at<l>();
    float* dst = gdst + MC::p_off(l);
#pragma unroll
    for (int s = 0; s < D::SLOTS; ++s) dw_flush_tile<D, N, K, MC::KW(l)>(st.acc[s], dst, s * THREADS + tid);
    if (has_bias && tid < N) dst[N * MC::KW(l) + tid] += st.db;
  });
}

// ------------------------------------------------------------------------------------------------
// MLP_bounds output maps (reference functions.py:9-34)
template <int BOUNDS>
__device__ __forceinline__ float bounds_fwd(float o, float lo, float hi) {
  if constexpr (BOUNDS == NM_BOUNDS_SIGMOID_SCALE) return (hi - lo) * (1.f / (1.f + expf(-o))) + lo;
  else if constexpr (BOUNDS == NM_BOUNDS_RELU_CLAMP) {
    float r = -o + lo; r = r > 0.f ? r : 0.f;
    const float x1 = o + r;
    float q = x1 - hi; q = q > 0.f ? q : 0.f;
    return x1 - q;
  } else return o;
}
template <int BOUNDS>
__device__ __forceinline__ float bounds_deriv(float o, float lo, float hi) {
  if constexpr (BOUNDS == NM_BOUNDS_SIGMOID_SCALE) {
    const float s = 1.f / (1.f + expf(-o));
    return (hi - lo) * (s * (1.f - s));
  } else if constexpr (BOUNDS == NM_BOUNDS_RELU_CLAMP) {
    const float r = -o + lo;
    const float d1 = 1.f - (r > 0.f ? 1.f : 0.f);
    const float x1 = o + (r > 0.f ? r : 0.f);
    const float d2 = 1.f - ((x1 - hi) > 0.f ? 1.f : 0.f);
    return d1 * d2;
  } else return 1.f;
}

// gather the policy input column of trajectory `tid` into a0[feature][tid]
template <class S>
__device__ __forceinline__ void build_policy_input(const NmPlan& p, const float* const* exo, float* __restrict__ a0, int tid,
                                                   int64_t b, int t, const float (&x)[S::NX],
                                                   const float (&y)[S::NY > 0 ? S::NY : 1]) {
  constexpr int LD = S::TILE + 4;
  int off = 0;
  static_for<0, S::NSEG>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    constexpr int W = S::seg_width(s);
    if constexpr (S::seg_kind(s) == NM_SEG_STATE) {
#pragma unroll
      for (int i = 0; i < W; ++i) a0[(off + i) * LD + tid] = x[i];
    } else if constexpr (S::seg_kind(s) == NM_SEG_OUTPUT) {
#pragma unroll
      for (int i = 0; i < W; ++i) a0[(off + i) * LD + tid] = y[i];
    } else if constexpr (S::seg_kind(s) == NM_SEG_PREVIEW) {
      // SystemPreview window  exo[e][:, t : t+1+L, :]  flattened time-major; entries past the signal are
      // right-padding of the FULL signal by torch.nn.functional.pad(mode) (reference system.py:341-345)
      constexpr int e = S::seg_index(s);
      constexpr int D = S::seg_sub(s);
      const int Tk = p.exo[e].steps;
      const float* __restrict__ base = exo[e] + b * Tk * D;
      if (t + W / D <= Tk) {
        // whole window inside the signal: W contiguous floats starting at entry t (time-major flatten)
        const float* __restrict__ src = base + t * D;
#pragma unroll 8
        for (int i = 0; i < W; ++i) a0[(off + i) * LD + tid] = __ldg(src + i);
      } else {
#pragma unroll 2
        for (int j = 0; j < W / D; ++j) {
          int ts = t + j;
          bool fill = false;
          if (ts >= Tk) {
            if (p.pad_mode == NM_PAD_REPLICATE) ts = Tk - 1;
            else if (p.pad_mode == NM_PAD_CIRCULAR) ts = (ts - Tk) % Tk;
            else if (p.pad_mode == NM_PAD_REFLECT) ts = 2 * (Tk - 1) - ts;
            else fill = true;
          }
#pragma unroll
          for (int dd = 0; dd < D; ++dd) a0[(off + j * D + dd) * LD + tid] = fill ? p.pad_value : __ldg(base + ts * D + dd);
        }
      }
    } else if constexpr (!S::HOIST) {                         // (hoisted: the exogenous segments enter through E)
      constexpr int e = S::seg_index(s);
      const float* __restrict__ src = exo[e] + (b * p.exo[e].steps + t) * W;
      if constexpr (W % 4 == 0) {
#pragma unroll 4
        for (int i = 0; i < W; i += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
          a0[(off + i + 0) * LD + tid] = v.x; a0[(off + i + 1) * LD + tid] = v.y;
          a0[(off + i + 2) * LD + tid] = v.z; a0[(off + i + 3) * LD + tid] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) a0[(off + i) * LD + tid] = __ldg(src + i);
      }
    }
    off += W;
  });
}

// stage the prepared weight block into shared memory with one bulk TMA copy (UBLKCP)
template <int FLOATS_A, int FLOATS_B>
__device__ __forceinline__ void stage_weights(float* __restrict__ dst_a, const float* __restrict__ src_a,
                                              float* __restrict__ dst_b, const float* __restrict__ src_b,
                                              uint64_t* bar, int tid) {
  if constexpr (FLOATS_A + FLOATS_B > 0) {
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, (FLOATS_A + FLOATS_B) * 4);
      if constexpr (FLOATS_A > 0) tma_bulk_g2s(dst_a, src_a, FLOATS_A * 4, bar);
      if constexpr (FLOATS_B > 0) tma_bulk_g2s(dst_b, src_b, FLOATS_B * 4, bar);
    }
    mbar_wait(bar, 0);
  }
}

// ================================================================================================
// forward kernel
template <class S>
__global__ void __launch_bounds__(S::THREADS, S::MINB_FWD) nm_fwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ FwdArgs args) {
  using KC = KCfg<S>;
  using PolCfg = typename KC::PolCfg;
  using RhsCfg = typename KC::RhsCfg;
  constexpr int NX = KC::NX, NU = KC::NU, NY = KC::NY, ND = KC::ND, NUA = KC::NUA, NYA = KC::NYA, NDA = KC::NDA;
  constexpr int TILE = KC::TILE, THREADS = KC::THREADS, LD = KC::LD;
  extern __shared__ __align__(128) float smem[];
  float* sm_w = smem;
  float* pp0 = sm_w + (KC::W_SMEM ? KC::W_FLOATS_FWD : 0);
  float* pp1 = pp0 + KC::FWD_PP0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(pp1 + KC::FWD_PP1);
  VarTable& vt = *reinterpret_cast<VarTable*>(pp1 + KC::FWD_PP1 + 16);
  const int tid = threadIdx.x;
  const NmPlan& p = plan;
  const VarTable vtc = make_vartable_regs(p, args, NX, NU, NY);      // compiled term code (see nm_terms.cuh)
  constexpr int NT = S::NTERMS > 0 ? S::NTERMS : 1;
  const bool fused = S::NTERMS > 0 && args.fused_terms;
  if (tid == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = p.nsteps + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = p.nsteps; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = p.nsteps; vt.D[NM_VAR_Y] = NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
  }
  float tsum[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) tsum[q] = 0.f;
  __syncthreads();

  // only the forward halves of the prepared weight blocks are needed here
  const float* Wpol;
  const float* Wrhs;
  if constexpr (KC::W_SMEM) {
    stage_weights<round_up(PolCfg::FWD_W, 4), round_up(RhsCfg::FWD_W, 4)>(
        sm_w, args.wprep + KC::POL_W_OFF, sm_w + KC::RHS_WF_OFF, args.wprep + KC::RHS_W_OFF, bar, tid);
    Wpol = sm_w;
    Wrhs = sm_w + KC::RHS_WF_OFF;
  } else {
    Wpol = args.wprep + KC::POL_W_OFF;
    Wrhs = args.wprep + KC::RHS_W_OFF;
  }

  const int T = p.nsteps;
  for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
    const int64_t b_raw = int64_t(tile) * TILE + tid;
    const bool owner = tid < TILE;
    const bool valid = owner && b_raw < args.B;
    const int64_t b = b_raw < args.B ? b_raw : args.B - 1;

    float x[NX];
    if (owner) {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = __ldg(args.x0 + b * NX + i);
      if (valid) {
#pragma unroll
        for (int i = 0; i < NX; ++i) args.x_traj[(b * (T + 1)) * NX + i] = x[i];
      }
    }
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float u[NUA], y[NYA], d[NDA];
#pragma unroll
      for (int j = 0; j < NUA; ++j) u[j] = 0.f;
#pragma unroll
      for (int j = 0; j < NYA; ++j) y[j] = 0.f;
#pragma unroll
      for (int j = 0; j < NDA; ++j) d[j] = 0.f;
      if (owner) {
        if constexpr (S::HAS_OUTPUT) {                      // y_t = C x_t  (grid_response cell 7)
#pragma unroll
          for (int j = 0; j < NY; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NX; ++i) s = fmaf(x[i], p.C[j * NX + i], s);
            y[j] = s;
            if (valid) args.y_traj[(b * T + t) * NY + j] = s;
          }
        }
        if constexpr (KC::HAS_POL) build_policy_input<S>(p, args.exo, pp0, tid, b, t, x, y);
      }
      if constexpr (KC::HAS_POL) {
        __syncthreads();
        mlp_forward<PolCfg, typename S::Pol, TILE, THREADS, false, KC::HOIST>(Wpol, pp0, pp1, nullptr, nullptr, p.policy.act_param, tid,
                                                                              PreAct{args.pre, int64_t(tile) * TILE, args.B, T, t});
        if (owner) {
          const float* outp = (PolCfg::NL & 1) ? pp1 : pp0;
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            const float o = outp[j * LD + tid];
            u[j] = bounds_fwd<S::Pol::BOUNDS>(o, p.policy.bmin[j], p.policy.bmax[j]);
            if (valid) args.u_traj[(b * T + t) * NU + j] = u[j];
          }
        }
      }
      if constexpr (ND > 0) {
        if (owner) {
          const float* src = args.exo[S::DIST_EXO] + (b * p.exo[S::DIST_EXO].steps + t) * ND;
#pragma unroll
          for (int j = 0; j < ND; ++j) d[j] = __ldg(src + j);
        }
      }
      if constexpr (!KC::HAS_POL && S::ACT_EXO >= 0) {       // open loop: the action comes from the data
        if (owner) {
          const float* src = args.exo[S::ACT_EXO] + (b * p.exo[S::ACT_EXO].steps + t) * NU;
#pragma unroll
          for (int j = 0; j < NU; ++j) { u[j] = __ldg(src + j); if (valid) args.u_traj[(b * T + t) * NU + j] = u[j]; }
        }
      }
      float xn[NX];
      if constexpr (!KC::HAS_RHS_MLP) {
        if (owner) Dyn<S>::step(p, x, u, d, xn);
      } else {
        // neural right-hand side: every stage is a CTA-cooperative MLP evaluation
        constexpr int NS = Tableau<S::INTEG>::NS;
        float k[NM_MAX_STAGES][NX];
        constexpr int RX = KC::RX;                           // width of the right-hand side's state input / output
        static_for<0, NS>([&](auto sc) {
          constexpr int St = decltype(sc)::value;
          float xs[RX];
          if (owner) {
            if constexpr (KC::SO) so_position<S::INTEG, St, NX>(p.h, x, k, xs);
            else stage_input<S::INTEG, St, NX>(p.h, x, k, xs);
#pragma unroll
            for (int i = 0; i < RX; ++i) pp0[i * LD + tid] = xs[i];
#pragma unroll
            for (int j = 0; j < NU; ++j) pp0[(RX + j) * LD + tid] = u[j];
          }
          __syncthreads();
          mlp_forward<RhsCfg, typename S::RhsM, TILE, THREADS, false>(Wrhs, pp0, pp1, nullptr, nullptr, p.rhs_mlp.act_param, tid);
          if (owner) {
            const float* outp = (RhsCfg::NL & 1) ? pp1 : pp0;
            if constexpr (KC::HYB) hybrid_eval<S>(p, xs, outp[tid], k[St]);
            else {
#pragma unroll
              for (int i = 0; i < RX; ++i) k[St][i] = outp[i * LD + tid];
            }
            if constexpr (St + 1 < NS) {
              if (valid && args.chk != nullptr) {
                float* cp = args.chk + ((b * T + t) * (NS - 1) + St) * NX;
#pragma unroll
                for (int i = 0; i < RX; ++i) cp[i] = k[St][i];
              }
            }
          }
          // the next writer of pp0/pp1 is this same thread's own column (stage input) or, behind a
          // barrier, the next layer GEMM: no extra barrier needed here
        });
        if (owner) {
          if constexpr (KC::SO) so_combine<S::INTEG, NX>(p.h, x, k, xn);
          else combine_stages<S::INTEG, NX>(p.h, x, k, xn);
        }
      }
      if (owner) {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = xn[i];
        if (valid) {
#pragma unroll
          for (int i = 0; i < NX; ++i) args.x_traj[(b * (T + 1) + t + 1) * NX + i] = x[i];
        }
      }
    }
    // PenaltyLoss scoring of this thread's trajectory with the compiled term structure (its entries
    // were stored by this same thread; plain loads are coherent)
    if constexpr (S::NTERMS > 0) {
      if (fused && valid) score_traj_s<S>(p, vtc, b, tsum);
    }
  }

  // block reduction of the term sums (double) -> this CTA's partials; with runtime-descriptor terms the
  // scoring runs in nm_term_score_kernel instead
  if constexpr (S::NTERMS > 0) {
    if (fused) {
      __syncthreads();
      double* scratch = reinterpret_cast<double*>(pp0);       // activations are dead now (>= 512 floats)
      static_for<0, S::NTERMS>([&](auto kc) {
        constexpr int K = decltype(kc)::value;
        double v = static_cast<double>(tsum[K]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) scratch[K * (THREADS / 32) + (tid >> 5)] = v;
      });
      __syncthreads();
      if (tid < S::NTERMS) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += scratch[tid * (THREADS / 32) + w];
        args.term_partials[size_t(blockIdx.x) * NM_MAX_TERMS + tid] = v;
      }
    }
  }
  (void)Wrhs; (void)Wpol;
}

// ================================================================================================
// adjoint kernel
template <class S>
__global__ void __launch_bounds__(S::THREADS, S::MINB_BWD) nm_bwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ BwdArgs args) {
  using KC = KCfg<S>;
  using PolCfg = typename KC::PolCfg;
  using RhsCfg = typename KC::RhsCfg;
  constexpr int NX = KC::NX, NU = KC::NU, NY = KC::NY, ND = KC::ND, NUA = KC::NUA, NYA = KC::NYA;
  constexpr int TILE = KC::TILE, THREADS = KC::THREADS, LD = KC::LD;
  constexpr bool SPILL = KC::DW_SPILL;
  extern __shared__ __align__(128) float smem[];
  float* sm_w = smem;
  float* keep_pol = sm_w + (KC::W_SMEM ? KC::W_FLOATS_ALL : 0);
  float* keep_rhs = keep_pol + PolCfg::KEEP_FLOATS;
  float* dbuf0 = keep_rhs + RhsCfg::KEEP_FLOATS;
  float* dbuf1 = dbuf0 + KC::DELTA_ROWS * LD;
  float* coef = dbuf1 + KC::DELTA_ROWS * LD;                   // [NM_MAX_TERMS]
  uint64_t* bar = reinterpret_cast<uint64_t*>(coef + 32);
  VarTable& vt = *reinterpret_cast<VarTable*>(coef + 48);
  const int tid = threadIdx.x;
  const NmPlan& p = plan;
  const VarTable vtc = make_vartable_regs(p, args, NX, NU, NY);      // compiled term code (see nm_terms.cuh)
  const int T = p.nsteps;
  const bool fused = S::NTERMS > 0 && args.fused_terms;
  const bool pre = !fused && args.G_x != nullptr;
  // row layout of the precomputed term gradients: G[b] = [ x(T+1, NX) | u(T, NU) | y(T, NY) ]
  const int g_row = (T + 1) * NX + T * (NU + NY), g_uoff = (T + 1) * NX, g_yoff = g_uoff + T * NU;
  // which of x / u / y carry time-broadcast atoms (those are not in G, see nm_terms.cuh)
  int tb_mask = 0;
  for (int k = 0; k < p.n_terms; ++k) {
    const NmTerm& tm = p.terms[k];
    const NmAtom* at[4] = {&tm.lhs.a, &tm.lhs.b, &tm.rhs.a, &tm.rhs.b};
    const bool live[4] = {true, tm.lhs.op != NM_OP_NONE, true, tm.rhs.op != NM_OP_NONE};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (live[i] && at[i]->var >= 0 && at[i]->var <= NM_VAR_Y && atom_is_tbcast(*at[i], tm)) tb_mask |= 1 << at[i]->var;
  }

  const float* Wpol;
  const float* Wrhs;
  if constexpr (KC::W_SMEM) {
    stage_weights<KC::W_FLOATS_ALL, 0>(sm_w, args.wprep, nullptr, nullptr, bar, tid);
    Wpol = sm_w + KC::POL_W_OFF;
    Wrhs = sm_w + KC::RHS_W_OFF;
  } else {
    Wpol = args.wprep + KC::POL_W_OFF;
    Wrhs = args.wprep + KC::RHS_W_OFF;
  }

  // effective upstream gradient of every fused term: d(term) + d(objective|penalty) + d(loss)
  if (tid < NM_MAX_TERMS) coef[tid] = tid < p.n_terms ? term_coef(p, tid, args.grad_scalars, args.b_total) : 0.f;
  // this CTA's private gradient block (summed over CTAs by nm_grad_finalize_kernel)
  float* gblock = args.grad_partials + size_t(blockIdx.x) * size_t(p.n_params > 0 ? p.n_params : 1);
  for (int i = tid; i < p.n_params; i += THREADS) gblock[i] = 0.f;
  if (tid == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = T + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = T; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = T; vt.D[NM_VAR_Y] = NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
  }
  __syncthreads();

  using PolStore = std::conditional_t<KC::HAS_POL && !SPILL, DwStore<PolCfg, TILE, THREADS>, DwNone>;
  using RhsStore = std::conditional_t<KC::HAS_RHS_MLP && !SPILL, DwStore<RhsCfg, TILE, THREADS>, DwNone>;
  PolStore pol_dw; RhsStore rhs_dw;
  pol_dw.zero(); rhs_dw.zero();
  float* gpol = gblock + (KC::HAS_POL ? p.policy.param_offset : 0);
  float* grhs = gblock + (KC::HAS_RHS_MLP ? p.rhs_mlp.param_offset : 0);
  const bool pol_train = KC::HAS_POL && p.policy.trainable;
  const bool rhs_train = KC::HAS_RHS_MLP && (p.rhs_mlp.trainable & 1);
  // trainable white-box constants (system identification): for a non-MLP right-hand side rhs_mlp.trainable / .param_offset
  // say that -- and where in the flat gradient -- d loss / d rhs_const[0 .. NC) is wanted
  // (hybrid right-hand sides: bit 0 of .trainable is the MLP's weights, bit 1 the constants, stored behind the MLP's parameters)
  constexpr int NC = (KC::HAS_RHS_MLP && !KC::HYB) ? 0 : rhs_num_const<S>();
  const bool theta_train = NC > 0 && (KC::HYB ? (p.rhs_mlp.trainable & 2) != 0 : p.rhs_mlp.trainable != 0);
  int theta_off = p.rhs_mlp.param_offset;
  if constexpr (KC::HYB) {
    for (int l = 0; l < p.rhs_mlp.n_layers; ++l) theta_off += p.rhs_mlp.sizes[l + 1] * (p.rhs_mlp.sizes[l] + (p.rhs_mlp.has_bias ? 1 : 0));
  }
  float gth[NM_MAX_CONST];
#pragma unroll
  for (int c = 0; c < NM_MAX_CONST; ++c) gth[c] = 0.f;

  for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
    const int64_t b_raw = int64_t(tile) * TILE + tid;
    const bool owner = tid < TILE;
    const bool valid = owner && b_raw < args.B;
    const int64_t b = b_raw < args.B ? b_raw : args.B - 1;

    float lam[NX];                                              // dL/dx_{t+1}
    float x_pre[NX];                                            // x_t fetched one step ahead (L2 latency hidden)
#pragma unroll
    for (int i = 0; i < NX; ++i) { lam[i] = 0.f; x_pre[i] = 0.f; }
    if (owner) {
#pragma unroll
      for (int i = 0; i < NX; ++i) x_pre[i] = __ldg(args.x_traj + (b * (T + 1) + (T - 1)) * NX + i);
    }
    if (valid) {
      if (pre) {
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = __ldg(args.G_x + b * g_row + T * NX + i);
        if (tb_mask & 1) term_grad_accum_ool<NX, 2>(p, vt, coef, b, NM_VAR_X, T, NX, lam);
      } else {
        if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_X, NX, NX>(p, vtc, coef, b, T, lam); }
        else term_grad_accum_ool<NX, 0>(p, vt, coef, b, NM_VAR_X, T, NX, lam);
        if (args.gx_ext != nullptr) {
#pragma unroll
          for (int i = 0; i < NX; ++i) lam[i] += __ldg(args.gx_ext + (b * (T + 1) + T) * NX + i);
        }
      }
    }
#pragma unroll 1
    for (int t = T - 1; t >= 0; --t) {
      float x[NX], u[NUA], y[NYA], o[NUA], gx[NX], gu[NUA], gy[NYA];
#pragma unroll
      for (int j = 0; j < NUA; ++j) { u[j] = 0.f; o[j] = 0.f; gu[j] = 0.f; }
#pragma unroll
      for (int j = 0; j < NYA; ++j) { y[j] = 0.f; gy[j] = 0.f; }
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = 0.f; gx[i] = 0.f; }
      if (owner) {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = x_pre[i];
        if (t > 0) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x_pre[i] = __ldg(args.x_traj + (b * (T + 1) + t - 1) * NX + i);
        }
        if constexpr (S::HAS_OUTPUT) {
#pragma unroll
          for (int j = 0; j < NY; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NX; ++i) s = fmaf(x[i], p.C[j * NX + i], s);
            y[j] = s;
          }
        }
        if constexpr (KC::HAS_POL) build_policy_input<S>(p, args.exo, keep_pol, tid, b, t, x, y);
      }
      if constexpr (KC::HAS_POL) {
        __syncthreads();
        mlp_forward<PolCfg, typename S::Pol, TILE, THREADS, true, KC::HOIST>(Wpol, nullptr, nullptr, keep_pol, dbuf0, p.policy.act_param, tid,
                                                                             PreAct{args.pre, int64_t(tile) * TILE, args.B, T, t});
        if (owner) {
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            o[j] = dbuf0[j * LD + tid];
            u[j] = bounds_fwd<S::Pol::BOUNDS>(o[j], p.policy.bmin[j], p.policy.bmax[j]);
          }
        }
      }
      if constexpr (!KC::HAS_POL && S::ACT_EXO >= 0) {       // open loop: the action comes from the data
        if (owner) {
          const float* src = args.exo[S::ACT_EXO] + (b * p.exo[S::ACT_EXO].steps + t) * NU;
#pragma unroll
          for (int j = 0; j < NU; ++j) u[j] = __ldg(src + j);
        }
      }
      // ---- dynamics adjoint: (gx, gu) = (d x_{t+1} / d (x_t, u_t))^T lam
      if constexpr (!KC::HAS_RHS_MLP) {
        if constexpr (NC > 0) {
          if (owner) {
            if (theta_train) Dyn<S>::step_vjp_theta(p, x, u, lam, gx, gu, gth);      // lam is zero for rows past the batch
            else Dyn<S>::step_vjp(p, x, u, lam, gx, gu);
          }
        } else {
          if (owner) Dyn<S>::step_vjp(p, x, u, lam, gx, gu);
        }
      } else {
        constexpr int NS = Tableau<S::INTEG>::NS, RX = KC::RX;
        const float h = p.h;
        float k[NM_MAX_STAGES][NX], gk[NM_MAX_STAGES][NX];
        // pass 1: stage derivatives k_0 .. k_{NS-2}
        static_for<0, NS>([&](auto sc) {
          constexpr int St = decltype(sc)::value;
#pragma unroll
          for (int i = 0; i < NX; ++i) gk[St][i] = 0.f;
          if constexpr (St + 1 < NS) {
           if (args.chk != nullptr) {                           // kept by the forward kernel: no recompute
            if (owner) {
              const float* cp = args.chk + ((b * T + t) * (NS - 1) + St) * NX;
#pragma unroll
              for (int i = 0; i < RX; ++i) k[St][i] = __ldg(cp + i);
            }
           } else {
            float xs[RX];
            if (owner) {
              if constexpr (KC::SO) so_position<S::INTEG, St, NX>(h, x, k, xs);
              else stage_input<S::INTEG, St, NX>(h, x, k, xs);
#pragma unroll
              for (int i = 0; i < RX; ++i) keep_rhs[i * LD + tid] = xs[i];
#pragma unroll
              for (int j = 0; j < NU; ++j) keep_rhs[(RX + j) * LD + tid] = u[j];
            }
            __syncthreads();
            mlp_forward<RhsCfg, typename S::RhsM, TILE, THREADS, true>(Wrhs, nullptr, nullptr, keep_rhs, dbuf0, p.rhs_mlp.act_param, tid);
            if (owner) {
              if constexpr (KC::HYB) hybrid_eval<S>(p, xs, dbuf0[tid], k[St]);
              else {
#pragma unroll
                for (int i = 0; i < RX; ++i) k[St][i] = dbuf0[i * LD + tid];
              }
            }
            __syncthreads();                                     // dbuf0 is rewritten by the next stage
           }
          }
        });
        // pass 2: stages in reverse, recomputing each stage's activations
#pragma unroll
        for (int i = 0; i < NX; ++i) gx[i] = lam[i];
        constexpr int SOH = KC::SO ? NX / 2 : 1;                // second-order schemes: running cotangents of position / velocity
        float so_gq[SOH], so_gv[SOH], so_ga[SOH];
        if constexpr (KC::SO) so_adj_init<S::INTEG, NX>(h, lam, so_gq, so_gv, so_ga);
        static_for_rev<0, NS>([&](auto sc) {
          constexpr int St = decltype(sc)::value;
          float xs[RX], gdir[RX];
          if (owner) {
            if constexpr (KC::SO) so_position<S::INTEG, St, NX>(h, x, k, xs);
            else stage_input<S::INTEG, St, NX>(h, x, k, xs);
#pragma unroll
            for (int i = 0; i < RX; ++i) keep_rhs[i * LD + tid] = xs[i];
#pragma unroll
            for (int j = 0; j < NU; ++j) keep_rhs[(RX + j) * LD + tid] = u[j];
          }
          __syncthreads();
          mlp_forward<RhsCfg, typename S::RhsM, TILE, THREADS, true>(Wrhs, nullptr, nullptr, keep_rhs, dbuf0, p.rhs_mlp.act_param, tid);
          if (owner) {
            float w[RX];
            if constexpr (KC::SO) so_adj_cot<S::INTEG, St, NX>(h, so_gv, so_ga, w);
            else stage_cotangent<S::INTEG, St, NX>(h, lam, gk, w);
            if constexpr (KC::HYB) {
              // grey-box: the stage cotangent goes through the white-box equations to the MLP's single output (rows past
              // the batch carry w = 0: nothing reaches the weight or constant gradients)
              float wm;
#pragma unroll
              for (int i = 0; i < RX; ++i) w[i] = valid ? w[i] : 0.f;
              if (theta_train) hybrid_vjp<S, true>(p, xs, dbuf0[tid], w, wm, gdir, gth);
              else hybrid_vjp<S, false>(p, xs, dbuf0[tid], w, wm, gdir, gth);
              dbuf0[tid] = wm;
            } else {
#pragma unroll
              for (int i = 0; i < RX; ++i) dbuf0[i * LD + tid] = valid ? w[i] : 0.f;
            }
          }
          __syncthreads();
          mlp_backward<RhsCfg, typename S::RhsM, TILE, THREADS, SPILL>(Wrhs, keep_rhs, dbuf0, dbuf1, rhs_dw, grhs,
                                                                        p.rhs_mlp.has_bias, rhs_train, p.rhs_mlp.act_param, tid);
          if (owner) {
            const float* gin = (RhsCfg::NL & 1) ? dbuf1 : dbuf0;
            float gxs[RX];
#pragma unroll
            for (int i = 0; i < RX; ++i) gxs[i] = gin[i * LD + tid];
            if constexpr (KC::HYB) {
#pragma unroll
              for (int i = 0; i < RX; ++i) gxs[i] += gdir[i];
            }
            if constexpr (KC::SO) so_adj_push<S::INTEG, St, NX>(h, gxs, so_gq, so_gv, so_ga);
            else stage_push<S::INTEG, St, NX>(h, gxs, gx, gk);
#pragma unroll
            for (int j = 0; j < NU; ++j) gu[j] += gin[(RX + j) * LD + tid];
          }
          __syncthreads();                                       // gin buffer is rewritten by the next stage
        });
        if constexpr (KC::SO) {
#pragma unroll
          for (int i = 0; i < NX / 2; ++i) { gx[i] = so_gq[i]; gx[NX / 2 + i] = so_gv[i]; }
        }
      }
      // ---- loss-term gradients w.r.t. u_t and y_t
      if (valid) {
        if (pre) {
#pragma unroll
          for (int j = 0; j < NU; ++j) gu[j] += __ldg(args.G_x + b * g_row + g_uoff + t * NU + j);
#pragma unroll
          for (int j = 0; j < NY; ++j) gy[j] += __ldg(args.G_x + b * g_row + g_yoff + t * NY + j);
          if constexpr (NU > 0) { if (tb_mask & 2) term_grad_accum_ool<NUA, 2>(p, vt, coef, b, NM_VAR_U, t, NU, gu); }
          if constexpr (NY > 0) { if (tb_mask & 4) term_grad_accum_ool<NYA, 2>(p, vt, coef, b, NM_VAR_Y, t, NY, gy); }
        } else {
          if constexpr (NU > 0) {
            if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_U, NU, NUA>(p, vtc, coef, b, t, gu); }
            else term_grad_accum_ool<NUA, 0>(p, vt, coef, b, NM_VAR_U, t, NU, gu);
            if (args.gu_ext != nullptr) {
#pragma unroll
              for (int j = 0; j < NU; ++j) gu[j] += __ldg(args.gu_ext + (b * T + t) * NU + j);
            }
          }
          if constexpr (NY > 0) {
            if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_Y, NY, NYA>(p, vtc, coef, b, t, gy); }
            else term_grad_accum_ool<NYA, 0>(p, vt, coef, b, NM_VAR_Y, t, NY, gy);
            if (args.gy_ext != nullptr) {
#pragma unroll
              for (int j = 0; j < NY; ++j) gy[j] += __ldg(args.gy_ext + (b * T + t) * NY + j);
            }
          }
        }
      }
      // ---- policy adjoint
      if constexpr (KC::HAS_POL) {
        if (owner) {
#pragma unroll
          for (int j = 0; j < NU; ++j)
            dbuf0[j * LD + tid] = valid ? gu[j] * bounds_deriv<S::Pol::BOUNDS>(o[j], p.policy.bmin[j], p.policy.bmax[j]) : 0.f;
        }
        __syncthreads();
        mlp_backward<PolCfg, typename S::Pol, TILE, THREADS, SPILL, KC::HOIST>(Wpol, keep_pol, dbuf0, dbuf1, pol_dw, gpol,
                                                                     p.policy.has_bias, pol_train, p.policy.act_param, tid,
                                                                     PreAct{args.dz1, int64_t(tile) * TILE, args.B, T, t});
        if (owner) {
          const float* gin = (PolCfg::NL & 1) ? dbuf1 : dbuf0;
          int off = 0;
          static_for<0, S::NSEG>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int W = S::seg_width(s);
            if constexpr (S::seg_kind(s) == NM_SEG_STATE) {
#pragma unroll
              for (int i = 0; i < W; ++i) gx[i] += gin[(off + i) * LD + tid];
            } else if constexpr (S::seg_kind(s) == NM_SEG_OUTPUT) {
#pragma unroll
              for (int i = 0; i < W; ++i) gy[i] += gin[(off + i) * LD + tid];
            }
            off += W;
          });
        }
        // the next step rewrites dbuf0 (raw policy output) only after its own barriers; the input
        // gradient buffer `gin` is rewritten by the next step's mlp_backward, also behind barriers
      }
      if (owner) {
        if constexpr (S::HAS_OUTPUT) {                           // y_t = C x_t
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NY; ++j) s = fmaf(p.C[j * NX + i], gy[j], s);
            gx[i] += s;
          }
        }
        if (valid) {
          if (pre) {
#pragma unroll
            for (int i = 0; i < NX; ++i) gx[i] += __ldg(args.G_x + b * g_row + t * NX + i);
            if (tb_mask & 1) term_grad_accum_ool<NX, 2>(p, vt, coef, b, NM_VAR_X, t, NX, gx);
          } else {
            if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_X, NX, NX>(p, vtc, coef, b, t, gx); }
            else term_grad_accum_ool<NX, 0>(p, vt, coef, b, NM_VAR_X, t, NX, gx);
            if (args.gx_ext != nullptr) {
#pragma unroll
              for (int i = 0; i < NX; ++i) gx[i] += __ldg(args.gx_ext + (b * (T + 1) + t) * NX + i);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = valid ? gx[i] : 0.f;
      }
    }
    if (valid && args.grad_x0 != nullptr) {
#pragma unroll
      for (int i = 0; i < NX; ++i) args.grad_x0[b * NX + i] = lam[i];
    }
  }
  // gradient of the white-box constants: fixed-order sum over the CTA's threads (deterministic) -> gradient block
  if constexpr (NC > 0) {
    if (theta_train) {
      __syncthreads();                                           // everybody is done with the staging buffers
#ifdef NM_HOST_EMU                                                // tests/emu: one block at a time, its threads are OS threads
      static float red[(THREADS / 32) * (NC > 0 ? NC : 1)];
#else
      __shared__ float red[(THREADS / 32) * (NC > 0 ? NC : 1)];   // [warp][constant]
#endif
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float v = gth[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[(tid >> 5) * NC + c] = v;
      }
      __syncthreads();
      if (tid < NC) {
        float v = 0.f;
        for (int w = 0; w < THREADS / 32; ++w) v += red[w * NC + tid];
        gblock[theta_off + tid] = v;
      }
      __syncthreads();
    }
  }
  // persistent accumulators -> this CTA's gradient block
  if constexpr (!SPILL) {
    __syncthreads();
    if constexpr (KC::HAS_POL) { if (pol_train) dw_writeout<PolCfg, TILE, THREADS>(pol_dw, gpol, p.policy.has_bias, tid); }
    if constexpr (KC::HAS_RHS_MLP) { if (rhs_train) dw_writeout<RhsCfg, TILE, THREADS>(rhs_dw, grhs, p.rhs_mlp.has_bias, tid); }
  }
  (void)ND; (void)Wrhs; (void)Wpol; (void)keep_rhs;
}

// ------------------------------------------------------------------------------------------------
// Hoist GEMMs (see HoistedMlp).  Rows r = b * T + t over the whole batch and horizon; a CTA takes tiles of TILE rows.
struct ExoPtrs { const float* p[NM_MAX_EXO]; };
template <class S> constexpr int exo_seg_off(int s) {        // offset of segment s within the exogenous part
  int o = 0;
  for (int i = 0; i < s; ++i) if (S::seg_kind(i) == NM_SEG_EXO) o += S::seg_width(i);
  return o;
}
template <class S> constexpr int exo_seg_wmax() {
  int m = 4;
  for (int i = 0; i < S::NSEG; ++i) if (S::seg_kind(i) == NM_SEG_EXO) m = cmax(m, S::seg_width(i));
  return m;
}
template <class S> constexpr size_t exo_gemm_smem_bytes() { return size_t(exo_seg_wmax<S>() + KCfg<S>::N0) * (S::TILE + 4) * 4; }

// gather rows r0 .. r0+TILE of exogenous segment `s` (width W) into in[c][row]
template <class S, int s>
__device__ __forceinline__ void exo_gather(const NmPlan& p, const ExoPtrs& exo, float* __restrict__ in, int tid, int64_t r0, int64_t R) {
  constexpr int W = S::seg_width(s), e = S::seg_index(s), LD = S::TILE + 4;
  if (tid < S::TILE) {
    const int64_t r = r0 + tid < R ? r0 + tid : R - 1;
    const int64_t b = r / p.nsteps;
    const int t = int(r - b * p.nsteps);
    const float* __restrict__ src = exo.p[e] + (b * p.exo[e].steps + t) * W;
#pragma unroll 4
    for (int i = 0; i < W; i += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
      in[(i + 0) * LD + tid] = v.x; in[(i + 1) * LD + tid] = v.y; in[(i + 2) * LD + tid] = v.z; in[(i + 3) * LD + tid] = v.w;
    }
  }
}

// E[r][n] = bias[n] + sum over exogenous segments, k:  W_0[n][KS + off + k] * exo_seg[r][k]
template <class S>
__global__ void __launch_bounds__(S::THREADS) nm_exo_gemm_fwd_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ wprep,
                                                                     const __grid_constant__ ExoPtrs exo, float* __restrict__ E, int64_t B) {
  using KC = KCfg<S>;
  constexpr int TILE = S::TILE, THREADS = S::THREADS, LD = TILE + 4, N0 = KC::N0, NP = KC::EXO_NP;
  using G = GemmCfg<N0, TILE, THREADS>;
  extern __shared__ __align__(128) float smem[];
  float* in = smem;                                          // [wmax][LD]
  float* out = smem + exo_seg_wmax<S>() * LD;                // [N0][LD]
  const int tid = threadIdx.x;
  const NmPlan& p = plan;
  const float* __restrict__ wexo = wprep + KC::EXO_W_OFF;
  const int64_t R = B * p.nsteps;
  const int64_t tiles = (R + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TILE;
    static_for<0, S::NSEG>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
        constexpr int W = S::seg_width(s), OFF = exo_seg_off<S>(s);
        __syncthreads();                                     // the previous GEMM has finished reading `in`
        exo_gather<S, s>(p, exo, in, tid, r0, R);
        __syncthreads();
        gemm_phase<W, N0, TILE, THREADS, LD>(wexo + OFF * NP, in, tid, [&](int ng, int traj0, float (&acc)[G::TMB][G::TN]) {
#pragma unroll
          for (int j = 0; j < G::TN; ++j) {
            const int n = ng + G::NG * j;
            if (n < N0) {
              float v[G::TMB];
              if constexpr (OFF == 0) {
#pragma unroll
                for (int i = 0; i < G::TMB; ++i) v[i] = acc[i][j];
              } else {
                ld_vec<G::TMB>(v, out + n * LD + traj0);    // same thread wrote it for the previous segment
#pragma unroll
                for (int i = 0; i < G::TMB; ++i) v[i] += acc[i][j];
              }
              st_vec<G::TMB>(out + n * LD + traj0, v);
            }
          }
        });
      }
    });
    __syncthreads();
    // consecutive threads -> consecutive outputs of one row: coalesced rows of N0 floats
    for (int idx = tid; idx < N0 * TILE; idx += THREADS) {
      const int row = idx / N0, n = idx - row * N0;
      if (r0 + row < R) E[(r0 + row) * N0 + n] = out[n * LD + row] + wexo[KC::KE * NP + g_col(N0, n)];
    }
  }
}

// dW_0[n][KS + off + k] += sum_r dZ_1[r][n] * exo_seg[r][k]   into this CTA's gradient block (after the adjoint kernel)
template <class S>
__global__ void __launch_bounds__(S::THREADS) nm_exo_gemm_bwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ ExoPtrs exo,
                                                                     const float* __restrict__ dZ1, float* __restrict__ grad_partials, int64_t B) {
  using KC = KCfg<S>;
  constexpr int TILE = S::TILE, THREADS = S::THREADS, LD = TILE + 4, N0 = KC::N0, KW = S::Pol::sz(0);
  extern __shared__ __align__(128) float smem[];
  float* in = smem;                                          // [wmax][LD]   exogenous rows, transposed
  float* delta = smem + exo_seg_wmax<S>() * LD;              // [N0][LD]     dZ_1 rows, transposed
  const int tid = threadIdx.x;
  const NmPlan& p = plan;
  const int64_t R = B * p.nsteps;
  const int64_t tiles = (R + TILE - 1) / TILE;
  float* gpol = grad_partials + size_t(blockIdx.x) * size_t(p.n_params) + p.policy.param_offset;
  static_for<0, S::NSEG>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
      constexpr int W = S::seg_width(s), OFF = exo_seg_off<S>(s);
      using D = DwCfg<N0, W, TILE, THREADS>;
      float acc[D::SLOTS][D::RJ][D::RK];
#pragma unroll
      for (int q = 0; q < D::SLOTS; ++q)
#pragma unroll
        for (int i = 0; i < D::RJ; ++i)
#pragma unroll
          for (int j = 0; j < D::RK; ++j) acc[q][i][j] = 0.f;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t r0 = tile * TILE;
        __syncthreads();                                     // the previous tile's products are done
        exo_gather<S, s>(p, exo, in, tid, r0, R);
        for (int idx = tid; idx < N0 * TILE; idx += THREADS) {
          const int row = idx / N0, n = idx - row * N0;
          delta[n * LD + row] = r0 + row < R ? __ldg(dZ1 + (r0 + row) * N0 + n) : 0.f;   // rows past the end contribute nothing
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < D::SLOTS; ++q) dw_tile<N0, W, TILE, THREADS, LD>(acc[q], delta, in, q * THREADS + tid);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < D::SLOTS; ++q) dw_flush_tile<D, N0, W, KW>(acc[q], gpol + KC::KS + OFF, q * THREADS + tid);
    }
  });
}


}  // namespace nm
