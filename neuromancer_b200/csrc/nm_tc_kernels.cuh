// nm_tc_kernels.cuh -- tensor-core (tcgen05 / TMEM) variant of the fused rollout for neural-ODE systems:
// System([Node(integrator(MLP))]) with an open-loop input from the data or none (reference
// examples/ODEs/Part_1_NODE.py; system.py:261-277, integrators.py:144-301, blocks.py:169-177).
//
// One CTA = 128 trajectories = the 128 lanes of tensor memory; thread m owns trajectory m of the tile, its
// state and stage derivatives live in registers for the whole horizon.  Per integrator stage:
//   layer 0        CUDA cores (fan-in = nx + nu), activation, tf32 hi/lo split -> the thread's TMEM lane
//   hidden layers  tcgen05.mma kind::tf32, A = activations in TMEM, B = weights in shared memory (staged once
//                  per CTA by one bulk TMA copy), 3-product split for fp32-level accuracy, accumulator in TMEM;
//                  epilogue (tcgen05.ld -> activation -> split -> tcgen05.st) by the owning thread
//   last layer     CUDA cores, fused into the epilogue of the last hidden layer
#pragma once
#include "nm_kernels.cuh"
#include "nm_tc.cuh"

namespace nm {

// Two kinds of systems run on this path:
//   RHS mode     System([Node(integrator(MLP))]): the MLP is the right-hand side, evaluated once per integrator stage
//   POLICY mode  System([Node(MLP policy), Node(integrator(white-box RHS) | SSM)]): the MLP is the closed-loop policy
//                (inputs: state and exogenous segments, output through MLP_bounds), evaluated once per step; the
//                dynamics are thread-per-trajectory register math (Dyn<S>).  The forward kernel keeps the RAW policy
//                output o_t as the checkpoint, so that the adjoint knows u_t = bounds(o_t) and the cotangent of o_t
//                before it starts the step's only MLP evaluation.
#ifndef NM_TUNE_HOIST
#define NM_TUNE_HOIST 0
#endif
// exogenous first-layer hoist (nm_kernels.cuh, HoistedMlp) for a policy: every exogenous segment follows the last
// state / output segment; returns the width of the leading (contracted) part, or 0 when the class is not eligible
template <class S>
constexpr int tc_hoist_ks() {
  if (S::Pol::NL < 2 || S::Pol::sz(1) % 4 != 0) return 0;
  int ks = 0, seen_exo = 0;
  for (int i = 0; i < S::NSEG; ++i) {
    if (S::seg_kind(i) == NM_SEG_PREVIEW) return 0;
    if (S::seg_kind(i) == NM_SEG_EXO) { seen_exo = 1; if (S::seg_width(i) % 4 != 0) return 0; }
    else { if (seen_exo) return 0; ks += S::seg_width(i); }
  }
  return (seen_exo && ks >= 1) ? ks : 0;
}
template <class S>
struct TcMode {
  static constexpr bool POLICY = S::RHS != NM_RHS_MLP && S::Pol::NL >= 3;
  static constexpr bool HOIST = POLICY && (NM_TUNE_HOIST != 0) && tc_hoist_ks<S>() > 0;
  using PolM = std::conditional_t<HOIST, HoistedMlp<typename S::Pol, (tc_hoist_ks<S>() > 0 ? tc_hoist_ks<S>() : 1)>, typename S::Pol>;
  using M = std::conditional_t<POLICY, PolM, typename S::RhsM>;
  static constexpr int N0 = M::sz(1);                                        // width of the first hidden layer (rows of E / dZ_1)
  static constexpr int NOUT = POLICY ? S::NU : S::NX;                       // width of the MLP output
  static constexpr int EVALS = POLICY ? 1 : Tableau<S::INTEG>::NS;          // MLP evaluations per step
  // checkpoint floats per trajectory-step; policy mode: [B*T rows of E (hoist only, N0 floats) | B*T raw outputs o_t (NU floats)]
  static constexpr int E_ROW = HOIST ? N0 : 0;
  static constexpr int CHK = POLICY ? S::NU + E_ROW : (Tableau<S::INTEG>::NS - 1) * S::NX;
  template <class T_> static __device__ __forceinline__ T_* o_base(T_* chk, int64_t B, int T) { return chk + int64_t(E_ROW) * B * T; }
  static __device__ __forceinline__ const NmMlp& mlp(const NmPlan& p) { if constexpr (POLICY) return p.policy; else return p.rhs_mlp; }
  // policy input = state, output and exogenous segments (no preview windows on this path)
  static constexpr bool segs_ok() {
    int w = 0;
    for (int i = 0; i < S::NSEG; ++i) {
      if (S::seg_kind(i) == NM_SEG_PREVIEW) return false;
      if (S::seg_kind(i) == NM_SEG_STATE && S::seg_width(i) != S::NX) return false;
      if (S::seg_kind(i) == NM_SEG_OUTPUT && (!S::HAS_OUTPUT || S::seg_width(i) != S::NY)) return false;
      w += S::seg_width(i);
    }
    return w == (POLICY ? S::Pol::sz(0) : M::sz(0));
  }
};

template <class S>
struct TcKCfg {
  using MD = TcMode<S>;
  using C = tc::TcCfg<typename MD::M>;
  static constexpr int NX = S::NX, NU = S::NU;
  static constexpr bool ELIGIBLE_RHS = S::RHS == NM_RHS_MLP && !second_order<S::INTEG>() && S::Pol::NL == 0 && !S::HAS_OUTPUT && S::ND == 0 && C::shape_ok() &&
                                       C::K(0) == NX + NU && C::N(C::NL - 1) == NX;
  // (policy mode also carries an output node y = C x and an additive disturbance: both are register math of the
  // trajectory's thread)
  static constexpr bool ELIGIBLE_POL = MD::POLICY && S::ACT_EXO < 0 && NU > 0 && C::shape_ok() &&
                                       MD::segs_ok() && C::N(C::NL - 1) == NU;
  static constexpr bool ELIGIBLE = MD::POLICY ? ELIGIBLE_POL : ELIGIBLE_RHS;
  // shared memory of the forward kernel: weight block + barriers/slots; padded so that no more CTAs become
  // resident than tensor memory has room for (512 columns per SM)
  static constexpr int MAX_CTAS = 512 / C::FWD_COLS;
  static constexpr size_t FWD_W_BYTES = size_t(round_up(C::FWD_W, 32)) * 4;
  static constexpr size_t FWD_SMEM_RAW = FWD_W_BYTES + 1024;
  static constexpr size_t FWD_SMEM_CAP = (227 * 1024) / (MAX_CTAS + 1) + 1024;
  static constexpr size_t FWD_SMEM_BYTES = FWD_SMEM_RAW > FWD_SMEM_CAP ? FWD_SMEM_RAW : FWD_SMEM_CAP;
  static constexpr bool FITS_FWD = FWD_SMEM_RAW <= 227 * 1024;
};

// policy input of trajectory b at step t: [x | exo[e][b, t, :] ...] in segment order (reference blocks.py:32-43)
template <class S, int K0>
__device__ __forceinline__ void tc_policy_input(const NmPlan& p, const float* const* exo, int64_t b, int t, const float (&x)[S::NX], float (&in)[K0]) {
  int off = 0;
  static_for<0, S::NSEG>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    constexpr int W = S::seg_width(s);
    if constexpr (S::seg_kind(s) == NM_SEG_STATE) {
#pragma unroll
      for (int i = 0; i < W; ++i) in[off + i] = x[i];
    } else if constexpr (S::seg_kind(s) == NM_SEG_OUTPUT) {
#pragma unroll
      for (int i = 0; i < W; ++i) in[off + i] = 0.f;                 // filled by tc_policy_state
    } else if constexpr (!TcMode<S>::HOIST) {                         // (hoisted: the exogenous segments enter through E)
      constexpr int e = S::seg_index(s);
      const float* __restrict__ src = exo[e] + (b * p.exo[e].steps + t) * W;
#pragma unroll
      for (int i = 0; i < W; ++i) in[off + i] = __ldg(src + i);
    }
    off += W;
  });
}

// overwrite the state / output segment(s) of a policy input whose exogenous entries were fetched one step ahead
template <class S, int K0>
__device__ __forceinline__ void tc_policy_state(const float (&x)[S::NX], const float (&y)[S::NY > 0 ? S::NY : 1], float (&in)[K0]) {
  int off = 0;
  static_for<0, S::NSEG>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (S::seg_kind(s) == NM_SEG_STATE) {
#pragma unroll
      for (int i = 0; i < S::seg_width(s); ++i) in[off + i] = x[i];
    } else if constexpr (S::seg_kind(s) == NM_SEG_OUTPUT) {
#pragma unroll
      for (int i = 0; i < S::seg_width(s); ++i) in[off + i] = y[i];
    }
    off += S::seg_width(s);
  });
}
// y = C x (reference: the output node of grid_response.ipynb cell 7)
template <class S>
__device__ __forceinline__ void tc_output(const NmPlan& p, const float (&x)[S::NX], float (&y)[S::NY > 0 ? S::NY : 1]) {
#pragma unroll
  for (int j = 0; j < (S::NY > 0 ? S::NY : 1); ++j) y[j] = 0.f;
  if constexpr (S::HAS_OUTPUT) {
#pragma unroll
    for (int j = 0; j < S::NY; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < S::NX; ++i) s = fmaf(x[i], p.C[j * S::NX + i], s);
      y[j] = s;
    }
  }
}

template <class S>
__global__ void nm_tc_prep_fwd_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, float* __restrict__ wprep) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  tc::prep_tc_fwd<typename TcMode<S>::M>(params + TcMode<S>::mlp(plan).param_offset, wprep, gtid, gthreads);
}

// One MLP evaluation on the tile: in[K0] (registers) -> out[NOUT] (registers).  All 128 threads call it.
//   `phase` is the parity of the MMA-completion mbarrier (flipped per tensor-core layer).
//   `pre` != nullptr (hoisted first layer): z_1 = pre[n] + W[:, :K0] in   (pre = E row: bias + exogenous part)
template <class M, int NOUT>
__device__ __forceinline__ void tc_mlp_forward(const float* __restrict__ W, uint32_t tmem, uint32_t lane_base, uint64_t* mma_bar,
                                               uint32_t& phase, float act_prm, const float (&in)[tc::TcCfg<M>::K(0)],
                                               float (&out)[NOUT], int tid, const float* __restrict__ pre = nullptr) {
  using C = tc::TcCfg<M>;
  static_assert(NOUT == C::N(C::NL - 1), "output width");
  constexpr int K0 = C::K(0), N0 = C::N(0), NL = C::NL;
  const uint32_t t_oph = tmem + lane_base + C::OPH, t_opl = tmem + lane_base + C::OPL, t_acc = tmem + lane_base + C::ACC;

  // ---- layer 0 on the CUDA cores: z[n] = b[n] + sum_i W[n][i] in[i]; operand of layer 1 -> tensor memory
  {
    constexpr int KP1 = C::kp(1), CH = tc::chunk_of(KP1);
    const float* __restrict__ W0 = W + C::first_off();
#pragma unroll
    for (int c0 = 0; c0 < KP1; c0 += CH) {
      float hi[CH], lo[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int n = c0 + i;
        float a = 0.f;
        if (n < N0) {
          float w[C::R0];
          ld_vec<C::R0>(w, W0 + n * C::R0);
          float z = pre != nullptr ? __ldg(pre + n) : (C::BIAS ? w[K0] : 0.f);
#pragma unroll
          for (int q = 0; q < K0; ++q) z = fmaf(w[q], in[q], z);
          a = act_fwd<M::ACT>(z, act_prm);
        } else if (n == N0 && C::BIAS) a = 1.f;
        hi[i] = tc::tf32_hi(a);
        lo[i] = a - hi[i];
      }
      tc::tmem_st<CH>(t_oph + c0, hi);
      tc::tmem_st<CH>(t_opl + c0, lo);
    }
  }
#pragma unroll
  for (int i = 0; i < NOUT; ++i) out[i] = 0.f;

  // ---- hidden layers on the tensor cores
  static_for<1, NL - 1>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int KP = C::kp(l), NP = C::np(l), Nl = C::N(l);
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    if (tid < 32) {                                       // warp 0, converged: one elected lane issues
      tc::fence_after();
      constexpr uint32_t idesc = tc::idesc_tf32(128, NP);
      constexpr uint32_t LBO = NP * 16, SBO = 128, KSTEP = 2 * NP * 16;
      const uint64_t wh = tc::smem_desc(smem_u32(W + C::img_off(l)), LBO, SBO);
      const uint64_t wl = tc::desc_add(wh, C::img_floats(l) * 4);
      if (tc::elect_one()) {
        // the two small products first: the accumulator truncates on every accumulate (see nm_tc.cuh)
        static_for<0, KP / 8>([&](auto kc) {
          constexpr int ks = decltype(kc)::value;
          tc::mma_ts_c<(ks > 0)>(tmem + C::ACC, tmem + C::OPL + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
          tc::mma_ts_c<true>(tmem + C::ACC, tmem + C::OPH + ks * 8, tc::desc_add(wl, ks * KSTEP), idesc);
        });
        static_for<0, KP / 8>([&](auto kc) {
          constexpr int ks = decltype(kc)::value;
          tc::mma_ts_c<true>(tmem + C::ACC, tmem + C::OPH + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
        });
        tc::commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc::fence_after();
    if constexpr (l < NL - 2) {
      // epilogue: activation, split, operand of the next tensor-core layer
      constexpr int KPN = C::kp(l + 1), CH = tc::chunk_of(KPN);
#pragma unroll
      for (int c0 = 0; c0 < KPN; c0 += CH) {
        float v[CH], hi[CH], lo[CH];
        tc::tmem_ld<CH>(t_acc + c0, v);
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int n = c0 + i;
          float a = 0.f;
          if (n < Nl) a = act_fwd<M::ACT>(v[i], act_prm);
          else if (n == Nl && C::BIAS) a = 1.f;
          hi[i] = tc::tf32_hi(a);
          lo[i] = a - hi[i];
        }
        tc::tmem_st<CH>(t_oph + c0, hi);
        tc::tmem_st<CH>(t_opl + c0, lo);
      }
    } else {
      // last hidden layer: activation, then the output layer on the CUDA cores
      constexpr int NLAST = C::N(NL - 1), CH = tc::chunk_of(C::np(l));
      const float* __restrict__ WL = W + C::last_off();
#pragma unroll
      for (int c0 = 0; c0 < round_up(Nl, CH); c0 += CH) {
        float v[CH];
        tc::tmem_ld<CH>(t_acc + c0, v);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = (c0 + i < Nl) ? act_fwd<M::ACT>(v[i], act_prm) : 0.f;
#pragma unroll
        for (int j = 0; j < NLAST; ++j) {
          float s = out[j];
#pragma unroll
          for (int i4 = 0; i4 < CH; i4 += 4) {
            if (c0 + i4 < C::RL) {
              const float4 w = *reinterpret_cast<const float4*>(WL + j * C::RL + c0 + i4);
              s = fmaf(w.x, v[i4], s); s = fmaf(w.y, v[i4 + 1], s); s = fmaf(w.z, v[i4 + 2], s); s = fmaf(w.w, v[i4 + 3], s);
            }
          }
          out[j] = s;
        }
      }
#pragma unroll
      for (int j = 0; j < NLAST; ++j) out[j] += W[C::lastb_off() + j];
    }
  });
}

// ================================================================================================
// forward (rollout) kernel, tensor-core variant
// as many resident CTAs as tensor memory allows (<= 4: 128 registers per thread -- measured on c2: forward 0.28 -> 0.23 ms)
#ifdef NM_TUNE_TCMINB
template <class S> constexpr int tc_fwd_minb() { return NM_TUNE_TCMINB; }
#else
template <class S> constexpr int tc_fwd_minb() { return TcKCfg<S>::MAX_CTAS < 4 ? TcKCfg<S>::MAX_CTAS : 4; }
#endif
template <class S>
__global__ void __launch_bounds__(128, tc_fwd_minb<S>()) nm_tc_fwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ FwdArgs args) {
  using MD = TcMode<S>;
  using M = typename MD::M;
  using C = tc::TcCfg<M>;
  using TK = TcKCfg<S>;
  constexpr int NX = S::NX, NU = S::NU, NUA = NU > 0 ? NU : 1, K0 = C::K(0);
  constexpr int TILE = 128, THREADS = 128;
  extern __shared__ __align__(128) float smem[];
  float* sm_w = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + round_up(C::FWD_W, 32));       // [0] weight TMA, [1] MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  VarTable& vt = *reinterpret_cast<VarTable*>(bars + 4);
  double* red = reinterpret_cast<double*>(smem);                                      // end-of-kernel reduction scratch (weights dead)
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  const VarTable vtc = make_vartable_regs(p, args, S::NX, S::NU, S::NY);
  constexpr int NT = S::NTERMS > 0 ? S::NTERMS : 1;
  const bool fused = S::NTERMS > 0 && args.fused_terms;
  if (tid == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = p.nsteps + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = p.nsteps; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = p.nsteps; vt.D[NM_VAR_Y] = S::NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<C::FWD_COLS>(tmem_slot);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], uint32_t(TK::FWD_W_BYTES));
    tma_bulk_g2s(sm_w, args.wprep, uint32_t(TK::FWD_W_BYTES), &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  uint32_t phase = 0;
  float tsum[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) tsum[q] = 0.f;
  const float act_prm = MD::mlp(p).act_param;

  const int T = p.nsteps;
  for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
    const int64_t b_raw = int64_t(tile) * TILE + tid;
    const bool valid = b_raw < args.B;
    const int64_t b = valid ? b_raw : args.B - 1;
    float x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = __ldg(args.x0 + b * NX + i);
    if (valid) {
#pragma unroll
      for (int i = 0; i < NX; ++i) args.x_traj[(b * (T + 1)) * NX + i] = x[i];
    }
    float in_pre[MD::POLICY ? K0 : 1];
    if constexpr (MD::POLICY) tc_policy_input<S, K0>(p, args.exo, b, 0, x, in_pre);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float u[NUA];
#pragma unroll
      for (int j = 0; j < NUA; ++j) u[j] = 0.f;
      float xn[NX];
      if constexpr (MD::POLICY) {
        // closed loop: u_t = bounds(policy([x_t | exo_t])), x_{t+1} = dynamics(x_t, u_t) in registers
        float in[K0], o[NUA], yv[S::NY > 0 ? S::NY : 1], dv_[S::ND > 0 ? S::ND : 1];
        tc_output<S>(p, x, yv);
        if constexpr (S::HAS_OUTPUT) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < S::NY; ++j) args.y_traj[(b * T + t) * S::NY + j] = yv[j];
          }
        }
#pragma unroll
        for (int j = 0; j < (S::ND > 0 ? S::ND : 1); ++j) dv_[j] = 0.f;
        if constexpr (S::ND > 0) {
          const float* dsrc = args.exo[S::DIST_EXO] + (b * p.exo[S::DIST_EXO].steps + t) * S::ND;
#pragma unroll
          for (int j = 0; j < S::ND; ++j) dv_[j] = __ldg(dsrc + j);
        }
#pragma unroll
        for (int q = 0; q < K0; ++q) in[q] = in_pre[q];
        tc_policy_state<S, K0>(x, yv, in);
        if (t + 1 < T) tc_policy_input<S, K0>(p, args.exo, b, t + 1, x, in_pre);   // exogenous inputs one step ahead
        tc_mlp_forward<M, MD::NOUT>(sm_w, tmem, lane_base, &bars[1], phase, act_prm, in, o, tid,
                                    MD::HOIST ? args.pre + (b * T + t) * int64_t(MD::N0) : nullptr);
#pragma unroll
        for (int j = 0; j < NU; ++j) u[j] = bounds_fwd<M::BOUNDS>(o[j], p.policy.bmin[j], p.policy.bmax[j]);
        if (valid) {
#pragma unroll
          for (int j = 0; j < NU; ++j) args.u_traj[(b * T + t) * NU + j] = u[j];
          if (args.chk != nullptr) {
#pragma unroll
            for (int j = 0; j < NU; ++j) MD::o_base(args.chk, args.B, T)[(b * T + t) * NU + j] = o[j];
          }
        }
        Dyn<S>::step(p, x, u, dv_, xn);
      } else {
        constexpr int NS = Tableau<S::INTEG>::NS;
        if constexpr (NU > 0 && S::ACT_EXO >= 0) {               // open loop: the action comes from the data
          const float* src = args.exo[S::ACT_EXO] + (b * p.exo[S::ACT_EXO].steps + t) * NU;
#pragma unroll
          for (int j = 0; j < NU; ++j) { u[j] = __ldg(src + j); if (valid) args.u_traj[(b * T + t) * NU + j] = u[j]; }
        }
        float k[NM_MAX_STAGES][NX];
        static_for<0, NS>([&](auto sc) {
          constexpr int St = decltype(sc)::value;
          float in[K0], xs[NX];
          stage_input<S::INTEG, St, NX>(p.h, x, k, xs);
#pragma unroll
          for (int i = 0; i < NX; ++i) in[i] = xs[i];
#pragma unroll
          for (int j = 0; j < NU; ++j) in[NX + j] = u[j];
          tc_mlp_forward<M, MD::NOUT>(sm_w, tmem, lane_base, &bars[1], phase, act_prm, in, k[St], tid);
          if constexpr (St + 1 < NS) {
            if (valid && args.chk != nullptr) {
              float* cp = args.chk + ((b * T + t) * (NS - 1) + St) * NX;
#pragma unroll
              for (int i = 0; i < NX; ++i) cp[i] = k[St][i];
            }
          }
        });
        combine_stages<S::INTEG, NX>(p.h, x, k, xn);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < NX; ++i) args.x_traj[(b * (T + 1) + t + 1) * NX + i] = x[i];
      }
    }
    if constexpr (S::NTERMS > 0) {
      if (fused && valid) score_traj_s<S>(p, vtc, b, tsum);
    }
  }

  if constexpr (S::NTERMS > 0) {
    if (fused) {
      __syncthreads();
      static_for<0, S::NTERMS>([&](auto kc) {
        constexpr int Kq = decltype(kc)::value;
        double v = static_cast<double>(tsum[Kq]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[Kq * (THREADS / 32) + (tid >> 5)] = v;
      });
      __syncthreads();
      if (tid < S::NTERMS) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += red[tid * (THREADS / 32) + w];
        args.term_partials[size_t(blockIdx.x) * NM_MAX_TERMS + tid] = v;
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<C::FWD_COLS>(tmem);
}



// ================================================================================================
// adjoint (BPTT) kernel, tensor-core variant
//
// 128 owner threads (thread m = trajectory m of the tile = TMEM lane m) + one control warp whose lane 0 streams
// the weight images of the hidden layers through a two-slot shared-memory ring (bulk TMA from the L2-resident
// prepared block) and issues every MMA.  Per integrator stage, in reverse stage order:
//   recompute     layer 0 on CUDA cores, hidden layers on the tensor cores (A = activations in TMEM, 3-product
//                 split); every layer's input is also written TRANSPOSED (tf32-rounded) into a swizzled shared
//                 memory tile  AT_l[feature][trajectory]  for the weight-gradient GEMM
//   backward      output layer on CUDA cores; data gradients of the hidden layers on the tensor cores
//                 (A = dZ in TMEM, B = transposed weight image); weight gradients
//                 dW_l[n][k] = sum_m dZ_{l+1}[m][n] a_l[m][k]  as M=128 (n) x N=k MMAs contracting over the 128
//                 trajectories, single-pass TF32 on round-to-nearest operands (gradient tolerance 1e-4; measured
//                 error <= 2e-5 at B = 127 and falling with B), bias gradient through a constant-one row of AT_l
//   accumulation  the TMEM accumulator truncates, so weight-gradient accumulators only run for FLUSH steps and are
//                 then added (fp32, round-to-nearest) to this CTA's private gradient block in L2
#ifdef NM_TUNE_TCPROF
#define TCP_DECL(n) long long n = 0
#define TCP_T0() const long long tcp_t0 = clock64()
#define TCP_ADD(n) n += clock64() - tcp_t0
#else
#define TCP_DECL(n)
#define TCP_T0()
#define TCP_ADD(n)
#endif
// column blocks of 16 alternate between the two threads of a trajectory; position of chunk `ci` (chunk size CH in
// {8, 16}) among the chunks its thread owns, and how many chunks a thread can own out of NCH
// (SOLO: one thread per trajectory owns every chunk)
constexpr int own_idx(int ci, int CH, bool SOLO) { return SOLO ? ci : (CH == 16 ? ci / 2 : (ci / 4) * 2 + (ci & 1)); }
constexpr int own_cnt(int NCH, int CH, bool SOLO) { return SOLO ? NCH : (CH == 16 ? (NCH + 1) / 2 : ((NCH + 3) / 4) * 2); }
template <class S>
struct TcBKCfg {
  using MD = TcMode<S>;
  using M = typename MD::M;
  using C = tc::TcCfg<M>;
  // policy mode always runs the weight-gradient GEMMs single-pass (one evaluation per step: with the 3-product
  // split the 48 weight-gradient MMAs per layer would outweigh everything else the step does)
  using B = tc::TcBwdCfg<M, !MD::POLICY>;
  static constexpr bool ELIGIBLE = TcKCfg<S>::ELIGIBLE && B::FITS;
  // Owner layout.  PAIR: two threads per trajectory (alternating 16-column blocks), one CTA per SM -- for classes whose
  // tiles fill the SM's shared memory (c3).  SOLO: one thread per trajectory and TWO CTAs per SM when shared memory
  // (<= 113 KB) and tensor memory (<= 256 columns) allow: the per-step phases are latency-bound and serialised by the
  // MMA round trips, so a second independent tile per SM nearly doubles the throughput, and the per-trajectory
  // scalar work (dynamics adjoint, loss-term gradients) is no longer duplicated across a thread pair.
#ifdef NM_TUNE_TCSOLO
  static constexpr bool SOLO = (NM_TUNE_TCSOLO != 0) && B::SMEM_BYTES + 2048 <= (227 * 1024) / 2 && B::COLS <= 256;
#else
  static constexpr bool SOLO = B::SMEM_BYTES + 2048 <= (227 * 1024) / 2 && B::COLS <= 256;
#endif
  static constexpr int OWNER_WARPS = SOLO ? 4 : 8;
  static constexpr int THREADS = OWNER_WARPS * 32 + 32;                    // owners + the control warp
#ifdef NM_TUNE_TCPAIR2
  static constexpr int CTAS_PER_SM = (SOLO || (NM_TUNE_TCPAIR2 != 0 && B::SMEM_BYTES + 2048 <= (227 * 1024) / 2 && B::COLS <= 256)) ? 2 : 1;
#else
  static constexpr int CTAS_PER_SM = SOLO ? 2 : 1;
#endif
  // weight-gradient accumulators live in tensor memory, whose accumulate truncates: bound the chain to ~320 accumulations
  static constexpr int MMAS_PER_WG = B::WG3 ? 48 : 16;
  static constexpr int FLUSH = (320 / (MD::EVALS * MMAS_PER_WG)) > 1 ? 320 / (MD::EVALS * MMAS_PER_WG) : 1;
  static constexpr bool WG3 = B::WG3;
  static constexpr size_t SMEM_BYTES = size_t(B::SMEM_BYTES) + 1024;      // + alignment slack
  static constexpr size_t W_BYTES = size_t(B::BWD_W) * 4;
};

template <class S>
__global__ void nm_tc_prep_bwd_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, float* __restrict__ wprep) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  tc::prep_tc_bwd<typename TcMode<S>::M>(params + TcMode<S>::mlp(plan).param_offset, wprep, gtid, gthreads);
}


// FUSED_ONLY: the instantiation for plans served by the class's COMPILED term structure -- no runtime-descriptor term code
// (pre-computed G rows, time-broadcast lanes, out-of-line evaluators) in the kernel at all.  The adjoint is sensitive to the
// mere presence of such code (profiles/r02c_barrier_ab.txt); the dispatcher picks the instantiation by args.fused_terms.
template <class S, bool FUSED_ONLY = false>
__global__ void __launch_bounds__(TcBKCfg<S>::THREADS, TcBKCfg<S>::CTAS_PER_SM) nm_tc_bwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ BwdArgs args) {
  using BK = TcBKCfg<S>;
  using MD = typename BK::MD;
  using M = typename BK::M;
  using C = typename BK::C;
  using B = typename BK::B;
  constexpr bool POLICY = MD::POLICY, SOLO = BK::SOLO;
  constexpr int NOW = BK::OWNER_WARPS;
  constexpr int NX = S::NX, NU = S::NU, NUA = NU > 0 ? NU : 1, K0 = C::K(0), N0 = C::N(0), NL = C::NL;
  constexpr int TILE = 128, NS = Tableau<S::INTEG>::NS, NEV = MD::EVALS, NI = B::NI, FLUSH = BK::FLUSH;
  constexpr int NLAST = C::N(NL - 1);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the swizzled tiles need a 1024-byte aligned base; the dynamic window starts right after the driver's 1 KB
  // (checked, not re-aligned: pointer arithmetic on an integer would turn every access into a generic one)
  uint8_t* sm = smem_raw;
  if ((smem_u32(sm) & 1023u) != 0u) __trap();
  const float* Wsm = reinterpret_cast<const float*>(sm + B::SMALL_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + B::BAR_OFF);
  // bars: 0 small-weights TMA, 1 ops_full (128 arrivals), 2 mma done, 3 wgrad done, 4/5 ring slot full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* coef = reinterpret_cast<float*>(bars + 10);                       // [NM_MAX_TERMS]
  VarTable& vt = *reinterpret_cast<VarTable*>(bars + 20);
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  const VarTable vtc = make_vartable_regs(p, args, S::NX, S::NU, S::NY);
  const int T = p.nsteps;
  constexpr bool FO = FUSED_ONLY && S::NTERMS > 0;
  const bool fused = FO || (S::NTERMS > 0 && args.fused_terms);
  const bool pre = !FO && !fused && args.G_x != nullptr;
  const int g_row = (T + 1) * NX + T * (NU + S::NY), g_yoff = (T + 1) * NX + T * NU;
  int tb_mask = 0;
  if constexpr (!FO) {
    for (int k = 0; k < p.n_terms; ++k) {
      const NmTerm& tm = p.terms[k];
      const NmAtom* at[4] = {&tm.lhs.a, &tm.lhs.b, &tm.rhs.a, &tm.rhs.b};
      const bool live[4] = {true, tm.lhs.op != NM_OP_NONE, true, tm.rhs.op != NM_OP_NONE};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (live[i] && at[i]->var >= 0 && at[i]->var <= NM_VAR_Y && atom_is_tbcast(*at[i], tm)) tb_mask |= 1 << at[i]->var;
    }
  }
  if (tid < NM_MAX_TERMS) coef[tid] = tid < p.n_terms ? term_coef(p, tid, args.grad_scalars, args.b_total) : 0.f;
  float* gblock = args.grad_partials + size_t(blockIdx.x) * size_t(p.n_params > 0 ? p.n_params : 1);
  for (int i = tid; i < p.n_params; i += BK::THREADS) gblock[i] = 0.f;
  float* grhs = gblock + MD::mlp(p).param_offset;
  if (tid == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = T + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = T; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = T; vt.D[NM_VAR_Y] = S::NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
    mbar_init(&bars[0], 1); mbar_init(&bars[1], NOW * 32); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1); mbar_init(&bars[5], 1);
    mbar_fence_init();
  }
  // constant rows of the transposed tiles: zero everything, then the constant-one (bias) row of every AT_l
  for (int i = tid; i < B::RING_OFF / 4; i += BK::THREADS) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  if (C::BIAS && tid < 128) {
    static_for<0, NL>([&](auto lc) {
      constexpr int l = decltype(lc)::value;
      *reinterpret_cast<float*>(sm + B::at_off(l) + tc::sw128_off(C::K(l), tid, B::ra(l))) = 1.f;
    });
  }
  if (warp == 0) tc::tmem_alloc<B::COLS>(tmem_slot);
  tc::fence_before();
  tc::fence_async_smem();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], uint32_t(B::SMALL * 4));
    tma_bulk_g2s(sm + B::SMALL_OFF, args.wprep, uint32_t(B::SMALL * 4), &bars[0]);
  }
  mbar_wait(&bars[0], 0);

  // number of MLP evaluations this CTA runs (both roles count the same way)
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) ++my_tiles;

  if (warp >= NOW) {
   if (warp == NOW) {
    // =============================================================== control thread: TMA producer + MMA issuer
    {
      const uint32_t sbase = smem_u32(sm);
      const uint64_t d_sw0 = tc::smem_desc_sw128(sbase);
      const long long total_imgs = (long long)my_tiles * T * NEV * NI;
      long long img_issued = 0;
      uint32_t ph_ops = 0, ph_slot[2] = {0, 0};
      TCP_DECL(c_wait_ops); TCP_DECL(c_wait_slot); TCP_DECL(c_issue); TCP_DECL(c_total);
#ifdef NM_TUNE_TCPROF
      const long long c_begin = clock64();
#endif
      auto issue_img = [&](int i) {                     // image i of the per-evaluation sequence -> ring slot i % 2
        static_for<0, NI>([&](auto ic) {
          constexpr int I = decltype(ic)::value;
          if (i == I) {
            constexpr uint32_t bytes = uint32_t(B::img_pair_floats(I)) * 4;
            if (tc::elect_one()) {
              mbar_expect_tx(&bars[4 + (I & 1)], bytes);
              tma_bulk_g2s(sm + B::RING_OFF + (I & 1) * B::slot_floats() * 4, args.wprep + B::gimg_off(I), bytes, &bars[4 + (I & 1)]);
            }
            __syncwarp();
          }
        });
        ++img_issued;
      };
      if (total_imgs > 0) issue_img(0);
      int step_cnt = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int t = T - 1; t >= 0; --t) {
          bool fresh = (step_cnt % FLUSH) == 0;         // accumulators were flushed before this step
          ++step_cnt;
          for (int st = NEV - 1; st >= 0; --st) {
            // one wgrad GEMM: D'[n][k] (+)= sum_m DT[n][m] AT_l[k][m]
            auto wgrad = [&](auto lc, auto dtoff_c, auto dtrows_c) {
              constexpr int l = decltype(lc)::value;
              constexpr uint32_t dt_off = decltype(dtoff_c)::value, dt_rows = decltype(dtrows_c)::value;
              constexpr uint32_t idesc = tc::idesc_tf32(128, B::wgc(l));
              const uint64_t da = tc::desc_add(d_sw0, dt_off), db = tc::desc_add(d_sw0, B::at_off(l));
              if (tc::elect_one()) {
                if constexpr (B::WG3) {
                  // remainder products first (see nm_tc.cuh: the accumulator truncates)
                  static_for<0, 16>([&](auto sc) {
                    constexpr uint32_t s = decltype(sc)::value;
                    constexpr uint32_t ao = (s >> 2) * (dt_rows * 128) + (s & 3) * 32, bo = (s >> 2) * (B::ra(l) * 128) + (s & 3) * 32;
                    tc::mma_ss(tmem + B::wg_off(l), tc::desc_add(da, ao + B::LO_OFF), tc::desc_add(db, bo), idesc, (s == 0 && fresh) ? 0u : 1u);
                    tc::mma_ss(tmem + B::wg_off(l), tc::desc_add(da, ao), tc::desc_add(db, bo + B::LO_OFF), idesc, 1u);
                  });
                  static_for<0, 16>([&](auto sc) {
                    constexpr uint32_t s = decltype(sc)::value;
                    tc::mma_ss(tmem + B::wg_off(l), tc::desc_add(da, (s >> 2) * (dt_rows * 128) + (s & 3) * 32),
                               tc::desc_add(db, (s >> 2) * (B::ra(l) * 128) + (s & 3) * 32), idesc, 1u);
                  });
                } else {
                  tc::mma_ss(tmem + B::wg_off(l), da, db, idesc, fresh ? 0u : 1u);
                  static_for<1, 16>([&](auto sc) {
                    constexpr uint32_t s = decltype(sc)::value;
                    tc::mma_ss(tmem + B::wg_off(l), tc::desc_add(da, (s >> 2) * (dt_rows * 128) + (s & 3) * 32),
                               tc::desc_add(db, (s >> 2) * (B::ra(l) * 128) + (s & 3) * 32), idesc, 1u);
                  });
                }
              }
              __syncwarp();
            };
            // ---- forward hidden layers
            static_for<1, NL - 1>([&](auto lc) {
              constexpr int l = decltype(lc)::value;
              constexpr int I = l - 1;
              { TCP_T0(); mbar_wait(&bars[1], ph_ops); ph_ops ^= 1; TCP_ADD(c_wait_ops); }
              tc::fence_after();
              if (img_issued < total_imgs) issue_img((I + 1) % NI);
              { TCP_T0(); mbar_wait(&bars[4 + (I & 1)], ph_slot[I & 1]); ph_slot[I & 1] ^= 1; TCP_ADD(c_wait_slot); }
              TCP_T0();
              constexpr int KP = C::kp(l), NP = C::np(l);
              constexpr uint32_t idesc = tc::idesc_tf32(128, NP);
              constexpr uint32_t LBO = NP * 16, SBO = 128, KSTEP = 2 * NP * 16;
              const uint64_t wh = tc::smem_desc(sbase + B::RING_OFF + (I & 1) * B::slot_floats() * 4, LBO, SBO);
              const uint64_t wl = tc::desc_add(wh, C::img_floats(l) * 4);
              if (tc::elect_one()) {
                static_for<0, KP / 8>([&](auto kc) {
                  constexpr int ks = decltype(kc)::value;
                  tc::mma_ts_c<(ks > 0)>(tmem + B::ACC, tmem + B::OPL + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
                  tc::mma_ts_c<true>(tmem + B::ACC, tmem + B::OPH + ks * 8, tc::desc_add(wl, ks * KSTEP), idesc);
                });
                static_for<0, KP / 8>([&](auto kc) {
                  constexpr int ks = decltype(kc)::value;
                  tc::mma_ts_c<true>(tmem + B::ACC, tmem + B::OPH + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
                });
                tc::commit(&bars[2]);
              }
              __syncwarp();
              TCP_ADD(c_issue);
            });
            // ---- backward hidden layers
            static_for_rev<1, NL - 1>([&](auto lc) {
              constexpr int l = decltype(lc)::value;
              constexpr int I = (NL - 2) + (NL - 2 - l);
              { TCP_T0(); mbar_wait(&bars[1], ph_ops); ph_ops ^= 1; TCP_ADD(c_wait_ops); }
              tc::fence_after();
              if (img_issued < total_imgs) issue_img((I + 1) % NI);
              { TCP_T0(); mbar_wait(&bars[4 + (I & 1)], ph_slot[I & 1]); ph_slot[I & 1] ^= 1; TCP_ADD(c_wait_slot); }
              TCP_T0();
              constexpr int KPd = B::dkp(l), NPd = B::dnp(l);
              constexpr uint32_t idesc = tc::idesc_tf32(128, NPd);
              constexpr uint32_t LBO = NPd * 16, SBO = 128, KSTEP = 2 * NPd * 16;
              const uint64_t wh = tc::smem_desc(sbase + B::RING_OFF + (I & 1) * B::slot_floats() * 4, LBO, SBO);
              const uint64_t wl = tc::desc_add(wh, B::dimg_floats(l) * 4);
              if (tc::elect_one()) {
                static_for<0, KPd / 8>([&](auto kc) {
                  constexpr int ks = decltype(kc)::value;
                  tc::mma_ts_c<(ks > 0)>(tmem + B::ACC, tmem + B::OPL + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
                  tc::mma_ts_c<true>(tmem + B::ACC, tmem + B::OPH + ks * 8, tc::desc_add(wl, ks * KSTEP), idesc);
                });
                static_for<0, KPd / 8>([&](auto kc) {
                  constexpr int ks = decltype(kc)::value;
                  tc::mma_ts_c<true>(tmem + B::ACC, tmem + B::OPH + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
                });
                tc::commit(&bars[2]);
              }
              __syncwarp();
              if constexpr (l == NL - 2) wgrad(IC<NL - 1>{}, IC<B::DTL_OFF>{}, IC<8>{});
              wgrad(IC<l>{}, IC<B::DT_OFF>{}, IC<B::RD>{});
              if (tc::elect_one()) tc::commit(&bars[3]);
              __syncwarp();
              TCP_ADD(c_issue);
            });
            // ---- weight gradient of layer 0
            { TCP_T0(); mbar_wait(&bars[1], ph_ops); ph_ops ^= 1; TCP_ADD(c_wait_ops); }
            tc::fence_after();
            { TCP_T0(); wgrad(IC<0>{}, IC<B::DT_OFF>{}, IC<B::RD>{});
            if (tc::elect_one()) tc::commit(&bars[3]);
            __syncwarp(); TCP_ADD(c_issue); }
            fresh = false;
          }
        }
      }
#ifdef NM_TUNE_TCPROF
      c_total = clock64() - c_begin;
      if (blockIdx.x == 0 && tid == NOW * 32) printf("[tcprof control] total %lld  wait_ops %lld  wait_slot %lld  issue %lld  (stages %lld)\n", c_total, c_wait_ops, c_wait_slot, c_issue, (long long)my_tiles * T * NEV);
#endif
    }
   }
  } else {
    // =============================================================== owners
    // two threads per trajectory: thread (m, hf) with m = tid & 127 owns TMEM lane m and the 16-column blocks of
    // every layer whose index has parity hf (two warps per scheduler instead of one: the epilogues are
    // latency-bound); per-trajectory scalars (state, stage derivatives, co-state) are kept by both
    const int m_ = tid & 127, hf = SOLO ? 0 : (tid >> 7);
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t t_oph = tmem + lane_base + B::OPH, t_opl = tmem + lane_base + B::OPL, t_acc = tmem + lane_base + B::ACC;
    uint32_t ph_mma = 0, ph_wg = 0;
    bool wg_pending = false;
    TCP_DECL(o_wait_mma); TCP_DECL(o_wait_wg); TCP_DECL(o_arrive); TCP_DECL(o_flush);
    TCP_DECL(o_l0); TCP_DECL(o_fwd); TCP_DECL(o_last); TCP_DECL(o_bwd); TCP_DECL(o_l1); TCP_DECL(o_xch);
    TCP_DECL(o_dyn); TCP_DECL(o_tu); TCP_DECL(o_pin); TCP_DECL(o_tx);
#ifdef NM_TUNE_TCPROF
    const long long o_begin = clock64();
#endif
    int step_cnt = 0;
    const float act_prm = MD::mlp(p).act_param;
    const float* __restrict__ W0 = Wsm + C::first_off();
    const float* __restrict__ WL = Wsm + C::last_off();
    float* xch = reinterpret_cast<float*>(sm + B::XCH_OFF);                 // [2][128][8] partial input gradients of the pair
    auto mine = [&](int c0) { return SOLO || ((c0 >> 4) & 1) == hf; };
    // per-thread pieces of the swizzled transposed-tile address: element (f, m) sits at
    //   block(m) * R*128 + f*128 + ((chunk ^ (f & 7)) << 4) + (m & 3)*4
    const uint32_t t_blk = uint32_t(m_ >> 5), t_chunk = uint32_t((m_ & 31) >> 2), t_in = uint32_t(m_ & 3) * 4;
    const uint32_t sm_u32 = smem_u32(sm);
    uint32_t t_x[8];                                                       // (chunk ^ r) << 4 for the 8 row residues
#pragma unroll
    for (int r = 0; r < 8; ++r) t_x[r] = sm_u32 + ((t_chunk ^ uint32_t(r)) << 4) + t_in;
    auto tstore = [&](uint32_t tile_off, int R, int f, float v_hi, float v_lo) {
      const uint32_t a = t_x[f & 7] + tile_off + t_blk * uint32_t(R * 128) + uint32_t(f * 128);
      tc::st_shared(a, v_hi);
      if constexpr (B::WG3) tc::st_shared(a + uint32_t(B::LO_OFF), v_lo);
    };
    auto arrive_ops = [&]() {
      TCP_T0();
      tc::wait_st();
      tc::fence_async_smem();
      tc::fence_before();
      mbar_arrive(&bars[1]);
      TCP_ADD(o_arrive);
    };
    auto wait_wg = [&]() { TCP_T0(); mbar_wait(&bars[3], ph_wg); ph_wg ^= 1; TCP_ADD(o_wait_wg); };
    // add this CTA's TMEM weight-gradient accumulators to its gradient block (fp32, round to nearest)
    auto flush = [&]() {
      static_for<0, NL>([&](auto lc) {
        constexpr int l = decltype(lc)::value;
        constexpr int Kl = C::K(l), Nl = C::N(l), KA = C::kaug(l), CH = 8;
        int off = 0;
        for (int i = 0; i < l; ++i) off += C::N(i) * C::KW(i) + (C::BIAS ? C::N(i) : 0);
        float* dst = grhs + off;
        // tcgen05.ld is warp-collective (.sync.aligned): every owner of the half executes it, rows n >= N(l) are dropped
#pragma unroll 1
        for (int c0 = 0; c0 < KA; c0 += CH) {
          if (!mine(c0)) continue;
          float v[CH];
          tc::tmem_ld<CH>(tmem + lane_base + B::wg_off(l) + c0, v);
          if (m_ < Nl) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int k = c0 + i;
              if (k < Kl) atomicAdd(dst + m_ * C::KW(l) + k, v[i]);        // result unused: RED, no round trip
              else if (k == Kl && C::BIAS) atomicAdd(dst + Nl * C::KW(l) + m_, v[i]);
            }
          }
        }
      });
    };
    // derivative of hidden activation a_j at column n (this thread's columns only)
    uint32_t masks[NL][4];
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) masks[i][j] = 0u;
    auto mask_set = [&](auto jc, int n, float a) {
      constexpr int j = decltype(jc)::value;
      masks[j][n >> 5] = (masks[j][n >> 5] & ~(1u << (n & 31))) | (a > 0.f ? (1u << (n & 31)) : 0u);
    };
    auto mask_d = [&](auto jc, int n) -> float {
      constexpr int j = decltype(jc)::value;
      if constexpr (M::ACT == NM_ACT_IDENTITY) return 1.f;
      return ((masks[j][n >> 5] >> (n & 31)) & 1u) ? 1.f : (M::ACT == NM_ACT_LEAKY ? act_prm : 0.f);
    };

    for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
      const int64_t b_raw = int64_t(tile) * TILE + m_;
      const bool valid = b_raw < args.B;
      const int64_t b = valid ? b_raw : args.B - 1;
      float lam[NX], x_pre[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) { lam[i] = 0.f; x_pre[i] = __ldg(args.x_traj + (b * (T + 1) + (T - 1)) * NX + i); }
      // policy mode: the raw policy output (checkpoint) and the exogenous policy inputs are fetched one step ahead too
      float o_pre[POLICY ? NUA : 1], in_pre[POLICY ? K0 : 1];
      if constexpr (POLICY) {
#pragma unroll
        for (int j = 0; j < NU; ++j) o_pre[j] = __ldg(MD::o_base(args.chk, args.B, T) + (b * T + (T - 1)) * NU + j);
        tc_policy_input<S, K0>(p, args.exo, b, T - 1, x_pre, in_pre);
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
        // flush point (mirrored by the control warp's `fresh`)
        if (step_cnt > 0 && (step_cnt % FLUSH) == 0) {
          if (wg_pending) { wait_wg(); wg_pending = false; tc::fence_after(); }
          { TCP_T0(); flush(); TCP_ADD(o_flush); }
        }
        ++step_cnt;
        constexpr int NYA = S::NY > 0 ? S::NY : 1;
        float x[NX], u[NUA], gx[NX], gtx[NX], yv[NYA], gty[NYA];
        float o_cur[POLICY ? NUA : 1], in_cur[POLICY ? K0 : 1];
#pragma unroll
        for (int i = 0; i < NX; ++i) { x[i] = x_pre[i]; gtx[i] = 0.f; }
#pragma unroll
        for (int j = 0; j < NYA; ++j) { yv[j] = 0.f; gty[j] = 0.f; }
        if constexpr (POLICY) {
#pragma unroll
          for (int j = 0; j < NU; ++j) o_cur[j] = o_pre[j];
#pragma unroll
          for (int q = 0; q < K0; ++q) in_cur[q] = in_pre[q];
          tc_output<S>(p, x, yv);
          tc_policy_state<S, K0>(x, yv, in_cur);
        }
        if (t > 0) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x_pre[i] = __ldg(args.x_traj + (b * (T + 1) + t - 1) * NX + i);
          if constexpr (POLICY) {
#pragma unroll
            for (int j = 0; j < NU; ++j) o_pre[j] = __ldg(MD::o_base(args.chk, args.B, T) + (b * T + t - 1) * NU + j);
            tc_policy_input<S, K0>(p, args.exo, b, t - 1, x, in_pre);     // (state entries are placeholders)
          }
        }
#pragma unroll
        for (int j = 0; j < NUA; ++j) u[j] = 0.f;
        if constexpr (!POLICY && NU > 0 && S::ACT_EXO >= 0) {
          const float* src = args.exo[S::ACT_EXO] + (b * p.exo[S::ACT_EXO].steps + t) * NU;
#pragma unroll
          for (int j = 0; j < NU; ++j) u[j] = __ldg(src + j);
        }
        const float h = p.h;
        float k[POLICY ? 1 : NM_MAX_STAGES][NX], gk[POLICY ? 1 : NM_MAX_STAGES][NX];
        float pol_cot[NLAST];                                    // policy mode: cotangent of the raw policy output o_t
#pragma unroll
        for (int j = 0; j < NLAST; ++j) pol_cot[j] = 0.f;
        // policy mode: dynamics adjoint and the loss-term gradient of u_t (-> co-state part gx and the cotangent of the raw
        // policy output).  Needs nothing from the step's MLP evaluation: runs while the first MMA batch is in flight.
        auto policy_dyn = [&](auto) {                           // generic: only instantiated in policy mode
          float o[NUA], gu[NUA];
          { TCP_T0();
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            o[j] = o_cur[j];
            u[j] = bounds_fwd<M::BOUNDS>(o[j], p.policy.bmin[j], p.policy.bmax[j]);
          }
          Dyn<S>::step_vjp(p, x, u, lam, gx, gu);
          TCP_ADD(o_dyn); }
          TCP_T0();
          if (valid) {
            if (pre) {
#pragma unroll
              for (int j = 0; j < NU; ++j) gu[j] += __ldg(args.G_x + b * g_row + (T + 1) * NX + t * NU + j);
              if (tb_mask & 2) term_grad_accum_ool<NUA, 2>(p, vt, coef, b, NM_VAR_U, t, NU, gu);
            } else {
              if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_U, NU, NUA>(p, vtc, coef, b, t, gu); }
              else term_grad_accum_ool<NUA, 0>(p, vt, coef, b, NM_VAR_U, t, NU, gu);
              if (args.gu_ext != nullptr) {
#pragma unroll
                for (int j = 0; j < NU; ++j) gu[j] += __ldg(args.gu_ext + (b * T + t) * NU + j);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < NU; ++j) pol_cot[j] = valid ? gu[j] * bounds_deriv<M::BOUNDS>(o[j], p.policy.bmin[j], p.policy.bmax[j]) : 0.f;
          TCP_ADD(o_tu);
        };
        if constexpr (POLICY) {
          (void)k; (void)gk;
        } else {
          static_for<0, NS>([&](auto sc) {
            constexpr int St = decltype(sc)::value;
#pragma unroll
            for (int i = 0; i < NX; ++i) { gk[St][i] = 0.f; k[St][i] = 0.f; }
            if constexpr (St + 1 < NS) {
              const float* cp = args.chk + ((b * T + t) * (NS - 1) + St) * NX;
#pragma unroll
              for (int i = 0; i < NX; ++i) k[St][i] = __ldg(cp + i);
            }
          });
#pragma unroll
          for (int i = 0; i < NX; ++i) gx[i] = lam[i];
        }
#pragma unroll 1
        for (int st = NEV - 1; st >= 0; --st) {
          float in[K0], w_cot[NLAST];
          if constexpr (POLICY) {
            TCP_T0();
#pragma unroll
            for (int q = 0; q < K0; ++q) in[q] = in_cur[q];
#pragma unroll
            for (int j = 0; j < NLAST; ++j) w_cot[j] = 0.f;
            TCP_ADD(o_pin);
          } else {
            float xs[NX];
            static_for<0, NS>([&](auto sc) {                       // only the tableau arithmetic is per-stage code
              constexpr int St = decltype(sc)::value;
              if (st == St) {
                stage_input<S::INTEG, St, NX>(h, x, k, xs);
                stage_cotangent<S::INTEG, St, NX>(h, lam, gk, w_cot);
              }
            });
#pragma unroll
            for (int i = 0; i < NX; ++i) { in[i] = xs[i]; w_cot[i] = valid ? w_cot[i] : 0.f; }
#pragma unroll
            for (int j = 0; j < NU; ++j) in[NX + j] = u[j];
          }
          // the previous evaluation's last weight-gradient GEMM still reads AT_0 / DT
          if (wg_pending) { wait_wg(); wg_pending = false; }
          // ---- layer 0 (CUDA cores): a_1, its derivative, operand of layer 1, transposed tiles AT_0 / AT_1
          if (hf == 0) {
#pragma unroll
            for (int i = 0; i < K0; ++i) { const float ih = tc::tf32_hi(in[i]); tstore(B::at_off(0), B::ra(0), i, ih, in[i] - ih); }
          }
          {
            TCP_T0();
            constexpr int KP1 = C::kp(1), CH = tc::chunk16_of(KP1);
            static_for<0, KP1 / CH>([&](auto cc) {
              constexpr int c0 = decltype(cc)::value * CH;
              if (mine(c0)) {
                float hi[CH], lo[CH], dv[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                  const int n = c0 + i;
                  float a = 0.f, d = 0.f;
                  if (n < N0) {
                    float w[C::R0];
                    ld_vec<C::R0>(w, W0 + n * C::R0);
                    float z = MD::HOIST ? __ldg(args.pre + (b * T + t) * int64_t(MD::N0) + n) : (C::BIAS ? w[K0] : 0.f);
#pragma unroll
                    for (int q = 0; q < K0; ++q) z = fmaf(w[q], in[q], z);
                    tc::act_ad<M::ACT>(z, act_prm, a, d);
                  } else if (n == N0 && C::BIAS) a = 1.f;
                  hi[i] = tc::tf32_hi(a);
                  lo[i] = a - hi[i];
                  dv[i] = d;
                  if (n < N0) {
                    tstore(B::at_off(1), B::ra(1), n, hi[i], lo[i]);
                    if constexpr (B::MASK) mask_set(IC<1>{}, n, a);
                  }
                }
                tc::tmem_st<CH>(t_oph + c0, hi);
                tc::tmem_st<CH>(t_opl + c0, lo);
                if constexpr (!B::MASK) {
                  if constexpr (c0 < B::der_w(1)) tc::tmem_st<CH>(tmem + lane_base + B::der_off(1) + c0, dv);
                }
                (void)dv;
              }
            });
            TCP_ADD(o_l0);
          }
          arrive_ops();
          // ---- hidden layers forward (tensor cores) and, after the last one, the output layer backward
          static_for<1, NL - 1>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            constexpr int Nl = C::N(l);
            if constexpr (l == 1) {
              // loss-term gradient w.r.t. x_t: depends on the stored trajectories only -- evaluated here, while the
              // step's first MMA batch is in flight
              if constexpr (POLICY) {
                policy_dyn(0);
#pragma unroll
                for (int j = 0; j < NLAST; ++j) w_cot[j] = pol_cot[j];
              }
              if (st == NEV - 1) {
                TCP_T0();
            if (valid) {
              if (pre) {
    #pragma unroll
                for (int i = 0; i < NX; ++i) gtx[i] += __ldg(args.G_x + b * g_row + t * NX + i);
                if (tb_mask & 1) term_grad_accum_ool<NX, 2>(p, vt, coef, b, NM_VAR_X, t, NX, gtx);
              } else {
                if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_X, NX, NX>(p, vtc, coef, b, t, gtx); }
                else term_grad_accum_ool<NX, 0>(p, vt, coef, b, NM_VAR_X, t, NX, gtx);
                if (args.gx_ext != nullptr) {
    #pragma unroll
                  for (int i = 0; i < NX; ++i) gtx[i] += __ldg(args.gx_ext + (b * (T + 1) + t) * NX + i);
                }
              }
            }
                if constexpr (S::NY > 0) {                    // loss-term gradient w.r.t. y_t (output node y = C x)
                  if (valid) {
                    if (pre) {
#pragma unroll
                      for (int j = 0; j < S::NY; ++j) gty[j] += __ldg(args.G_x + b * g_row + g_yoff + t * S::NY + j);
                      if (tb_mask & 4) term_grad_accum_ool<NYA, 2>(p, vt, coef, b, NM_VAR_Y, t, S::NY, gty);
                    } else {
                      if (fused) { if constexpr (S::NTERMS > 0) term_grad_s<S, NM_VAR_Y, S::NY, NYA>(p, vtc, coef, b, t, gty); }
                      else term_grad_accum_ool<NYA, 0>(p, vt, coef, b, NM_VAR_Y, t, S::NY, gty);
                      if (args.gy_ext != nullptr) {
#pragma unroll
                        for (int j = 0; j < S::NY; ++j) gty[j] += __ldg(args.gy_ext + (b * T + t) * S::NY + j);
                      }
                    }
                  }
                }
                TCP_ADD(o_tx);
              }
            }
            { TCP_T0(); mbar_wait(&bars[2], ph_mma); ph_mma ^= 1; TCP_ADD(o_wait_mma); }
            tc::fence_after();
            if constexpr (l < NL - 2) {
              TCP_T0();
              constexpr int KPN = C::kp(l + 1), CH = tc::chunk16_of(KPN);
              static_for<0, KPN / CH>([&](auto cc) {
                constexpr int c0 = decltype(cc)::value * CH;
                if (mine(c0)) {
                  float v[CH], hi[CH], lo[CH], dv[CH];
                  tc::tmem_ld<CH>(t_acc + c0, v);
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    float a = 0.f, d = 0.f;
                    if (n < Nl) tc::act_ad<M::ACT>(v[i], act_prm, a, d);
                    else if (n == Nl && C::BIAS) a = 1.f;
                    hi[i] = tc::tf32_hi(a);
                    lo[i] = a - hi[i];
                    dv[i] = d;
                    if (n < Nl) {
                      tstore(B::at_off(l + 1), B::ra(l + 1), n, hi[i], lo[i]);
                      if constexpr (B::MASK) mask_set(IC<l + 1>{}, n, a);
                    }
                  }
                  tc::tmem_st<CH>(t_oph + c0, hi);
                  tc::tmem_st<CH>(t_opl + c0, lo);
                  if constexpr (!B::MASK) {
                    if constexpr (c0 < B::der_w(l + 1)) tc::tmem_st<CH>(tmem + lane_base + B::der_off(l + 1) + c0, dv);
                  }
                  (void)dv;
                }
              });
              TCP_ADD(o_fwd);
              arrive_ops();
            } else {
              TCP_T0();
              // last hidden layer: a_{NL-1} (transposed tile only), then straight into the backward pass:
              // output layer on CUDA cores  dA[k] = sum_j w[j] WL[j][k],  dZ = dA * act'
              if (hf == 0) {
#pragma unroll
                for (int j = 0; j < NLAST; ++j) { const float wh_ = tc::tf32_hi(w_cot[j]); tstore(B::DTL_OFF, 8, j, wh_, w_cot[j] - wh_); }
              }
              constexpr int KPD = B::dkp(l), CH = tc::chunk16_of(KPD);      // operand width of the data-gradient GEMM of layer l
              static_for<0, KPD / CH>([&](auto cc) {
                constexpr int c0 = decltype(cc)::value * CH;
                if (mine(c0)) {
                  float v[CH], hi[CH], lo[CH];
                  float wl4[NLAST][4];
                  tc::tmem_ld<CH>(t_acc + c0, v);
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    float dz = 0.f;
                    if ((i & 3) == 0) {                                      // output-layer weights of 4 columns: one LDS.128 per row
#pragma unroll
                      for (int j = 0; j < NLAST; ++j) {
                        if (n < C::RL) ld_vec<4>(wl4[j], WL + j * C::RL + n);
                      }
                    }
                    if (n < Nl) {
                      float a, d;
                      tc::act_ad<M::ACT>(v[i], act_prm, a, d);
                      { const float ah_ = tc::tf32_hi(a); tstore(B::at_off(NL - 1), B::ra(NL - 1), n, ah_, a - ah_); }
                      float da = 0.f;
#pragma unroll
                      for (int j = 0; j < NLAST; ++j) da = fmaf(w_cot[j], wl4[j][i & 3], da);
                      dz = da * d;
                    }
                    hi[i] = tc::tf32_hi(dz);
                    lo[i] = dz - hi[i];
                    if (n < Nl) tstore(B::DT_OFF, B::RD, n, hi[i], lo[i]);
                  }
                  tc::tmem_st<CH>(t_oph + c0, hi);
                  tc::tmem_st<CH>(t_opl + c0, lo);
                }
              });
              TCP_ADD(o_last);
              arrive_ops();
            }
          });
          // ---- hidden layers backward
          float d_in[K0];
#pragma unroll
          for (int i = 0; i < K0; ++i) d_in[i] = 0.f;
          static_for_rev<1, NL - 1>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            constexpr int Kl = C::K(l);                                    // width of dA_l = hidden width of a_l
            { TCP_T0(); mbar_wait(&bars[2], ph_mma); ph_mma ^= 1; TCP_ADD(o_wait_mma); }
            tc::fence_after();
            if constexpr (l > 1) {
              TCP_T0();
              constexpr int KPD = B::dkp(l - 1), CH = tc::chunk16_of(KPD), NCH = KPD / CH;
              // this thread's dZ_l columns: operand (tensor memory) first, the transposed copy only after the
              // weight-gradient GEMMs that still read DT have completed
              float keep_hi[own_cnt(NCH, CH, SOLO)][CH], keep_lo[B::WG3 ? own_cnt(NCH, CH, SOLO) : 1][CH];
              static_for<0, NCH>([&](auto cc) {
                constexpr int ci = decltype(cc)::value, c0 = ci * CH;
                if (mine(c0)) {
                  float v[CH], hi[CH], lo[CH], dv[CH];
                  tc::tmem_ld<CH>(t_acc + c0, v);
                  if constexpr (!B::MASK) tc::tmem_ld<CH>(tmem + lane_base + B::der_off(l) + c0, dv);
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    float dz = 0.f;
                    if (n < Kl) {
                      float d;
                      if constexpr (B::MASK) d = mask_d(IC<l>{}, n); else d = dv[i];
                      dz = v[i] * d;
                    }
                    hi[i] = tc::tf32_hi(dz);
                    lo[i] = dz - hi[i];
                    keep_hi[own_idx(ci, CH, SOLO)][i] = hi[i];
                    if constexpr (B::WG3) keep_lo[own_idx(ci, CH, SOLO)][i] = lo[i];
                  }
                  tc::tmem_st<CH>(t_oph + c0, hi);
                  tc::tmem_st<CH>(t_opl + c0, lo);
                }
              });
              TCP_ADD(o_bwd);
              wait_wg();                                                   // weight-gradient GEMMs reading DT are done
              static_for<0, NCH>([&](auto cc) {
                constexpr int ci = decltype(cc)::value, c0 = ci * CH;
                if (mine(c0)) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    if (n < Kl) tstore(B::DT_OFF, B::RD, n, keep_hi[own_idx(ci, CH, SOLO)][i], B::WG3 ? keep_lo[own_idx(ci, CH, SOLO)][i] : 0.f);
                  }
                }
              });
              arrive_ops();
            } else {
              // l == 1: dZ_1, then layer 0 backward on CUDA cores
              TCP_T0();
              constexpr int CH = tc::chunk16_of(round_up(Kl, 8)), NCH = round_up(Kl, 8) / CH;
              float keep_dz[own_cnt(NCH, CH, SOLO)][CH];
              static_for<0, NCH>([&](auto cc) {
                constexpr int ci = decltype(cc)::value, c0 = ci * CH;
                if (mine(c0)) {
                  float v[CH], dv[CH];
                  tc::tmem_ld<CH>(t_acc + c0, v);
                  if constexpr (!B::MASK) tc::tmem_ld<CH>(tmem + lane_base + B::der_off(1) + c0, dv);
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    float dz = 0.f;
                    if (n < Kl) {
                      float d;
                      if constexpr (B::MASK) d = mask_d(IC<1>{}, n); else d = dv[i];
                      dz = v[i] * d;
                      float w[C::R0];
                      ld_vec<C::R0>(w, W0 + n * C::R0);
#pragma unroll
                      for (int q = 0; q < K0; ++q) d_in[q] = fmaf(dz, w[q], d_in[q]);
                    }
                    keep_dz[own_idx(ci, CH, SOLO)][i] = dz;
                  }
                  (void)dv;
                }
              });
              TCP_ADD(o_l1);
              wait_wg();                                                   // weight-gradient GEMMs reading DT are done
              static_for<0, NCH>([&](auto cc) {
                constexpr int ci = decltype(cc)::value, c0 = ci * CH;
                if (mine(c0)) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const int n = c0 + i;
                    if (n < Kl) {
                      const float dz = keep_dz[own_idx(ci, CH, SOLO)][i]; const float dh_ = tc::tf32_hi(dz); tstore(B::DT_OFF, B::RD, n, dh_, dz - dh_);
                      if constexpr (MD::HOIST) { if (valid) args.dz1[(b * T + t) * int64_t(MD::N0) + n] = dz; }
                    }
                  }
                }
              });
              arrive_ops();
              wg_pending = true;
            }
          });
          // the pair's partial input gradients meet in shared memory
          TCP_T0();
          float gin[K0];
          if constexpr (SOLO) {
#pragma unroll
            for (int q = 0; q < K0; ++q) gin[q] = d_in[q];
          } else {
#pragma unroll
            for (int q = 0; q < K0; ++q) xch[(hf * 128 + m_) * 8 + q] = d_in[q];
            named_bar_sync64(1 + (warp & 3));
#pragma unroll
            for (int q = 0; q < K0; ++q) gin[q] = xch[m_ * 8 + q] + xch[(128 + m_) * 8 + q];
          }
          if constexpr (POLICY) {
            // gradient of the policy input: the state segment feeds back into the co-state
            TCP_ADD(o_xch);
            int off = 0;
            static_for<0, S::NSEG>([&](auto sc) {
              constexpr int sg = decltype(sc)::value;
              if constexpr (S::seg_kind(sg) == NM_SEG_STATE) {
#pragma unroll
                for (int i = 0; i < NX; ++i) gx[i] += gin[off + i];
              } else if constexpr (S::seg_kind(sg) == NM_SEG_OUTPUT) {
#pragma unroll
                for (int i = 0; i < S::NY; ++i) gty[i] += gin[off + i];
              }
              off += S::seg_width(sg);
            });
          } else {
            float gxs[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) gxs[i] = gin[i];
            TCP_ADD(o_xch);
            static_for<0, NS>([&](auto sc) {
              constexpr int St = decltype(sc)::value;
              if (st == St) stage_push<S::INTEG, St, NX>(h, gxs, gx, gk);
            });
          }
        }
        // ---- co-state of step t: dynamics / policy part + the loss-term gradient computed in the MMA wait window
        if constexpr (S::HAS_OUTPUT) {                                     // y_t = C x_t
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            float sC = 0.f;
#pragma unroll
            for (int j = 0; j < S::NY; ++j) sC = fmaf(p.C[j * NX + i], gty[j], sC);
            gx[i] += sC;
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = valid ? gx[i] + gtx[i] : 0.f;
      }
      if (valid && hf == 0 && args.grad_x0 != nullptr) {
#pragma unroll
        for (int i = 0; i < NX; ++i) args.grad_x0[b * NX + i] = lam[i];
      }
    }
    if (step_cnt > 0) {
      if (wg_pending) { wait_wg(); wg_pending = false; tc::fence_after(); }
      flush();
    }
#ifdef NM_TUNE_TCPROF
    if (blockIdx.x == 0 && tid == 0)
      printf("[tcprof owner0] total %lld  wait_mma %lld  wait_wg %lld  arrive %lld  flush %lld | l0 %lld fwd %lld last %lld bwd %lld l1 %lld xch %lld | dyn %lld terms_u %lld pol_in %lld terms_x %lld\n", clock64() - o_begin, o_wait_mma, o_wait_wg, o_arrive, o_flush, o_l0, o_fwd, o_last, o_bwd, o_l1, o_xch, o_dyn, o_tu, o_pin, o_tx);
#endif
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<B::COLS>(tmem);
}

// ================================================================================================
// First-layer hoist GEMM on the tensor cores:  E[r][n] = bias[n] + sum_k exo_cat[r][k] * W_0[n][KS + k]
// (rows r = b * T + t over the whole batch and horizon; see HoistedMlp in nm_kernels.cuh).  A plain GEMM that streams
// the exogenous rows once: per exogenous segment the 128 row-owning threads copy their row slice into a K-major
// interleaved operand tile  A[k/4][row][k%4]  (raw fp32 = the tf32 "hi" part once the tensor core has truncated it)
// and its exact remainder tile, one elected lane issues the 3-product MMAs (remainder products first) into a
// 32-column accumulator, and the epilogue writes the row back with the bias added.
// exogenous segments in plan order: the first one, and the one after segment s (-1: none)
template <class S> constexpr int exo_first_seg() { for (int i = 0; i < S::NSEG; ++i) if (S::seg_kind(i) == NM_SEG_EXO) return i; return -1; }
template <class S> constexpr int exo_next_seg(int s) { for (int i = s + 1; i < S::NSEG; ++i) if (S::seg_kind(i) == NM_SEG_EXO) return i; return -1; }
// this thread's row slice of exogenous segment s -> registers (all loads in flight at once; the row slice is 4 * PF bytes)
template <class S, int s, int PF>
__device__ __forceinline__ void exo_row_prefetch(const NmPlan& p, const ExoPtrs& exo, int64_t b, int t, float4 (&pre)[PF]) {
  constexpr int W = S::seg_width(s), e = S::seg_index(s);
  const float* __restrict__ src = exo.p[e] + (b * p.exo[e].steps + t) * W;
#pragma unroll
  for (int k4 = 0; k4 < W / 4; ++k4) pre[k4] = __ldg(reinterpret_cast<const float4*>(src + k4 * 4));
}

template <class S>
struct ExoTcCfg {
  using KC = KCfg<S>;
  static constexpr int N0 = KC::N0, NP = round_up(N0, 16), KE = KC::KE;
  static constexpr int wmax() { return round_up(exo_seg_wmax<S>(), 8); }
  static constexpr int IMG_FLOATS = round_up(KE, 8) * NP;                 // one weight image [k/4][NP][4]
  static constexpr int W_FLOATS = 2 * IMG_FLOATS + round_up(NP, 32);      // hi | lo | bias
  static constexpr int A_FLOATS = wmax() * 128;                           // one operand tile [k/4][128][4]
  static constexpr int NRAW = 3, RAW_FLOATS = wmax() * 128;               // ring of raw 128-row slices as they lie in global memory
  static constexpr int nsx() { int n = 0; for (int i = 0; i < S::NSEG; ++i) if (S::seg_kind(i) == NM_SEG_EXO) ++n; return n > 0 ? n : 1; }
  static constexpr int ACC_COLS = tc::pow2_ceil(2 * NP) < 32 ? 32 : tc::pow2_ceil(2 * NP);   // two accumulators (alternating tiles)
  static constexpr size_t SMEM_BYTES = size_t(W_FLOATS + 4 * A_FLOATS + NRAW * RAW_FLOATS) * 4 + 256;
  static constexpr bool OK = KC::HOIST && NP <= 256 && SMEM_BYTES <= 227 * 1024 && KE % 8 == 0;
};

template <class S>
__global__ void nm_exo_tc_prep_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, float* __restrict__ wexo) {
  using X = ExoTcCfg<S>;
  using KC = KCfg<S>;
  constexpr int KW = S::Pol::sz(0), KS = KC::KS;
  const float* __restrict__ Wg = params + plan.policy.param_offset;
  const float* __restrict__ bg = Wg + X::N0 * KW;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  for (int idx = gtid; idx < X::IMG_FLOATS; idx += gthreads) {
    const int kc = idx / (X::NP * 4), n = (idx / 4) % X::NP, k = kc * 4 + (idx & 3);
    const float v = (n < X::N0 && k < X::KE) ? Wg[n * KW + KS + k] : 0.f;
    const float h = tc::tf32_hi(v);
    wexo[idx] = h; wexo[X::IMG_FLOATS + idx] = v - h;
  }
  for (int idx = gtid; idx < round_up(X::NP, 32); idx += gthreads) wexo[2 * X::IMG_FLOATS + idx] = (idx < X::N0 && plan.policy.has_bias) ? bg[idx] : 0.f;
}

// TMA-fed: the 128 x W row slices of a tile are CONTIGUOUS in global memory (rows r = b * T + t of a [B][T][W] tensor), so
// one elected thread streams them as 1-D bulk copies (cp.async.bulk, SASS UBLKCP) into a ring of raw tiles, NRAW items
// (tile, segment) ahead, completion on an mbarrier.  The 128 threads turn a raw tile into the K-major operand tile and
// its remainder tile (bank-conflict-free: thread = row, the k-chunk rotated by the row), double-buffered, so the MMAs of
// item i run while item i + 1 is converted; two accumulators alternate between tiles, the epilogue of a tile overlaps
// the next tile's first MMAs.  (The register-prefetch version before it sat at 40 % of the DRAM peak on long-scoreboard
// stalls with 8 warps per SM: profiles/r02c_c4_ncu_summary.txt.)
template <class S>
__global__ void __launch_bounds__(288, 1) nm_exo_gemm_tc_fwd_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ wexo,
                                                                    const __grid_constant__ ExoPtrs exo, float* __restrict__ E, int64_t B) {
  using X = ExoTcCfg<S>;
  constexpr int N0 = X::N0, NP = X::NP, NRAW = X::NRAW, NSX = X::nsx();
  extern __shared__ __align__(128) float smem[];
  float* sm_w = smem;                                           // hi | lo | bias
  float* a_hi = smem + X::W_FLOATS;                             // [2][A_FLOATS]
  float* a_lo = a_hi + 2 * X::A_FLOATS;                         // [2][A_FLOATS]
  float* raw = a_lo + 2 * X::A_FLOATS;                          // [NRAW][RAW_FLOATS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + NRAW * X::RAW_FLOATS);
  uint64_t* mma_done = bars + 1;                                // [2]  MMAs that read operand buffer 0 / 1 have completed (tcgen05.commit)
  uint64_t* a_ready = bars + 3;                                 // [2]  the 128 converters have written operand buffer 0 / 1 (and drained its raw tile)
  uint64_t* full = bars + 5;                                    // [NRAW]  raw tile landed (TMA transaction bytes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + NRAW);
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1);
    mbar_init(&a_ready[0], 128); mbar_init(&a_ready[1], 128);
    for (int i = 0; i < NRAW; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<X::ACC_COLS>(tmem_slot);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  const int T = p.nsteps;
  const int64_t R = B * T;
  const int64_t tiles = (R + 127) / 128;
  constexpr uint32_t idesc = tc::idesc_tf32(128, NP);

  if (warp == 8) {
    // ---------------------------------------------------------------------------------------- producer + MMA issuer
    // One lane streams the raw tiles (NRAW items ahead) and issues the MMAs of every item as soon as its operand tiles are
    // written; its state advances incrementally (no divisions in the loop).
    if (tc::elect_one()) {
      int segw[NSX], sege[NSX], segoff[NSX];
      {
        int n = 0;
        static_for<0, S::NSEG>([&](auto sc) {
          constexpr int sg = decltype(sc)::value;
          if constexpr (S::seg_kind(sg) == NM_SEG_EXO) { segw[n] = S::seg_width(sg); sege[n] = S::seg_index(sg); segoff[n] = exo_seg_off<S>(sg); ++n; }
        });
      }
      int64_t p_tile = blockIdx.x;
      int p_sx = 0, p_slot = 0;
      const int64_t stride_rows = int64_t(gridDim.x) * 128;
      const int64_t p_db = stride_rows / T;
      const int p_dt = int(stride_rows - p_db * T);
      int64_t p_b = (int64_t(blockIdx.x) * 128) / T;
      int p_t = int(int64_t(blockIdx.x) * 128 - p_b * T);
      auto issue_next = [&]() {
        if (p_tile < tiles) {
          const int W = segw[p_sx], e = sege[p_sx];
          float* dst = raw + p_slot * X::RAW_FLOATS;
          const int64_t left = R - p_tile * 128;
          const int rows = left < 128 ? int(left) : 128;
          mbar_expect_tx(&full[p_slot], uint32_t(rows) * uint32_t(W) * 4u);
          const int64_t steps_e = p.exo[e].steps;
          int rr = 0, tt = p_t;
          int64_t bb = p_b;
          while (rr < rows) {                                   // one copy per trajectory piece (rows of one trajectory are contiguous)
            const int n = (T - tt) < (rows - rr) ? (T - tt) : (rows - rr);
            tma_bulk_g2s(dst + rr * W, exo.p[e] + (bb * steps_e + tt) * W, uint32_t(n) * uint32_t(W) * 4u, &full[p_slot]);
            rr += n; ++bb; tt = 0;
          }
        }
        p_slot = p_slot + 1 == NRAW ? 0 : p_slot + 1;
        if (++p_sx == NSX) {
          p_sx = 0; p_tile += gridDim.x;
          p_b += p_db; p_t += p_dt;
          if (p_t >= T) { p_t -= T; ++p_b; }
        }
      };
      mbar_expect_tx(&bars[0], uint32_t(X::W_FLOATS) * 4);
      tma_bulk_g2s(sm_w, wexo, uint32_t(X::W_FLOATS) * 4, &bars[0]);
      for (int i = 0; i < NRAW; ++i) issue_next();
      mbar_wait(&bars[0], 0);
      constexpr uint32_t A_LBO = 128 * 16, A_KSTEP = 2 * 128 * 16, B_LBO = NP * 16, B_KSTEP = 2 * NP * 16;
      const uint64_t dh0 = tc::smem_desc(smem_u32(a_hi), A_LBO, 128), dl0 = tc::smem_desc(smem_u32(a_lo), A_LBO, 128);
      const uint64_t wbase = tc::smem_desc(smem_u32(sm_w), B_LBO, 128);
      int64_t it = 0;
      int tl = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
        const uint32_t acc = tmem + uint32_t((tl & 1) * NP);
        static_for<0, S::NSEG>([&](auto sc) {
          constexpr int s = decltype(sc)::value;
          if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
            constexpr int W = S::seg_width(s), OFF = exo_seg_off<S>(s), WP = round_up(W, 8);
            const int ab = int(it & 1);
            mbar_wait(&a_ready[ab], uint32_t(it >> 1) & 1u);    // operand tiles of this item written, its raw tile drained
            tc::fence_after();
            issue_next();                                       // item it + NRAW into the ring slot just drained
            const uint64_t dh = tc::desc_add(dh0, ab * X::A_FLOATS * 4), dl = tc::desc_add(dl0, ab * X::A_FLOATS * 4);
            const uint64_t wh = tc::desc_add(wbase, (OFF / 4) * NP * 16);
            const uint64_t wl = tc::desc_add(wh, X::IMG_FLOATS * 4);
            static_for<0, WP / 8>([&](auto kc) {                // remainder products first (the accumulator truncates)
              constexpr int ks = decltype(kc)::value;
              tc::mma_ss(acc, tc::desc_add(dl, ks * A_KSTEP), tc::desc_add(wh, ks * B_KSTEP), idesc, (OFF == 0 && ks == 0) ? 0u : 1u);
              tc::mma_ss(acc, tc::desc_add(dh, ks * A_KSTEP), tc::desc_add(wl, ks * B_KSTEP), idesc, 1u);
            });
            static_for<0, WP / 8>([&](auto kc) {
              constexpr int ks = decltype(kc)::value;
              tc::mma_ss(acc, tc::desc_add(dh, ks * A_KSTEP), tc::desc_add(wh, ks * B_KSTEP), idesc, 1u);
            });
            tc::commit(&mma_done[ab]);
            ++it;
          }
        });
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------- converters (thread = row) + epilogue
    // Two groups of 128 threads: group g converts the items with parity g into operand buffer g -- two items in flight on
    // the CUDA cores while the tensor pipe works on the previous ones; the group of a tile's last item writes its rows of E.
    const int grp = warp >> 2, row = tid & 127;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    mbar_wait(&bars[0], 0);                                     // the bias is read in the epilogue
    int64_t it = 0;
    int tl = 0, c_slot = 0;
    uint32_t c_phase = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tmem + uint32_t((tl & 1) * NP);
      int last_ab = 0;
      int64_t last_it = 0;
      static_for<0, S::NSEG>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
          constexpr int W = S::seg_width(s), WP = round_up(W, 8), NK4 = W / 4;
          const int st = c_slot, ab = int(it & 1);
          if (ab == grp) {
          mbar_wait(&full[st], c_phase);                        // the raw tile has landed
          if (it >= 2) { mbar_wait(&mma_done[ab], uint32_t((it >> 1) - 1) & 1u); tc::fence_after(); }   // operand buffer ab is free again
          const float* __restrict__ rw = raw + st * X::RAW_FLOATS + row * W;
          float* __restrict__ ah = a_hi + ab * X::A_FLOATS;
          float* __restrict__ al = a_lo + ab * X::A_FLOATS;
          int k4 = row % NK4;                                    // rotated start: conflict-free reads (row stride W floats) and writes
#pragma unroll
          for (int j = 0; j < NK4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(rw + 4 * k4);
            // the tensor core reads the hi tile truncated to tf32: the remainder is taken against the truncated value
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(ah + (k4 * 128 + row) * 4) = v;
            *reinterpret_cast<float4*>(al + (k4 * 128 + row) * 4) = l;
            k4 = k4 + 1 < NK4 ? k4 + 1 : 0;
          }
          if constexpr (WP > W) {                                // zero columns up to the MMA's K granularity
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(ah + (NK4 * 128 + row) * 4) = z;
            *reinterpret_cast<float4*>(al + (NK4 * 128 + row) * 4) = z;
          }
          tc::fence_async_smem();                                // generic-proxy writes -> visible to the tensor core / TMA (async proxy)
          tc::fence_before();
          mbar_arrive(&a_ready[ab]);
          }
          last_ab = ab; last_it = it;
          ++it;
          if (++c_slot == NRAW) { c_slot = 0; c_phase ^= 1u; }
        }
      });
      // epilogue: the tile's MMAs are done (they complete in order) -> this thread's accumulator row + bias -> E.  The next
      // tile accumulates into the OTHER accumulator; this one is overwritten two tiles later, behind this thread's
      // fence_before + arrive of the items in between.
      if (last_ab != grp) continue;
      mbar_wait(&mma_done[last_ab], uint32_t(last_it >> 1) & 1u);
      tc::fence_after();
      const int64_t r = tile * 128 + row;
      const bool valid = r < R;
      static_for<0, NP / 16>([&](auto cc) {
        constexpr int c0 = decltype(cc)::value * 16;
        float v[16];
        tc::tmem_ld<16>(acc + lane_base + c0, v);
        if (valid) {
#pragma unroll
          for (int i4 = 0; i4 < 16; i4 += 4) {
            if (c0 + i4 < N0) {
              const float4 bb = *reinterpret_cast<const float4*>(sm_w + 2 * X::IMG_FLOATS + c0 + i4);
              *reinterpret_cast<float4*>(E + r * N0 + c0 + i4) = make_float4(v[i4] + bb.x, v[i4 + 1] + bb.y, v[i4 + 2] + bb.z, v[i4 + 3] + bb.w);
            }
          }
        }
      });
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<X::ACC_COLS>(tmem);
}


// ================================================================================================
// First-layer hoist, backward GEMM on the tensor cores:  dW_0[n][KS + k] += sum_r dZ_1[r][n] * exo_cat[r][k]
// A contraction over ROWS (r = b * T + t), i.e. the weight-gradient form of nm_tc_bwd_kernel: both operands are
// "K-major in rows" -- transposed, tf32-rounded (round to nearest: unbiased, averages out over the rows) tiles
// DT[n][row] and XT[k][row] in the 128-byte swizzle, written by the row-owning threads; M = 128 accumulator lanes = n
// (the first N0 are real), N = segment width, 16 MMAs per 128-row tile and segment.  The tensor-memory accumulator
// truncates, so it is flushed (RED.ADD into this CTA's gradient block, fp32 round-to-nearest) every 20 tiles.
// Single-pass TF32: the dispatcher uses it from 2^20 rows up, the exact FFMA GEMM below that.
template <class S>
struct ExoTcBwdCfg {
  using KC = KCfg<S>;
  static constexpr int N0 = KC::N0, KE = KC::KE;
  static constexpr int RD = round_up(N0, 8);                                // rows of DT
  static constexpr int rx() { return round_up(exo_seg_wmax<S>(), 8); }      // rows of XT (widest segment)
  static constexpr int DT_BYTES = RD * 512, XT_BYTES = rx() * 512;
  // the MMA reads 128 rows of the A operand: the window has to cover them (rows >= RD are garbage lanes, dropped at the flush)
  static constexpr int A_SPAN = 3 * RD * 128 + 16 * 1024;
  static constexpr int TILE_BYTES = cmax(DT_BYTES + XT_BYTES, A_SPAN);
  static constexpr size_t SMEM_BYTES = size_t(TILE_BYTES) + 256 + 1024;
  static constexpr int acc_cols(int w) { return round_up(w, 16); }
  static constexpr int acc_off(int s) { int o = 0; for (int i = 0; i < s; ++i) if (S::seg_kind(i) == NM_SEG_EXO) o += acc_cols(S::seg_width(i)); return o; }
  static constexpr int COLS = tc::pow2_ceil(acc_off(S::NSEG));
  static constexpr int FLUSH_TILES = 20;                                    // 20 tiles x 16 MMAs = 320 accumulations
  static constexpr bool OK = KC::HOIST && N0 <= 128 && COLS <= 512 && SMEM_BYTES <= 227 * 1024;
};

template <class S>
__global__ void __launch_bounds__(128, 2) nm_exo_gemm_tc_bwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ ExoPtrs exo,
                                                                    const float* __restrict__ dZ1, float* __restrict__ grad_partials, int64_t B) {
  using X = ExoTcBwdCfg<S>;
  using KC = KCfg<S>;
  constexpr int N0 = X::N0, KW = S::Pol::sz(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  if ((smem_u32(sm) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + X::TILE_BYTES);         // [0] MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  for (int i = tid; i < X::TILE_BYTES / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc<X::COLS>(tmem_slot);
  tc::fence_before();
  tc::fence_async_smem();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  const uint32_t sbase = smem_u32(sm);
  const uint64_t d_sw0 = tc::smem_desc_sw128(sbase);
  float* gpol = grad_partials + size_t(blockIdx.x) * size_t(p.n_params) + p.policy.param_offset;
  uint32_t phase = 0;
  const int64_t R = B * p.nsteps;
  const int64_t tiles = (R + 127) / 128;
  // per-thread pieces of the swizzled address of element (row f, trajectory-slot tid) -- see tc::sw128_off
  const uint32_t t_blk = uint32_t(tid >> 5), t_chunk = uint32_t((tid & 31) >> 2), t_in = uint32_t(tid & 3) * 4;
  auto tstore = [&](uint32_t tile_off, int Rr, int f, float v) {
    tc::st_shared(sbase + tile_off + t_blk * uint32_t(Rr * 128) + uint32_t(f * 128) + ((t_chunk ^ uint32_t(f & 7)) << 4) + t_in, tc::tf32_hi(v));
  };
  auto flush = [&]() {                                                    // lane n of the accumulators -> row n of dW_0[:, exo]
    static_for<0, S::NSEG>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
        constexpr int W = S::seg_width(s), OFF = exo_seg_off<S>(s);
#pragma unroll 1
        for (int c0 = 0; c0 < X::acc_cols(W); c0 += 8) {
          float v[8];
          tc::tmem_ld<8>(tmem + lane_base + X::acc_off(s) + c0, v);
          if (tid < N0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c0 + i < W) atomicAdd(gpol + tid * KW + KC::KS + OFF + c0 + i, v[i]);
          }
        }
      }
    });
  };
  int since_flush = 0;
  // software pipeline (as in the forward GEMM): the next segment's row slice and the next tile's dZ_1 row are loaded into
  // registers while the tensor pipe contracts the current one
  constexpr int PF = X::rx() / 4, S0 = exo_first_seg<S>(), PD = N0 / 4;
  static_assert(N0 % 4 == 0, "hoisted first layer: N0 % 4 == 0 (KCfg::hoist_ok)");
  float4 pre[PF], pre_d[PD];
  auto dz_prefetch = [&](int64_t tile_) {
    const bool ok = tile_ * 128 + tid < R;
#pragma unroll
    for (int n4 = 0; n4 < PD; ++n4) pre_d[n4] = ok ? __ldg(reinterpret_cast<const float4*>(dZ1 + (tile_ * 128 + tid) * N0 + n4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (int64_t(blockIdx.x) < tiles) {
    const int64_t r0 = int64_t(blockIdx.x) * 128 + tid < R ? int64_t(blockIdx.x) * 128 + tid : R - 1;
    dz_prefetch(blockIdx.x);
    exo_row_prefetch<S, S0, PF>(p, exo, r0 / p.nsteps, int(r0 % p.nsteps), pre);
  }
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const bool valid = tile * 128 + tid < R;
    const int64_t r = valid ? tile * 128 + tid : R - 1;
    const int64_t b = r / p.nsteps;
    const int t = int(r - b * p.nsteps);
    const bool fresh = since_flush == 0;
    // DT <- dZ_1 rows (zero for rows past the end: they contribute nothing)
#pragma unroll
    for (int n4 = 0; n4 < PD; ++n4) {
      const float4 v = pre_d[n4];
      tstore(0, X::RD, n4 * 4 + 0, v.x); tstore(0, X::RD, n4 * 4 + 1, v.y); tstore(0, X::RD, n4 * 4 + 2, v.z); tstore(0, X::RD, n4 * 4 + 3, v.w);
    }
    if (tile + gridDim.x < tiles) dz_prefetch(tile + gridDim.x);
    static_for<0, S::NSEG>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
        constexpr int W = S::seg_width(s), RX = X::rx();
#pragma unroll
        for (int k4 = 0; k4 < W / 4; ++k4) {
          const float4 v = pre[k4];
          tstore(X::DT_BYTES, RX, k4 * 4 + 0, v.x); tstore(X::DT_BYTES, RX, k4 * 4 + 1, v.y); tstore(X::DT_BYTES, RX, k4 * 4 + 2, v.z); tstore(X::DT_BYTES, RX, k4 * 4 + 3, v.w);
        }
        if constexpr (exo_next_seg<S>(s) >= 0) exo_row_prefetch<S, exo_next_seg<S>(s), PF>(p, exo, b, t, pre);
        else if (tile + gridDim.x < tiles) {
          const int64_t rn = (tile + gridDim.x) * 128 + tid < R ? (tile + gridDim.x) * 128 + tid : R - 1;
          exo_row_prefetch<S, S0, PF>(p, exo, rn / p.nsteps, int(rn % p.nsteps), pre);
        }
        tc::fence_async_smem();
        tc::fence_before();
        __syncthreads();
        if (tid < 32) {
          tc::fence_after();
          constexpr uint32_t idesc = tc::idesc_tf32(128, X::acc_cols(W));
          const uint64_t da = d_sw0, db = tc::desc_add(d_sw0, X::DT_BYTES);
          if (tc::elect_one()) {
            static_for<0, 16>([&](auto kc) {
              constexpr uint32_t q = decltype(kc)::value;
              tc::mma_ss(tmem + X::acc_off(s), tc::desc_add(da, (q >> 2) * (X::RD * 128) + (q & 3) * 32),
                         tc::desc_add(db, (q >> 2) * (RX * 128) + (q & 3) * 32), idesc, (q == 0 && fresh) ? 0u : 1u);
            });
            tc::commit(&bars[0]);
          }
          __syncwarp();
        }
        mbar_wait(&bars[0], phase);                             // XT (and, after the last segment, DT) may be overwritten
        phase ^= 1;
        tc::fence_after();
      }
    });
    if (++since_flush == X::FLUSH_TILES) { flush(); since_flush = 0; tc::fence_before(); __syncthreads(); tc::fence_after(); }
  }
  if (since_flush > 0) flush();
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<X::COLS>(tmem);
}

}  // namespace nm
