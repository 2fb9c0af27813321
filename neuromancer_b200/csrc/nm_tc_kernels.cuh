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

template <class S>
struct TcKCfg {
  using C = tc::TcCfg<typename S::RhsM>;
  static constexpr int NX = S::NX, NU = S::NU;
  static constexpr bool ELIGIBLE = S::RHS == NM_RHS_MLP && S::Pol::NL == 0 && !S::HAS_OUTPUT && S::ND == 0 && C::shape_ok() &&
                                   C::K(0) == NX + NU && C::N(C::NL - 1) == NX;
  // shared memory of the forward kernel: weight block + barriers/slots; padded so that no more CTAs become
  // resident than tensor memory has room for (512 columns per SM)
  static constexpr int MAX_CTAS = 512 / C::FWD_COLS;
  static constexpr size_t FWD_W_BYTES = size_t(round_up(C::FWD_W, 32)) * 4;
  static constexpr size_t FWD_SMEM_RAW = FWD_W_BYTES + 1024;
  static constexpr size_t FWD_SMEM_CAP = (227 * 1024) / (MAX_CTAS + 1) + 1024;
  static constexpr size_t FWD_SMEM_BYTES = FWD_SMEM_RAW > FWD_SMEM_CAP ? FWD_SMEM_RAW : FWD_SMEM_CAP;
  static constexpr bool FITS_FWD = FWD_SMEM_RAW <= 227 * 1024;
};

template <class S>
__global__ void nm_tc_prep_fwd_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, float* __restrict__ wprep) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  tc::prep_tc_fwd<typename S::RhsM>(params + plan.rhs_mlp.param_offset, wprep, gtid, gthreads);
}

// One MLP evaluation on the tile: in[K0] (registers) -> out[NX] (registers).  All 128 threads call it.
//   `phase` is the parity of the MMA-completion mbarrier (flipped per tensor-core layer).
template <class S>
__device__ __forceinline__ void tc_mlp_forward(const float* __restrict__ W, uint32_t tmem, uint32_t lane_base, uint64_t* mma_bar,
                                               uint32_t& phase, float act_prm, const float (&in)[tc::TcCfg<typename S::RhsM>::K(0)],
                                               float (&out)[S::NX], int tid) {
  using M = typename S::RhsM;
  using C = tc::TcCfg<M>;
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
          float z = C::BIAS ? w[K0] : 0.f;
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
  for (int i = 0; i < S::NX; ++i) out[i] = 0.f;

  // ---- hidden layers on the tensor cores
  static_for<1, NL - 1>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int KP = C::kp(l), NP = C::np(l), Nl = C::N(l);
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after();
      const uint32_t wh = smem_u32(W + C::img_off(l)), wl = wh + C::img_floats(l) * 4;
      constexpr uint32_t idesc = tc::idesc_tf32(128, NP);
      constexpr uint32_t LBO = NP * 16, SBO = 128, KSTEP = 2 * NP * 16;
      // the two small products first: the accumulator truncates on every accumulate (see nm_tc.cuh)
#pragma unroll
      for (int ks = 0; ks < KP / 8; ++ks) {
        tc::mma_ts(tmem + C::ACC, tmem + C::OPL + ks * 8, tc::smem_desc(wh + ks * KSTEP, LBO, SBO), idesc, ks > 0);
        tc::mma_ts(tmem + C::ACC, tmem + C::OPH + ks * 8, tc::smem_desc(wl + ks * KSTEP, LBO, SBO), idesc, 1);
      }
#pragma unroll
      for (int ks = 0; ks < KP / 8; ++ks)
        tc::mma_ts(tmem + C::ACC, tmem + C::OPH + ks * 8, tc::smem_desc(wh + ks * KSTEP, LBO, SBO), idesc, 1);
      tc::commit(mma_bar);
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
template <class S>
__global__ void __launch_bounds__(128, 1) nm_tc_fwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ FwdArgs args) {
  using C = tc::TcCfg<typename S::RhsM>;
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
  constexpr int NT = S::NTERMS > 0 ? S::NTERMS : 1;
  const bool fused = S::NTERMS > 0 && args.fused_terms;
  if (tid == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = p.nsteps + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = p.nsteps; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = p.nsteps; vt.D[NM_VAR_Y] = 0;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
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

  const int T = p.nsteps;
  constexpr int NS = Tableau<S::INTEG>::NS;
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
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float u[NUA];
#pragma unroll
      for (int j = 0; j < NUA; ++j) u[j] = 0.f;
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
        tc_mlp_forward<S>(sm_w, tmem, lane_base, &bars[1], phase, p.rhs_mlp.act_param, in, k[St], tid);
        if constexpr (St + 1 < NS) {
          if (valid && args.chk != nullptr) {
            float* cp = args.chk + ((b * T + t) * (NS - 1) + St) * NX;
#pragma unroll
            for (int i = 0; i < NX; ++i) cp[i] = k[St][i];
          }
        }
      });
      float xn[NX];
      combine_stages<S::INTEG, NX>(p.h, x, k, xn);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = xn[i];
      if (valid) {
#pragma unroll
        for (int i = 0; i < NX; ++i) args.x_traj[(b * (T + 1) + t + 1) * NX + i] = x[i];
      }
    }
    if constexpr (S::NTERMS > 0) {
      if (fused && valid) score_traj_s<S>(p, vt, b, tsum);
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

}  // namespace nm
