// nm_device.cuh -- device building blocks of the fused DPC rollout (sm_100a).
//
// Layout conventions inside a CTA (batch tile of TILE trajectories):
//   * activations live in shared memory as  act[feature][LD]  with the trajectory index contiguous
//     (LD = TILE + 4 floats: rows start 4 banks apart, so the row-parallel accesses of the
//     weight-gradient GEMM and the interleaved-row stores of the layer GEMMs are conflict-free);
//   * layer GEMMs are CTA-cooperative register-tiled FFMA GEMMs:  out[n][traj] = sum_c W[c][n] in[c][traj]
//     with a TMB x TN accumulator tile per thread (exact fp32 -- the 1e-5 trajectory tolerance rules
//     out single-pass TF32, and the measured mma.sync 3xTF32 rate on B200 is only 1.28x the FFMA
//     peak, see profiles/r01_pipe_rates_microbench.txt);
//   * dynamics, bounds and loss terms are thread-per-trajectory math in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "nm_rollout.h"

namespace nm {

// ------------------------------------------------------------------------------------------------
// compile-time helpers
template <int I> using IC = std::integral_constant<int, I>;

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(static_cast<F&&>(f));
  }
}
template <int B, int E, class F>
__device__ __forceinline__ void static_for_rev(F&& f) {   // E-1, E-2, ..., B
  if constexpr (B < E) {
    f(IC<E - 1>{});
    static_for_rev<B, E - 1>(static_cast<F&&>(f));
  }
}

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int pow2_floor(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; }
constexpr int round_up(int a, int b) { return cdiv(a, b) * b; }

// ------------------------------------------------------------------------------------------------
// activations: torch.nn semantics incl. the subgradient conventions autograd uses
__device__ __forceinline__ float ex2_approx(float x) {
#ifdef NM_HOST_EMU                 // tests/emu: the kernel templates compiled for the host (logic checks without a GPU)
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// Standard normal CDF Phi(z) = 0.5*erfc(-z/sqrt(2)) and exp(-z^2/2), branch-free.
//   erfc(t) = 2^(-t*Q(t)) on [0, 4.2] (degree-9 minimax fit of the exponent, tools/fit_erfc.py: fit error
//   1.3e-8, fp32 evaluation error ~7e-8 = 1 ulp of 1.0; erfc(4.2) < 3e-9 so larger arguments are clamped).
//   The exponent form is used because a rounding error d in the exponent costs only erfc(t)*d absolute.
//   torch's exact-erf GELU (nn.GELU()) is reproduced to ~1e-7 absolute, far inside the 1e-5 budget;
//   CUDA's erff costs ~4x the instructions (two code paths, both executed by a diverged warp).
__device__ __forceinline__ void normal_cdf_pdf(float z, float& cdf, float& gauss) {
  const float x = z * 0.70710678118654752440f;
  const float t = fminf(fabsf(x), 4.2f);
  float q = -1.476634237e-05f;
  q = fmaf(q, t, 1.797868946e-04f);
  q = fmaf(q, t, -9.401043775e-04f);
  q = fmaf(q, t, 2.438597105e-03f);
  q = fmaf(q, t, -2.483013238e-04f);
  q = fmaf(q, t, -2.763307886e-02f);
  q = fmaf(q, t, 1.482808647e-01f);
  q = fmaf(q, t, 9.184465932e-01f);
  q = fmaf(q, t, 1.627907100e+00f);
  const float half_erfc = 0.5f * ex2_approx(-(q * t));            // 0.5*erfc(|x|)
  cdf = 0.5f + copysignf(0.5f - half_erfc, x);
  gauss = ex2_approx(z * z * -0.72134752044448170368f);           // exp(-z^2/2)
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z, float prm) {
  if constexpr (ACT == NM_ACT_IDENTITY) return z;
  else if constexpr (ACT == NM_ACT_RELU) return z > 0.f ? z : 0.f;
  else if constexpr (ACT == NM_ACT_GELU) {
    float cdf, g;
    normal_cdf_pdf(z, cdf, g);
    return z * cdf;
  }
  else if constexpr (ACT == NM_ACT_LEAKY) return z > 0.f ? z : z * prm;
  else if constexpr (ACT == NM_ACT_TANH) return tanhf(z);
  else if constexpr (ACT == NM_ACT_ELU) return z > 0.f ? z : prm * expm1f(z);
  else if constexpr (ACT == NM_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  else if constexpr (ACT == NM_ACT_SOFTPLUS) return z > 20.f ? z : log1pf(expf(z));
  else return z;
}
// activations whose derivative cannot be recovered from the output keep a derivative buffer
template <int ACT> constexpr bool act_needs_dbuf() { return ACT == NM_ACT_GELU || ACT == NM_ACT_SOFTPLUS; }

// activation and its derivative from the pre-activation in one go (shares the transcendental work)
template <int ACT>
__device__ __forceinline__ void act_pair_z(float z, float prm, float& a, float& d) {
  if constexpr (ACT == NM_ACT_GELU) {
    float cdf, g;
    normal_cdf_pdf(z, cdf, g);
    a = z * cdf;
    d = fmaf(z * 0.39894228040143267794f, g, cdf);
  } else if constexpr (ACT == NM_ACT_SOFTPLUS) {
    a = z > 20.f ? z : log1pf(expf(z));
    d = z > 20.f ? 1.f : 1.f / (1.f + expf(-z));
  } else {
    a = act_fwd<ACT>(z, prm);
    d = 1.f;
  }
}

template <int ACT>
__device__ __forceinline__ float act_deriv_a(float a, float prm) {   // derivative from the output a = act(z)
  if constexpr (ACT == NM_ACT_IDENTITY) return 1.f;
  else if constexpr (ACT == NM_ACT_RELU) return a > 0.f ? 1.f : 0.f;           // relu'(0) = 0
  else if constexpr (ACT == NM_ACT_LEAKY) return a > 0.f ? 1.f : prm;          // slope at z <= 0
  else if constexpr (ACT == NM_ACT_TANH) return 1.f - a * a;
  else if constexpr (ACT == NM_ACT_ELU) return a > 0.f ? 1.f : a + prm;        // alpha*exp(z)
  else if constexpr (ACT == NM_ACT_SIGMOID) return a * (1.f - a);
  else return 1.f;
}

// ------------------------------------------------------------------------------------------------
// vectorised access of V contiguous floats (V in {1,2,3} or a multiple of 4)
template <int V>
__device__ __forceinline__ void ld_vec(float (&dst)[V], const float* __restrict__ src) {
  if constexpr (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
      dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  } else if constexpr (V == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) dst[i] = src[i];
  }
}
template <int V>
__device__ __forceinline__ void st_vec(float* __restrict__ dst, const float (&src)[V]) {
  if constexpr (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; ++i)
      *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(src[0], src[1]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) dst[i] = src[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled CTA GEMM:   out[n][traj] = sum_{c < KC} W[c][col(n)] * in[c][traj]
//
//   thread (ng, tg) owns outputs  n = ng + NG*j (j < TN)   x   traj = tg*TM + mb*TMB + i.
//   The interleaved n assignment makes the epilogue's row stores hit rows 1 apart across lanes
//   (4 banks apart with LD = TILE+4) -- conflict-free; the weight matrix is staged already
//   permuted so that a thread's TN weights are contiguous (one LDS.128).
constexpr int g_tn(int n) { return n >= 4 ? 4 : (n < 1 ? 1 : n); }
constexpr int g_ng(int n) { return cdiv(n < 1 ? 1 : n, g_tn(n)); }
constexpr int g_np(int n) { return n < 1 ? 0 : round_up(g_ng(n) * g_tn(n), 4); }   // row pitch (floats)
constexpr int g_col(int n_total, int n) { return (n % g_ng(n_total)) * g_tn(n_total) + n / g_ng(n_total); }

template <int N, int TILE, int THREADS>
struct GemmCfg {
  static constexpr int TN = g_tn(N);
  static constexpr int NG = g_ng(N);
  static constexpr int NP = g_np(N);
  static_assert(NG <= THREADS, "layer too wide for this CTA size");
  static constexpr int TG = pow2_floor(cmin(THREADS / NG, TILE));     // trajectory groups
  static constexpr int TM = TILE / TG;                                 // trajectories per thread
  static constexpr int TMB = cmin(TM, 8);                              // ... per register block
  static constexpr int MB = TM / TMB;
  static constexpr int ACTIVE = NG * TG;
  static_assert(TILE % TG == 0 && TM % TMB == 0, "tile shape");
};

#ifndef NM_TUNE_GEMM_UNROLL
#define NM_TUNE_GEMM_UNROLL 4
#endif
constexpr int kGemmUnroll = NM_TUNE_GEMM_UNROLL;   // k-loop unroll of the layer GEMMs (tuning knob)

// epi(ng, traj0, acc[TMB][TN]) is invoked once per register block; output j of the thread is
// n = ng + NG*j, stored weight column ng*TN + j.
template <int KC, int N, int TILE, int THREADS, int LD, class Epi>
__device__ __forceinline__ void gemm_phase(const float* __restrict__ W, const float* __restrict__ in,
                                           int tid, Epi&& epi) {
  using G = GemmCfg<N, TILE, THREADS>;
  if (tid >= G::ACTIVE) return;
  const int ng = tid % G::NG, tg = tid / G::NG;
  const float* __restrict__ wp = W + ng * G::TN;
#pragma unroll 1
  for (int mb = 0; mb < G::MB; ++mb) {
    const int traj0 = tg * G::TM + mb * G::TMB;
    const float* __restrict__ ip = in + traj0;
    float acc[G::TMB][G::TN];
#pragma unroll
    for (int i = 0; i < G::TMB; ++i)
#pragma unroll
      for (int j = 0; j < G::TN; ++j) acc[i][j] = 0.f;
#pragma unroll kGemmUnroll
    for (int c = 0; c < KC; ++c) {
      float w[G::TN], a[G::TMB];
      ld_vec<G::TN>(w, wp + c * G::NP);
      ld_vec<G::TMB>(a, ip + c * LD);
#pragma unroll
      for (int i = 0; i < G::TMB; ++i)
#pragma unroll
        for (int j = 0; j < G::TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    epi(ng, traj0, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight-gradient accumulation:  dW[n][k] += sum_traj delta[n][traj] * a[k][traj]
//
//   thread tile RJ x RK with interleaved rows  n = jt + NJ*i,  k = kt + NKT*i'  (the tile index
//   varies fastest along k across lanes, so the 8 lanes of a quarter-warp read 8 consecutive `a`
//   rows and one broadcast `delta` row: conflict-free with LD = TILE+4).  With fewer tiles than
//   threads the trajectory range is split across SPLIT thread groups; the partial sums meet in the
//   end-of-kernel reduction.
template <int N, int K, int TILE, int THREADS>
struct DwCfg {
  static constexpr int RJ = cmin(4, N), RK = cmin(4, K);
  static constexpr int NJ = cdiv(N, RJ), NKT = cdiv(K, RK);
  static constexpr int TILES = NJ * NKT;
  static constexpr int SLOTS = cdiv(TILES, THREADS);
  static constexpr int SPLIT = TILES >= THREADS ? 1 : pow2_floor(cmin(THREADS / TILES, TILE / 4));
  static constexpr int TRAJ_PER = TILE / SPLIT;      // trajectories each group sums over
  static constexpr int REGS = SLOTS * RJ * RK;
  static_assert(TRAJ_PER % 4 == 0, "trajectory split");
};

// one (slot) tile over this thread-group's trajectory range, accumulated into acc[RJ][RK]
template <int N, int K, int TILE, int THREADS, int LD>
__device__ __forceinline__ void dw_tile(float (&acc)[DwCfg<N, K, TILE, THREADS>::RJ][DwCfg<N, K, TILE, THREADS>::RK],
                                        const float* __restrict__ delta, const float* __restrict__ a,
                                        int unit) {
  using D = DwCfg<N, K, TILE, THREADS>;
  const int tile = unit % D::TILES;
  const int grp = unit / D::TILES;
  if (grp >= D::SPLIT) return;
  const int kt = tile % D::NKT, jt = tile / D::NKT;
  const float* dp[D::RJ];
  const float* ap[D::RK];
  // rows beyond N / K are clamped: they read valid shared memory and are dropped at write-out
#pragma unroll
  for (int i = 0; i < D::RJ; ++i) dp[i] = delta + min(jt + D::NJ * i, N - 1) * LD + grp * D::TRAJ_PER;
#pragma unroll
  for (int i = 0; i < D::RK; ++i) ap[i] = a + min(kt + D::NKT * i, K - 1) * LD + grp * D::TRAJ_PER;
#pragma unroll 2
  for (int t4 = 0; t4 < D::TRAJ_PER; t4 += 4) {
    float4 dv[D::RJ], av[D::RK];
#pragma unroll
    for (int i = 0; i < D::RJ; ++i) dv[i] = *reinterpret_cast<const float4*>(dp[i] + t4);
#pragma unroll
    for (int i = 0; i < D::RK; ++i) av[i] = *reinterpret_cast<const float4*>(ap[i] + t4);
#pragma unroll
    for (int i = 0; i < D::RJ; ++i)
#pragma unroll
      for (int j = 0; j < D::RK; ++j) {
        float v = acc[i][j];
        v = fmaf(dv[i].x, av[j].x, v);
        v = fmaf(dv[i].y, av[j].y, v);
        v = fmaf(dv[i].z, av[j].z, v);
        v = fmaf(dv[i].w, av[j].w, v);
        acc[i][j] = v;
      }
  }
}

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarriers, 1-D bulk TMA copy global -> shared with completion on an mbarrier (SASS: UBLKCP), named
// barriers.  (NM_HOST_EMU: tests/emu replaces them with host models to run the kernel templates without a GPU.)
#ifdef NM_HOST_EMU
}  // namespace nm
#include "ptx_emu.h"
namespace nm {
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                   "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// bar.sync <id>, 64: the two warps of a thread pair in the PAIR owner layout of the tensor-core adjoint
__device__ __forceinline__ void named_bar_sync64(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }
#endif

}  // namespace nm
