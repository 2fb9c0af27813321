// nm_wide_kernels.cuh -- tcgen05 kernels for closed-loop systems with a WIDE policy MLP (C5: 24-128-128-128-128-6 in
// front of a 12-state RK4 plant; reference blocks.py:169-177, 611-619, system.py:261-277, integrators.py:196-211).
//
// Why a second tensor-core family (measured, profiles/r02_probes_fp16_tcgen05_l2stream.txt):
//   * three 128x128 hidden layers as tf32 hi/lo images are 418 KB -- they do not fit an SM.  As FP16 hi/lo images they
//     are 192 KB and stay RESIDENT in shared memory; kind::f16 MMAs cost the same cycles as kind::tf32 ones but
//     contract K = 16 per instruction, and fp16 has the same 11-bit significand as tf32: the 3-product split
//     A_hi*W_lo + A_lo*W_hi + A_hi*W_hi keeps 22 bits (the weight remainders are stored scaled by 2^11 so they stay in
//     the normal range: their product is accumulated first and folded in through scale-input-d = 2^-11; the activation
//     remainders are stored as they are -- see split2).  Range: |activation|, |weight| < 65504 (beyond that the result
//     is inf/NaN, never wrong).
//   * fp16 operands may be K-major or MN-major: ONE weight image serves the forward contraction (K-major) and the
//     data-gradient contraction (the same bytes read MN-major) -- no transposed copies.
//   * every layer of the policy runs on the tensor pipe (layer 0: K padded to 16, output layer: N padded to 16); the
//     CUDA cores only do activations, operand splitting, dynamics and loss terms.
// Thread = trajectory = TMEM lane (as in nm_tc_kernels.cuh); one CTA per SM holds TWO 128-trajectory tiles in
// tensor memory (2 x [A_hi 64 | A_lo 64 | D 128] columns) that ping-pong between the tensor pipe and their owner
// warps; the first warp of each slot issues the slot's MMAs once all 128 operand rows have arrived (a ninth warp
// would cost every thread 56 registers: the register file is allocated per 4 warps).
//
// Adjoint: recompute (3-product) + data gradients (3-product, cotangents carried scaled by a power of two `gs` so that
// they sit in fp16 range) in nm_wide_bwd_kernel; the WEIGHT gradients are a separate streaming GEMM
// (nm_wide_wgrad_kernel): the adjoint writes the fp16 "hi" halves of every layer input a_l and every scaled
// pre-activation cotangent dZ_l -- the very words it hands to the tensor pipe -- into a staging buffer laid out as
// MN-major operand tiles X[feature/8][trajectory][8], and the GEMM contracts them over trajectories with TMEM
// accumulators (432 + 64 columns: they cannot share tensor memory with the rollout pipeline).  Single-pass fp16 =
// 11-bit operands, unbiased rounding noise that averages out over the B*nsteps products of every entry -- the rule of
// nm_tc_kernels.cuh: the dispatcher runs the exact FFMA adjoint below 2^20 products.  The horizon is swept in chunks
// of CS steps (staging = tiles * CS * 268 KB), the co-state crossing chunk boundaries through a [B, nx] buffer.
#pragma once
#include "nm_tc_kernels.cuh"
#ifndef NM_HOST_EMU
#include <cuda_fp16.h>

#ifdef NM_TUNE_WPROF
#define WP_DECL(n) long long n = 0
#define WP_T0() const long long wp_t0 = clock64()
#define WP_ADD(n) n += clock64() - wp_t0
#else
#define WP_DECL(n)
#define WP_T0()
#define WP_ADD(n)
#endif
namespace nm {
namespace wd {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);   // D f32, A/B f16
}
// MODE 0: D = A*B   1: D += A*B   2: D = A*B + D * 2^-11
template <int MODE>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  if constexpr (MODE == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
  else if constexpr (MODE == 1)
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <int W> __device__ __forceinline__ void tmem_stu(uint32_t taddr, const uint32_t (&v)[W]) {
  static_assert(W == 8 || W == 16, "tmem_stu: 8 or 16 words");
  if constexpr (W == 16) tmem_st16u(taddr, v); else tmem_st8u(taddr, v);
}
// L2 residency: the derivative stash is rewritten and re-read every step (evict_last keeps it out of DRAM: 256-bit
// accesses carry the priority as a qualifier -- a policy operand would cost two R2UR per access); the staging tiles are
// written once and read by a later kernel (streaming stores: they must not push the stash out)
#ifndef NM_TUNE_WSTPOL
#define NM_TUNE_WSTPOL 0
#endif
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#if NM_TUNE_WSTPOL == 0
  asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#elif NM_TUNE_WSTPOL == 1
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#else
  asm volatile("st.global.cg.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#endif
}
struct alignas(32) float8 { float v[8]; };
__device__ __forceinline__ void st_stash8(float8* p, const float* v) {
  asm volatile("st.global.L2::evict_last.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_stash8(const float8* p, float* v) {
#ifdef NM_TUNE_WNOSTASH
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.5f;
  return;
#endif
  asm volatile("ld.global.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}
// fire-and-forget add of two adjacent floats (8-byte aligned): the flush of the weight-gradient accumulators does not wait
// for the old values (a read-modify-write exposed an L2 round trip per group of eight)
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

// Packed fp32 pairs: Blackwell's FFMA2 / FMUL2 / FADD2 work on two fp32 lanes per instruction -- one issue slot instead of
// two (the FMA pipe is busy just as long).  The epilogues below are issue-bound, so everything element-wise runs in pairs.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b) { f2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// fp16 split of two values: hi = rn16(a), lo = rn16(a - hi), packed (first value in the low half).  The remainder is
// NOT scaled: it is at most 2^-11 |a|, and where that falls below fp16's normal range (|a| < 2^-3) the subnormal
// spacing 2^-24 bounds its error by 2^-25 absolute -- never worse than the 2^-22 relative of a scaled remainder.  (The
// WEIGHT remainders stay scaled by 2^11: weights are small and constant, their images are prepared once.)
__device__ __forceinline__ void split2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  float d0, d1;
  up2(sub2(pk2(a0, a1), pk2(hf.x, hf.y)), d0, d1);
  const __half2 l = __floats2half2_rn(d0, d1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Activation epilogue on an accumulator entry `acc` (= z - bias).  GELU (exact erf form, nn.GELU()) works on
// x = z / sqrt(2) = fma(acc, 1/sqrt2, bias/sqrt2) -- the bias comes pre-scaled, its add is free -- with the erfc fit of
// nm_device.cuh (erfc(t) = 2^(-t Q(t)), |t| <= 4.2):   e = erfc(|x|) / sqrt2 = 2^(-t Q - 1/2),
//   gelu(z)  = z Phi(z)         = x / sqrt2 + |x| (1/sqrt2 - e)
//   gelu'(z) = Phi(z) + z phi(z) = 1/2 + sign(x) (1/sqrt2 - e) / sqrt2 + x exp(-x^2) / sqrt(pi)
template <int ACT> constexpr float bias_scale() { return ACT == NM_ACT_GELU ? 0.70710678118654752440f : 1.f; }
__device__ __forceinline__ float gelu_core(float x, float& ax) {          // returns s = 1/sqrt2 - e
  ax = fabsf(x);
  const float t = fminf(ax, 4.2f);
  float q = -1.476634237e-05f;
  q = fmaf(q, t, 1.797868946e-04f);
  q = fmaf(q, t, -9.401043775e-04f);
  q = fmaf(q, t, 2.438597105e-03f);
  q = fmaf(q, t, -2.483013238e-04f);
  q = fmaf(q, t, -2.763307886e-02f);
  q = fmaf(q, t, 1.482808647e-01f);
  q = fmaf(q, t, 9.184465932e-01f);
  q = fmaf(q, t, 1.627907100e+00f);
  return 0.70710678118654752440f - ex2_approx(fmaf(-q, t, -0.5f));
}
template <int ACT>
__device__ __forceinline__ float act_acc(float acc, float bs, float prm) {
  if constexpr (ACT == NM_ACT_GELU) {
    const float x = fmaf(acc, 0.70710678118654752440f, bs);
    float ax;
    const float s = gelu_core(x, ax);
    return fmaf(ax, s, x * 0.70710678118654752440f);
  } else return act_fwd<ACT>(acc + bs, prm);
}
template <int ACT>
__device__ __forceinline__ void act_acc_pair(float acc, float bs, float prm, float& a, float& d) {
  if constexpr (ACT == NM_ACT_GELU) {
    const float x = fmaf(acc, 0.70710678118654752440f, bs);
    float ax;
    const float s = gelu_core(x, ax);
    a = fmaf(ax, s, x * 0.70710678118654752440f);
    const float g = ex2_approx(x * (x * -1.44269504088896340736f));        // exp(-x^2)
    const float cdf = fmaf(copysignf(0.70710678118654752440f, x), s, 0.5f);
    d = fmaf(x * 0.56418958354775628695f, g, cdf);
  } else tc::act_ad<ACT>(acc + bs, prm, a, d);
}

// The same two epilogues on a PAIR of accumulator entries (bias pair pre-scaled like `bs`)
__device__ __forceinline__ f2_t gelu_core2(f2_t x, float& ax0, float& ax1) {        // s = 1/sqrt2 - e, per lane
  float x0, x1;
  up2(x, x0, x1);
  ax0 = fabsf(x0); ax1 = fabsf(x1);
  const float t0 = fminf(ax0, 4.2f), t1 = fminf(ax1, 4.2f);
  const f2_t t = pk2(t0, t1);
  f2_t q = bc2(-1.476634237e-05f);
  q = fma2(q, t, bc2(1.797868946e-04f));
  q = fma2(q, t, bc2(-9.401043775e-04f));
  q = fma2(q, t, bc2(2.438597105e-03f));
  q = fma2(q, t, bc2(-2.483013238e-04f));
  q = fma2(q, t, bc2(-2.763307886e-02f));
  q = fma2(q, t, bc2(1.482808647e-01f));
  q = fma2(q, t, bc2(9.184465932e-01f));
  q = fma2(q, t, bc2(1.627907100e+00f));
  float g0, g1;
  up2(fma2(q, pk2(-t0, -t1), bc2(-0.5f)), g0, g1);
  return sub2(bc2(0.70710678118654752440f), pk2(ex2_approx(g0), ex2_approx(g1)));
}
template <int ACT>
__device__ __forceinline__ void act_acc2(float& v0, float& v1, float bs0, float bs1, float prm) {
  if constexpr (ACT == NM_ACT_GELU) {
    const f2_t c = bc2(0.70710678118654752440f);
    const f2_t x = fma2(pk2(v0, v1), c, pk2(bs0, bs1));
    float ax0, ax1;
    const f2_t s = gelu_core2(x, ax0, ax1);
    up2(fma2(pk2(ax0, ax1), s, mul2(x, c)), v0, v1);
  } else { v0 = act_fwd<ACT>(v0 + bs0, prm); v1 = act_fwd<ACT>(v1 + bs1, prm); }
}
template <int ACT>
__device__ __forceinline__ void act_acc_pair2(float& v0, float& v1, float bs0, float bs1, float prm, float& d0, float& d1) {
  if constexpr (ACT == NM_ACT_GELU) {
    const f2_t c = bc2(0.70710678118654752440f);
    const f2_t x = fma2(pk2(v0, v1), c, pk2(bs0, bs1));
    float ax0, ax1, x0, x1, g0, g1;
    const f2_t s = gelu_core2(x, ax0, ax1);
    up2(fma2(pk2(ax0, ax1), s, mul2(x, c)), v0, v1);
    up2(x, x0, x1);
    up2(mul2(x, mul2(x, bc2(-1.44269504088896340736f))), g0, g1);
    const f2_t g = pk2(ex2_approx(g0), ex2_approx(g1));                    // exp(-x^2)
    const f2_t cdf = fma2(pk2(copysignf(0.70710678118654752440f, x0), copysignf(0.70710678118654752440f, x1)), s, bc2(0.5f));
    up2(fma2(mul2(x, bc2(0.56418958354775628695f)), g, cdf), d0, d1);
  } else { tc::act_ad<ACT>(v0 + bs0, prm, v0, d0); tc::act_ad<ACT>(v1 + bs1, prm, v1, d1); }
}

// The epilogue of 2 NP consecutive accumulator entries v[O .. O + 2 NP), written "vertically": every step of the evaluation
// runs across all NP pairs before the next one starts, so NP independent FFMA2 chains are in flight (written pair after
// pair, the compiler interleaved two chains and every Horner step waited out the FFMA2 latency).  DERIV: also act'(z) -> d.
template <int ACT, int O, int NP, bool DERIV, int N>
__device__ __forceinline__ void act_group(float (&v)[N], float (&d)[N], const float* __restrict__ bias, float prm) {
  if constexpr (ACT == NM_ACT_GELU) {
    const f2_t c = bc2(0.70710678118654752440f);
    f2_t x[NP], t[NP], q[NP];
#pragma unroll
    for (int p = 0; p < NP; p += 2) {
      const float4 bv = *reinterpret_cast<const float4*>(bias + O + 2 * p);
      x[p] = fma2(pk2(v[O + 2 * p], v[O + 2 * p + 1]), c, pk2(bv.x, bv.y));
      x[p + 1] = fma2(pk2(v[O + 2 * p + 2], v[O + 2 * p + 3]), c, pk2(bv.z, bv.w));
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float x0, x1;
      up2(x[p], x0, x1);
      t[p] = pk2(fminf(fabsf(x0), 4.2f), fminf(fabsf(x1), 4.2f));
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(bc2(-1.476634237e-05f), t[p], bc2(1.797868946e-04f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(-9.401043775e-04f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(2.438597105e-03f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(-2.483013238e-04f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(-2.763307886e-02f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(1.482808647e-01f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(9.184465932e-01f));
#pragma unroll
    for (int p = 0; p < NP; ++p) q[p] = fma2(q[p], t[p], bc2(1.627907100e+00f));
#pragma unroll
    for (int p = 0; p < NP; ++p) {                      // s = 1/sqrt2 - 2^(-q t - 1/2)
      float t0, t1, g0, g1;
      up2(t[p], t0, t1);
      up2(fma2(q[p], pk2(-t0, -t1), bc2(-0.5f)), g0, g1);
      q[p] = sub2(c, pk2(ex2_approx(g0), ex2_approx(g1)));
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float x0, x1;
      up2(x[p], x0, x1);
      up2(fma2(pk2(fabsf(x0), fabsf(x1)), q[p], mul2(x[p], c)), v[O + 2 * p], v[O + 2 * p + 1]);
    }
    if constexpr (DERIV) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float x0, x1, g0, g1;
        up2(x[p], x0, x1);
        up2(mul2(x[p], mul2(x[p], bc2(-1.44269504088896340736f))), g0, g1);
        const f2_t g = pk2(ex2_approx(g0), ex2_approx(g1));                  // exp(-x^2)
        const f2_t cdf = fma2(pk2(copysignf(0.70710678118654752440f, x0), copysignf(0.70710678118654752440f, x1)), q[p], bc2(0.5f));
        up2(fma2(mul2(x[p], bc2(0.56418958354775628695f)), g, cdf), d[O + 2 * p], d[O + 2 * p + 1]);
      }
    }
  } else {
    const float is = 1.f / bias_scale<ACT>();
#pragma unroll
    for (int i = O; i < O + 2 * NP; ++i) {
      if constexpr (DERIV) tc::act_ad<ACT>(v[i] + bias[i] * is, prm, v[i], d[i]);
      else v[i] = act_fwd<ACT>(v[i] + bias[i] * is, prm);
    }
  }
}

// ------------------------------------------------------------------------------------------------ static layout
struct ScaleState { float gs; float inv_gs; unsigned int max_bits; unsigned int pad; };

template <class S>
struct WideCfg {
  using MD = TcMode<S>;
  using M = typename S::Pol;
  static constexpr int NL = M::NL;
  static constexpr int K0 = M::sz(0), H = M::sz(1), NOUT = M::sz(NL > 0 ? NL : 0);
  static constexpr int K0P = round_up(K0, 16), NOP = 16;
  static constexpr bool BIAS = M::BIAS;
  static constexpr int NX = S::NX, NU = S::NU;
  // gradient is needed for the policy inputs up to the end of the last state segment
  static constexpr int gw0() {
    int off = 0, gw = 0;
    for (int i = 0; i < S::NSEG; ++i) { off += S::seg_width(i); if (S::seg_kind(i) == NM_SEG_STATE) gw = off; }
    return gw;
  }
  static constexpr int GWP = round_up(gw0() > 0 ? gw0() : 1, 16);
  static constexpr bool uniform() { for (int l = 1; l < NL; ++l) if (M::sz(l) != H) return false; return true; }
  static constexpr bool ELIGIBLE = MD::POLICY && !MD::HOIST && NL >= 3 && NL <= 6 && uniform() && H == 128 && K0 > 8 && K0P <= 64 && GWP <= K0P &&
                                   NOUT <= NOP && NOUT == NU && !S::HAS_OUTPUT && S::ND == 0 && S::ACT_EXO < 0 && MD::segs_ok();
  static constexpr int NHID = NL - 2;                                      // H x H layers (1 .. NL-2)
  // ---- prepared weight block = shared-memory image (bytes); every fp16 image is [k/8][rows][8], hi then lo (x 2^11)
  static constexpr int IMG0_BYTES = K0P * H * 2, IMGH_BYTES = H * H * 2, IMGO_BYTES = H * NOP * 2;
  static constexpr int img0_off() { return 0; }
  static constexpr int imgh_off(int l) { return 2 * IMG0_BYTES + (l - 1) * 2 * IMGH_BYTES; }   // l = 1 .. NL-2
  static constexpr int imgo_off() { return 2 * IMG0_BYTES + NHID * 2 * IMGH_BYTES; }
  static constexpr int F32_OFF = imgo_off() + 2 * IMGO_BYTES;
  // fp32 tail: bias[l][H] for l = 0..NL-2, bias_out[NOP], W_out[NOUT][H] (output layer backward on the CUDA cores)
  static constexpr int bias_off(int l) { return F32_OFF + l * H * 4; }
  static constexpr int biaso_off() { return F32_OFF + (NL - 1) * H * 4; }
  static constexpr int wout_off() { return biaso_off() + NOP * 4; }
  static constexpr int W_BYTES = round_up(wout_off() + NOUT * H * 4, 128);
  static constexpr int BAR_OFF = W_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 1024;
  static constexpr bool FITS = SMEM_BYTES + 1024 <= 227 * 1024;
  // parameter layout of the policy (reference nn.Linear: W[out][in] then bias[out])
  static constexpr int kin(int l) { return l == 0 ? K0 : H; }
  static constexpr int nout(int l) { return l == NL - 1 ? NOUT : H; }
  static constexpr int p_off(int l) { int o = 0; for (int i = 0; i < l; ++i) o += nout(i) * kin(i) + (BIAS ? nout(i) : 0); return o; }
  // ---- tensor memory: two tile slots of [A_hi H/2 | A_lo H/2 | D H] columns
  static constexpr int SLOT_COLS = 2 * H, A_HI = 0, A_LO = H / 2, DCOL = H;
  // ---- launch shape
  static constexpr int THREADS = 256;                                     // 2 tile slots x 4 owner warps (register file: allocation is per 4 warps)
#ifdef NM_TUNE_WFT
  static constexpr int FWD_TPT = NM_TUNE_WFT;
#else
  static constexpr int FWD_TPT = 2;                                       // forward kernel: threads per trajectory (16 warps: the epilogues hide each other's stalls)
#endif
  static constexpr int FWD_THREADS = 256 * FWD_TPT;
#ifdef NM_TUNE_WBT
  static constexpr int BWD_TPT = NM_TUNE_WBT;
#else
  static constexpr int BWD_TPT = 1;                                       // adjoint kernel: threads per trajectory (2 measured SLOWER, twice: 19.0 vs 15.6 ms at B = 37,888 with 384 B of spills; 13.8 vs 12.7 ms after the 16-column epilogue chunks and L2-hint prefetch cut them to 200 B -- the sweep is bound by its per-slot chain and memory round trips, not by warps per scheduler)
#endif
  static constexpr int BWD_THREADS = 256 * BWD_TPT;
  // ---- adjoint: staging item (one tile-step), byte offsets of the operand tiles X[feature/8][128][8] halves
  static constexpr int GRP = 128 * 16;                                    // one 8-feature group of a tile
  static constexpr int a_off(int l) { return l == 0 ? 0 : (K0P / 8) * GRP + (l - 1) * (H / 8) * GRP; }       // a_l, l = 0..NL-1
  static constexpr int z_off(int l) { return a_off(NL - 1) + (H / 8) * GRP + l * (H / 8) * GRP; }             // dZ_l, l = 0..NL-2
  static constexpr int c_off() { return z_off(NL - 2) + (H / 8) * GRP; }                                      // cotangent of the raw output
  static constexpr int ITEM_BYTES = c_off() + (NOP / 8) * GRP;
  // derivative stash of one tile slot: act'(z_l) for l = 0..NL-3, [l][k/8][128][8] floats
  static constexpr int STASH_SLOT_BYTES = (NL - 2) * H * 128 * 4;
  // ---- weight-gradient GEMM: tensor-memory columns
  static constexpr int wg_col(int l) { return l == 0 ? 0 : (l < NL - 1 ? K0P + (l - 1) * H : K0P + NHID * H); }
  static constexpr int wgb_col(int l) { return K0P + NHID * H + NOP + 16 * l; }                              // bias sums of layer l < NL-1
  static constexpr int WG_COLS = wgb_col(NL - 1);
  static constexpr int WG_RING = 3, WG_SLOT_BYTES = 2 * (H / 8) * GRP;
  static constexpr int WG_ONES_OFF = WG_RING * WG_SLOT_BYTES, WG_BAR_OFF = WG_ONES_OFF + 2 * GRP;
  static constexpr int WG_SMEM_BYTES = WG_BAR_OFF + 256;
  static constexpr int WG_THREADS = 192, WG_FLUSH_ITEMS = 16;
  static constexpr bool WG_OK = WG_COLS <= 512 && WG_SMEM_BYTES + 1024 <= 227 * 1024;
};

// ------------------------------------------------------------------------------------------------ weight preparation
template <class S>
__global__ void nm_wide_prep_kernel(const __grid_constant__ NmPlan plan, const float* __restrict__ params, uint8_t* __restrict__ wprep) {
  using C = WideCfg<S>;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const float* P = params + plan.policy.param_offset;
  auto image = [&](int l, int off, int rows, int kp) {               // [kp/8][rows][8]: element (n, k) = W_l[n][k]
    const int Kr = C::kin(l), Nr = C::nout(l);
    const float* W = P + C::p_off(l);
    __half* hi = reinterpret_cast<__half*>(wprep + off);
    __half* lo = hi + rows * kp;
    for (int idx = gtid; idx < rows * kp; idx += gthreads) {
      const int kc = idx / (rows * 8), n = (idx / 8) % rows, k = kc * 8 + (idx & 7);
      const float v = (n < Nr && k < Kr) ? W[n * Kr + k] : 0.f;
      const __half h = __float2half_rn(v);
      hi[idx] = h;
      lo[idx] = __float2half_rn((v - __half2float(h)) * 2048.f);
    }
  };
  image(0, C::img0_off(), C::H, C::K0P);
  for (int l = 1; l <= C::NL - 2; ++l) image(l, C::imgh_off(l), C::H, C::H);
  image(C::NL - 1, C::imgo_off(), C::NOP, C::H);
  for (int l = 0; l <= C::NL - 2; ++l) {
    float* b = reinterpret_cast<float*>(wprep + C::bias_off(l));
    const float* src = P + C::p_off(l) + C::nout(l) * C::kin(l);
    for (int i = gtid; i < C::H; i += gthreads) b[i] = C::BIAS ? src[i] * bias_scale<C::M::ACT>() : 0.f;
  }
  {
    float* b = reinterpret_cast<float*>(wprep + C::biaso_off());
    const float* src = P + C::p_off(C::NL - 1) + C::NOUT * C::H;
    for (int i = gtid; i < C::NOP; i += gthreads) b[i] = (C::BIAS && i < C::NOUT) ? src[i] : 0.f;
    float* w = reinterpret_cast<float*>(wprep + C::wout_off());
    const float* ws = P + C::p_off(C::NL - 1);
    for (int i = gtid; i < C::NOUT * C::H; i += gthreads) w[i] = ws[i];
  }
}

// ------------------------------------------------------------------------------------------------ MMA batches (issuer lane)
// D[128][N] = A[128][KP] * W^T, K-major image [KP/8][ROWS = N][8]: hi at `img`, lo at `img + ROWS*KP*2`
template <int KP, int N>
__device__ __forceinline__ void issue_kmajor(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t img) {
  constexpr uint32_t idesc = idesc_f16(128, N, 0, 0);
  constexpr uint32_t LBO = N * 16, SBO = 128, KSTEP = 2 * N * 16;
  const uint64_t wh = tc::smem_desc(img, LBO, SBO);
  const uint64_t wl = tc::desc_add(wh, N * KP * 2);
  // smallest products first (the TMEM accumulator truncates): A_hi W_lo at scale 2^11, folded in by the first A_lo W_hi
  static_for<0, KP / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<(ks > 0 ? 1 : 0)>(d, a_hi + ks * 8, tc::desc_add(wl, ks * KSTEP), idesc);
  });
  static_for<0, KP / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<(ks > 0 ? 1 : 2)>(d, a_lo + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
  });
  static_for<0, KP / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<1>(d, a_hi + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
  });
}
// D[128][NP] = A[128][ROWS] * W[:, 0:NP]: the same image [KPI/8][ROWS][8] read MN-major (contraction over its rows)
template <int ROWS, int KPI, int NP>
__device__ __forceinline__ void issue_mnmajor(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t img) {
  constexpr uint32_t idesc = idesc_f16(128, NP, 0, 1);
  constexpr uint32_t LBO = 128, SBO = ROWS * 16, KSTEP = 256;
  const uint64_t wh = tc::smem_desc(img, LBO, SBO);
  const uint64_t wl = tc::desc_add(wh, ROWS * KPI * 2);
  static_for<0, ROWS / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<(ks > 0 ? 1 : 0)>(d, a_hi + ks * 8, tc::desc_add(wl, ks * KSTEP), idesc);
  });
  static_for<0, ROWS / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<(ks > 0 ? 1 : 2)>(d, a_lo + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
  });
  static_for<0, ROWS / 16>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    mma_ts<1>(d, a_hi + ks * 8, tc::desc_add(wh, ks * KSTEP), idesc);
  });
}

// N contiguous floats of a row whose start is aligned like its length (N % 4 == 0: 16 B, N % 2 == 0: 8 B)
template <int N>
__device__ __forceinline__ void ld_row(float (&dst)[N], const float* __restrict__ src) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) { const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i); dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w; }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { const float2 v = __ldg(reinterpret_cast<const float2*>(src) + i); dst[2 * i] = v.x; dst[2 * i + 1] = v.y; }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = __ldg(src + i);
  }
}
// exogenous segments of the policy input of trajectory b at step t (state entries of `in` are left untouched)
template <class S, int K0>
__device__ __forceinline__ void policy_exo(const NmPlan& p, const float* const* exo, int64_t b, int t, float (&in)[K0]) {
  int off = 0;
  static_for<0, S::NSEG>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    constexpr int W = S::seg_width(s);
    if constexpr (S::seg_kind(s) == NM_SEG_EXO) {
      constexpr int e = S::seg_index(s);
      float row[W];
      ld_row<W>(row, exo[e] + (b * p.exo[e].steps + t) * W);
#pragma unroll
      for (int i = 0; i < W; ++i) in[off + i] = row[i];
    }
    off += W;
  });
}

// policy input -> the thread's TMEM lane (hi | lo, K0P halves each); returns the packed hi words for staging
template <class C>
__device__ __forceinline__ void put_input(uint32_t t_hi, uint32_t t_lo, const float (&in)[C::K0], uint32_t (&hi)[C::K0P / 2]) {
  uint32_t lo[C::K0P / 2];
#pragma unroll
  for (int i = 0; i < C::K0P / 2; ++i) {
    const float a0 = 2 * i < C::K0 ? in[2 * i < C::K0 ? 2 * i : 0] : 0.f;
    const float a1 = 2 * i + 1 < C::K0 ? in[2 * i + 1 < C::K0 ? 2 * i + 1 : 0] : 0.f;
    split2(a0, a1, hi[i], lo[i]);
  }
#pragma unroll
  for (int c = 0; c < C::K0P / 2; c += 8) {
    uint32_t h8[8], l8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h8[i] = hi[c + i]; l8[i] = lo[c + i]; }
    tmem_st8u(t_hi + c, h8);
    tmem_st8u(t_lo + c, l8);
  }
}

// number of tiles slot `slot` of this CTA processes
__device__ __forceinline__ int slot_rounds(int num_tiles, int slot) {
  int n = 0;
  for (int r = 0;; ++r) { if ((r * int(gridDim.x) + int(blockIdx.x)) * 2 + slot >= num_tiles) break; ++n; }
  return n;
}

}  // namespace wd

// ================================================================================================ forward
// (PenaltyLoss scoring is NOT fused here: read back per trajectory it is a chain of L2-latency loads during which both
// tile slots and the tensor pipe idle -- measured 45 % of this kernel; the streaming nm_term_score_kernel does it at HBM rate)
template <class S>
__global__ void __launch_bounds__(wd::WideCfg<S>::FWD_THREADS, 1) nm_wide_fwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ FwdArgs args) {
  using C = wd::WideCfg<S>;
  using M = typename C::M;
  constexpr int NX = S::NX, NU = S::NU, K0 = C::K0, H = C::H, NL = C::NL, TILE = 128, TPT = C::FWD_TPT;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::BAR_OFF);           // 0 weights, 1/2 operands ready (slot), 3/4 MMAs done (slot)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 128 * TPT); mbar_init(&bars[2], 128 * TPT);
    mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], uint32_t(C::W_BYTES));
    tma_bulk_g2s(sm, args.wprep, uint32_t(C::W_BYTES), &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  const int T = p.nsteps;

  {
    // ------------------------------------------------------------------ owners: TPT threads per trajectory (= TMEM lane); thread
    // `half` of a trajectory owns the columns [half * H / TPT, (half + 1) * H / TPT) of every layer.  The per-trajectory
    // scalars (state, dynamics) are kept by every thread of the trajectory; thread 0 writes operands and trajectories.
    const int slot = warp / (4 * TPT), half = (warp >> 2) % TPT, m_ = (warp & 3) * 32 + (tid & 31);
    constexpr int HC = H / TPT;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t cb = tmem + lane_base + slot * C::SLOT_COLS;
    const uint32_t t_hi = cb + C::A_HI, t_lo = cb + C::A_LO, t_d = cb + C::DCOL;
    uint64_t* bar_ops = &bars[1 + slot];
    uint64_t* bar_done = &bars[3 + slot];
    uint32_t ph_done = 0, ph_ops = 0;
    const float act_prm = p.policy.act_param;
    const bool issuer = (warp % (4 * TPT)) == 0;
    const uint32_t sbase = smem_u32(sm), cbs = tmem + slot * C::SLOT_COLS;
    // operands of this thread are in tensor memory; the slot's first warp waits for all 128 rows and issues MMA batch `bi`
    // (0: layer 0, 1 .. NL-2: hidden layer, NL-1: output layer)
    auto arrive_ops = [&](int bi) {
      tc::wait_st(); tc::fence_before(); mbar_arrive(bar_ops);
      if (issuer) {
        mbar_wait(bar_ops, ph_ops); ph_ops ^= 1;
        tc::fence_after();
        if (tc::elect_one()) {
          if (bi == 0) wd::issue_kmajor<C::K0P, H>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::img0_off());
          else if (bi < NL - 1) wd::issue_kmajor<H, H>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::imgh_off(1) + (bi - 1) * 2 * C::IMGH_BYTES);
          else wd::issue_kmajor<H, C::NOP>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::imgo_off());
          tc::commit(bar_done);
        }
        __syncwarp();
      }
    };
    WP_DECL(w_arr); WP_DECL(w_wait); WP_DECL(w_epi); WP_DECL(w_out); WP_DECL(w_in);
#ifdef NM_TUNE_WPROF
    const long long w_begin = clock64();
    long long tl[24];
    int tli = 0;
#define WP_MARK() do { if (r == 0 && (t == 100 || t == 101) && tli < 24) tl[tli++] = clock64() - w_begin; } while (0)
#else
#define WP_MARK()
#endif
    for (int r = 0;; ++r) {
      const int tile = (r * int(gridDim.x) + int(blockIdx.x)) * 2 + slot;
      if (tile >= args.num_tiles) break;
      const int64_t b_raw = int64_t(tile) * TILE + m_;
      const bool valid = b_raw < args.B;
      const int64_t b = valid ? b_raw : args.B - 1;
      float x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = __ldg(args.x0 + b * NX + i);
      if (valid && half == 0) st_vec<NX>(args.x_traj + (b * (T + 1)) * NX, x);
      float in_pre[K0];
#pragma unroll
      for (int q = 0; q < K0; ++q) in_pre[q] = 0.f;
      if (half == 0) wd::policy_exo<S, K0>(p, args.exo, b, 0, in_pre);
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        float in[K0], yv[1] = {0.f};
#pragma unroll
        for (int q = 0; q < K0; ++q) in[q] = in_pre[q];
        { WP_T0();
        if (half == 0) {                          // the trajectory's scalars (state, policy input, dynamics) live on its first thread only
          uint32_t hi0[C::K0P / 2];
          tc_policy_state<S, K0>(x, yv, in);
          wd::put_input<C>(t_hi, t_lo, in, hi0);
          (void)hi0;
        }
        WP_ADD(w_in); }
        { WP_T0(); arrive_ops(0); WP_ADD(w_arr); }
        if (half == 0 && t + 1 < T) wd::policy_exo<S, K0>(p, args.exo, b, t + 1, in_pre);     // exogenous inputs one step ahead
        // hidden activations a_1 .. a_{NL-1}
#pragma unroll 1
        for (int l = 0; l < NL - 1; ++l) {
          const float* __restrict__ bias = reinterpret_cast<const float*>(sm + C::bias_off(0)) + l * H;
          { WP_T0(); mbar_wait(bar_done, ph_done); ph_done ^= 1; WP_ADD(w_wait); }
          tc::fence_after();
          WP_MARK();
          WP_T0();
#pragma unroll 1
          for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
            float v[32];
            uint32_t hi[16], lo[16];
            tc::tmem_ld<32>(t_d + c0, v);
            wd::act_group<M::ACT, 0, 8, false>(v, v, bias + c0, act_prm);
            wd::act_group<M::ACT, 16, 8, false>(v, v, bias + c0, act_prm);
#pragma unroll
            for (int i = 0; i < 16; ++i) wd::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            wd::tmem_st16u(t_hi + c0 / 2, hi);
            wd::tmem_st16u(t_lo + c0 / 2, lo);
          }
          WP_ADD(w_epi);
          WP_MARK();
          { WP_T0(); arrive_ops(l + 1); WP_ADD(w_arr); }
        }
        // output layer
        { WP_T0(); mbar_wait(bar_done, ph_done); ph_done ^= 1; WP_ADD(w_wait); }
        tc::fence_after();
        WP_MARK();
        WP_T0();
        if (half == 0) {
          float o16[16], u[NU];
          tc::tmem_ld<16>(t_d, o16);
          const float* __restrict__ bo = reinterpret_cast<const float*>(sm + C::biaso_off());
#pragma unroll
          for (int j = 0; j < NU; ++j) {
            o16[j] += bo[j];
            u[j] = bounds_fwd<M::BOUNDS>(o16[j], p.policy.bmin[j], p.policy.bmax[j]);
          }
          if (valid) {
            st_vec<NU>(args.u_traj + (b * T + t) * NU, u);
            if (args.chk != nullptr) {
              float oc[NU];
#pragma unroll
              for (int j = 0; j < NU; ++j) oc[j] = o16[j];
              st_vec<NU>(args.chk + (b * T + t) * NU, oc);
            }
          }
          float xn[NX], dv_[1] = {0.f};
          Dyn<S>::step(p, x, u, dv_, xn);
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = xn[i];
          if (valid) st_vec<NX>(args.x_traj + (b * (T + 1) + t + 1) * NX, x);
        }
        WP_ADD(w_out);
        WP_MARK();
      }
    }
#ifdef NM_TUNE_WPROF
    if (blockIdx.x == 0 && (tid == 0 || tid == 32 || tid == 128 || tid == 256))
      printf("[wprof fwd tid %d] total %lld | input %lld  arrive+issue %lld  wait_mma %lld  epilogue %lld  out+dyn %lld (steps %d)\n", tid,
             clock64() - w_begin, w_in, w_arr, w_wait, w_epi, w_out, T);
    if (blockIdx.x == 0 && (tid == 32 || tid == 288)) {
      // per step: [epilogue start, epilogue end] x (NL - 1), output-layer result ready, step end
      printf("[wprof fwd timeline tid %d]", tid);
      for (int i = 0; i < tli; ++i) printf(" %lld", tl[i] - tl[0]);
      printf("\n");
    }
#endif
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// PenaltyLoss scoring with the compiled term structure, thread per (trajectory, time element): streams the
// trajectories once at HBM rate (block partial sums in double, same layout as nm_term_score_kernel's)
template <class S>
__global__ void __launch_bounds__(256) nm_wide_score_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ FwdArgs args, int tmax) {
  constexpr int NT = S::NTERMS > 0 ? S::NTERMS : 1;
  __shared__ double red[NT][8];
  const NmPlan& p = plan;
  const VarTable vtc = make_vartable_regs(p, args, S::NX, S::NU, S::NY);
  float ts[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) ts[k] = 0.f;
  const int64_t n = args.B * tmax;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx / tmax;
    const int te = int(idx - b * tmax);
    static_for<0, S::NTERMS>([&](auto kc) {
      constexpr int K = decltype(kc)::value;
      const NmTerm& tm = p.terms[K];
      if (te < tm.t_len) {
        float s = 0.f;
        for (int de = 0; de < tm.d_len; ++de) s += term_penalty_s<S, K>(p, term_value_s<S, K>(tm, vtc, b, te, de));
        ts[K] += s;
      }
    });
  }
  static_for<0, S::NTERMS>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    double v = static_cast<double>(ts[K]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[K][threadIdx.x >> 5] = v;
  });
  __syncthreads();
  if (threadIdx.x < S::NTERMS) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    args.term_partials[size_t(blockIdx.x) * NM_MAX_TERMS + threadIdx.x] = v;
  }
}

// ================================================================================================ adjoint
struct WideBwdArgs {
  BwdArgs b;
  uint8_t* staging;              // [tile][step in chunk] items of WideCfg::ITEM_BYTES
  float* stash;                  // [grid * 2] slots of WideCfg::STASH_SLOT_BYTES
  float* lam_carry;              // [B][NX] co-state at the chunk boundary
  wd::ScaleState* scale;         // cotangent scale of this chunk (read), largest scaled cotangent seen (atomic max)
  const float* G;                // loss-term gradients of this chunk's GROUP of chunks [B][g_nt][NX + NU], row tt = time g_lo + tt
                                 // (nm_wide_termgrad_kernel: one launch per group -- long contiguous reads of every trajectory's rows)
  int g_lo, g_nt;
  int t_lo, t_hi;                // this launch sweeps steps t_hi-1 .. t_lo  (t_hi == nsteps: the first chunk of the pass)
};

// Loss-term gradients of one chunk, thread per (trajectory, time): G[b][tt][0:NX] = d loss / d x[b, t_lo + tt, :] for
// tt <= t_hi - t_lo (the last row only matters when t_hi == nsteps), G[b][tt][NX:] = d loss / d u[b, t_lo + tt, :] for
// tt < t_hi - t_lo.  Time-broadcast atoms (x[:, [-1], :] against a whole range) are left to the adjoint sweep, where all
// lanes sit at the same time index (see term_grad_entry).  External gradients (Python-evaluated terms) are added here.
// (FUSED_ONLY: the instantiation for plans served by the compiled term structure -- no runtime-descriptor code, no VarTable in
// shared memory; picked by the launcher from args.fused_terms)
template <class S, bool FUSED_ONLY = false>
__global__ void __launch_bounds__(256) nm_wide_termgrad_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ BwdArgs args,
                                                               float* __restrict__ G, int t_lo, int t_hi) {
  constexpr int NX = S::NX, NU = S::NU, RW = NX + NU;
  __shared__ float coef[NM_MAX_TERMS];
  __shared__ VarTable vt;
  const NmPlan& p = plan;
  const int T = p.nsteps;
  if (threadIdx.x < NM_MAX_TERMS) coef[threadIdx.x] = threadIdx.x < p.n_terms ? term_coef(p, threadIdx.x, args.grad_scalars, args.b_total) : 0.f;
  if (threadIdx.x == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = T + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = T; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = T; vt.D[NM_VAR_Y] = S::NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
  }
  __syncthreads();
  const VarTable vtc = make_vartable_regs(p, args, S::NX, S::NU, S::NY);
  constexpr bool FO = FUSED_ONLY && S::NTERMS > 0;
  const bool fused = FO || (S::NTERMS > 0 && args.fused_terms);
  const int NT = t_hi - t_lo + 1;
  const int NTW = t_hi == T ? NT : NT - 1;              // rows worth computing: the x row of t_hi only matters at the end of the horizon
  const int64_t n = args.B * NTW;
  for (int64_t idx0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx0 < n; idx0 += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = idx0 / NTW;
    const int tt = int(idx0 - b * NTW), t = t_lo + tt;
    const int64_t idx = b * NT + tt;
    // the term code reads its atoms entry by entry (4-byte loads): pull the rows of this (trajectory, time) into L1 first
    asm volatile("prefetch.global.L1 [%0];\n" ::"l"(args.x_traj + (b * (T + 1) + t) * NX));
    if (t < T) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(args.u_traj + (b * T + t) * NU));
#pragma unroll
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      if (e < p.n_exo) {
        const int te = t < p.exo[e].steps ? t : p.exo[e].steps - 1;
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(args.exo[e] + (b * p.exo[e].steps + te) * p.exo[e].width));
      }
    }
    float gx[NX], gu[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) gx[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NU; ++j) gu[j] = 0.f;
    const bool want_x = t < t_hi || t_hi == T, want_u = t < t_hi;
    if (fused) {
      if constexpr (S::NTERMS > 0) {
        if (want_x) term_grad_s<S, NM_VAR_X, NX, NX, 1>(p, vtc, coef, b, t, gx);
        if (want_u) term_grad_s<S, NM_VAR_U, NU, NU, 1>(p, vtc, coef, b, t, gu);
      }
    } else if constexpr (!FO) {
      if (want_x) term_grad_accum<NX, 1>(p, vt, coef, b, NM_VAR_X, t, NX, gx);
      if (want_u) term_grad_accum<NU, 1>(p, vt, coef, b, NM_VAR_U, t, NU, gu);
    }
    if (want_x && args.gx_ext != nullptr) {
#pragma unroll
      for (int i = 0; i < NX; ++i) gx[i] += __ldg(args.gx_ext + (b * (T + 1) + t) * NX + i);
    }
    if (want_u && args.gu_ext != nullptr) {
#pragma unroll
      for (int j = 0; j < NU; ++j) gu[j] += __ldg(args.gu_ext + (b * T + t) * NU + j);
    }
    float* dst = G + idx * RW;
#pragma unroll
    for (int i = 0; i < NX; ++i) dst[i] = gx[i];
#pragma unroll
    for (int j = 0; j < NU; ++j) dst[NX + j] = gu[j];
  }
}

// Time-broadcast atoms (a single time entry compared against a whole range, e.g. x[:, [-1], :] > r - 0.01): the
// gradient of the anchored entry is a sum over the range -- one warp per trajectory, range strided over the lanes, fixed
// shuffle order (deterministic), added to the rows nm_wide_termgrad_kernel wrote.  `anchor` lists the anchored times.
struct WideTbAnchors { int n; int var[8]; int t[8]; };
template <class S, bool FUSED_ONLY = false>
__global__ void __launch_bounds__(256) nm_wide_termgrad_tb_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ BwdArgs args,
                                                                  const __grid_constant__ WideTbAnchors anchors, float* __restrict__ G, int t_lo, int t_hi) {
  constexpr int NX = S::NX, NU = S::NU, RW = NX + NU, MAXD = NX > NU ? NX : NU;
  __shared__ float coef[NM_MAX_TERMS];
  __shared__ VarTable vt;
  const NmPlan& p = plan;
  const int T = p.nsteps;
  if (threadIdx.x < NM_MAX_TERMS) coef[threadIdx.x] = threadIdx.x < p.n_terms ? term_coef(p, threadIdx.x, args.grad_scalars, args.b_total) : 0.f;
  if (threadIdx.x == 0) {
    vt.base[NM_VAR_X] = args.x_traj; vt.T[NM_VAR_X] = T + 1; vt.D[NM_VAR_X] = NX;
    vt.base[NM_VAR_U] = args.u_traj; vt.T[NM_VAR_U] = T; vt.D[NM_VAR_U] = NU;
    vt.base[NM_VAR_Y] = args.y_traj; vt.T[NM_VAR_Y] = T; vt.D[NM_VAR_Y] = S::NY;
    for (int e = 0; e < NM_MAX_EXO; ++e) {
      vt.base[NM_VAR_EXO0 + e] = args.exo[e]; vt.T[NM_VAR_EXO0 + e] = p.exo[e].steps; vt.D[NM_VAR_EXO0 + e] = p.exo[e].width;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, NT = t_hi - t_lo + 1;
  const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const VarTable vtc = make_vartable_regs(p, args, S::NX, S::NU, S::NY);
  constexpr bool FO = FUSED_ONLY && S::NTERMS > 0;
  const bool fused = FO || (S::NTERMS > 0 && args.fused_terms);
  for (int64_t b = warp0; b < args.B; b += nwarps) {
    for (int a = 0; a < anchors.n; ++a) {
      const int var = anchors.var[a], t = anchors.t[a], dim = var == NM_VAR_X ? NX : NU;
      float g[MAXD];
#pragma unroll
      for (int d = 0; d < MAXD; ++d) g[d] = 0.f;
      if (fused) {
        if constexpr (S::NTERMS > 0) {
          if (var == NM_VAR_X) term_grad_s<S, NM_VAR_X, NX, MAXD, 2>(p, vtc, coef, b, t, g, lane, 32);
          else term_grad_s<S, NM_VAR_U, NU, MAXD, 2>(p, vtc, coef, b, t, g, lane, 32);
        }
      } else if constexpr (!FO) term_grad_accum_tb_lanes<MAXD>(p, vt, coef, b, var, t, dim, lane, g);
#pragma unroll
      for (int d = 0; d < MAXD; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g[d] += __shfl_xor_sync(0xffffffffu, g[d], o);
      }
      if (lane == 0) {
        float* dst = G + (b * NT + (t - t_lo)) * RW + (var == NM_VAR_X ? 0 : NX);
        for (int d = 0; d < dim; ++d) dst[d] += g[d];
      }
    }
  }
}

template <class S>
__global__ void __launch_bounds__(wd::WideCfg<S>::BWD_THREADS, 1) nm_wide_bwd_kernel(const __grid_constant__ NmPlan plan, const __grid_constant__ WideBwdArgs wargs) {
  using C = wd::WideCfg<S>;
  using M = typename C::M;
  constexpr int NX = S::NX, NU = S::NU, K0 = C::K0, H = C::H, NL = C::NL, TILE = 128, NOUT = C::NOUT, TPT = C::BWD_TPT, HC = H / TPT;
  const BwdArgs& args = wargs.b;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::BAR_OFF);           // 0 weights, 1/2 operands ready (slot), 3/4 MMAs done (slot)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* dbred = reinterpret_cast<float*>(bars + 10);                       // [8 warps][NOP] output-bias gradient partials
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  const int T = p.nsteps;
  constexpr int RW = NX + NU;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 128 * TPT); mbar_init(&bars[2], 128 * TPT);
    mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], uint32_t(C::W_BYTES));
    tma_bulk_g2s(sm, args.wprep, uint32_t(C::W_BYTES), &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  const int t_lo = wargs.t_lo, t_hi = wargs.t_hi, CSn = t_hi - t_lo;
  const float gs = wargs.scale->gs, inv_gs = wargs.scale->inv_gs;

  {
    // ------------------------------------------------------------------ owners: TPT threads per trajectory (= TMEM lane).  Thread
    // `half` owns the columns [half * HC, (half + 1) * HC) of every layer epilogue; thread 0 ("primary") alone carries the
    // trajectory's scalars: co-state, dynamics adjoint, loss-term gradients, the policy input operand and the cotangent of
    // the raw policy output, which it hands to its partner through shared memory (the output-layer image region: unused here).
    const int slot = warp / (4 * TPT), half = (warp >> 2) % TPT, m_ = (warp & 3) * 32 + (tid & 31);
    const bool primary = half == 0;
    float* cotx = reinterpret_cast<float*>(sm + C::imgo_off()) + (slot * 128 + m_) * 8;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t cb = tmem + lane_base + slot * C::SLOT_COLS;
    const uint32_t t_hi_ = cb + C::A_HI, t_lo_ = cb + C::A_LO, t_d = cb + C::DCOL;
    uint64_t* bar_ops = &bars[1 + slot];
    uint64_t* bar_done = &bars[3 + slot];
    uint32_t ph_done = 0, ph_ops = 0;
    const float act_prm = p.policy.act_param;
    const bool issuer = (warp % (4 * TPT)) == 0;
    const uint32_t sbase = smem_u32(sm), cbs = tmem + slot * C::SLOT_COLS;
    // operands of this thread are in tensor memory; the slot's first warp waits for all 128 rows and issues MMA batch `bi`:
    //   0: recompute layer 0    1 .. NL-2: recompute hidden layer bi    NL-1 .. 2NL-4: data gradient of hidden layer
    //   l = 2NL-3-bi (dA_l = dZ_l W_l, the forward image read MN-major)    2NL-3: data gradient of layer 0 (state columns)
    auto arrive_ops = [&](int bi) {
      tc::wait_st(); tc::fence_before(); mbar_arrive(bar_ops);
      if (issuer) {
        mbar_wait(bar_ops, ph_ops); ph_ops ^= 1;
        tc::fence_after();
        if (tc::elect_one()) {
          if (bi == 0) wd::issue_kmajor<C::K0P, H>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::img0_off());
          else if (bi < NL - 1) wd::issue_kmajor<H, H>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::imgh_off(1) + (bi - 1) * 2 * C::IMGH_BYTES);
          else if (bi < 2 * NL - 3) wd::issue_mnmajor<H, H, H>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::imgh_off(1) + (2 * NL - 3 - bi - 1) * 2 * C::IMGH_BYTES);
          else wd::issue_mnmajor<H, C::K0P, C::GWP>(cbs + C::DCOL, cbs + C::A_HI, cbs + C::A_LO, sbase + C::img0_off());
          tc::commit(bar_done);
        }
        __syncwarp();
      }
    };
    wd::float8* stash = reinterpret_cast<wd::float8*>(reinterpret_cast<uint8_t*>(wargs.stash) + size_t(blockIdx.x * 2 + slot) * C::STASH_SLOT_BYTES) + m_;
    const float* __restrict__ Wout = reinterpret_cast<const float*>(sm + C::wout_off());
    float dbo[NOUT];                                  // output-bias gradient (scaled), summed over this thread's trajectory-steps
#pragma unroll
    for (int j = 0; j < NOUT; ++j) dbo[j] = 0.f;
    float dzmax = 0.f;
    WP_DECL(w_in); WP_DECL(w_dyn); WP_DECL(w_wait_rc); WP_DECL(w_epi_rc); WP_DECL(w_wait_dg); WP_DECL(w_epi_dg); WP_DECL(w_wait_in); WP_DECL(w_arr); WP_DECL(w_tile);
#ifdef NM_TUNE_WPROF
    const long long w_begin = clock64();
#endif
    // Epilogue chunk: CW accumulator columns at a time -- 32 with one thread per trajectory, 16 with two (128 registers:
    // v / d / hi / lo / the stash double buffer of a 32-column chunk are 128 registers by themselves)
    constexpr int CW = TPT == 1 ? 32 : 16;
    // staging store of CW consecutive features (CW / 2 packed words) of operand tile `toff`
    auto stage_cw = [&](uint8_t* item, int toff, int c0, const uint32_t (&hi)[CW / 2]) {
#ifndef NM_TUNE_WNOSTAGE                               // (timing experiment: the adjoint without its staging stores -- results are wrong)
      uint8_t* q = item + toff + (c0 / 8) * C::GRP + m_ * 16;
#pragma unroll
      for (int g = 0; g < CW / 8; ++g) wd::st_global_v4(q + g * C::GRP, hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
#endif
    };

    // Everything a step reads from HBM is requested one step ahead -- across tile boundaries too (with short chunks a
    // tile is only CSn steps long): rows of step t-1 of this trajectory, or the first rows of the next tile's trajectory
    // One thread per trajectory has the registers to hold the next step's rows (REGPF); with two threads per trajectory
    // (128 registers) the rows are only pulled into L2 ahead of time and loaded at the top of their step.
    constexpr bool REGPF = TPT == 1;
    float lam_pre[NX] = {}, x_pre[NX] = {}, o_pre[NU] = {}, in_pre[K0] = {}, g_pre[RW] = {};
    auto l2_hint = [&](const float* q, int n) {
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q));
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q + n - 1));
    };
    auto load_rows = [&](int64_t bq, int tq) {
      wd::ld_row<NX>(x_pre, args.x_traj + (bq * (T + 1) + tq) * NX);
      wd::ld_row<NU>(o_pre, args.chk + (bq * T + tq) * NU);
      wd::ld_row<RW>(g_pre, wargs.G + (bq * int64_t(wargs.g_nt) + (tq - wargs.g_lo)) * RW);
      wd::policy_exo<S, K0>(p, args.exo, bq, tq, in_pre);
    };
    auto load_lam = [&](int64_t bq) {
      if (t_hi == T) {
        float gT[RW];
        wd::ld_row<RW>(gT, wargs.G + (bq * int64_t(wargs.g_nt) + (T - wargs.g_lo)) * RW);
#pragma unroll
        for (int i = 0; i < NX; ++i) lam_pre[i] = gT[i];
      } else wd::ld_row<NX>(lam_pre, wargs.lam_carry + bq * NX);
    };
    auto prefetch_rows = [&](int64_t bq, int tq) {
      if constexpr (REGPF) load_rows(bq, tq);
      else {
        l2_hint(args.x_traj + (bq * (T + 1) + tq) * NX, NX);
        l2_hint(args.chk + (bq * T + tq) * NU, NU);
        l2_hint(wargs.G + (bq * int64_t(wargs.g_nt) + (tq - wargs.g_lo)) * RW, RW);
        static_for<0, S::NSEG>([&](auto sc) {
          constexpr int sg = decltype(sc)::value;
          if constexpr (S::seg_kind(sg) == NM_SEG_EXO) {
            constexpr int e = S::seg_index(sg), W = S::seg_width(sg);
            l2_hint(args.exo[e] + (bq * p.exo[e].steps + tq) * W, W);
          }
        });
      }
    };
    auto prefetch_tile = [&](int tile_q) {                 // first step of a tile + its incoming co-state
      const int64_t bq_raw = int64_t(tile_q) * TILE + m_;
      const int64_t bq = bq_raw < args.B ? bq_raw : args.B - 1;
      prefetch_rows(bq, t_hi - 1);
      if constexpr (REGPF) load_lam(bq);
      else l2_hint(t_hi == T ? wargs.G + (bq * int64_t(wargs.g_nt) + (T - wargs.g_lo)) * RW : wargs.lam_carry + bq * NX, NX);
    };
    if (primary && int(blockIdx.x) * 2 + slot < args.num_tiles) prefetch_tile(int(blockIdx.x) * 2 + slot);
    for (int r = 0;; ++r) {
      const int tile = (r * int(gridDim.x) + int(blockIdx.x)) * 2 + slot;
      if (tile >= args.num_tiles) break;
      const int tile_next = tile + 2 * int(gridDim.x);
      const int64_t b_raw = int64_t(tile) * TILE + m_;
      const bool valid = b_raw < args.B;
      const int64_t b = valid ? b_raw : args.B - 1;
      if constexpr (!REGPF) { if (primary) load_lam(b); }
      float lam[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) lam[i] = valid ? lam_pre[i] : 0.f;
#pragma unroll 1
      for (int t = t_hi - 1; t >= t_lo; --t) {
        if constexpr (!REGPF) { if (primary) load_rows(b, t); }
#ifdef NM_TUNE_WSTWRAP                                 // (timing experiment: staging wrapped into an L2-sized window -- results are wrong)
        uint8_t* item = wargs.staging + ((size_t(tile) * CSn + size_t(t - t_lo)) % NM_TUNE_WSTWRAP) * C::ITEM_BYTES;
#else
        uint8_t* item = wargs.staging + (size_t(tile) * CSn + size_t(t - t_lo)) * C::ITEM_BYTES;
#endif
        float x[NX], o_cur[NU], in[K0], g_cur[RW], yv[1] = {0.f};
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = x_pre[i];
#pragma unroll
        for (int j = 0; j < NU; ++j) o_cur[j] = o_pre[j];
#pragma unroll
        for (int q = 0; q < K0; ++q) in[q] = in_pre[q];
#pragma unroll
        for (int q = 0; q < RW; ++q) g_cur[q] = g_pre[q];
        tc_policy_state<S, K0>(x, yv, in);
        if (primary) {
          uint32_t hi0[C::K0P / 2];
          wd::put_input<C>(t_hi_, t_lo_, in, hi0);
          uint8_t* q = item + C::a_off(0) + m_ * 16;
#pragma unroll
          for (int g = 0; g < C::K0P / 8; ++g) wd::st_global_v4(q + g * C::GRP, hi0[4 * g], hi0[4 * g + 1], hi0[4 * g + 2], hi0[4 * g + 3]);
        }
        { WP_T0(); arrive_ops(0); WP_ADD(w_arr); }   // -> recompute of layer 0
        WP_T0();
        float gx[NX], gtx[NX], cot[NOUT];
#pragma unroll
        for (int i = 0; i < NX; ++i) { gx[i] = 0.f; gtx[i] = 0.f; }
#pragma unroll
        for (int j = 0; j < NOUT; ++j) cot[j] = 0.f;
        if (primary) {
        if (t > t_lo) prefetch_rows(b, t - 1);
        else if (tile_next < args.num_tiles) prefetch_tile(tile_next);
        // ---- dynamics adjoint, loss-term gradients, cotangent of the raw policy output (all while the MMAs run)
        float u[NU], gu[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) u[j] = bounds_fwd<M::BOUNDS>(o_cur[j], p.policy.bmin[j], p.policy.bmax[j]);
        Dyn<S>::step_vjp(p, x, u, lam, gx, gu);
#pragma unroll
        for (int i = 0; i < NX; ++i) gtx[i] = g_cur[i];
#pragma unroll
        for (int j = 0; j < NU; ++j) gu[j] += g_cur[NX + j];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) {
          cot[j] = valid ? gu[j] * bounds_deriv<M::BOUNDS>(o_cur[j], p.policy.bmin[j], p.policy.bmax[j]) * gs : 0.f;
          dbo[j] += cot[j];
          dzmax = fmaxf(dzmax, fabsf(cot[j]));
        }
        {                                             // cotangent tile (NOP halves): weight gradient of the output layer
          uint32_t ch[C::NOP / 2];
#pragma unroll
          for (int i = 0; i < C::NOP / 2; ++i) {
            const float c0 = 2 * i < NOUT ? cot[2 * i < NOUT ? 2 * i : 0] : 0.f, c1 = 2 * i + 1 < NOUT ? cot[2 * i + 1 < NOUT ? 2 * i + 1 : 0] : 0.f;
            const __half2 h2 = __floats2half2_rn(c0, c1);
            ch[i] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          uint8_t* q = item + C::c_off() + m_ * 16;
#pragma unroll
          for (int g = 0; g < C::NOP / 8; ++g) wd::st_global_v4(q + g * C::GRP, ch[4 * g], ch[4 * g + 1], ch[4 * g + 2], ch[4 * g + 3]);
        }
        if constexpr (TPT > 1) {                      // -> partner thread (read after the batch-1 arrivals: release / acquire through bar_ops)
#pragma unroll
          for (int j = 0; j < NOUT; ++j) cotx[j] = cot[j];
        }
        }
        WP_ADD(w_dyn);
        // ---- recompute: a_{l+1} = act(z_l), derivative stash; the last one goes straight into the backward pass
#pragma unroll 1
        for (int l = 0; l < NL - 1; ++l) {
          const float* __restrict__ bias = reinterpret_cast<const float*>(sm + C::bias_off(0)) + l * H;
          const bool last = l == NL - 2;
          { WP_T0(); mbar_wait(bar_done, ph_done); ph_done ^= 1; WP_ADD(w_wait_rc); }
          tc::fence_after();
          WP_T0();
#pragma unroll 1
          for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += CW) {
            float v[CW], d[CW];
            uint32_t hi[CW / 2], lo[CW / 2];
            tc::tmem_ld<CW>(t_d + c0, v);
            wd::act_group<M::ACT, 0, 8, true>(v, d, bias + c0, act_prm);
            if constexpr (CW == 32) wd::act_group<M::ACT, 16, 8, true>(v, d, bias + c0, act_prm);
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) wd::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            stage_cw(item, C::a_off(1) + l * (H / 8) * C::GRP, c0, hi);                   // a_{l+1}
            if (!last) {
              wd::tmem_stu<CW / 2>(t_hi_ + c0 / 2, hi);
              wd::tmem_stu<CW / 2>(t_lo_ + c0 / 2, lo);
#ifndef NM_TUNE_WNOSTASH                               // (timing experiment: no derivative stash traffic -- results are wrong)
#pragma unroll
              for (int i = 0; i < CW; i += 8) wd::st_stash8(stash + (l * (H / 8) + (c0 + i) / 8) * 128, d + i);
#endif
            } else {
              // output layer backward on the CUDA cores: dA[k] = sum_j cot[j] W_out[j][k];  dZ_{NL-2} = dA * act'
#pragma unroll
              for (int i = 0; i < CW; i += 4) {
                float da0 = 0.f, da1 = 0.f, da2 = 0.f, da3 = 0.f;
#pragma unroll
                for (int j = 0; j < NOUT; ++j) {
                  const float4 w = *reinterpret_cast<const float4*>(Wout + j * H + c0 + i);
                  da0 = fmaf(cot[j], w.x, da0); da1 = fmaf(cot[j], w.y, da1); da2 = fmaf(cot[j], w.z, da2); da3 = fmaf(cot[j], w.w, da3);
                }
                v[i] = da0 * d[i]; v[i + 1] = da1 * d[i + 1]; v[i + 2] = da2 * d[i + 2]; v[i + 3] = da3 * d[i + 3];
              }
              if (c0 == half * HC) {
#pragma unroll
                for (int i = 0; i < CW; ++i) dzmax = fmaxf(dzmax, fabsf(v[i]));
              }
#pragma unroll
              for (int i = 0; i < CW / 2; ++i) wd::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
              stage_cw(item, C::z_off(NL - 2), c0, hi);                                  // dZ_{NL-2}
              wd::tmem_stu<CW / 2>(t_hi_ + c0 / 2, hi);
              wd::tmem_stu<CW / 2>(t_lo_ + c0 / 2, lo);
            }
          }
          WP_ADD(w_epi_rc);
          { WP_T0(); arrive_ops(l + 1); WP_ADD(w_arr); }   // -> recompute of layer l+1, or (last: bi = NL-1) data gradient of layer NL-2
          if constexpr (TPT > 1) {
            if (l == 0 && !primary) {                 // all arrivals of batch 1 (odd phase of bar_ops: 2 (NL - 1) batches per step) are in:
              mbar_wait(bar_ops, 1);                  // the primary wrote the cotangent before it arrived
#pragma unroll
              for (int j = 0; j < NOUT; ++j) cot[j] = cotx[j];
            }
          }
        }
        // ---- data gradients: D = dA_l (cotangent of a_l), l = NL-2 .. 1;  dZ_{l-1} = dA_l * act'(z_{l-1})
#pragma unroll 1
        for (int l = NL - 2; l >= 1; --l) {
          // the stashed derivatives come from L2: two CW-column chunks are always in flight -- chunks 0 and 1 are
          // requested before the wait for the MMAs, chunk c+2 when chunk c has been consumed
          float dq[2][CW];
          const wd::float8* __restrict__ sl = stash + ((l - 1) * (H / 8) + half * (HC / 8)) * 128;
#pragma unroll
          for (int i = 0; i < CW / 8; ++i) wd::ld_stash8(sl + i * 128, dq[0] + 8 * i);
#pragma unroll
          for (int i = 0; i < CW / 8; ++i) wd::ld_stash8(sl + (CW / 8 + i) * 128, dq[1] + 8 * i);
          { WP_T0(); mbar_wait(bar_done, ph_done); ph_done ^= 1; WP_ADD(w_wait_dg); }
          tc::fence_after();
          WP_T0();
#pragma unroll 1
          for (int c2 = 0; c2 < HC; c2 += 2 * CW) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              const int cl = c2 + hb * CW, c0 = half * HC + cl;      // column within this thread's range / within the layer
              float v[CW];
              uint32_t hi[CW / 2], lo[CW / 2];
              tc::tmem_ld<CW>(t_d + c0, v);
#pragma unroll
              for (int i = 0; i < CW; ++i) v[i] *= dq[hb][i];
              if (cl + 2 * CW < HC) {
#pragma unroll
                for (int i = 0; i < CW / 8; ++i) wd::ld_stash8(sl + ((cl + 2 * CW) / 8 + i) * 128, dq[hb] + 8 * i);
              }
              if (cl == 0) {
#pragma unroll
                for (int i = 0; i < CW; ++i) dzmax = fmaxf(dzmax, fabsf(v[i]));
              }
#pragma unroll
              for (int i = 0; i < CW / 2; ++i) wd::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
              stage_cw(item, C::z_off(l - 1), c0, hi);                                    // dZ_{l-1}
              wd::tmem_stu<CW / 2>(t_hi_ + c0 / 2, hi);
              wd::tmem_stu<CW / 2>(t_lo_ + c0 / 2, lo);
            }
          }
          WP_ADD(w_epi_dg);
          { WP_T0(); arrive_ops(2 * NL - 2 - l); WP_ADD(w_arr); }   // -> data gradient of layer l-1
        }
        // ---- gradient of the policy input (state segment) closes the loop
        { WP_T0(); mbar_wait(bar_done, ph_done); ph_done ^= 1; WP_ADD(w_wait_in); }
        tc::fence_after();
        if (primary) {
          float gin[C::GWP];
#pragma unroll
          for (int c0 = 0; c0 < C::GWP; c0 += 16) {
            float g16[16];
            tc::tmem_ld<16>(t_d + c0, g16);
#pragma unroll
            for (int i = 0; i < 16; ++i) gin[c0 + i] = g16[i] * inv_gs;
          }
          int off = 0;
          static_for<0, S::NSEG>([&](auto sc) {
            constexpr int sg = decltype(sc)::value;
            if constexpr (S::seg_kind(sg) == NM_SEG_STATE) {
#pragma unroll
              for (int i = 0; i < NX; ++i) gx[i] += gin[off + i];
            }
            off += S::seg_width(sg);
          });
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) lam[i] = valid ? gx[i] + gtx[i] : 0.f;
      }
      if (valid && primary) {
        if (t_lo == 0) {
          if (args.grad_x0 != nullptr) {
#pragma unroll
            for (int i = 0; i < NX; ++i) args.grad_x0[b * NX + i] = lam[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < NX; ++i) wargs.lam_carry[b * NX + i] = lam[i];
        }
      }
    }
#ifdef NM_TUNE_WPROF
    if (blockIdx.x == 0 && (tid == 0 || tid == 128 || tid == 32))
      printf("[wprof bwd tid %d] total %lld | arrive+issue %lld  dyn/terms %lld  wait_rc %lld  epi_rc %lld  wait_dg %lld  epi_dg %lld  wait_in %lld (steps %d)\n", tid,
             clock64() - w_begin, w_arr, w_dyn, w_wait_rc, w_epi_rc, w_wait_dg, w_epi_dg, w_wait_in, CSn);
#endif
    // ---- output-bias gradient and the scale statistics of this launch (fixed order: deterministic)
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      float v = dbo[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0 && primary) dbred[(slot * 4 + (warp & 3)) * C::NOP + j] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dzmax = fmaxf(dzmax, __shfl_xor_sync(0xffffffffu, dzmax, o));
    if ((tid & 31) == 0) atomicMax(&wargs.scale->max_bits, __float_as_uint(dzmax));
    wd::named_bar_sync(1, C::BWD_THREADS);
    if (tid < NOUT && C::BIAS && p.policy.trainable) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += dbred[w * C::NOP + tid];
      float* gb = args.grad_partials + size_t(blockIdx.x) * size_t(p.n_params) + p.policy.param_offset + C::p_off(NL - 1) + NOUT * H + tid;
      *gb += s * inv_gs;
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// next chunk's cotangent scale: the largest scaled cotangent seen goes to ~2^6 (fp16 overflows at 2^16)
static __global__ void nm_wide_scale_update_kernel(wd::ScaleState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float m = __uint_as_float(st->max_bits);
    if (m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);                                  // m = f * 2^e, f in [0.5, 1)
      int shift = 6 - e;
      shift = shift > 40 ? 40 : (shift < -40 ? -40 : shift);
      const float g = ldexpf(st->gs, shift);
      if (g > 1e-30f && g < 1e30f) { st->gs = g; st->inv_gs = 1.f / g; }
    }
    st->max_bits = 0u;
  }
}
static __global__ void nm_wide_scale_init_kernel(wd::ScaleState* st, float gs0) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { st->gs = gs0; st->inv_gs = 1.f / gs0; st->max_bits = 0u; st->pad = 0u; }
}

// ================================================================================================ weight-gradient GEMM
// One work item = one staged tile-step.  warp 0: bulk-TMA producer (per layer unit: the dZ tile and the input tile into a
// 3-slot ring); warp 1: MMA issuer, D_l[n][k] += sum_m dZ_l[m][n] a_l[m][k] (both operands MN-major, contraction over the
// 128 trajectories, 8 MMAs of K = 16) plus the bias sums against a constant tile whose feature 0 is 1; warps 2..5: every
// WG_FLUSH_ITEMS items (the TMEM accumulator truncates: chains stay short) add the accumulators, un-scaled, to this
// CTA's private gradient block (plain read-modify-write: nobody else touches it).
template <class S>
__global__ void __launch_bounds__(wd::WideCfg<S>::WG_THREADS, 1) nm_wide_wgrad_kernel(const __grid_constant__ NmPlan plan, const uint8_t* __restrict__ staging,
                                                                                      int64_t n_items, const wd::ScaleState* __restrict__ scale,
                                                                                      float* __restrict__ grad_partials) {
  using C = wd::WideCfg<S>;
  constexpr int NL = C::NL, H = C::H, FL = C::WG_FLUSH_ITEMS;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::WG_BAR_OFF);        // 0..2 full, 3..5 empty, 6 accumulators ready, 7 accumulators free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const NmPlan& p = plan;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[3 + i], 1); }
    mbar_init(&bars[6], 1); mbar_init(&bars[7], 128);
    mbar_fence_init();
  }
  // constant tile [2 groups][128][8] halves: feature 0 = 1
  for (int i = tid; i < 2 * C::GRP / 4; i += C::WG_THREADS) reinterpret_cast<uint32_t*>(sm + C::WG_ONES_OFF)[i] = 0u;
  __syncthreads();
  if (tid < 128) *reinterpret_cast<__half*>(sm + C::WG_ONES_OFF + tid * 16) = __float2half_rn(1.f);
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  int64_t my_items = 0;
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) ++my_items;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t ph_empty[3] = {0, 0, 0};
    int64_t q = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint8_t* item = staging + size_t(it) * C::ITEM_BYTES;
      static_for<0, NL>([&](auto lc) {
        constexpr int l = decltype(lc)::value;
        const int slot = int(q % 3);
        if (q >= 3) { mbar_wait(&bars[3 + slot], ph_empty[slot]); ph_empty[slot] ^= 1; }
        uint8_t* dst = sm + slot * C::WG_SLOT_BYTES;
        if (tc::elect_one()) {
          if constexpr (l < NL - 1) {
            constexpr uint32_t za = (H / 8) * C::GRP, ab = (l == 0 ? C::K0P / 8 : H / 8) * C::GRP;
            mbar_expect_tx(&bars[slot], za + ab);
            tma_bulk_g2s(dst, item + C::z_off(l), za, &bars[slot]);
            tma_bulk_g2s(dst + (H / 8) * C::GRP, item + C::a_off(l), ab, &bars[slot]);
          } else {
            constexpr uint32_t aa = (H / 8) * C::GRP, cb = (C::NOP / 8) * C::GRP;
            mbar_expect_tx(&bars[slot], aa + cb);
            tma_bulk_g2s(dst, item + C::a_off(NL - 1), aa, &bars[slot]);
            tma_bulk_g2s(dst + (H / 8) * C::GRP, item + C::c_off(), cb, &bars[slot]);
          }
        }
        __syncwarp();
        ++q;
      });
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t ph_full[3] = {0, 0, 0}, ph_free = 0;
    int64_t q = 0, done = 0;
    const uint32_t sbase = smem_u32(sm);
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint32_t fresh = (done % FL) == 0 ? 0u : 1u;
      static_for<0, NL>([&](auto lc) {
        constexpr int l = decltype(lc)::value;
        const int slot = int(q % 3);
        mbar_wait(&bars[slot], ph_full[slot]); ph_full[slot] ^= 1;
        tc::fence_after();
        const uint32_t a0 = sbase + slot * C::WG_SLOT_BYTES, b0 = a0 + (H / 8) * C::GRP;
        // MN-major tiles X[f/8][128][8]: 8-feature groups GRP apart (SBO), 8-trajectory groups 128 B apart (LBO); K step = 16 trajectories = 256 B
        const uint64_t da = tc::smem_desc(a0, 128, C::GRP), db = tc::smem_desc(b0, 128, C::GRP);
        const uint64_t dones = tc::smem_desc(sbase + C::WG_ONES_OFF, 128, C::GRP);
        if (tc::elect_one()) {
          constexpr int N = l == 0 ? C::K0P : (l < NL - 1 ? H : C::NOP);
          constexpr uint32_t idesc = wd::idesc_f16(128, N, 1, 1), idesc_b = wd::idesc_f16(128, 16, 1, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) wd::mma_ss(tmem + C::wg_col(l), tc::desc_add(da, ks * 256), tc::desc_add(db, ks * 256), idesc, ks > 0 ? 1u : fresh);
          if constexpr (l < NL - 1 && C::BIAS) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) wd::mma_ss(tmem + C::wgb_col(l), tc::desc_add(da, ks * 256), tc::desc_add(dones, ks * 256), idesc_b, ks > 0 ? 1u : fresh);
          }
          tc::commit(&bars[3 + slot]);
        }
        __syncwarp();
        ++q;
      });
      ++done;
      if ((done % FL) == 0 || done == my_items) {
        if (tc::elect_one()) tc::commit(&bars[6]);
        __syncwarp();
        mbar_wait(&bars[7], ph_free); ph_free ^= 1;
        tc::fence_after();
      }
    }
  } else {
    // ------------------------------------------------------------------ flush warps: thread = accumulator lane
    const int q4 = warp & 3, lane_n = q4 * 32 + (tid & 31);
    const uint32_t lane_base = uint32_t(q4 * 32) << 16;
    uint32_t ph_rdy = 0;
    const float inv_gs = scale->inv_gs;
    float* gblock = grad_partials + size_t(blockIdx.x) * size_t(p.n_params) + p.policy.param_offset;
    const bool train = p.policy.trainable != 0;
    int64_t done = 0;
    while (done < my_items) {
      done = done + FL < my_items ? done + FL : my_items;
      mbar_wait(&bars[6], ph_rdy); ph_rdy ^= 1;
      tc::fence_after();
      if (train) {
        static_for<0, NL - 1>([&](auto lc) {
          constexpr int l = decltype(lc)::value;
          constexpr int Kl = C::kin(l);
          float* dst = gblock + C::p_off(l) + lane_n * Kl;
          // this CTA's gradient block is private and every address has one owner thread: the order of the adds -- and the
          // result -- is fixed whether they are read-modify-writes or reductions at the L2
#ifdef NM_TUNE_WNORED
          const bool pairs = false;
#else
          const bool pairs = Kl % 2 == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
#endif
#pragma unroll 1
          for (int c0 = 0; c0 < Kl; c0 += 8) {
            float v[8];
            tc::tmem_ld<8>(tmem + lane_base + C::wg_col(l) + c0, v);
            if (pairs) {
#pragma unroll
              for (int i = 0; i < 8; i += 2) if (c0 + i < Kl) wd::red_add2(dst + c0 + i, v[i] * inv_gs, v[i + 1] * inv_gs);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (c0 + i < Kl) dst[c0 + i] += v[i] * inv_gs;
            }
          }
          if constexpr (C::BIAS) {
            float v[8];
            tc::tmem_ld<8>(tmem + lane_base + C::wgb_col(l), v);
            gblock[C::p_off(l) + H * Kl + lane_n] += v[0] * inv_gs;
          }
        });
        {   // output layer: accumulator lanes = input features k, columns = outputs j
          float v[16];
          tc::tmem_ld<16>(tmem + lane_base + C::wg_col(NL - 1), v);
          float* dst = gblock + C::p_off(NL - 1);
#pragma unroll
          for (int j = 0; j < C::NOUT; ++j) dst[j * H + lane_n] += v[j] * inv_gs;
        }
      }
      tc::fence_before();
      mbar_arrive(&bars[7]);
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

}  // namespace nm
#endif  // NM_HOST_EMU
