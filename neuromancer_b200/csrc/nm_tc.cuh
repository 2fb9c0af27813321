// nm_tc.cuh -- tcgen05 (5th-gen tensor core) building blocks of the neural-ODE rollout: MLP right-hand
// sides whose hidden layers are wide enough for the tensor pipe to pay.
//
// Measured on B200 before this was designed (tools/microbench/tc_*.cu, profiles/r01_tcgen05_findings.md):
//   * kind::tf32 MMAs read 32-bit operands and ignore the low 13 mantissa bits; the fp32 accumulator in
//     tensor memory TRUNCATES on every accumulate (error grows linearly, always toward zero).  fp32-level
//     results need the 3-product split  A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  with the two small products issued
//     FIRST (while the accumulator is small their truncation is free): 2e-7 relative for K = 64.
//   * one M=128 MMA costs max(44, N/2) cycles whatever K-step it is, from shared or tensor memory: layers
//     narrower than ~48 outputs cannot beat the FFMA path, so those layers stay on the CUDA cores.
//   * MN-major tf32 operands are only accepted in the 128B_BASE32B swizzle, K-major ones in the classic
//     formats: one shared-memory tile can never serve both a feature contraction and a trajectory
//     contraction.  Here the row operand (activations of this thread's trajectory) lives in TENSOR MEMORY
//     (thread = TMEM lane = trajectory; written with tcgen05.st, consumed as the A operand), the weights in
//     shared memory as K-major "interleaved" (no-swizzle) tiles  W[k/4][n][k%4].
//
// Layer classes of an MLP with sizes s_0 .. s_NL (NL >= 3):
//   layer 0      (tiny fan-in: state + input)   CUDA cores, thread-per-trajectory, weights broadcast from smem
//   layers 1..NL-2                               tensor cores, M = 128 trajectories, N = s_{l+1} (pad 16), K = s_l + 1 (bias column, pad 8)
//   layer NL-1   (tiny fan-out: state width)     CUDA cores, fused into the epilogue of the layer before
#pragma once
#include "nm_device.cuh"

namespace nm {
namespace tc {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;                                   // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);   // D f32, A/B tf32, both K-major
}
// descriptor of the same tile family `byte_off` bytes further (the start-address field holds address >> 4)
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t byte_off) { return d + uint64_t(byte_off >> 4); }
// (NM_HOST_EMU: tests/emu/tc_ptx_emu.h models tensor memory, the MMAs and the barriers on the host)
#ifdef NM_HOST_EMU
}  // namespace tc
}  // namespace nm
#include "tc_ptx_emu.h"
namespace nm {
namespace tc {
#else
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]; compile-time accumulate flag (no setp in the issue loop)
template <bool ACC>
__device__ __forceinline__ void mma_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void st_shared(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }
// one lane of a converged warp (the tensor-core instructions live on the uniform datapath: issuing them from
// `if (tid == k)` makes the compiler wrap each one in a per-thread vote loop -- measured 94 instead of 46 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot_smem)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(COLS));
}

// CH consecutive 32-bit columns of this thread's TMEM lane  <->  registers   (CH in {8, 16, 32})
template <int CH> __device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[CH]);
template <int CH> __device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[CH]);
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  wait_ld();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <> __device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
template <> __device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
               "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
template <> __device__ __forceinline__ void tmem_st<32>(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::
          "r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
      "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])),
      "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
      "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}

// round-to-nearest tf32 part of v and the exact remainder (the tensor core truncates the remainder to tf32)
__device__ __forceinline__ float tf32_hi(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

#endif

constexpr int pow2_ceil(int v) { int p = 32; while (p < v) p *= 2; return p; }
constexpr int chunk_of(int w) { return w % 32 == 0 ? 32 : (w % 16 == 0 ? 16 : 8); }
constexpr int chunk16_of(int w) { return w % 16 == 0 ? 16 : 8; }     // adjoint kernel: smaller register footprint

// ------------------------------------------------------------------------------------------------
// Static layout of one MLP on the tensor-core path
template <class M>
struct TcCfg {
  static constexpr int NL = M::NL;
  static constexpr bool BIAS = M::BIAS;
  static constexpr int K(int l) { return M::sz(l); }
  static constexpr int N(int l) { return M::sz(l + 1); }
  static constexpr int KW(int l) { return l == 0 ? KW0Of<M>::value : M::sz(l); }   // fan-in in the parameter layout (HoistedMlp)
  static constexpr bool is_tc(int l) { return l >= 1 && l <= NL - 2; }
  static constexpr int kaug(int l) { return K(l) + (BIAS ? 1 : 0); }      // + constant-one column carrying the bias
  static constexpr int kp(int l) { return round_up(kaug(l), 8); }          // MMA K steps of 8
  static constexpr int np(int l) { return round_up(N(l), 16); }            // MMA N (M = 128 needs N % 16 == 0)
  static constexpr int kp_max() { int m = 8; for (int l = 1; l <= NL - 2; ++l) m = cmax(m, kp(l)); return m; }
  static constexpr int np_max() { int m = 16; for (int l = 1; l <= NL - 2; ++l) m = cmax(m, np(l)); return m; }
  // eligibility: tiny first fan-in / last fan-out, hidden layers within the tensor-memory budget
  static constexpr bool shape_ok() {
    if (NL < 3 || K(0) > 8 || N(NL - 1) > 8) return false;
    for (int l = 1; l <= NL - 2; ++l) if (K(l) < 8 || N(l) < 8 || kp(l) > 136 || np(l) > 128) return false;
    return true;
  }
  static constexpr int max_hidden() { int m = 0; for (int l = 1; l < NL; ++l) m = cmax(m, K(l)); return m; }
  // tensor-memory columns (forward kernel): operand hi | operand lo | accumulator (wide enough to be read as
  // the next operand: kp of the following layer)
  static constexpr int OPH = 0, OPL = kp_max(), ACC = 2 * kp_max();
  static constexpr int ACC_COLS = cmax(np_max(), kp_max());
  static constexpr int FWD_COLS = pow2_ceil(ACC + ACC_COLS);
  // shared-memory weight block of the forward kernel (floats):
  //   first layer  rows [N(0)][R0]   R0 = round_up(K(0)+1, 4): weights then bias
  //   last layer   rows [N(NL-1)][RL] RL = round_up(K(NL-1), 4), then bias[round_up(N, 4)]
  //   TC layer l   hi image [kp/4][np][4], lo image, 128-byte aligned
  static constexpr int R0 = round_up(K(0) + 1, 4);
  static constexpr int RL = round_up(K(NL - 1), 4);
  static constexpr int first_off() { return 0; }
  static constexpr int last_off() { return round_up(N(0) * R0, 32); }
  static constexpr int lastb_off() { return last_off() + N(NL - 1) * RL; }
  static constexpr int img_floats(int l) { return np(l) * kp(l); }
  static constexpr int img_off(int l) {                                    // hi image of TC layer l (lo follows)
    int o = round_up(lastb_off() + round_up(N(NL - 1), 4), 32);
    for (int i = 1; i < l; ++i) o += 2 * img_floats(i);
    return o;
  }
  static constexpr int FWD_W = img_off(NL - 1);                            // total floats
};

// ------------------------------------------------------------------------------------------------
// weight preparation for the forward kernel (one launch per call, tiny)
template <class M>
__device__ void prep_tc_fwd(const float* __restrict__ params, float* __restrict__ dst, int gtid, int gthreads) {
  using C = TcCfg<M>;
  auto p_off = [](int l) { int o = 0; for (int i = 0; i < l; ++i) o += C::N(i) * C::KW(i) + (C::BIAS ? C::N(i) : 0); return o; };
  {   // first layer
    const float* W = params + p_off(0);
    const float* b = W + C::N(0) * C::KW(0);
    for (int idx = gtid; idx < C::N(0) * C::R0; idx += gthreads) {
      const int n = idx / C::R0, c = idx % C::R0;
      dst[C::first_off() + idx] = c < C::K(0) ? W[n * C::KW(0) + c] : (c == C::K(0) && C::BIAS ? b[n] : 0.f);
    }
  }
  {   // last layer
    constexpr int l = C::NL - 1;
    const float* W = params + p_off(l);
    const float* b = W + C::N(l) * C::K(l);
    for (int idx = gtid; idx < C::N(l) * C::RL; idx += gthreads) {
      const int n = idx / C::RL, c = idx % C::RL;
      dst[C::last_off() + idx] = c < C::K(l) ? W[n * C::K(l) + c] : 0.f;
    }
    for (int idx = gtid; idx < round_up(C::N(l), 4); idx += gthreads) dst[C::lastb_off() + idx] = (idx < C::N(l) && C::BIAS) ? b[idx] : 0.f;
  }
  static_for<1, C::NL - 1>([&](auto lc) {
    constexpr int l = decltype(lc)::value;
    constexpr int Kl = C::K(l), Nl = C::N(l), KP = C::kp(l), NP = C::np(l);
    const float* W = params + p_off(l);
    const float* b = W + Nl * Kl;
    float* hi = dst + C::img_off(l);
    float* lo = hi + C::img_floats(l);
    for (int idx = gtid; idx < NP * KP; idx += gthreads) {
      // image index -> (k chunk, n, k % 4)
      const int kc = idx / (NP * 4), n = (idx / 4) % NP, k = kc * 4 + (idx & 3);
      float v = 0.f;
      if (n < Nl) v = k < Kl ? W[n * Kl + k] : ((k == Kl && C::BIAS) ? b[n] : 0.f);
      const float h = tf32_hi(v);
      hi[idx] = h; lo[idx] = v - h;
    }
  });
}

}  // namespace tc
}  // namespace nm

// ================================================================================================
// adjoint kernel support
namespace nm {
namespace tc {

// K-major SWIZZLE_128B descriptor (tiles [feature row][128 trajectories], 128-byte rows, 4 blocks of 32 trajectories)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;                                   // leading byte offset: unused for swizzled K-major
  d |= uint64_t(1024 >> 4) << 32;                           // stride between 8-row groups
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;                                   // SWIZZLE_128B
  return d;
}
// byte offset of element (row f, trajectory m) in such a tile with R rows
__host__ __device__ constexpr uint32_t sw128_off(int f, int m, int R) {
  return uint32_t((m >> 5) * (R * 128) + f * 128 + ((((m & 31) >> 2) ^ (f & 7)) << 4) + (m & 3) * 4);
}

template <int ACT> constexpr bool act_is_mask() { return ACT == NM_ACT_RELU || ACT == NM_ACT_LEAKY || ACT == NM_ACT_IDENTITY; }
// activation and derivative from the pre-activation
template <int ACT>
__device__ __forceinline__ void act_ad(float z, float prm, float& a, float& d) {
  if constexpr (act_needs_dbuf<ACT>()) act_pair_z<ACT>(z, prm, a, d);
  else { a = act_fwd<ACT>(z, prm); d = act_deriv_a<ACT>(a, prm); }
}

template <class M, bool ALLOW_WG3 = true>
struct TcBwdCfg {
  using C = TcCfg<M>;
  static constexpr int NL = C::NL;
  static constexpr int NI = 2 * (NL - 2);                                  // streamed weight images per MLP evaluation
  // data-gradient GEMM of tensor-core layer l:  dA[m][k] = sum_n dZ[m][n] W[n][k]
  static constexpr int dnp(int l) { return round_up(C::K(l), 16); }        // MMA N (rows of the transposed image)
  static constexpr int dkp(int l) { return round_up(C::N(l), 8); }         // contraction extent
  static constexpr int dimg_floats(int l) { return dnp(l) * dkp(l); }
  static constexpr int slot_floats() {
    int m = 256;
    for (int l = 1; l <= NL - 2; ++l) m = cmax(m, cmax(2 * C::img_floats(l), 2 * dimg_floats(l)));
    return round_up(m, 256);
  }
  // image i of the per-evaluation sequence: forward layers 1..NL-2, then data-gradient layers NL-2..1
  static constexpr int img_layer(int i) { return i < NL - 2 ? i + 1 : (NL - 2) - (i - (NL - 2)); }
  static constexpr bool img_is_fwd(int i) { return i < NL - 2; }
  static constexpr int img_pair_floats(int i) { return img_is_fwd(i) ? 2 * C::img_floats(img_layer(i)) : 2 * dimg_floats(img_layer(i)); }
  // prepared global block: [first rows | last rows | last bias] (same offsets as the forward block) then the images
  static constexpr int SMALL = round_up(C::lastb_off() + round_up(C::N(NL - 1), 4), 32);
  static constexpr int gimg_off(int i) { int o = SMALL; for (int j = 0; j < i; ++j) o += img_pair_floats(j); return o; }
  static constexpr int BWD_W = gimg_off(NI);
  // tensor-memory columns
  static constexpr int opw() { int m = C::kp_max(); for (int l = 1; l <= NL - 2; ++l) m = cmax(m, dkp(l)); return m; }
  // accumulator width: read back as the next layer's operand (kp wide) only when there is a next tensor-core layer
  static constexpr int accw() { int m = NL > 3 ? cmax(C::np_max(), C::kp_max()) : C::np_max(); for (int l = 1; l <= NL - 2; ++l) m = cmax(m, cmax(dnp(l), dkp(l))); return m; }
  static constexpr int OPH = 0, OPL = opw(), ACC = 2 * opw();
  static constexpr int wgc(int l) { return round_up(C::kaug(l), 16); }     // weight-gradient accumulator of layer l: lanes n, columns k (+ bias)
  static constexpr int wg_off(int l) { int o = ACC + accw(); for (int i = 0; i < l; ++i) o += wgc(i); return o; }
  static constexpr bool MASK = act_is_mask<M::ACT>();
  // derivative stash for hidden layers 1 .. NL-2 (the last hidden layer's derivative is used on the spot)
  static constexpr int der_w(int j) { return round_up(C::K(j), chunk16_of(C::kp(j))); }   // j = index of the activation a_j (whole store chunks)
  static constexpr int der_off(int j) { int o = wg_off(NL); for (int i = 1; i < j; ++i) o += der_w(i); return o; }
  static constexpr int COLS_USED = MASK ? wg_off(NL) : der_off(NL - 1);
  static constexpr int COLS = pow2_ceil(COLS_USED);
  // shared memory (bytes): DT | DTL | AT[0..NL-1] | ring (2 slots) | small | barriers
  static constexpr int max_hidden() { return C::max_hidden(); }
  static constexpr int RD = round_up(max_hidden(), 8);
  static constexpr int ra(int l) { return round_up(C::kaug(l), 8); }
  static constexpr int DT_OFF = 0, DTL_OFF = RD * 512;
  static constexpr int at_off(int l) { int o = DTL_OFF + 8 * 512; for (int i = 0; i < l; ++i) o += ra(i) * 512; return o; }
  // weight-gradient GEMMs: 3-product split (needs the remainder tiles: a second copy of the tile region) when it
  // fits, else single-pass TF32 on round-to-nearest operands (noise ~ 2^-12 / sqrt(#products): large batches only)
  static constexpr int TILES_BYTES = at_off(NL);
  static constexpr int TAIL_BYTES = 2 * slot_floats() * 4 + SMALL * 4 + 2 * 128 * 8 * 4 + 512;
  static constexpr bool WG3 = ALLOW_WG3 && 2 * TILES_BYTES + TAIL_BYTES + 1024 <= 227 * 1024;
  static constexpr int LO_OFF = TILES_BYTES;                              // remainder tile = same offset + LO_OFF
  static constexpr int RING_OFF = (WG3 ? 2 : 1) * TILES_BYTES;
  static constexpr int SMALL_OFF = RING_OFF + 2 * slot_floats() * 4;
  static constexpr int XCH_OFF = SMALL_OFF + SMALL * 4;                   // [2][128][8] floats: partial input gradients of a thread pair
  static constexpr int BAR_OFF = XCH_OFF + 2 * 128 * 8 * 4;
  static constexpr int SMEM_BYTES = BAR_OFF + 512;
  static constexpr bool FITS = SMEM_BYTES + 1024 <= 227 * 1024 && COLS_USED <= 512 && C::N(NL - 1) <= 8 && max_hidden() <= 128;
};

// weight preparation for the adjoint kernel
template <class M>
__device__ void prep_tc_bwd(const float* __restrict__ params, float* __restrict__ dst, int gtid, int gthreads) {
  using C = TcCfg<M>;
  using B = TcBwdCfg<M>;                       // the prepared block does not depend on the weight-gradient mode
  auto p_off = [](int l) { int o = 0; for (int i = 0; i < l; ++i) o += C::N(i) * C::KW(i) + (C::BIAS ? C::N(i) : 0); return o; };
  {
    const float* W = params + p_off(0);
    const float* b = W + C::N(0) * C::KW(0);
    for (int idx = gtid; idx < C::N(0) * C::R0; idx += gthreads) {
      const int n = idx / C::R0, c = idx % C::R0;
      dst[C::first_off() + idx] = c < C::K(0) ? W[n * C::KW(0) + c] : (c == C::K(0) && C::BIAS ? b[n] : 0.f);
    }
  }
  {
    constexpr int l = C::NL - 1;
    const float* W = params + p_off(l);
    const float* b = W + C::N(l) * C::K(l);
    for (int idx = gtid; idx < C::N(l) * C::RL; idx += gthreads) {
      const int n = idx / C::RL, c = idx % C::RL;
      dst[C::last_off() + idx] = c < C::K(l) ? W[n * C::K(l) + c] : 0.f;
    }
    for (int idx = gtid; idx < round_up(C::N(l), 4); idx += gthreads) dst[C::lastb_off() + idx] = (idx < C::N(l) && C::BIAS) ? b[idx] : 0.f;
  }
  static_for<0, B::NI>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int l = B::img_layer(i);
    constexpr int Kl = C::K(l), Nl = C::N(l);
    const float* W = params + p_off(l);
    const float* b = W + Nl * Kl;
    float* hi = dst + B::gimg_off(i);
    if constexpr (B::img_is_fwd(i)) {
      constexpr int KP = C::kp(l), NP = C::np(l);
      float* lo = hi + C::img_floats(l);
      for (int idx = gtid; idx < NP * KP; idx += gthreads) {
        const int kc = idx / (NP * 4), n = (idx / 4) % NP, k = kc * 4 + (idx & 3);
        float v = 0.f;
        if (n < Nl) v = k < Kl ? W[n * Kl + k] : ((k == Kl && C::BIAS) ? b[n] : 0.f);
        const float h = tf32_hi(v);
        hi[idx] = h; lo[idx] = v - h;
      }
    } else {
      // transposed image: rows k (MMA N), contraction n:  T[n/4][k][n%4] = W[n][k]
      constexpr int NPd = B::dnp(l), KPd = B::dkp(l);
      float* lo = hi + B::dimg_floats(l);
      for (int idx = gtid; idx < NPd * KPd; idx += gthreads) {
        const int nc = idx / (NPd * 4), k = (idx / 4) % NPd, n = nc * 4 + (idx & 3);
        const float v = (k < Kl && n < Nl) ? W[n * Kl + k] : 0.f;
        const float h = tf32_hi(v);
        hi[idx] = h; lo[idx] = v - h;
      }
    }
  });
}

}  // namespace tc
}  // namespace nm
