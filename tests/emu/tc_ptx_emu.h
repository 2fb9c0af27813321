// tc_ptx_emu.h -- TEST INFRASTRUCTURE ONLY: host models of the tcgen05 / tensor-memory wrappers of nm_tc.cuh (included
// by it under NM_HOST_EMU).  One CTA at a time: tensor memory is a 128-lane x 512-column float array, an MMA executes
// synchronously in the issuing thread (so tcgen05.commit arrives immediately), operands are truncated to tf32 like the
// hardware does, accumulation is plain fp32 (the hardware truncates the accumulator; not modelled -- this checks
// LOGIC: layouts, descriptors, column maps, the barrier protocol between owners and the control warp).
#pragma once
#include <cstdint>
#include <cstring>

namespace nm {
namespace tc {

inline float g_tmem[128][512];

inline float tf32_trunc(float v) { uint32_t u; std::memcpy(&u, &v, 4); u &= 0xFFFFE000u; std::memcpy(&v, &u, 4); return v; }

struct EmuDesc {
  uint32_t start, lbo, sbo, layout;
  explicit EmuDesc(uint64_t d) : start(uint32_t(d & 0x3FFF) << 4), lbo(uint32_t((d >> 16) & 0x3FFF) << 4), sbo(uint32_t((d >> 32) & 0x3FFF) << 4), layout(uint32_t(d >> 61) & 7u) {}
  // element (row r, k) of a K-major operand tile, k in [0, 8) (one MMA K-step of 32-bit elements)
  float at(int r, int k) const {
    uint32_t off;
    if (layout == 0) off = start + uint32_t(k / 4) * lbo + uint32_t(r / 8) * sbo + uint32_t(r % 8) * 16u + uint32_t(k % 4) * 4u;   // interleaved, no swizzle
    else {                                                                // SWIZZLE_128B: 128-byte rows, 16-byte chunks XOR-ed with the row index
      off = start + uint32_t(r / 8) * sbo + uint32_t(r % 8) * 128u + uint32_t(k) * 4u;
      off ^= ((off >> 7) & 7u) << 4;
    }
    float v; std::memcpy(&v, smem_at(off), 4); return v;
  }
};
inline int idesc_n(uint32_t idesc) { return int((idesc >> 17) & 0x3Fu) << 3; }
inline int idesc_m(uint32_t idesc) { return int((idesc >> 24) & 0x1Fu) << 4; }

// D[tmem] (+)= A[smem] * B[smem]
inline void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const EmuDesc A(adesc), B(bdesc);
  const int M = idesc_m(idesc), N = idesc_n(idesc), dc = int(d_tmem & 0xFFFFu);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = accumulate ? g_tmem[m][dc + n] : 0.f;
      for (int k = 0; k < 8; ++k) s += tf32_trunc(A.at(m, k)) * tf32_trunc(B.at(n, k));
      g_tmem[m][dc + n] = s;
    }
}
// D[tmem] (+)= A[tmem] * B[smem]
template <bool ACC>
inline void mma_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  const EmuDesc B(bdesc);
  const int M = idesc_m(idesc), N = idesc_n(idesc), dc = int(d_tmem & 0xFFFFu), ac = int(a_tmem & 0xFFFFu);
  float out[256];
  for (int m = 0; m < M; ++m) {                                          // (D and A may overlap column ranges of one lane: read A first)
    float a[8];
    for (int k = 0; k < 8; ++k) a[k] = tf32_trunc(g_tmem[m][ac + k]);
    for (int n = 0; n < N; ++n) {
      float s = ACC ? g_tmem[m][dc + n] : 0.f;
      for (int k = 0; k < 8; ++k) s += a[k] * tf32_trunc(B.at(n, k));
      out[n] = s;
    }
    for (int n = 0; n < N; ++n) g_tmem[m][dc + n] = out[n];
  }
}
inline void st_shared(uint32_t addr, float v) { std::memcpy(smem_at(addr), &v, 4); }
inline bool elect_one() { return (threadIdx.x & 31u) == 0u; }
inline void commit(uint64_t* bar) { mbar_arrive(bar); }                  // the MMAs issued before it have already run
inline void fence_before() {}
inline void fence_after() {}
inline void wait_ld() {}
inline void wait_st() {}
inline void fence_async_smem() {}
template <int COLS> inline void tmem_alloc(uint32_t* slot_smem) { if ((threadIdx.x & 31u) == 0u) *slot_smem = 0u; }   // warp-collective: one write
template <int COLS> inline void tmem_dealloc(uint32_t) {}
// 32x32b shapes: thread i of the warp accesses lane (lane base of the address) + i
template <int CH> inline void tmem_ld(uint32_t taddr, float (&v)[CH]) {
  const int lane = int(taddr >> 16) + int(threadIdx.x & 31u), col = int(taddr & 0xFFFFu);
  for (int i = 0; i < CH; ++i) v[i] = g_tmem[lane][col + i];
}
template <int CH> inline void tmem_st(uint32_t taddr, const float (&v)[CH]) {
  const int lane = int(taddr >> 16) + int(threadIdx.x & 31u), col = int(taddr & 0xFFFFu);
  for (int i = 0; i < CH; ++i) g_tmem[lane][col + i] = v[i];
}
// cvt.rna.tf32.f32: round to nearest, ties away from zero, to a 10-bit mantissa
inline float tf32_hi(float v) { uint32_t u; std::memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xFFFFE000u; std::memcpy(&v, &u, 4); return v; }

}  // namespace tc
}  // namespace nm
