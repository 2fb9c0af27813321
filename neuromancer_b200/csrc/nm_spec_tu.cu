// nm_spec_tu.cu -- compiled once per shape class:  nvcc -DNM_SPEC_HEADER='"specs/<name>.h"' ...
// Instantiates the fused kernels for the generated spec and registers their launchers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include NM_SPEC_HEADER   // defines struct NM_SPEC_NAME (the "Shape") and the NM_TUNE_* knobs: before the kernel headers
#include "nm_kernels.cuh"
#include "nm_tc_kernels.cuh"
#include "nm_wide_kernels.cuh"
#include "nm_registry.h"
#include <cmath>

namespace {

using namespace nm;

// optional per-shape-class tuning emitted by build.py into the generated header
#ifndef NM_TUNE_MINB_FWD
#define NM_TUNE_MINB_FWD 1
#endif
#ifndef NM_TUNE_MINB_BWD
#define NM_TUNE_MINB_BWD 1
#endif
// exogenous first-layer hoist (nm_kernels.cuh, HoistedMlp): compiled in with NM_TUNE_HOIST=1.  Written without GPU access
// at the end of round 1 -- it compiles, the kernels of every class built without the flag are unchanged (identical SASS),
// and it has NOT run on a GPU yet: off by default until the parity suite has seen it.
#ifndef NM_TUNE_HOIST
#define NM_TUNE_HOIST 0
#endif
template <class Shape, int TILE_, int THREADS_, bool W_SMEM_>
struct Tiled : Shape {
  static constexpr int TILE = TILE_, THREADS = THREADS_;
  static constexpr bool W_SMEM = W_SMEM_;
  static constexpr bool HOIST = NM_TUNE_HOIST != 0;
  static constexpr int MINB_FWD = NM_TUNE_MINB_FWD, MINB_BWD = NM_TUNE_MINB_BWD;   // __launch_bounds__ min CTAs / SM
};

// widest layer decides the CTA size; the forward and the adjoint kernel are tiled independently (the
// trajectory layout in HBM does not depend on the tile): each takes the largest batch tile whose buffers fit
// in 227 KB, preferring weights staged in shared memory over weights streamed from L2.
constexpr int shape_max_width() {
  int m = 1;
  for (int l = 0; l <= NM_SPEC_NAME::Pol::NL; ++l) m = cmax(m, NM_SPEC_NAME::Pol::sz(l));
  for (int l = 0; l <= NM_SPEC_NAME::RhsM::NL; ++l) m = cmax(m, NM_SPEC_NAME::RhsM::sz(l));
  return m;
}
constexpr int shape_max_out() {                     // widest layer OUTPUT (the fan-in of the first layer does not count)
  int m = 1;
  for (int l = 1; l <= NM_SPEC_NAME::Pol::NL; ++l) m = cmax(m, NM_SPEC_NAME::Pol::sz(l));
  for (int l = 1; l <= NM_SPEC_NAME::RhsM::NL; ++l) m = cmax(m, NM_SPEC_NAME::RhsM::sz(l));
  return m;
}
#ifdef NM_TUNE_THREADS
constexpr int kThreadsF = NM_TUNE_THREADS, kThreadsB = NM_TUNE_THREADS;
#else
// adjoint kernel: 256 threads for wide layers, and whenever the keep-all activation buffers leave room for only
// ONE resident CTA per SM (measured on c3: 8 warps/SM instead of 4 -> adjoint 57 ms instead of 69 ms)
constexpr bool kOneCtaPerSm = KCfg<Tiled<NM_SPEC_NAME, 128, 128, true>>::BWD_SMEM_BYTES > 113 * 1024 + 512;
#ifdef NM_TUNE_THREADS_BWD
constexpr int kThreadsB = NM_TUNE_THREADS_BWD;
#else
// 128-wide layers with weights streamed from L2 (c5): 512 threads hide the L2 latency of the weight rows (measured:
// adjoint 95 -> 76 ms)
constexpr int kThreadsB = shape_max_out() >= 128 ? 512 : ((shape_max_width() >= 96 || kOneCtaPerSm) ? 256 : 128);
#endif
#ifdef NM_TUNE_THREADS_FWD
constexpr int kThreadsF = NM_TUNE_THREADS_FWD;
#else
constexpr int kThreadsF = shape_max_width() >= 96 ? 256 : 128;
#endif
#endif
template <int TILE, int THREADS, bool WS> constexpr bool fits_f() { return KCfg<Tiled<NM_SPEC_NAME, TILE, THREADS, WS>>::FITS_FWD; }
template <int TILE, int THREADS, bool WS> constexpr bool fits_b() { return KCfg<Tiled<NM_SPEC_NAME, TILE, THREADS, WS>>::FITS_BWD; }
#ifdef NM_TUNE_TILE
constexpr bool kWSmemF = fits_f<NM_TUNE_TILE, kThreadsF, true>(), kWSmemB = fits_b<NM_TUNE_TILE, kThreadsB, true>();
constexpr int kTileF = NM_TUNE_TILE, kTileB = NM_TUNE_TILE;
#else
#ifdef NM_TUNE_TILE_FWD
constexpr bool kWSmemF = fits_f<NM_TUNE_TILE_FWD, kThreadsF, true>();
constexpr int kTileF = NM_TUNE_TILE_FWD;
#else
constexpr bool kWSmemF = fits_f<128, kThreadsF, true>() || fits_f<64, kThreadsF, true>();
// weights streamed from L2 (they do not fit): measured on c5, a 64-trajectory forward tile with more resident
// CTAs hides the L2 latency better than one 128-trajectory tile (13.4 vs 17.3 ms)
constexpr int kTileF = kWSmemF ? (fits_f<128, kThreadsF, true>() ? 128 : 64)
                               : (fits_f<64, kThreadsF, false>() ? 64 : 32);
#endif
constexpr bool kWSmemB = fits_b<128, kThreadsB, true>() || fits_b<64, kThreadsB, true>();
constexpr int kTileB = kWSmemB ? (fits_b<128, kThreadsB, true>() ? 128 : 64)
                               : (fits_b<128, kThreadsB, false>() ? 128 : (fits_b<64, kThreadsB, false>() ? 64 : 32));
#endif
using SF = Tiled<NM_SPEC_NAME, kTileF, kThreadsF, kWSmemF>;     // forward (rollout) kernel
using SB = Tiled<NM_SPEC_NAME, kTileB, kThreadsB, kWSmemB>;     // adjoint kernel
using S = SB;
using KCF = KCfg<SF>;
using KC = KCfg<SB>;
static_assert(KCF::FITS_FWD && KC::FITS_BWD, "shape class does not fit in shared memory even with a 32-trajectory tile");

// tensor-core (tcgen05) forward kernel for neural-ODE shape classes whose hidden layers are wide enough for the
// tensor pipe to pay (one M=128 MMA costs >= 44 cycles whatever its N: measured, profiles/r01_tcgen05_findings.md)
using TK = TcKCfg<NM_SPEC_NAME>;
#ifdef NM_TUNE_TC
constexpr bool kTcFwd = TK::ELIGIBLE && TK::FITS_FWD && (NM_TUNE_TC != 0) && (KC::HOIST == TK::MD::HOIST);
#else
// right-hand-side mode: from ~48 hidden units up (below, the 44-cycle MMA floor loses to FFMA over the 4+ evaluations
// of a step); policy mode: from 16 up -- one evaluation per step, and what the tensor pipe removes is not pipe time
// but issue slots (FFMA + LDS are ~60 % of the FFMA kernels' instructions on the narrow policies)
// (with the first-layer hoist compiled in, both kernel families have to agree on it: the FFMA adjoint is the fallback)
constexpr bool kTcFwd = TK::ELIGIBLE && TK::FITS_FWD && TK::C::max_hidden() >= (TK::MD::POLICY ? 16 : 48) && (KC::HOIST == TK::MD::HOIST);
#endif

// launch helpers as class templates so that the tensor-core kernels are only instantiated for eligible classes
template <class Sp, bool ON> struct TcFwd {
  static cudaError_t setup(int*) { return cudaSuccess; }
  static void launch(const NmPlan&, const FwdArgs&, const float*, float*, int, cudaStream_t) {}
};
template <class Sp> struct TcFwd<Sp, true> {
  static cudaError_t setup(int* occ) {
    cudaError_t e = cudaFuncSetAttribute(nm_tc_fwd_kernel<Sp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcKCfg<Sp>::FWD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, nm_tc_fwd_kernel<Sp>, 128, TcKCfg<Sp>::FWD_SMEM_BYTES);
    if (e == cudaSuccess) {
      // the occupancy query is conservative for kernels that allocate tensor memory (it answered 1 where two
      // CTAs fit): recompute from shared memory and registers, then cap by the 512 TMEM columns
      cudaFuncAttributes fa{};
      e = cudaFuncGetAttributes(&fa, nm_tc_fwd_kernel<Sp>);
      if (e == cudaSuccess) {
        const int by_smem = int((227 * 1024) / (TcKCfg<Sp>::FWD_SMEM_BYTES + 1024));
        const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * 128;
        const int by_regs = regs_per_cta > 0 ? 65536 / regs_per_cta : 1;
        int manual = by_smem < by_regs ? by_smem : by_regs;
        if (getenv("NM_DEBUG")) fprintf(stderr, "[nm] tc fwd occupancy: api %d, smem %d, regs %d (numRegs %d)\n", *occ, by_smem, by_regs, fa.numRegs);
        if (manual > *occ) *occ = manual;
      }
    }
    if (*occ > TcKCfg<Sp>::MAX_CTAS) *occ = TcKCfg<Sp>::MAX_CTAS;
    if (*occ < 1) *occ = 1;
    return e;
  }
  static void launch(const NmPlan& p, const FwdArgs& a, const float* params, float* wprep, int grid, cudaStream_t st) {
    nm_tc_prep_fwd_kernel<Sp><<<8, 256, 0, st>>>(p, params, wprep);
    nm_tc_fwd_kernel<Sp><<<grid, 128, TcKCfg<Sp>::FWD_SMEM_BYTES, st>>>(p, a);
  }
};

template <class Sp, bool ON> struct TcBwd {
  static constexpr size_t W_BYTES = 0;
  static constexpr bool WG3 = false;
  static constexpr int EVALS = 1, CTAS = 1, THREADS = 0;
  static cudaError_t setup() { return cudaSuccess; }
  static void launch(const NmPlan&, const BwdArgs&, const float*, float*, int, cudaStream_t) {}
};
template <class Sp> struct TcBwd<Sp, true> {
  static constexpr size_t W_BYTES = TcBKCfg<Sp>::W_BYTES;
  static constexpr bool WG3 = TcBKCfg<Sp>::WG3;
  static constexpr int EVALS = TcMode<Sp>::EVALS, CTAS = TcBKCfg<Sp>::CTAS_PER_SM, THREADS = TcBKCfg<Sp>::THREADS;
  static constexpr bool HAS_FO = Sp::NTERMS > 0;          // a second instantiation without any runtime-descriptor term code
  static cudaError_t setup() {
    cudaError_t e = cudaFuncSetAttribute(nm_tc_bwd_kernel<Sp, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcBKCfg<Sp>::SMEM_BYTES);
    if constexpr (HAS_FO) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_tc_bwd_kernel<Sp, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcBKCfg<Sp>::SMEM_BYTES);
    }
    return e;
  }
  static void launch(const NmPlan& p, const BwdArgs& a, const float* params, float* wprep, int grid, cudaStream_t st) {
    nm_tc_prep_bwd_kernel<Sp><<<16, 256, 0, st>>>(p, params, wprep);
    if constexpr (HAS_FO) {
      if (a.fused_terms) { nm_tc_bwd_kernel<Sp, true><<<grid, TcBKCfg<Sp>::THREADS, TcBKCfg<Sp>::SMEM_BYTES, st>>>(p, a); return; }
    }
    nm_tc_bwd_kernel<Sp, false><<<grid, TcBKCfg<Sp>::THREADS, TcBKCfg<Sp>::SMEM_BYTES, st>>>(p, a);
  }
};
template <class Sp, bool EL> struct TcBwdEligible { static constexpr bool value = false; };
template <class Sp> struct TcBwdEligible<Sp, true> { static constexpr bool value = TcBKCfg<Sp>::ELIGIBLE; };
// the tensor-core adjoint needs the stage checkpoints written by the forward kernel; without them the FFMA adjoint runs
constexpr bool kTcBwd = kTcFwd && TcBwdEligible<NM_SPEC_NAME, kTcFwd>::value;

// ------------------------------------------------------------------------------------------------
// wide-policy tensor-core family (nm_wide_kernels.cuh): fp16 3-product MMAs, weights resident in shared memory,
// weight gradients as a separate streaming GEMM over a staging buffer, horizon swept in chunks
template <class Sp, bool ON> struct Wide {
  static constexpr size_t W_BYTES = 0;
  static cudaError_t setup() { return cudaSuccess; }
  static int chunk_steps(int64_t, int) { return 1; }
  static size_t bwd_extra(int64_t, int, int) { return 0; }
  static int launch_fwd(const NmPlan&, const FwdArgs&, const float*, void*, int, cudaStream_t) { return 0; }
  static int launch_bwd(const NmPlan&, const BwdArgs&, const float*, void*, char*, int, int, cudaStream_t) { return 0; }
  static void info(char*, size_t, int, int, size_t, int, int) {}
};
template <class Sp> struct Wide<Sp, true> {
  using C = wd::WideCfg<Sp>;
  static constexpr size_t W_BYTES = C::W_BYTES;
  static cudaError_t setup() {
    cudaError_t e = cudaFuncSetAttribute(nm_wide_fwd_kernel<Sp>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_wide_bwd_kernel<Sp>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_wide_wgrad_kernel<Sp>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::WG_SMEM_BYTES);
    if (e == cudaSuccess) {
      // L2 set-aside for the evict_last derivative stash (58 MB on 148 SMs).  Measured on B200 (C5, B = 4,194,304,
      // profiles/r02_l2_carveout_sweep.txt): the adjoint window takes 1.60 s with a 48-56 MB carve-out, 1.72 s with 64 MB
      // and 1.90 s from 72 MB up (the maximum, 126 MB, leaves the streaming staging stores and the row loads no room).
      // Re-swept after the grouped term gradients (profiles/r02c_l2_carveout_sweep.txt): 44 MB 1.574 s, 52 MB 1.530-1.534 s,
      // 56 MB 1.514 s, 60 MB 1.510-1.516 s, 64 MB 1.604 s -> 56 MB, clear of the cliff at 64.
      int dev = 0, maxp = 0;
      size_t want = size_t(56) << 20;
      if (const char* sv = getenv("NM_WIDE_L2_PERSIST_MB")) want = size_t(atol(sv)) << 20;        // experiments: 0 = no carve-out
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess && maxp > 0)
        (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want < size_t(maxp) ? want : size_t(maxp));
      (void)cudaGetLastError();
    }
    return e;
  }
  // steps per adjoint launch: the staging buffer (tiles * CS items of ITEM_BYTES) is capped at NM_WIDE_STAGING_GB (default 24)
  static int chunk_steps(int64_t B, int T) {
    double cap_gb = 24.0;
    if (const char* s = getenv("NM_WIDE_STAGING_GB")) { const double v = atof(s); if (v > 0.0) cap_gb = v; }
    const double per_step = double((B + 127) / 128) * double(C::ITEM_BYTES);
    int64_t cs = int64_t(cap_gb * 1073741824.0 / per_step);
    if (cs < 1) cs = 1;
    if (cs > T) cs = T;
    return int(cs);
  }
  static void info(char* buf, size_t cap, int tile, int threads, size_t smem, int spill, int nterms) {
    std::snprintf(buf, cap, "fwd[tcgen05 3xFP16 wide tile=2x128 threads=%d smem=%d weights-resident] bwd[tcgen05 3xFP16 wide threads=%d + wgrad GEMM 1xFP16-rn item=%d B; below 2^20 products or without checkpoints: ffma tile=%d threads=%d smem=%zu dw_spill=%d] compiled_terms=%d",
                  C::THREADS, C::SMEM_BYTES, C::THREADS, C::ITEM_BYTES, tile, threads, smem, spill, nterms);
  }
  static constexpr size_t a256(size_t v) { return (v + 255) & ~size_t(255); }
  static size_t staging_bytes(int64_t B, int T) { return a256(size_t((B + 127) / 128) * size_t(chunk_steps(B, T)) * size_t(C::ITEM_BYTES)); }
  static size_t stash_bytes(int sms) { return a256(size_t(sms) * 2 * size_t(C::STASH_SLOT_BYTES)); }
  static size_t carry_bytes(int64_t B) { return a256(size_t(B) * Sp::NX * sizeof(float)); }
  // steps per term-gradient launch: a multiple of the chunk length, G capped at NM_WIDE_G_GB (default 6).  The trajectories are
  // [B][T][nx] (the reference's layout): a launch that covers few steps reads 48-byte rows 9.6 KB apart and pays DRAM page
  // misses for every one of them (measured 6x its HBM floor at 2 steps per launch) -- long groups read long contiguous runs
  static int group_steps(int64_t B, int T) {
    const int cs = chunk_steps(B, T);
    double cap_gb = 6.0;
    if (const char* s = getenv("NM_WIDE_G_GB")) { const double v = atof(s); if (v > 0.0) cap_gb = v; }
    const double per_step = double(B) * double(Sp::NX + Sp::NU) * sizeof(float);
    int64_t k = int64_t(cap_gb * 1073741824.0 / per_step - 1.0) / cs;
    if (k < 1) k = 1;
    int64_t gs = k * cs;
    if (gs > T) gs = (T + cs - 1) / cs * cs;
    return int(gs);
  }
  static size_t g_bytes(int64_t B, int T) { return a256(size_t(B) * size_t(group_steps(B, T) + 1) * size_t(Sp::NX + Sp::NU) * sizeof(float)); }
  // [scale state 256 B | co-state carry | derivative stash | chunk term gradients | staging]
  static size_t bwd_extra(int64_t B, int T, int sms) { return 256 + carry_bytes(B) + stash_bytes(sms) + g_bytes(B, T) + staging_bytes(B, T); }
  static int grid_for(int64_t B, int sms) { const int64_t pairs = ((B + 127) / 128 + 1) / 2; return int(pairs < sms ? pairs : sms); }
  static int score_grid(const NmPlan& p, int64_t B, int sms) {
    int tmax = 1;
    for (int k = 0; k < p.n_terms; ++k) tmax = p.terms[k].t_len > tmax ? p.terms[k].t_len : tmax;
    const int64_t g = (B * tmax + 255) / 256;
    return int(g < int64_t(sms) * 8 ? (g < 1 ? 1 : g) : int64_t(sms) * 8);
  }
  // returns the number of blocks that wrote term partials (0: the caller scores the terms)
  static int launch_fwd(const NmPlan& p, const FwdArgs& a, const float* params, void* wprep, int sms, cudaStream_t st) {
    wd::nm_wide_prep_kernel<Sp><<<32, 256, 0, st>>>(p, params, static_cast<uint8_t*>(wprep));
    nm_wide_fwd_kernel<Sp><<<grid_for(a.B, sms), C::FWD_THREADS, C::SMEM_BYTES, st>>>(p, a);
    if (Sp::NTERMS > 0 && a.fused_terms) {
      int tmax = 1;
      for (int k = 0; k < p.n_terms; ++k) tmax = p.terms[k].t_len > tmax ? p.terms[k].t_len : tmax;
      const int g = score_grid(p, a.B, sms);
      nm_wide_score_kernel<Sp><<<g, 256, 0, st>>>(p, a, tmax);
      return g;
    }
    return 0;
  }
  // loss-term gradients of the chunk [t_lo, t_hi) into G; returns the number of kernels launched
  static int launch_termgrad(const NmPlan& p, const BwdArgs& a, float* G, int t_lo, int t_hi, int sms, cudaStream_t st) {
    const int T = p.nsteps;
    int launches = 1;
    const int64_t n = a.B * int64_t(t_hi - t_lo + 1);
    int64_t g = (n + 255) / 256;
    if (g > int64_t(sms) * 16) g = int64_t(sms) * 16;
    if (Sp::NTERMS > 0 && a.fused_terms) nm_wide_termgrad_kernel<Sp, true><<<int(g), 256, 0, st>>>(p, a, G, t_lo, t_hi);
    else nm_wide_termgrad_kernel<Sp, false><<<int(g), 256, 0, st>>>(p, a, G, t_lo, t_hi);
    // time-broadcast atoms anchored inside this chunk: one warp per trajectory sums the range
    WideTbAnchors an{};
    for (int k = 0; k < p.n_terms && an.n < 8; ++k) {
      const NmTerm& tm = p.terms[k];
      const NmAtom* at[4] = {&tm.lhs.a, &tm.lhs.b, &tm.rhs.a, &tm.rhs.b};
      const bool live[4] = {true, tm.lhs.op != NM_OP_NONE, true, tm.rhs.op != NM_OP_NONE};
      for (int i = 0; i < 4 && an.n < 8; ++i) {
        if (!live[i] || !(at[i]->var == NM_VAR_X || at[i]->var == NM_VAR_U) || !(at[i]->t_len == 1 && tm.t_len > 1)) continue;
        const int t0 = at[i]->t_start;
        const bool in_chunk = at[i]->var == NM_VAR_X ? (t0 >= t_lo && (t0 < t_hi || (t0 == t_hi && t_hi == T))) : (t0 >= t_lo && t0 < t_hi);
        bool seen = false;
        for (int q = 0; q < an.n; ++q) seen = seen || (an.var[q] == at[i]->var && an.t[q] == t0);
        if (in_chunk && !seen) { an.var[an.n] = at[i]->var; an.t[an.n] = t0; ++an.n; }
      }
    }
    if (an.n > 0) {
      int64_t gw = (a.B * 32 + 255) / 256;
      if (gw > int64_t(sms) * 16) gw = int64_t(sms) * 16;
      if (Sp::NTERMS > 0 && a.fused_terms) nm_wide_termgrad_tb_kernel<Sp, true><<<int(gw), 256, 0, st>>>(p, a, an, G, t_lo, t_hi);
      else nm_wide_termgrad_tb_kernel<Sp, false><<<int(gw), 256, 0, st>>>(p, a, an, G, t_lo, t_hi);
      ++launches;
    }
    return launches;
  }
  // returns the number of kernels launched; `extra` is the wide region of the backward workspace, `nblocks` gradient blocks are zeroed
  static int launch_bwd(const NmPlan& p, const BwdArgs& a, const float* params, void* wprep, char* extra, int sms, int nblocks, cudaStream_t st) {
    const int T = p.nsteps;
    const int CS = chunk_steps(a.B, T);
    const int64_t tiles = (a.B + 127) / 128;
    wd::ScaleState* scale = reinterpret_cast<wd::ScaleState*>(extra);
    float* carry = reinterpret_cast<float*>(extra + 256);
    float* stash = reinterpret_cast<float*>(extra + 256 + carry_bytes(a.B));
    float* G = reinterpret_cast<float*>(extra + 256 + carry_bytes(a.B) + stash_bytes(sms));
    uint8_t* staging = reinterpret_cast<uint8_t*>(extra + 256 + carry_bytes(a.B) + stash_bytes(sms) + g_bytes(a.B, T));
    // Initial cotangent scale.  A term of weight w puts w / (B_total * t_len * d_len) on each of its t_len time entries --
    // all on ONE entry when its atom is time-broadcast (terminal constraints), otherwise accumulating in the co-state
    // along the sweep -- so sum_k |w_k| / (B_total * d_len_k) estimates the largest co-state of the first chunk; it goes to
    // ~2^9: 2^7 of headroom below fp16's 65504 for quadratic penalties and growing dynamics, 23 binades of normal range
    // below it (the per-chunk update then steers by the largest cotangent actually seen).  The former rule scaled the
    // largest single coefficient to 2^4 and overflowed on a whole-horizon chunk with terminal terms (B = 300, 200 steps).
    double mass = 0.0;
    for (int k = 0; k < p.n_terms; ++k) mass += std::fabs(double(p.terms[k].weight)) / (double(a.b_total) * double(p.terms[k].d_len));
    int e0 = mass > 0.0 ? 9 - int(std::floor(std::log2(mass))) : int(std::floor(std::log2(double(a.b_total)))) + 10;
    if (const char* s = getenv("NM_WIDE_GS0_LOG2")) e0 = atoi(s);
    if (e0 > 100) e0 = 100;
    if (e0 < -100) e0 = -100;
    cudaMemsetAsync(a.grad_partials, 0, size_t(nblocks) * size_t(p.n_params) * sizeof(float), st);
    nm_wide_scale_init_kernel<<<1, 32, 0, st>>>(scale, float(std::ldexp(1.0, e0)));
    wd::nm_wide_prep_kernel<Sp><<<32, 256, 0, st>>>(p, params, static_cast<uint8_t*>(wprep));
    int launches = 3;
    const int grid = grid_for(a.B, sms);
    // (Measured and dropped: the term-gradient kernels of chunk i+1 on a side stream next to the weight-gradient GEMM of
    // chunk i -- both want DRAM bandwidth, the step got 3 % slower.)
    const int GS = group_steps(a.B, T);
    int g_lo = T, g_hi = T;                                  // term gradients of [g_lo, g_hi] are in G
    for (int t_hi = T; t_hi > 0; t_hi -= CS) {
      const int t_lo = t_hi - CS > 0 ? t_hi - CS : 0;
      if (t_lo < g_lo) {                                       // next group of chunks (GS is a multiple of CS: chunks never straddle groups)
        g_hi = t_hi; g_lo = g_hi - GS > 0 ? g_hi - GS : 0;
        launches += launch_termgrad(p, a, G, g_lo, g_hi, sms, st);
      }
      WideBwdArgs w{};
      w.b = a; w.staging = staging; w.stash = stash; w.lam_carry = carry; w.scale = scale; w.G = G; w.t_lo = t_lo; w.t_hi = t_hi;
      w.g_lo = g_lo; w.g_nt = g_hi - g_lo + 1;
      nm_wide_bwd_kernel<Sp><<<grid, C::BWD_THREADS, C::SMEM_BYTES, st>>>(p, w);
      const int64_t n_items = tiles * int64_t(t_hi - t_lo);
      const int wgrid = int(n_items < sms ? n_items : sms);
      nm_wide_wgrad_kernel<Sp><<<wgrid, C::WG_THREADS, C::WG_SMEM_BYTES, st>>>(p, staging, n_items, scale, a.grad_partials);
      nm_wide_scale_update_kernel<<<1, 32, 0, st>>>(scale);
      launches += 3;
    }
    return launches;
  }
};
template <class Sp, bool EL> struct WideEligible { static constexpr bool value = false; };
template <class Sp> struct WideEligible<Sp, true> { static constexpr bool value = wd::WideCfg<Sp>::ELIGIBLE && wd::WideCfg<Sp>::FITS && wd::WideCfg<Sp>::WG_OK; };
#ifndef NM_TUNE_WIDE
#define NM_TUNE_WIDE 1
#endif
constexpr bool kWide = !kTcFwd && (NM_TUNE_WIDE != 0) && !KC::HOIST && NM_SPEC_NAME::RHS != NM_RHS_MLP && NM_SPEC_NAME::Pol::NL >= 3 &&
                       WideEligible<NM_SPEC_NAME, (NM_SPEC_NAME::RHS != NM_RHS_MLP && NM_SPEC_NAME::Pol::NL >= 3)>::value;
using WD = Wide<NM_SPEC_NAME, kWide>;

constexpr bool kHoist = KC::HOIST && KCF::HOIST;
constexpr size_t hoist_bytes(const NmPlan* p, int64_t B) { return kHoist ? ((size_t(B) * size_t(p->nsteps) * size_t(KC::N0) * sizeof(float) + 255) & ~size_t(255)) : 0; }
// forward hoist GEMM on the tensor cores (nm_exo_gemm_tc_fwd_kernel) when the class allows it; its weight images
// live behind the other prepared blocks
#ifndef NM_TUNE_HOIST_TCGEMM
#define NM_TUNE_HOIST_TCGEMM 1
#endif
template <bool ON> constexpr bool exo_tc_ok() { if constexpr (ON) return ExoTcCfg<SB>::OK; else return false; }
constexpr bool kExoTc = exo_tc_ok<kHoist>() && (NM_TUNE_HOIST_TCGEMM != 0);
template <bool ON> constexpr size_t exo_tc_wbytes() { if constexpr (ON) return (size_t(ExoTcCfg<SB>::W_FLOATS) * 4 + 255) & ~size_t(255); else return 0; }
template <bool ON> constexpr size_t exo_tc_smem() { if constexpr (ON) return ExoTcCfg<SB>::SMEM_BYTES; else return 0; }
template <bool ON> constexpr bool exo_tc_bwd_ok() { if constexpr (ON) return ExoTcBwdCfg<SB>::OK; else return false; }
constexpr bool kExoTcBwd = exo_tc_bwd_ok<kHoist>() && (NM_TUNE_HOIST_TCGEMM != 0);
template <bool ON> constexpr size_t exo_tc_bwd_smem() { if constexpr (ON) return ExoTcBwdCfg<SB>::SMEM_BYTES; else return 0; }
constexpr size_t kExoSmem = exo_gemm_smem_bytes<SB>();
// the hoisted configuration addresses the same parameter block as the plain one
template <class K_, class Tl> constexpr bool hoist_layout_ok() {       // (dependent on K_: only instantiated for hoisted classes)
  if constexpr (K_::HOIST) {
    using Full = MlpCfg<NM_SPEC_NAME::Pol, Tl::TILE, Tl::THREADS, K_::pol_gw0()>;
    using Pc = typename K_::PolCfg;
    return Pc::N_PARAMS == Full::N_PARAMS && K_::KS + K_::KE == NM_SPEC_NAME::Pol::sz(0) && K_::KS >= 1 && K_::KE >= 4 &&
           Pc::p_off(1) == Full::p_off(1) && Pc::K(0) == K_::KS && Pc::KW(0) == NM_SPEC_NAME::Pol::sz(0);
  } else return true;
}
static_assert(hoist_layout_ok<KC, SB>() && hoist_layout_ok<KCF, SF>(), "hoist: parameter layout of the restricted first layer");
static_assert(!kHoist || kExoSmem <= 200 * 1024, "hoist: GEMM tiles do not fit in shared memory");

struct DeviceInfo { int sms = 0; int occ_fwd = 0; int occ_bwd = 0; bool ready = false; };
DeviceInfo g_dev[64];
std::mutex g_mu;

int device_info(DeviceInfo** out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { nm_set_error("cudaGetDevice failed"); return NM_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.ready) {
    cudaError_t e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_fwd_kernel<SF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KCF::FWD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KC::BWD_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_fwd, nm_fwd_kernel<SF>, SF::THREADS, KCF::FWD_SMEM_BYTES);
    if (kTcFwd && e == cudaSuccess) e = TcFwd<NM_SPEC_NAME, kTcFwd>::setup(&d.occ_fwd);
    if (kTcBwd && e == cudaSuccess) e = TcBwd<NM_SPEC_NAME, kTcBwd>::setup();
    if (kWide && e == cudaSuccess) e = WD::setup();
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_bwd, nm_bwd_kernel<S>, S::THREADS, KC::BWD_SMEM_BYTES);
    if constexpr (kHoist) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_exo_gemm_fwd_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExoSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_exo_gemm_bwd_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExoSmem);
      if constexpr (kExoTc) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_exo_gemm_tc_fwd_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exo_tc_smem<kExoTc>());
      }
      if constexpr (kExoTcBwd) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nm_exo_gemm_tc_bwd_kernel<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exo_tc_bwd_smem<kExoTcBwd>());
      }
    }
    if (e != cudaSuccess) { nm_set_error(cudaGetErrorString(e)); return NM_ERR_CUDA; }
    if (d.occ_fwd < 1 || d.occ_bwd < 1) { nm_set_error("kernel does not fit on an SM (occupancy 0)"); return NM_ERR_CUDA; }
    d.ready = true;
  }
  *out = &d;
  return NM_OK;
}

constexpr size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
constexpr size_t kWprepFfma = size_t(KC::W_FLOATS_ALL + KC::EXO_W_FLOATS > 0 ? KC::W_FLOATS_ALL + KC::EXO_W_FLOATS : 4) * 4;
constexpr size_t kWprepTc0 = TcBwd<NM_SPEC_NAME, kTcBwd>::W_BYTES > (kTcFwd ? TK::FWD_W_BYTES : 0) ? TcBwd<NM_SPEC_NAME, kTcBwd>::W_BYTES : (kTcFwd ? TK::FWD_W_BYTES : 0);
constexpr size_t kWprepTc = kWprepTc0 > WD::W_BYTES ? kWprepTc0 : WD::W_BYTES;
// tensor-core class with the hoist: the FFMA-layout block (exogenous rows for the hoist GEMM) lives BEHIND the
// tensor-core images while the tensor-core kernels run; every other class keeps one region
constexpr size_t kFfmaRegionOff = (kHoist && kTcFwd) ? align256(kWprepTc) : 0;
constexpr size_t kBaseWprepBytes = (kHoist && kTcFwd) ? align256(kWprepTc) + align256(kWprepFfma) : align256(kWprepTc > kWprepFfma ? kWprepTc : kWprepFfma);
constexpr size_t kExoTcOff = kBaseWprepBytes;
constexpr size_t kWprepBytes = kBaseWprepBytes + exo_tc_wbytes<kExoTc>();
constexpr int kFwdTile = (kTcFwd || kWide) ? 128 : SF::TILE;

int check_plan(const NmPlan* p) {
  if (p->n_params < KC::N_PARAMS) { nm_set_error("plan.n_params smaller than the shape class expects"); return NM_ERR_BAD_PLAN; }
  if (KC::HAS_POL && (p->policy.has_bias != 0) != S::Pol::BIAS) { nm_set_error("policy bias flag mismatch"); return NM_ERR_BAD_PLAN; }
  if (KC::HAS_RHS_MLP && (p->rhs_mlp.has_bias != 0) != S::RhsM::BIAS) { nm_set_error("rhs bias flag mismatch"); return NM_ERR_BAD_PLAN; }
  return NM_OK;
}
int check_fused(const NmPlan* p, int fused) {
  if (fused && p->n_terms != S::NTERMS) { nm_set_error("fused term count mismatch"); return NM_ERR_BAD_PLAN; }
  return NM_OK;
}

int workspace_fn(const NmPlan* p, int64_t B, size_t* fwd, size_t* bwd) {
  DeviceInfo* d = nullptr;
  // sizes must be computable without a device too (host-only tests): assume a B200 (148 SMs, <= 8 CTAs/SM)
  int sms = 148, occf = 8, occb = 8, ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && device_info(&d) == NM_OK) { sms = d->sms; occf = d->occ_fwd; occb = d->occ_bwd; }
  else { (void)cudaGetLastError(); }
  const int64_t tiles_f = (B + kFwdTile - 1) / kFwdTile, tiles_b = (B + SB::TILE - 1) / SB::TILE;
  const int64_t gf = tiles_f < int64_t(sms) * occf ? tiles_f : int64_t(sms) * occf;
  int64_t gb = tiles_b < int64_t(sms) * occb ? tiles_b : int64_t(sms) * occb;
  if (kTcBwd) {                                     // the tensor-core adjoint launches its own grid (128-trajectory tiles)
    const int64_t tiles_tc = (B + 127) / 128, slots = int64_t(sms) * TcBwd<NM_SPEC_NAME, kTcBwd>::CTAS;
    const int64_t gtc = tiles_tc < slots ? tiles_tc : slots;
    if (gtc > gb) gb = gtc;
  }
  // (hoist: E[B][T][N0] in the forward workspace when the caller gives no checkpoint buffer; E recomputed / dZ_1 in the backward one)
  if (kWide && gb < sms) gb = sms;                   // the wide adjoint / weight-gradient GEMM own one gradient block per SM
  int64_t gfw = gf;
  if (kWide && gfw < int64_t(sms) * 8) gfw = int64_t(sms) * 8;   // blocks of the streaming scoring kernel
  *fwd = kWprepBytes + align256(size_t(gfw > 0 ? gfw : 1) * NM_MAX_TERMS * sizeof(double)) + hoist_bytes(p, B);
  *bwd = kWprepBytes + align256(size_t(gb > 0 ? gb : 1) * size_t(p->n_params > 0 ? p->n_params : 1) * sizeof(float)) + hoist_bytes(p, B);
  if (kWide) *bwd = align256(*bwd) + WD::bwd_extra(B, p->nsteps, sms);
  return NM_OK;
}

int launch_prep(const NmPlan* p, const float* params, float* wprep, cudaStream_t st) {
  if constexpr (KC::W_FLOATS_ALL + KC::EXO_W_FLOATS > 0) {
    const int blocks = (KC::W_FLOATS_ALL + KC::EXO_W_FLOATS + 4 * 256 - 1) / (4 * 256);
    nm_prep_kernel<S><<<blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks), 256, 0, st>>>(*p, params, wprep);
  }
  return NM_OK;
}

int forward_fn(const NmPlan* p, NmFwdCall* c) {
  int rc = check_plan(p); if (rc != NM_OK) return rc;
  rc = check_fused(p, c->fused_terms); if (rc != NM_OK) return rc;
  DeviceInfo* d = nullptr; rc = device_info(&d); if (rc != NM_OK) return rc;
  size_t need_f = 0, need_b = 0; workspace_fn(p, c->B, &need_f, &need_b);
  if (c->workspace_bytes < need_f || (reinterpret_cast<uintptr_t>(c->workspace) & 255)) { nm_set_error("forward workspace too small or misaligned"); return NM_ERR_WORKSPACE; }
  const int tiles = int((c->B + kFwdTile - 1) / kFwdTile);
  int grid = d->sms * d->occ_fwd; if (grid > tiles) grid = tiles; if (grid < 1) grid = 1;
  char* ws = static_cast<char*>(c->workspace);
  float* wprep = reinterpret_cast<float*>(ws);
  c->grid = grid;
  c->term_partials = reinterpret_cast<double*>(ws + kWprepBytes);
  float* wprep_ffma = reinterpret_cast<float*>(ws + kFfmaRegionOff);
  if (kWide) { grid = d->sms; if (grid > (tiles + 1) / 2) grid = (tiles + 1) / 2; if (grid < 1) grid = 1; c->grid = grid; }
  if ((!kTcFwd && !kWide) || kHoist) launch_prep(p, c->params, wprep_ffma, c->stream);
  FwdArgs a{};
  a.term_partials = c->term_partials; a.fused_terms = c->fused_terms;
  a.params = c->params; a.wprep = wprep; a.x0 = c->x0;
  for (int e = 0; e < NM_MAX_EXO; ++e) a.exo[e] = e < p->n_exo ? c->exo[e] : nullptr;
  a.x_traj = c->x_traj; a.u_traj = c->u_traj; a.y_traj = c->y_traj; a.chk = c->checkpoints;
  a.B = c->B; a.num_tiles = tiles;
  if constexpr (kHoist) {
    // E = bias + exogenous part of the first layer for every (trajectory, step): into the caller's checkpoint buffer
    // when there is one (the adjoint then reads it back), else into the workspace tail
    float* E = c->checkpoints != nullptr ? c->checkpoints
                                         : reinterpret_cast<float*>(ws + (need_f - hoist_bytes(p, c->B)));
    ExoPtrs ep{};
    for (int e = 0; e < NM_MAX_EXO; ++e) ep.p[e] = a.exo[e];
    const int64_t row_tiles = (c->B * p->nsteps + SB::TILE - 1) / SB::TILE;
    int ggrid = d->sms * 2; if (ggrid > row_tiles) ggrid = int(row_tiles); if (ggrid < 1) ggrid = 1;
    if constexpr (kExoTc) {
      float* wexo = reinterpret_cast<float*>(ws + kExoTcOff);
      const int64_t tiles128 = (c->B * p->nsteps + 127) / 128;
      int tgrid = d->sms; if (tgrid > tiles128) tgrid = int(tiles128); if (tgrid < 1) tgrid = 1;      // one CTA per SM (220 KB: operand double buffer + raw ring)
      nm_exo_tc_prep_kernel<SB><<<8, 256, 0, c->stream>>>(*p, c->params, wexo);
      nm_exo_gemm_tc_fwd_kernel<SB><<<tgrid, 288, exo_tc_smem<kExoTc>(), c->stream>>>(*p, wexo, ep, E, c->B);
      c->extra_launches = 2;
    } else {
      nm_exo_gemm_fwd_kernel<SB><<<ggrid, SB::THREADS, kExoSmem, c->stream>>>(*p, wprep_ffma, ep, E, c->B);
      c->extra_launches = 1;
    }
    a.pre = E;
  }
  c->scored = c->fused_terms ? 1 : 0;
  if (kWide) {                                       // scoring: a streaming kernel behind the rollout (compiled term structure), else the API's
    const int sg = WD::launch_fwd(*p, a, c->params, wprep, d->sms, c->stream);
    c->scored = sg > 0 ? 1 : 0;
    if (sg > 0) { c->grid = sg; c->extra_launches += 1; }
  }
  else if (kTcFwd) TcFwd<NM_SPEC_NAME, kTcFwd>::launch(*p, a, c->params, wprep, grid, c->stream);
  else nm_fwd_kernel<SF><<<grid, SF::THREADS, KCF::FWD_SMEM_BYTES, c->stream>>>(*p, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { nm_set_error(cudaGetErrorString(e)); return NM_ERR_CUDA; }
  return NM_OK;
}

int backward_fn(const NmPlan* p, NmBwdCall* c) {
  int rc = check_plan(p); if (rc != NM_OK) return rc;
  rc = check_fused(p, c->fused_terms); if (rc != NM_OK) return rc;
  DeviceInfo* d = nullptr; rc = device_info(&d); if (rc != NM_OK) return rc;
  size_t need_f = 0, need_b = 0; workspace_fn(p, c->B, &need_f, &need_b);
  if (c->workspace_bytes < need_b || (reinterpret_cast<uintptr_t>(c->workspace) & 255)) { nm_set_error("backward workspace too small or misaligned"); return NM_ERR_WORKSPACE; }
  const int tiles = int((c->B + S::TILE - 1) / S::TILE);
  int grid = d->sms * d->occ_bwd; if (grid > tiles) grid = tiles; if (grid < 1) grid = 1;
  char* ws = static_cast<char*>(c->workspace);
  float* wprep = reinterpret_cast<float*>(ws);
  c->grad_partials = reinterpret_cast<float*>(ws + kWprepBytes);
  c->grid = grid;
  // Shape classes whose remainder tiles do not fit run the weight-gradient GEMMs as single-pass TF32 on
  // round-to-nearest operands: unbiased noise of 2^-12 per product that averages out over the B * nsteps * stages
  // products of every gradient entry (measured on c3: 2.2e-5 at 1.2e5 products, tests/test_parity_gpu.py).  Below
  // 2^20 products the exact FFMA adjoint runs instead.  NM_TC_BWD=0 / 1 forces the choice (tests, experiments).
  bool tc_bwd = kTcBwd && c->checkpoints != nullptr;
  if (tc_bwd) {
    using TB = TcBwd<NM_SPEC_NAME, kTcBwd>;
    const char* force = getenv("NM_TC_BWD");
    if (force != nullptr) tc_bwd = force[0] != '0';
    else if (!TB::WG3) tc_bwd = double(c->B) * double(p->nsteps) * double(TB::EVALS) >= 1048576.0;
  }
  // trainable white-box constants (rhs_mlp.trainable on a non-MLP right-hand side): only the exact FFMA adjoint
  // accumulates d loss / d rhs_const -- it runs whatever the batch size and whatever NM_TC_BWD says
  const bool theta_train = !KC::HAS_RHS_MLP && rhs_num_const<NM_SPEC_NAME>() > 0 && p->rhs_mlp.trainable != 0;
  if (theta_train) tc_bwd = false;
  // wide-policy classes: same rule (single-pass fp16 weight-gradient operands), NM_TC_BWD forces
  bool wide_bwd = kWide && c->checkpoints != nullptr && !theta_train;
  if (wide_bwd) {
    const char* force = getenv("NM_TC_BWD");
    if (force != nullptr) wide_bwd = force[0] != '0';
    else wide_bwd = double(c->B) * double(p->nsteps) >= 1048576.0;
  }
  if (wide_bwd) {
    BwdArgs a{};
    a.params = c->params; a.wprep = wprep;
    for (int e = 0; e < NM_MAX_EXO; ++e) a.exo[e] = e < p->n_exo ? c->exo[e] : nullptr;
    a.x_traj = c->x_traj; a.u_traj = c->u_traj; a.y_traj = c->y_traj; a.chk = c->checkpoints;
    a.grad_scalars = c->grad_scalars; a.gx_ext = c->gx_ext; a.gu_ext = c->gu_ext; a.gy_ext = c->gy_ext;
    a.G_x = c->G_x; a.G_u = c->G_u; a.G_y = c->G_y; a.fused_terms = c->fused_terms;
    a.grad_partials = c->grad_partials; a.grad_x0 = c->grad_x0;
    a.B = c->B; a.b_total = c->b_total; a.num_tiles = int((c->B + 127) / 128);
    char* extra = ws + (need_b - WD::bwd_extra(c->B, p->nsteps, d->sms));
    c->grid = d->sms;
    const int n = WD::launch_bwd(*p, a, c->params, wprep, extra, d->sms, d->sms, c->stream);
    c->extra_launches += n - 2;                      // the caller counts weight prep + adjoint
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { nm_set_error(cudaGetErrorString(e)); return NM_ERR_CUDA; }
    return NM_OK;
  }
  if (tc_bwd) { grid = d->sms * TcBwd<NM_SPEC_NAME, kTcBwd>::CTAS; if (grid > int((c->B + 127) / 128)) grid = int((c->B + 127) / 128); if (grid < 1) grid = 1; c->grid = grid; }
  else launch_prep(p, c->params, wprep, c->stream);
  BwdArgs a{};
  a.params = c->params; a.wprep = wprep;
  for (int e = 0; e < NM_MAX_EXO; ++e) a.exo[e] = e < p->n_exo ? c->exo[e] : nullptr;
  a.x_traj = c->x_traj; a.u_traj = c->u_traj; a.y_traj = c->y_traj; a.chk = c->checkpoints;
  a.grad_scalars = c->grad_scalars; a.gx_ext = c->gx_ext; a.gu_ext = c->gu_ext; a.gy_ext = c->gy_ext;
  a.G_x = c->G_x; a.G_u = c->G_u; a.G_y = c->G_y; a.fused_terms = c->fused_terms;
  a.grad_partials = c->grad_partials; a.grad_x0 = c->grad_x0;
  a.B = c->B; a.b_total = c->b_total; a.num_tiles = tc_bwd ? int((c->B + 127) / 128) : tiles;
  ExoPtrs ep{};
  for (int e = 0; e < NM_MAX_EXO; ++e) ep.p[e] = a.exo[e];
  if constexpr (kHoist) {
    // workspace tail G1[B][T][N0]: dZ_1 written by the adjoint; without a checkpoint buffer E is recomputed into it first
    // (the adjoint reads each entry of E before it overwrites it with dZ_1)
    float* G1 = reinterpret_cast<float*>(ws + (need_b - hoist_bytes(p, c->B)));
    if (c->checkpoints != nullptr) a.pre = c->checkpoints;
    else {
      const int64_t row_tiles = (c->B * p->nsteps + SB::TILE - 1) / SB::TILE;
      int ggrid = d->sms * 2; if (ggrid > row_tiles) ggrid = int(row_tiles); if (ggrid < 1) ggrid = 1;
      nm_exo_gemm_fwd_kernel<SB><<<ggrid, SB::THREADS, kExoSmem, c->stream>>>(*p, wprep, ep, G1, c->B);
      a.pre = G1;
      c->extra_launches += 1;
    }
    a.dz1 = G1;
  }
  if (tc_bwd) TcBwd<NM_SPEC_NAME, kTcBwd>::launch(*p, a, c->params, wprep, grid, c->stream);
  else nm_bwd_kernel<S><<<grid, S::THREADS, KC::BWD_SMEM_BYTES, c->stream>>>(*p, a);
  if constexpr (kHoist) {
    if (p->policy.trainable) {
      // single-pass TF32 on the tensor cores from 2^20 rows up (unbiased rounding noise, averaged over the rows of every
      // entry -- the rule of the tensor-core adjoint's weight gradients); the exact FFMA GEMM below that.  NM_TC_BWD forces.
      bool tc_gemm = kExoTcBwd && double(c->B) * double(p->nsteps) >= 1048576.0;
      if (kExoTcBwd) { const char* force = getenv("NM_TC_BWD"); if (force != nullptr) tc_gemm = force[0] != '0'; }
      if constexpr (kExoTcBwd) {
        if (tc_gemm) nm_exo_gemm_tc_bwd_kernel<SB><<<grid, 128, exo_tc_bwd_smem<kExoTcBwd>(), c->stream>>>(*p, ep, a.dz1, c->grad_partials, c->B);
      }
      if (!tc_gemm) nm_exo_gemm_bwd_kernel<SB><<<grid, SB::THREADS, kExoSmem, c->stream>>>(*p, ep, a.dz1, c->grad_partials, c->B);
      c->extra_launches += 1;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { nm_set_error(cudaGetErrorString(e)); return NM_ERR_CUDA; }
  return NM_OK;
}

char g_info[512];
NmKernelEntry g_entry;

struct Registrar {
  Registrar() {
    if (kWide)
      WD::info(g_info, sizeof(g_info), SB::TILE, SB::THREADS, KC::BWD_SMEM_BYTES, int(KC::DW_SPILL), S::NTERMS);
    else if (kTcFwd && kTcBwd)
      std::snprintf(g_info, sizeof(g_info), "fwd[tcgen05 3xTF32 tile=128 threads=128 tmem_cols=%d smem=%zu ctas/sm<=%d] bwd[tcgen05 tile=128 threads=%d ctas/sm=%d wgrad=%s; below 2^20 products or without checkpoints: ffma tile=%d threads=%d smem=%zu] compiled_terms=%d",
                    TK::C::FWD_COLS, TK::FWD_SMEM_BYTES, TK::MAX_CTAS, TcBwd<NM_SPEC_NAME, kTcBwd>::THREADS, TcBwd<NM_SPEC_NAME, kTcBwd>::CTAS,
                    TcBwd<NM_SPEC_NAME, kTcBwd>::WG3 ? "3xTF32" : "1xTF32-rn", SB::TILE, SB::THREADS,
                    KC::BWD_SMEM_BYTES, S::NTERMS);
    else if (kTcFwd)
      std::snprintf(g_info, sizeof(g_info), "fwd[tcgen05 3xTF32 tile=128 threads=128 tmem_cols=%d smem=%zu ctas/sm<=%d] bwd[tile=%d threads=%d w_smem=%d smem=%zu dw_spill=%d] minb=%d/%d compiled_terms=%d wprep_floats=%d",
                    TK::C::FWD_COLS, TK::FWD_SMEM_BYTES, TK::MAX_CTAS, SB::TILE, SB::THREADS, int(SB::W_SMEM), KC::BWD_SMEM_BYTES,
                    int(KC::DW_SPILL), S::MINB_FWD, S::MINB_BWD, S::NTERMS, KC::W_FLOATS_ALL);
    else
    std::snprintf(g_info, sizeof(g_info), "fwd[tile=%d threads=%d w_smem=%d smem=%zu] bwd[tile=%d threads=%d w_smem=%d smem=%zu dw_spill=%d] minb=%d/%d compiled_terms=%d wprep_floats=%d",
                  SF::TILE, SF::THREADS, int(SF::W_SMEM), KCF::FWD_SMEM_BYTES, SB::TILE, SB::THREADS, int(SB::W_SMEM), KC::BWD_SMEM_BYTES,
                  int(KC::DW_SPILL), S::MINB_FWD, S::MINB_BWD, S::NTERMS, KC::W_FLOATS_ALL);
    g_entry.signature = NM_SPEC_NAME::signature();
    g_entry.term_signature = NM_SPEC_NAME::term_signature();
    g_entry.workspace = workspace_fn;
    g_entry.forward = forward_fn;
    g_entry.backward = backward_fn;
    g_entry.info = g_info;
    g_entry.chk_floats_per_step = kWide ? NM_SPEC_NAME::NU : (kTcFwd ? TcMode<NM_SPEC_NAME>::CHK : (kHoist ? KC::N0 : -1));   // -1: the generic rule of nm_api.cu
    g_entry.next = nullptr;
    nm_register_kernel(&g_entry);
  }
} g_registrar;

}  // namespace
