// tc_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h, tc_ptx_emu.h).  Runs the tcgen05 rollout / adjoint kernel
// templates of nm_tc_kernels.cuh for ONE shape class on the host, launched the way nm_spec_tu.cu launches them.
// Same file formats as ffma_emu.cpp.
#include "cuda_emu.h"
EMU_DEFINE_SMEM
#include NM_SPEC_HEADER
#include "nm_kernels.cuh"
#include "nm_tc_kernels.cuh"

#ifndef EMU_GRID
#define EMU_GRID 2
#endif
#ifndef NM_TUNE_HOIST
#define NM_TUNE_HOIST 0
#endif

using namespace nm;
using S = NM_SPEC_NAME;
// FFMA-side view of the class (weight preparation and the hoist GEMMs run through it, as in nm_spec_tu.cu)
template <class Shape>
struct Tiled : Shape {
  static constexpr int TILE = 128, THREADS = 256;
  static constexpr bool W_SMEM = false;
  static constexpr int MINB_FWD = 1, MINB_BWD = 1;
  static constexpr bool HOIST = NM_TUNE_HOIST != 0;
};
using SB = Tiled<NM_SPEC_NAME>;
using KC = KCfg<SB>;
using TK = TcKCfg<S>;
using BK = TcBKCfg<S>;
using MD = TcMode<S>;
static_assert(TK::ELIGIBLE && TK::FITS_FWD && BK::ELIGIBLE, "shape class is not on the tensor-core path");
static_assert(KC::HOIST == MD::HOIST, "the FFMA and the tensor-core kernels disagree on the first-layer hoist");
static_assert(BK::SMEM_BYTES <= 4 * 64 * 1024 && TK::FWD_SMEM_BYTES <= 4 * 64 * 1024, "emulated shared memory window too small");

template <class T> static std::vector<T> read_vec(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int64_t B = 0;
  if (fread(&B, sizeof(B), 1, f) != 1) return 2;
  NmPlan plan;
  if (fread(&plan, sizeof(plan), 1, f) != 1) return 2;
  const int T = plan.nsteps, nx = plan.nx, nu = plan.nu, ny = plan.ny;
  auto params = read_vec<float>(f, plan.n_params);
  auto x0 = read_vec<float>(f, size_t(B) * nx);
  std::vector<std::vector<float>> exo(plan.n_exo);
  for (int e = 0; e < plan.n_exo; ++e) exo[e] = read_vec<float>(f, size_t(B) * plan.exo[e].steps * plan.exo[e].width);
  fclose(f);

  const int tiles = int((B + 127) / 128);
  const int grid = std::min(tiles, EMU_GRID);
  const size_t wfloats = std::max(size_t(TK::FWD_W_BYTES), size_t(BK::W_BYTES)) / 4 + 64;
  std::vector<float> wprep(wfloats, 0.f);
  std::vector<float> x(size_t(B) * (T + 1) * nx), u(size_t(B) * T * std::max<int>(nu, 1)), y(size_t(B) * T * std::max<int>(ny, 1));
  std::vector<float> chk(size_t(B) * T * std::max<int>(MD::CHK, 1), 0.f);
  std::vector<double> partials(size_t(grid) * NM_MAX_TERMS, 0.0);
  std::vector<float> gpart(size_t(grid) * std::max<int64_t>(plan.n_params, 1), 0.f);

  // first-layer hoist: E = bias + exogenous part, for every (trajectory, step), into the head of the checkpoint buffer
  std::vector<float> wprep_ffma(size_t(KC::W_FLOATS_ALL + KC::EXO_W_FLOATS) + 16, 0.f), G1(MD::HOIST ? size_t(B) * T * MD::N0 : 1, 0.f);
  ExoPtrs ep{};
  for (int e = 0; e < plan.n_exo; ++e) ep.p[e] = exo[e].data();
  if constexpr (MD::HOIST) {
    emu::launch(2, 256, [&] { nm_prep_kernel<SB>(plan, params.data(), wprep_ffma.data()); });
#ifdef EMU_TC_GEMM
    // the tensor-core version of the forward hoist GEMM
    using X = ExoTcCfg<SB>;
    static_assert(X::OK, "tensor-core hoist GEMM not available for this class");
    std::vector<float> wexo(size_t(X::W_FLOATS) + 64, 0.f);
    emu::launch(4, 256, [&] { nm_exo_tc_prep_kernel<SB>(plan, params.data(), wexo.data()); });
    emu::launch(grid, 288, [&] { nm_exo_gemm_tc_fwd_kernel<SB>(plan, wexo.data(), ep, chk.data(), B); });
#else
    emu::launch(grid, SB::THREADS, [&] { nm_exo_gemm_fwd_kernel<SB>(plan, wprep_ffma.data(), ep, chk.data(), B); });
#endif
  }

  FwdArgs fa{};
  fa.pre = chk.data();
  fa.params = params.data(); fa.wprep = wprep.data(); fa.x0 = x0.data();
  for (int e = 0; e < plan.n_exo; ++e) fa.exo[e] = exo[e].data();
  fa.x_traj = x.data(); fa.u_traj = u.data(); fa.y_traj = y.data(); fa.chk = chk.data();
  fa.term_partials = partials.data(); fa.fused_terms = 1; fa.B = B; fa.num_tiles = tiles;
  emu::launch(8, 256, [&] { nm_tc_prep_fwd_kernel<S>(plan, params.data(), wprep.data()); });
  emu::launch(grid, 128, [&] { nm_tc_fwd_kernel<S>(plan, fa); });

  BwdArgs ba{};
  ba.params = params.data(); ba.wprep = wprep.data();
  for (int e = 0; e < plan.n_exo; ++e) ba.exo[e] = exo[e].data();
  ba.x_traj = x.data(); ba.u_traj = u.data(); ba.y_traj = y.data(); ba.chk = chk.data();
  ba.grad_scalars = nullptr; ba.grad_partials = gpart.data(); ba.grad_x0 = nullptr; ba.fused_terms = 1;
  ba.B = B; ba.b_total = B; ba.num_tiles = tiles;
  ba.pre = chk.data(); ba.dz1 = G1.data();
  emu::launch(16, 256, [&] { nm_tc_prep_bwd_kernel<S>(plan, params.data(), wprep.data()); });
  emu::launch(grid, BK::THREADS, [&] { nm_tc_bwd_kernel<S>(plan, ba); });
#ifdef EMU_TC_GEMM
  if constexpr (MD::HOIST) {
    static_assert(ExoTcBwdCfg<SB>::OK, "tensor-core backward hoist GEMM not available for this class");
    emu::launch(grid, 128, [&] { nm_exo_gemm_tc_bwd_kernel<SB>(plan, ep, G1.data(), gpart.data(), B); });
  }
#else
  if constexpr (MD::HOIST) emu::launch(grid, SB::THREADS, [&] { nm_exo_gemm_bwd_kernel<SB>(plan, ep, G1.data(), gpart.data(), B); });
#endif

  std::vector<double> tsum(std::max<int>(plan.n_terms, 1), 0.0);
  for (int k = 0; k < plan.n_terms; ++k)
    for (int c = 0; c < grid; ++c) tsum[k] += partials[size_t(c) * NM_MAX_TERMS + k];
  std::vector<float> grad(std::max<int64_t>(plan.n_params, 1), 0.f);
  for (int i = 0; i < plan.n_params; ++i)
    for (int c = 0; c < grid; ++c) grad[i] += gpart[size_t(c) * plan.n_params + i];

  FILE* o = fopen(argv[2], "wb");
  fwrite(x.data(), 4, size_t(B) * (T + 1) * nx, o);
  fwrite(u.data(), 4, size_t(B) * T * nu, o);
  fwrite(y.data(), 4, size_t(B) * T * ny, o);
  fwrite(tsum.data(), 8, plan.n_terms, o);
  fwrite(grad.data(), 4, plan.n_params, o);
  fclose(o);
  return 0;
}
