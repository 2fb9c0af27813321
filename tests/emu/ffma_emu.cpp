// ffma_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  Runs the FFMA rollout / adjoint kernel templates of
// nm_kernels.cuh (and, with -DNM_TUNE_HOIST=1, the first-layer hoist GEMMs) for ONE shape class on the host, the way
// nm_spec_tu.cu launches them on the device, and writes trajectories, term sums and the parameter gradient to a file.
//   build:  g++ -std=c++20 -O1 -DNM_SPEC_HEADER='"<spec>.h"' [-DNM_TUNE_HOIST=1] -DEMU_TILE=128 -DEMU_THREADS=256 ...
//   run:    ffma_emu in.bin out.bin
// in.bin : int64 B | NmPlan bytes | params[n_params] | x0[B*nx] | exo_e[B*steps_e*width_e] ...
// out.bin: x[B*(T+1)*nx] | u[B*T*nu] | y[B*T*ny] | term_sums[n_terms] (double) | grad[n_params]
#include "cuda_emu.h"
EMU_DEFINE_SMEM
#include NM_SPEC_HEADER
#include "nm_kernels.cuh"

#ifndef NM_TUNE_HOIST
#define NM_TUNE_HOIST 0
#endif
#ifndef EMU_TILE
#define EMU_TILE 128
#endif
#ifndef EMU_THREADS
#define EMU_THREADS 256
#endif
#ifndef EMU_GRID
#define EMU_GRID 2            // fewer CTAs than tiles: exercises the persistent tile loop and the per-CTA gradient blocks
#endif

using namespace nm;

template <class Shape>
struct Tiled : Shape {
  static constexpr int TILE = EMU_TILE, THREADS = EMU_THREADS;
  static constexpr bool W_SMEM = false;                      // weights read from the prepared global block (no TMA staging)
  static constexpr int MINB_FWD = 1, MINB_BWD = 1;
  static constexpr bool HOIST = NM_TUNE_HOIST != 0;
};
using S = Tiled<NM_SPEC_NAME>;
using KC = KCfg<S>;
static_assert(KC::FITS_FWD && KC::FITS_BWD, "tile does not fit");
static_assert(KC::BWD_SMEM_FLOATS <= 64 * 1024 && KC::FWD_SMEM_FLOATS <= 64 * 1024, "emulated shared memory window too small");

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

  const int tiles = int((B + S::TILE - 1) / S::TILE);
  const int grid = std::min(tiles, EMU_GRID);
  std::vector<float> wprep(size_t(KC::W_FLOATS_ALL + KC::EXO_W_FLOATS) + 16, 0.f);
  std::vector<float> x(size_t(B) * (T + 1) * nx), u(size_t(B) * T * std::max<int>(nu, 1)), y(size_t(B) * T * std::max<int>(ny, 1));
  std::vector<double> partials(size_t(grid) * NM_MAX_TERMS, 0.0);
  std::vector<float> gpart(size_t(grid) * std::max<int64_t>(plan.n_params, 1), 0.f);
  std::vector<float> E(KC::HOIST ? size_t(B) * T * KC::N0 : 1, 0.f), G1(KC::HOIST ? size_t(B) * T * KC::N0 : 1, 0.f);
  ExoPtrs ep{};
  for (int e = 0; e < plan.n_exo; ++e) ep.p[e] = exo[e].data();

  // weight preparation (nm_spec_tu.cu: launch_prep)
  emu::launch(2, 256, [&] { nm_prep_kernel<S>(plan, params.data(), wprep.data()); });
  if constexpr (KC::HOIST) emu::launch(grid, S::THREADS, [&] { nm_exo_gemm_fwd_kernel<S>(plan, wprep.data(), ep, E.data(), B); });

  FwdArgs fa{};
  fa.params = params.data(); fa.wprep = wprep.data(); fa.x0 = x0.data();
  for (int e = 0; e < plan.n_exo; ++e) fa.exo[e] = exo[e].data();
  fa.x_traj = x.data(); fa.u_traj = u.data(); fa.y_traj = y.data(); fa.chk = nullptr;
  fa.term_partials = partials.data(); fa.fused_terms = 1; fa.B = B; fa.num_tiles = tiles; fa.pre = E.data();
  emu::launch(grid, S::THREADS, [&] { nm_fwd_kernel<S>(plan, fa); });

  BwdArgs ba{};
  ba.params = params.data(); ba.wprep = wprep.data();
  for (int e = 0; e < plan.n_exo; ++e) ba.exo[e] = exo[e].data();
  ba.x_traj = x.data(); ba.u_traj = u.data(); ba.y_traj = y.data(); ba.chk = nullptr;
  ba.grad_scalars = nullptr; ba.grad_partials = gpart.data(); ba.grad_x0 = nullptr; ba.fused_terms = 1;
  ba.B = B; ba.b_total = B; ba.num_tiles = tiles;
#ifdef EMU_ALIAS_DZ1
  // the no-checkpoint configuration of nm_spec_tu.cu: E recomputed into the workspace tail, dZ_1 overwrites it in place
  ba.pre = E.data(); ba.dz1 = E.data();
#else
  ba.pre = E.data(); ba.dz1 = G1.data();
#endif
  emu::launch(grid, S::THREADS, [&] { nm_bwd_kernel<S>(plan, ba); });
  if constexpr (KC::HOIST) emu::launch(grid, S::THREADS, [&] { nm_exo_gemm_bwd_kernel<S>(plan, ep, ba.dz1, gpart.data(), B); });

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
