// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A minimal host emulation of the CUDA execution model (one OS thread per CUDA
// thread of ONE block at a time, std::barrier for __syncthreads) so that the FFMA kernel templates of
// neuromancer_b200/csrc/nm_kernels.cuh can be compiled with g++ and their LOGIC (index math, layouts, staging order)
// checked on a machine without a GPU.  Nothing in the product imports, links or ships this.
#pragma once
#include <atomic>
#include <barrier>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>
#include <algorithm>

#include <cuda_runtime.h>          // host mode: vector types (float4, ...), make_float4, cudaStream_t

#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#undef __forceinline__
#define __forceinline__ inline
#undef __device__
#define __device__
#undef __global__
#define __global__
#undef __host__
#define __host__
#undef __shared__
#define __shared__
#undef __noinline__
#define __noinline__
#undef __align__
#define __align__(n)

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
inline thread_local Idx t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline std::barrier<>* g_block_barrier = nullptr;
inline double g_xchg[2048];
inline std::unique_ptr<std::barrier<>> g_warp_bar[64];
inline std::mutex g_named_mu;
inline std::unique_ptr<std::barrier<>> g_named[16];
inline std::barrier<>* named_barrier(int id, int n) {                  // bar.sync id, n
  std::lock_guard<std::mutex> l(g_named_mu);
  if (!g_named[id]) g_named[id] = std::make_unique<std::barrier<>>(n);
  return g_named[id].get();
}
}  // namespace emu
#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim

// the dynamic shared memory window of the (single) running block: the kernels declare `extern __shared__ float smem[]`
// at function scope inside namespace nm, which here is an ordinary extern declaration of this array
namespace nm { extern float smem[]; }
#define EMU_DEFINE_SMEM                                                                                  \
  namespace nm {                                                                                         \
  alignas(1024) float smem[64 * 1024];                                                                   \
  /* the tensor-core adjoint declares its window as `uint8_t smem_raw[]`: the same memory */             \
  extern uint8_t smem_raw[] __attribute__((alias("_ZN2nm4smemE")));                                      \
  }

inline void __syncthreads() { emu::g_block_barrier->arrive_and_wait(); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline void __trap() { std::abort(); }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }   // (the PTX helpers that use it are never called here)
// warp shuffle, emulated block-wide (every thread of the block calls it in the same program order in these kernels)
inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  const unsigned t = emu::t_threadIdx.x;
  emu::g_xchg[t] = v;
  emu::g_block_barrier->arrive_and_wait();
  const double r = emu::g_xchg[t ^ unsigned(lane_mask)];
  emu::g_block_barrier->arrive_and_wait();
  return r;
}
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
// a real barrier among the 32 threads of the warp: on the hardware the lanes of a converged warp advance in lockstep; here
// it keeps the non-elected lanes of a control warp within one step of the elected one (otherwise a slow lane can miss
// two phase flips of an mbarrier and wait forever)
inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_warp_bar[emu::t_threadIdx.x >> 5]->arrive_and_wait(); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
using std::min;
using std::max;

namespace emu {
// run `fn()` as a grid of `grid` blocks of `threads` threads, one block after the other
template <class F>
void launch(int grid, int threads, F&& fn) {
  for (int b = 0; b < grid; ++b) {
    std::barrier<> bar(threads);
    g_block_barrier = &bar;
    for (auto& nb : g_named) nb.reset();
    for (int w = 0; w < (threads + 31) / 32; ++w) g_warp_bar[w] = std::make_unique<std::barrier<>>(std::min(32, threads - 32 * w));
    std::vector<std::thread> th;
    th.reserve(threads);
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&, t] {
        t_threadIdx.x = unsigned(t); t_blockIdx.x = unsigned(b); t_blockDim.x = unsigned(threads); t_gridDim.x = unsigned(grid);
        fn();
      });
    for (auto& x : th) x.join();
  }
}
}  // namespace emu
