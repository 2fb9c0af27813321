// In-kernel all-reduce of the flat [gradient | loss scalars] buffer over NVLink / NVSwitch peer memory.
//
// The data-parallel exchange of the path is ONE sum over ranks of the few-hundred-KB buffer the adjoint's finalise kernel
// wrote (reference: Lightning DDP's gradient averaging, trainer.py:49,112-116).  Instead of handing it to NCCL, the
// finalise kernel writes it straight into this rank's half of a SYMMETRIC buffer (same allocation on every rank, mapped
// into every peer's address space and -- on NVSwitch systems -- behind one multicast address), and one small kernel
//   1. publishes "my half of epoch e is written" to every peer (system-scope release store into the peer's flag word)
//      and waits for every peer's flag (system-scope acquire loads),
//   2. reads the sum: `multimem.ld_reduce.add.v4.f32` on the multicast address -- the reduction happens INSIDE the switch
//      (NVLS), one load per 16 bytes whatever the number of ranks -- or, without multicast support, plain peer loads
//      summed in rank order (bitwise identical on every rank),
//   3. writes it to the caller's output (which may be the parameter-gradient arena itself).
// Halves alternate with the epoch parity, so no trailing barrier is needed: a rank can only be one epoch ahead (it cannot
// pass the flag wait of epoch e+1 before every peer has arrived there, i.e. has finished reading epoch e).
// Flags are monotonic epoch numbers (no reset race).  All spin loops are bounded: a peer that never arrives sets the
// status word instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nm {

constexpr int kArMaxRanks = 16;
constexpr int kArBlocks = 16, kArThreads = 256;
// symmetric buffer layout (floats): [half 0: npad][half 1: npad][kArFlagWords 32-bit words], npad = n rounded up to 64
//   word q < kArMaxRanks: arrival flag of rank q;  word kArMaxRanks: block release flag;  word kArMaxRanks + 1: status
constexpr int kArFlagWords = 64;
constexpr long long kArTimeoutCycles = 60000000000LL;     // ~30 s (first-step skew between ranks can be seconds): a missing peer is an error, not a hang

struct ArPeers { float* buf[kArMaxRanks]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}

// wraparound-safe "flag has reached epoch"
__device__ __forceinline__ bool ar_reached(uint32_t flag, uint32_t epoch) { return int32_t(flag - epoch) >= 0; }

__global__ void __launch_bounds__(kArThreads) nm_grad_allreduce_kernel(const __grid_constant__ ArPeers peers, const float* __restrict__ mc,
                                                                      float* __restrict__ out, int64_t n, int64_t npad, int rank, int world,
                                                                      uint32_t epoch, float scale_tail, int64_t tail_from) {
  const int tid = threadIdx.x;
  float* mine = peers.buf[rank];
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine + 2 * npad);
  __shared__ int ok_s;
  if (tid == 0) ok_s = 1;
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < world) {
      // this rank's half (written by the finalise kernel before this launch) is ordered before the flag for every peer
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(peers.buf[tid] + 2 * npad) + rank, epoch);
      const long long t0 = clock64();
      while (!ar_reached(ld_acquire_sys(my_flags + tid), epoch)) {
        if (clock64() - t0 > kArTimeoutCycles) { ok_s = 0; break; }
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (!ok_s) my_flags[kArMaxRanks + 1] = 1u;                 // status: a peer did not arrive
      __threadfence();
      st_release_gpu(my_flags + kArMaxRanks, ok_s ? epoch : epoch | 0x80000000u);
    }
  } else {
    if (tid == 0) {
      const long long t0 = clock64();
      uint32_t f;
      while (true) {
        f = ld_acquire_gpu(my_flags + kArMaxRanks);
        if ((f & 0x7fffffffu) == (epoch & 0x7fffffffu)) break;
        if (clock64() - t0 > 2 * kArTimeoutCycles) { f = 0x80000000u; break; }
      }
      if (f & 0x80000000u) ok_s = 0;
    }
    __syncthreads();
  }
  if (!ok_s) return;
  const int64_t off = int64_t(epoch & 1u) * npad;
  const int64_t n4 = npad / 4;
  for (int64_t i = int64_t(blockIdx.x) * kArThreads + tid; i < n4; i += int64_t(gridDim.x) * kArThreads) {
    float4 v;
    if (mc != nullptr) {
      v = multimem_ld_reduce_add_v4(mc + off + 4 * i);           // summed inside the switch
    } else {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < world; ++q) {                          // rank order: bitwise identical on every rank
        const float4 w = ld_relaxed_sys_v4(peers.buf[q] + off + 4 * i);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
    }
    float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t idx = 4 * i + j;
      if (idx < n) out[idx] = idx >= tail_from ? e[j] * scale_tail : e[j];     // loss scalars are averaged, gradients summed
    }
  }
}

}  // namespace nm
