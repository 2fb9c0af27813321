// tcgen05.mma issue/throughput for the small shapes of the fused rollout: back-to-back kind::tf32 MMAs with
// precomputed descriptors (no per-MMA address math), one commit at the end.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4); d |= uint64_t((lbo >> 4) & 0x3FFF) << 16; d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46; d |= uint64_t(layout & 7) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// MODE 0: A smem interleaved K-major;  1: A in TMEM;  2: both SW128 K-major (wgrad form)
template <int MODE, int M, int N, int NMMA>
__global__ void __launch_bounds__(128) rate_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40960; i += 128) reinterpret_cast<float*>(sm)[i] = 0.001f * (i & 63);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tb)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tb;
  constexpr uint32_t idesc = make_idesc(M, N);
  uint32_t phase = 0;
  long long best = 1ll << 60, lat = 0;
  for (int rep = 0; rep < 6; ++rep) {
    long long t0 = clock64(), t1 = 0;
    if (tid == 0) {
      const uint32_t abase = smem_u32(sm), bbase = smem_u32(sm) + 65536;
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        const int ks = i & 3;
        if constexpr (MODE == 0) mma_ss(tmem, make_desc(abase + ks * 2 * 128 * 16, 128 * 16, 128, 0), make_desc(bbase + ks * 2 * N * 16, N * 16, 128, 0), idesc, i > 0);
        else if constexpr (MODE == 1) mma_ts(tmem, tmem + 256 + ks * 8, make_desc(bbase + ks * 2 * N * 16, N * 16, 128, 0), idesc, i > 0);
        else mma_ss(tmem, make_desc(abase + ks * 32, 16, 1024, 2), make_desc(bbase + ks * 32, 16, 1024, 2), idesc, i > 0);
      }
      t1 = clock64();
      tc_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    long long t2 = clock64();
    if (tid == 0 && t2 - t0 < best) { best = t2 - t0; lat = t1 - t0; }
    __syncthreads();
  }
  if (tid == 0) { out[0] = best; out[1] = lat; }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}
template <int MODE, int M, int N>
void run() {
  long long* d; CK(cudaMalloc(&d, 16));
  long long h8[2], h64[2];
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(rate_kernel<MODE, M, N, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate_kernel<MODE, M, N, 72>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rate_kernel<MODE, M, N, 8><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h8, d, 16, cudaMemcpyDeviceToHost));
  rate_kernel<MODE, M, N, 72><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h64, d, 16, cudaMemcpyDeviceToHost));
  printf("mode %d (%s) M=%3d N=%3d:  8 MMAs %5lld clk (issue %4lld)   72 MMAs %5lld clk (issue %4lld)   => %.1f clk/MMA  (ideal %.1f)\n", MODE,
         MODE == 0 ? "A smem interleaved" : MODE == 1 ? "A tmem" : "SW128 K-major both", M, N, h8[0], h8[1], h64[0], h64[1], double(h64[0] - h8[0]) / 64.0, double(M) * N * 8 / 1443.0);
  cudaFree(d);
}
int main() {
  run<0, 128, 16>(); run<0, 128, 32>(); run<0, 128, 48>(); run<0, 128, 64>(); run<0, 128, 96>(); run<0, 128, 128>(); run<0, 128, 256>();
  run<1, 128, 16>(); run<1, 128, 32>(); run<1, 128, 48>(); run<1, 128, 64>(); run<1, 128, 128>();
  run<2, 128, 16>(); run<2, 128, 32>(); run<2, 128, 48>(); run<2, 128, 64>(); run<2, 128, 96>(); run<2, 128, 128>();
  run<2, 64, 16>(); run<2, 64, 32>(); run<2, 64, 64>(); run<2, 64, 96>();
  run<0, 64, 32>(); run<0, 64, 64>();
  return 0;
}
