// Round-2 probe (written at the end of round 1, NOT YET RUN): is the 44-cycle floor of a small tcgen05.mma an ISSUE limit
// of the one issuing thread or a minimum occupancy of the tensor pipe?  profiles/r01_tcgen05_mma_rate.txt shows the
// issuing thread itself needing 44 cycles per MMA for N <= 64.  Here 1 or 2 warps (one elected lane each) issue
// back-to-back MMAs into DIFFERENT accumulators; if two issuers finish 2 x NMMA MMAs in the time one issuer needs for
// NMMA, narrow layers (c1 / c2: N = 32, c3: N = 64) can halve their MMA time by splitting the K loop over two issuers.
// Also: kind::f16 (K = 16 per instruction) next to kind::tf32 (K = 8), N up to 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_rate2 tc_rate2.cu && ./tc_rate2
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4); d |= uint64_t((lbo >> 4) & 0x3FFF) << 16; d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
// D f32; A/B format: 2 = tf32, 0 = f16; both K-major
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) { return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int KIND, int N, int NMMA, int ISSUERS>
__global__ void __launch_bounds__(128) rate2_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2]; __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40960; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;          // zeros are valid in every format
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tb)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tb;
  constexpr uint32_t idesc = make_idesc(KIND == 0 ? 2u : 0u, 128, N);
  uint32_t phase = 0;
  long long best = 1ll << 60;
  for (int rep = 0; rep < 6; ++rep) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp < ISSUERS) {
      // A: [k/chunk][128 rows][16 B], B: [k/chunk][N rows][16 B]  (interleaved K-major, no swizzle); 4 K-steps reused
      const uint32_t abase = smem_u32(sm), bbase = smem_u32(sm) + 65536;
      const uint32_t acc_col = tmem + uint32_t(warp) * 256u;
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          const int ks = i & 3;
          mma_ss<KIND>(acc_col, make_desc(abase + ks * 2 * 128 * 16, 128 * 16, 128), make_desc(bbase + ks * 2 * N * 16, N * 16, 128), idesc, i > 0);
        }
        tc_commit(&bar[warp]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[0], phase);
    if (ISSUERS > 1) mbar_wait(&bar[1], phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const long long t2 = clock64();
    if (tid == 0 && t2 - t0 < best) best = t2 - t0;
  }
  if (tid == 0) out[0] = best;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

template <int KIND, int N, int ISSUERS>
void run() {
  long long* d; CK(cudaMalloc(&d, 16));
  long long h8 = 0, h72 = 0;
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(rate2_kernel<KIND, N, 8, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate2_kernel<KIND, N, 72, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rate2_kernel<KIND, N, 8, ISSUERS><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h8, d, 8, cudaMemcpyDeviceToHost));
  rate2_kernel<KIND, N, 72, ISSUERS><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h72, d, 8, cudaMemcpyDeviceToHost));
  printf("kind::%s N=%3d issuers %d:  %d x 8 MMAs %5lld clk   %d x 72 MMAs %5lld clk   => %.1f clk per MMA of EACH issuer, %.1f clk per MMA overall\n",
         KIND == 0 ? "tf32" : "f16 ", N, ISSUERS, ISSUERS, h8, ISSUERS, h72, double(h72 - h8) / 64.0, double(h72 - h8) / 64.0 / ISSUERS);
  cudaFree(d);
}
int main() {
  run<0, 32, 1>(); run<0, 32, 2>(); run<0, 64, 1>(); run<0, 64, 2>(); run<0, 128, 1>(); run<0, 128, 2>(); run<0, 256, 1>(); run<0, 256, 2>();
  run<1, 32, 1>(); run<1, 32, 2>(); run<1, 64, 1>(); run<1, 64, 2>(); run<1, 128, 1>(); run<1, 128, 2>(); run<1, 256, 1>(); run<1, 256, 2>();
  return 0;
}
