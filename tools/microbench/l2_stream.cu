// Round-2 probe (written at the end of round 1, NOT YET RUN): budgets that decide how C5's 128-wide policy can reach the
// tensor cores.
//   (a) L2 -> shared-memory streaming rate per SM when EVERY SM streams the same L2-resident weight block through a
//       two-slot ring with bulk TMA copies (cp.async.bulk + mbarrier, the mechanism of nm_tc_bwd_kernel's weight ring):
//       3 x 139 KB of 3xTF32 images per 128-trajectory tile-step is 37 B/cycle/SM if a step takes 12 k cycles.
//   (b) RED.ADD.F32 throughput per SM into a CTA-private, L2-resident block (what flushing per-step weight-gradient
//       tiles of 53.5 k floats would cost).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu && ./l2_stream
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (a) every CTA streams `total` bytes of `src` (wrapping) through two `slot`-byte shared-memory slots
__global__ void __launch_bounds__(128) stream_kernel(const uint8_t* __restrict__ src, size_t src_bytes, uint32_t slot, int copies, long long* cycles, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  uint32_t ph[2] = {0, 0};
  size_t off = (size_t(blockIdx.x) * slot) % (src_bytes - slot);          // CTAs start at different chunks of the block
  off &= ~size_t(127);
  float acc = 0.f;
  const long long t0 = clock64();
  if (tid == 0) { mbar_expect_tx(&bar[0], slot); tma_bulk_g2s(sm, src + off, slot, &bar[0]); }
  for (int c = 0; c < copies; ++c) {
    const int s = c & 1;
    if (tid == 0 && c + 1 < copies) {                                     // next chunk into the other slot (its readers passed the barrier below)
      off += slot; if (off + slot > src_bytes) off = 0;
      mbar_expect_tx(&bar[s ^ 1], slot);
      tma_bulk_g2s(sm + size_t(s ^ 1) * slot, src + off, slot, &bar[s ^ 1]);
    }
    mbar_wait(&bar[s], ph[s]); ph[s] ^= 1;
    acc += reinterpret_cast<const float*>(sm + size_t(s) * slot)[tid];   // touch the slot
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

// (b) every thread fires `n` RED.ADD into its CTA's private block of `block_floats` floats
__global__ void __launch_bounds__(256) red_kernel(float* __restrict__ blocks, int block_floats, int n, long long* cycles) {
  float* mine = blocks + size_t(blockIdx.x) * block_floats;
  const long long t0 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < n; ++i) {
    atomicAdd(mine + idx, 1.0f);                                          // result unused -> RED
    idx += blockDim.x; if (idx >= block_floats) idx -= block_floats;
  }
  __threadfence();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const size_t src_bytes = 448 * 1024;                                    // ~ the 3xTF32 images of C5's three hidden layers
  uint8_t* src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 0, src_bytes));
  long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * sms * 2));
  float* sink; CK(cudaMalloc(&sink, 4));
  long long* h = (long long*)malloc(sizeof(long long) * sms * 2);
  for (uint32_t slot : {16384u, 32768u, 65536u, 98304u}) {
    for (int ctas : {1, sms}) {
      const int copies = 256;
      CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * slot)));
      stream_kernel<<<ctas, 128, 2 * slot>>>(src, src_bytes, slot, copies, cyc, sink);   // warm-up: block becomes L2 resident
      stream_kernel<<<ctas, 128, 2 * slot>>>(src, src_bytes, slot, copies, cyc, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bpc = double(slot) * copies / double(mx);
      printf("stream  slot %6u B  %3d CTAs: %.1f B/cycle/SM  = %.2f TB/s aggregate at %.0f MHz\n", slot, ctas, bpc, bpc * ctas * khz * 1e3 / 1e12, khz / 1e3);
    }
  }
  const int block_floats = 53510;                                         // C5's parameter count
  float* blocks; CK(cudaMalloc(&blocks, sizeof(float) * size_t(block_floats) * sms)); CK(cudaMemset(blocks, 0, sizeof(float) * size_t(block_floats) * sms));
  for (int ctas : {1, sms}) {
    const int n = 2048;
    red_kernel<<<ctas, 256>>>(blocks, block_floats, n, cyc);
    red_kernel<<<ctas, 256>>>(blocks, block_floats, n, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("red.add %3d CTAs x 256 threads x %d: %.2f RED/cycle/SM  (a 53.5 k-float gradient flush = %.0f cycles)\n", ctas, n, 256.0 * n / double(mx), 53510.0 / (256.0 * n / double(mx)));
  }
  return 0;
}
