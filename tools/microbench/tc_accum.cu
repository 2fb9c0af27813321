// How does the tcgen05 fp32 accumulator round?  Accumulate R cycles of a 16-k-step TF32 product (inputs
// exactly representable in tf32) into one TMEM accumulator and compare with the exact sum: error growing
// linearly in R (always the same sign) = truncation, ~sqrt(R) = round-to-nearest.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
constexpr int K = 128, N = 16;
// A [128][K], B [N][K] interleaved K-major images
__global__ void __launch_bounds__(128) accum_kernel(const float* __restrict__ aimg, const float* __restrict__ bimg, float* __restrict__ D, int cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* sa = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sb = sa + 128 * K;
  __shared__ uint64_t bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) sa[i] = aimg[i];
  for (int i = tid; i < N * K; i += 128) sb[i] = bimg[i];
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n" ::"r"(smem_u32(&tb)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tb;
  if (tid == 0) {
    for (int c = 0; c < cycles; ++c)
#pragma unroll
      for (int ks = 0; ks < K / 8; ++ks)
        mma_ss(tmem, make_desc(smem_u32(sa) + ks * 2 * 128 * 16, 128 * 16, 128, 0), make_desc(smem_u32(sb) + ks * 2 * N * 16, N * 16, 128, 0), make_idesc(128, N), (c | ks) != 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(tmem + (uint32_t(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  for (int i = 0; i < 16; ++i) D[tid * N + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;\n" ::"r"(tmem));
}
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
static inline int il(int r, int c, int R) { return (c >> 2) * (R * 4) + r * 4 + (c & 3); }
int main() {
  for (int positive = 0; positive < 2; ++positive) {
    std::vector<float> A(128 * K), B(N * K), ai(128 * K), bi(N * K);
    srand(11);
    for (auto& v : A) v = tf32_rna(positive ? rand() / float(RAND_MAX) : (rand() / float(RAND_MAX)) * 2.f - 1.f);
    for (auto& v : B) v = tf32_rna(positive ? rand() / float(RAND_MAX) : (rand() / float(RAND_MAX)) * 2.f - 1.f);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < K; ++c) ai[il(r, c, 128)] = A[r * K + c];
    for (int r = 0; r < N; ++r) for (int c = 0; c < K; ++c) bi[il(r, c, N)] = B[r * K + c];
    float *da, *db, *dD;
    CK(cudaMalloc(&da, ai.size() * 4)); CK(cudaMalloc(&db, bi.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(da, ai.data(), ai.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, bi.data(), bi.size() * 4, cudaMemcpyHostToDevice));
    const int smem = (128 * K + N * K) * 4 + 2048;
    CK(cudaFuncSetAttribute(accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<double> one(128 * N);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += double(A[m * K + k]) * B[n * K + k]; one[m * N + n] = s; }
    for (int cycles : {1, 16, 256, 4096, 65536}) {
      accum_kernel<<<1, 128, smem>>>(da, db, dD, cycles);
      CK(cudaDeviceSynchronize());
      std::vector<float> D(128 * N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double sum_rel = 0, max_rel = 0, sum_signed = 0; int cnt = 0;
      for (int i = 0; i < 128 * N; ++i) {
        const double ex = one[i] * cycles;
        if (fabs(one[i]) < 1.0) continue;                   // look at well-conditioned entries only
        const double rel = (D[i] - ex) / fabs(ex);
        sum_rel += fabs(rel); sum_signed += rel * (ex > 0 ? 1 : -1); max_rel = fmax(max_rel, fabs(rel)); ++cnt;
      }
      printf("%s inputs, %6d cycles x 16 accumulations: mean |rel err| %.3e  mean signed (toward +inf of |x|) %.3e  max %.3e   (%d entries; fp32 ulp 6e-8)\n",
             positive ? "positive" : "signed  ", cycles, sum_rel / cnt, sum_signed / cnt, max_rel, cnt);
    }
    cudaFree(da); cudaFree(db); cudaFree(dD);
  }
  return 0;
}
