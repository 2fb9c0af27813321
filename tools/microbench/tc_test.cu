// Standalone validation of the tcgen05 building block planned for the MLP layers:
//   D[m][n] = sum_k A[k][m] * W[n][k]     (m = trajectory in the CTA tile, M = 128)
// operands in shared memory as  [feature row][32-element block]  tiles with 128-byte rows and the
// 128B swizzle (MN-major for this GEMM), fp32-accurate through the 3xTF32 split
//   D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo,
// accumulator in TMEM, read back with tcgen05.ld (thread = TMEM lane = trajectory).
// Also tests the weight-gradient form  G[n][k] = sum_m Dl[n][m] * A[k][m]  with the SAME physical
// tiles interpreted K-major.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- tile layout: element (row r, column c) of a [R][C] matrix, C split in blocks of 32 floats (128 B rows),
// block j holds rows 0..R-1 consecutively; 16-byte chunk index XORed with (row % 8)   (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ int tile_off(int r, int c, int R) {
  const int blk = c >> 5, cc = c & 31;
  return blk * (R * 32) + r * 32 + ((((cc >> 2) ^ (r & 7)) << 2) | (cc & 3));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);                  // start address, bits [0,14)
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;         // leading byte offset, bits [16,30)
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;         // stride byte offset, bits [32,46)
  d |= uint64_t(1) << 46;                                 // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;                                 // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)            // D format F32
       | (2u << 7)            // A format TF32
       | (2u << 10)           // B format TF32
       | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16)
       | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int M = 128;   // trajectories per CTA tile
// MODE 0: forward form   D[m][n] = sum_k A[k][m] W[n][k]      A tile [K][128] MN-major, B = Wt tile [K][N] MN-major
// MODE 1: weight-grad    G[n][k] = sum_m Dl[n][m] A[k][m]      A-op = Dl tile [128 (n padded)][128 m] K-major, B-op = A tile [K][128 m] K-major
template <int MODE, int K, int N>
__global__ void __launch_bounds__(128) tc_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ Dg, int split) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int AR = MODE == 0 ? K : 128;            // rows of operand-A tile
  constexpr int AC = 128;                            // columns (trajectories)
  constexpr int BR = K;                              // rows of operand-B tile
  constexpr int BC = MODE == 0 ? N : 128;
  float* a_hi = sm;
  float* a_lo = a_hi + AR * AC;
  float* b_hi = a_lo + AR * AC;
  float* b_lo = b_hi + BR * BC;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // stage operands (hi/lo split) into the swizzled tiles
  for (int i = tid; i < AR * AC; i += 128) {
    const int r = i / AC, c = i % AC;
    float v = 0.f;
    if (MODE == 0) v = Ag[r * 128 + c]; else v = (r < N) ? Ag[r * 128 + c] : 0.f;   // MODE 1: Dl [N][128], rows >= N zero
    const float h = split ? tf32_hi(v) : v;
    a_hi[tile_off(r, c, AR)] = h;
    a_lo[tile_off(r, c, AR)] = split ? tf32_hi(v - h) : 0.f;
  }
  for (int i = tid; i < BR * BC; i += 128) {
    const int r = i / BC, c = i % BC;
    const float v = MODE == 0 ? Bg[c * K + r] /* Wt[k][n] = W[n][k] */ : Bg[r * 128 + c] /* A [K][128] */;
    const float h = split ? tf32_hi(v) : v;
    b_hi[tile_off(r, c, BR)] = h;
    b_lo[tile_off(r, c, BR)] = split ? tf32_hi(v - h) : 0.f;
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    if (MODE == 0) {
      constexpr uint32_t idesc = make_idesc(128, N, 1, 1);
      // MN-major SW128: LBO = stride between 32-element blocks along MN, SBO = stride between 8-row K groups
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t koff = ks * 8 * 128;                                  // 8 rows of 128 B
        const uint64_t ah = make_desc(smem_u32(a_hi) + koff, AR * 128, 1024), al = make_desc(smem_u32(a_lo) + koff, AR * 128, 1024);
        const uint64_t bh = make_desc(smem_u32(b_hi) + koff, BR * 128, 1024), bl = make_desc(smem_u32(b_lo) + koff, BR * 128, 1024);
        mma_tf32(tmem, al, bh, idesc, ks > 0);
        mma_tf32(tmem, ah, bl, idesc, 1);
        mma_tf32(tmem, ah, bh, idesc, 1);
      }
    } else {
      constexpr uint32_t idesc = make_idesc(128, K, 0, 0);                   // D[128 (n)][K]
      // K-major SW128: SBO = stride between 8-row groups (1024 B), LBO unused (1); K advances 32 B inside the
      // swizzled row, then jumps to the next 32-trajectory block
      for (int ks = 0; ks < 128 / 8; ++ks) {
        const uint32_t a_off = (ks >> 2) * (AR * 128) + (ks & 3) * 32;
        const uint32_t b_off = (ks >> 2) * (BR * 128) + (ks & 3) * 32;
        const uint64_t ah = make_desc(smem_u32(a_hi) + a_off, 16, 1024), al = make_desc(smem_u32(a_lo) + a_off, 16, 1024);
        const uint64_t bh = make_desc(smem_u32(b_hi) + b_off, 16, 1024), bl = make_desc(smem_u32(b_lo) + b_off, 16, 1024);
        mma_tf32(tmem, al, bh, idesc, ks > 0);
        mma_tf32(tmem, ah, bl, idesc, 1);
        mma_tf32(tmem, ah, bh, idesc, 1);
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  constexpr int NCOL = MODE == 0 ? N : K;
  for (int c0 = 0; c0 < NCOL; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) Dg[tid * NCOL + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem));
}

template <int MODE, int K, int N>
int run(const char* name, int split) {
  const int a_elems = MODE == 0 ? K * 128 : N * 128, b_elems = MODE == 0 ? N * K : K * 128;
  const int ncol = MODE == 0 ? N : K;
  std::vector<float> A(a_elems), B(b_elems), D(128 * ncol);
  srand(7);
  for (auto& v : A) v = (rand() / float(RAND_MAX)) * 2.f - 1.f;
  for (auto& v : B) v = (rand() / float(RAND_MAX)) * 2.f - 1.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, a_elems * 4)); CK(cudaMalloc(&dB, b_elems * 4)); CK(cudaMalloc(&dD, 128 * ncol * 4));
  CK(cudaMemcpy(dA, A.data(), a_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), b_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * ncol * 4));
  const size_t smem = (size_t(MODE == 0 ? K : 128) * 128 * 2 + size_t(K) * (MODE == 0 ? N : 128) * 2) * 4 + 1024;
  CK(cudaFuncSetAttribute(tc_kernel<MODE, K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_kernel<MODE, K, N><<<1, 128, smem>>>(dA, dB, dD, split);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, 128 * ncol * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < ncol; ++c) {
      double ref = 0;
      if (MODE == 0) { for (int k = 0; k < K; ++k) ref += double(A[k * 128 + r]) * double(B[c * K + k]); }
      else { if (r < N) for (int m = 0; m < 128; ++m) ref += double(A[r * 128 + m]) * double(B[c * 128 + m]); }
      maxerr = fmax(maxerr, fabs(ref - D[r * ncol + c]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("%-34s split=%d  max|err| %.3e  max|ref| %.3e  rel %.3e\n", name, split, maxerr, maxref, maxerr / maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr / maxref < (split ? 2e-6 : 2e-3) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<0, 32, 32>("fwd  K=32  N=32", 0);
  bad += run<0, 32, 32>("fwd  K=32  N=32", 1);
  bad += run<0, 64, 64>("fwd  K=64  N=64", 1);
  bad += run<0, 64, 128>("fwd  K=64  N=128", 1);
  bad += run<0, 128, 32>("fwd  K=128 N=32", 1);
  bad += run<0, 8, 64>("fwd  K=8   N=64", 1);
  bad += run<1, 32, 32>("dW   n=32  k=32 (K-major)", 1);
  bad += run<1, 64, 60>("dW   n=60  k=64 (K-major)", 1);
  bad += run<1, 64, 128>("dW   n=128 k=64 (K-major)", 1);
  printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
  return bad;
}
