// Microbenchmark: issue / pipe rates of the instructions of the wide kernels' epilogues on sm_100a -- FFMA vs FFMA2 (packed
// fp32 pair), MUFU.EX2, F2FP (fp32 pair -> half2), HADD2.F32 (half -> fp32) -- alone and mixed, at 4 / 8 / 16 warps per SM.
// Prints warp-instructions per cycle per SM (4 schedulers: 4.0 is the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma_v(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2_v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp_v(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float h2f_v(uint32_t h) { float r; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(r) : "r"(h)); return r; }

// MODE 0: 16 independent FFMA chains   1: 8 independent FFMA2 chains (same flops)   2: FFMA2 + FFMA alternating
//      3: MUFU.EX2 x16   4: F2FP x8 + HADD2.F32 x16   5: 8 FFMA2 + 2 MUFU (the GELU mix)
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float acc[16];
  f2_t p[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = pk2(acc[2 * i], acc[2 * i + 1]); h[i] = 0x3c003c00u + i; }
  const f2_t a2 = pk2(a, a), b2 = pk2(b, b);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = ffma_v(acc[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], a2, b2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], a2, b2); acc[i] = ffma_v(acc[i], a, b); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = ex2_v(acc[i]);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { h[i] = f2fp_v(acc[2 * i], acc[2 * i + 1]); acc[2 * i] = h2f_v(h[i]); acc[2 * i + 1] = h2f_v(h[i] >> 16); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], a2, b2);
      acc[0] = ex2_v(acc[0]); acc[1] = ex2_v(acc[1]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float x, y; up2(p[i], x, y); s += x + y + __uint_as_float(h[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char* name, double inst_per_iter, float* out, int sms, int mhz) {
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<sms, warps * 32>>>(out, 100, 0.999f, 0.001f);
    CK(cudaEventRecord(e0));
    k<MODE><<<sms, warps * 32>>>(out, iters, 0.999f, 0.001f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = ms * 1e-3 * mhz * 1e6;
    printf("%-34s %2d warps/SM: %.3f warp-inst/cycle/SM\n", name, warps, inst_per_iter * iters * warps / cyc);
  }
  return 0;
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  int mhz = 0;
  CK(cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0));
  mhz /= 1000;
  printf("%s, %d SMs, %d MHz (nominal; rates assume it)\n", pr.name, pr.multiProcessorCount, mhz);
  float* out;
  CK(cudaMalloc(&out, size_t(pr.multiProcessorCount) * 1024 * 4));
  run<0>("FFMA (16 chains)", 16, out, pr.multiProcessorCount, mhz);
  run<1>("FFMA2 (8 chains)", 8, out, pr.multiProcessorCount, mhz);
  run<2>("FFMA2 + FFMA 1:1", 16, out, pr.multiProcessorCount, mhz);
  run<3>("MUFU.EX2", 16, out, pr.multiProcessorCount, mhz);
  run<4>("F2FP + 2 HADD2.F32", 24, out, pr.multiProcessorCount, mhz);
  run<5>("8 FFMA2 + 2 MUFU", 10, out, pr.multiProcessorCount, mhz);
  return 0;
}
