// Microbenchmark: CUDA-core FFMA vs legacy mma.sync (tf32 / bf16) issue rates on sm_100a.
// Decides whether the fused rollout's MLP layers go on the tensor pipe (3xTF32) or on FFMA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int NACC>
__global__ void k_mma(float* out, int iters) {
  float d[NACC][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f; }
  a[0] = a[1] = a[2] = a[3] = 0x3f800000u + threadIdx.x; b[0] = b[1] = 0x3f000000u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0) mma_tf32(d[i], a, b); else mma_bf16(d[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory broadcast LDS.128 + FMA mix as in a thread-per-trajectory layer: per k, NJ/4 LDS.128 + NJ*RT FMAs
template <int NJ, int RT>
__global__ void k_layer(float* out, int iters) {
  __shared__ __align__(16) float W[32 * NJ];
  for (int i = threadIdx.x; i < 32 * NJ; i += blockDim.x) W[i] = 1e-3f * (i % 17);
  __syncthreads();
  float acc[RT][NJ];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;
  float in[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) in[r] = threadIdx.x * 1e-3f + r;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
#pragma unroll
      for (int j4 = 0; j4 < NJ / 4; ++j4) {
        float4 w = *reinterpret_cast<const float4*>(&W[k * NJ + j4 * 4]);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          acc[r][j4 * 4 + 0] = fmaf(in[r], w.x, acc[r][j4 * 4 + 0]);
          acc[r][j4 * 4 + 1] = fmaf(in[r], w.y, acc[r][j4 * 4 + 1]);
          acc[r][j4 * 4 + 2] = fmaf(in[r], w.z, acc[r][j4 * 4 + 2]);
          acc[r][j4 * 4 + 3] = fmaf(in[r], w.w, acc[r][j4 * 4 + 3]);
        }
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) in[r] += 1e-6f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += acc[r][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 64 * 1024));
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {         // warps per SM
    int threads = 128, blocks = sms * wps / 4;
    float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf("ffma      warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
  }
  for (int wps : {4, 8, 16, 32}) {
    int threads = 128, blocks = sms * wps / 4;
    float ms = time_ms([&] { k_mma<0, 8><<<blocks, threads>>>(out, iters / 4); });
    double fl = 2.0 * 16 * 8 * 8 * 8 * (iters / 4) * (double)blocks * (threads / 32);
    printf("mma tf32 m16n8k8  warps/SM %2d : %8.3f ms  %8.2f TFLOP/s (dense tf32)\n", wps, ms, fl / ms * 1e-9);
  }
  for (int wps : {4, 8, 16, 32}) {
    int threads = 128, blocks = sms * wps / 4;
    float ms = time_ms([&] { k_mma<1, 8><<<blocks, threads>>>(out, iters / 4); });
    double fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)blocks * (threads / 32);
    printf("mma bf16 m16n8k16 warps/SM %2d : %8.3f ms  %8.2f TFLOP/s (dense bf16)\n", wps, ms, fl / ms * 1e-9);
  }
  for (int wps : {4, 8, 16}) {
    int threads = 128, blocks = sms * wps / 4;
    float ms = time_ms([&] { k_layer<32, 1><<<blocks, threads>>>(out, 400); });
    double fl = 2.0 * 32 * 32 * 1 * 400 * (double)blocks * threads;
    printf("layer32 RT1 (8 LDS.128 : 32 FMA) warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
    ms = time_ms([&] { k_layer<32, 2><<<blocks, threads>>>(out, 400); });
    fl = 2.0 * 32 * 32 * 2 * 400 * (double)blocks * threads;
    printf("layer32 RT2 (8 LDS.128 : 64 FMA) warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
    ms = time_ms([&] { k_layer<16, 4><<<blocks, threads>>>(out, 400); });
    fl = 2.0 * 32 * 16 * 4 * 400 * (double)blocks * threads;
    printf("layer16 RT4 (4 LDS.128 : 64 FMA) warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
