// Round-2 probe: the tcgen05 facts the wide-policy (C5) kernels are built on.  kind::f16, fp32 accumulate, M = 128.
//   T1  A in TENSOR MEMORY as packed fp16 pairs (column c of lane m = elements k = 2c, 2c+1), B in shared memory as a
//       K-major interleaved (no-swizzle) image  W[k/8][n][k%8]:                     D[m][n] = sum_k A[m][k] W[n][k]
//   T2  the SAME image read as an MN-major B operand (contraction over n):          D[m][k] = sum_n A[m][n] W[n][k]
//   T3  A from shared memory (K-major interleaved tile [k/8][m][k%8]) x K-major B
//   T4  scale-input-d:  D = A*B + D * 2^-s  (only run when compiled with -DWITH_SCALE_D)
//   T5  weight-gradient form: A and B both MN-major shared-memory tiles  X[f/8][m][f%8]  (m = contraction):
//                                                                                    D[n][k] = sum_m dZ[m][n] Act[m][k]
//   T6  mixed operand formats: A = bf16 (tensor memory), B = fp16 (shared memory)
//   T7  issue/completion cost per MMA for the shapes the kernels use
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o h16_probe h16_probe.cu && ./h16_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4); d |= uint64_t((lbo >> 4) & 0x3FFF) << 16; d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
// D f32; afmt / bfmt: 0 = f16, 1 = bf16; amaj / bmaj: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(uint32_t afmt, uint32_t bfmt, uint32_t amaj, uint32_t bmaj, uint32_t M, uint32_t N) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (amaj << 15) | (bmaj << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#ifdef WITH_SCALE_D
__device__ __forceinline__ void mma_ts_scale(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {   // D = A*B + D * 2^-11
  asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}
#endif
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
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------- one generic test kernel
// mode 0 (T1): TS, B K-major, N=128, K=64          mode 1 (T2): TS, B MN-major (same image), contraction 128, N=64
// mode 2 (T3): SS, A K-major tile, K=64            mode 3 (T4): scale-d     mode 4 (T5): SS, both MN-major, K=128
// mode 5 (T6): TS, A bf16, B fp16, K=64
struct Args {
  const uint16_t* a;     // A operand source: [128][KA] halves (row-major) -- written to TMEM or a smem tile by the kernel
  const uint16_t* b;     // B image, already in its shared-memory layout (bytes copied verbatim), 32 KB max
  float* d;              // D[128][N]
  int mode;
};
__global__ void __launch_bounds__(128) probe_kernel(Args g) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* smA = sm;                 // 32 KB
  uint8_t* smB = sm + 32768;         // 32 KB
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smB)[i] = reinterpret_cast<const uint32_t*>(g.b)[i];
  const int mode = g.mode;
  const int KA = (mode == 1 || mode == 4) ? 128 : 64;
  if (mode == 2) {            // A tile K-major interleaved: element (m, k) at (k/8)*(128*16) + m*16 + (k%8)*2
    for (int k = 0; k < KA; ++k) *reinterpret_cast<uint16_t*>(smA + (k / 8) * 2048 + tid * 16 + (k % 8) * 2) = g.a[tid * KA + k];
  }
  if (mode == 4) {            // A tile MN-major: source a = dZ[m][n] (m = tid), stored X[n/8][m][n%8]
    for (int n = 0; n < 128; ++n) *reinterpret_cast<uint16_t*>(smA + (n / 8) * 2048 + tid * 16 + (n % 8) * 2) = g.a[tid * 128 + n];
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tb)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tb, lane_base = uint32_t(warp * 32) << 16;
  const uint32_t t_a = tmem + 256, t_d = tmem;
  if (mode == 0 || mode == 1 || mode == 3 || mode == 5) {      // A -> tensor memory, packed pairs (even k in the low half)
    for (int c = 0; c < KA / 2; ++c) {
      const uint32_t v = uint32_t(g.a[tid * KA + 2 * c]) | (uint32_t(g.a[tid * KA + 2 * c + 1]) << 16);
      tmem_st1(t_a + lane_base + c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  if (mode == 3) {                                             // preload D with 2048.0 everywhere
    for (int c = 0; c < 128; ++c) tmem_st1(t_d + lane_base + c, __float_as_uint(2048.f));
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  int N = 128;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (elect_one()) {
      if (mode == 0 || mode == 5) {          // B K-major image [k/8][128][8]: LBO = 128*16 (k chunks), SBO = 128 (8-row groups); K step = 2 chunks
        const uint32_t idesc = make_idesc(mode == 5 ? 1u : 0u, 0, 0, 0, 128, 128);
        for (int ks = 0; ks < 4; ++ks) mma_ts(t_d, t_a + ks * 8, make_desc(smem_u32(smB) + ks * 2 * 2048, 2048, 128), idesc, ks > 0);
      } else if (mode == 1) {                // same image [k/8][n][8], k < 64, n < 128, read MN-major: N' = k (64), contraction n (128)
        N = 64;                              //   MN groups (k/8) stride SBO = 2048, contraction groups (n/8) stride LBO = 128; K step (16 n) = 256 B
        const uint32_t idesc = make_idesc(0, 0, 0, 1, 128, 64);
        for (int ks = 0; ks < 8; ++ks) mma_ts(t_d, t_a + ks * 8, make_desc(smem_u32(smB) + ks * 256, 128, 2048), idesc, ks > 0);
      } else if (mode == 2) {
        const uint32_t idesc = make_idesc(0, 0, 0, 0, 128, 128);
        for (int ks = 0; ks < 4; ++ks) mma_ss(t_d, make_desc(smem_u32(smA) + ks * 2 * 2048, 2048, 128), make_desc(smem_u32(smB) + ks * 2 * 2048, 2048, 128), idesc, ks > 0);
      } else if (mode == 3) {
#ifdef WITH_SCALE_D
        const uint32_t idesc = make_idesc(0, 0, 0, 0, 128, 128);
        mma_ts_scale(t_d, t_a, make_desc(smem_u32(smB), 2048, 128), idesc);
        for (int ks = 1; ks < 4; ++ks) mma_ts(t_d, t_a + ks * 8, make_desc(smem_u32(smB) + ks * 2 * 2048, 2048, 128), idesc, 1);
#endif
      } else if (mode == 4) {                // both tiles X[f/8][m][8]: MN groups stride SBO = 2048, contraction groups (m/8) stride LBO = 128
        const uint32_t idesc = make_idesc(0, 0, 1, 1, 128, 128);
        for (int ks = 0; ks < 8; ++ks) mma_ss(t_d, make_desc(smem_u32(smA) + ks * 256, 128, 2048), make_desc(smem_u32(smB) + ks * 256, 128, 2048), idesc, ks > 0);
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  N = (mode == 1) ? 64 : 128;
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c = 0; c < N; ++c) g.d[tid * N + c] = __uint_as_float(tmem_ld1(t_d + lane_base + c));
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

static float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }
static uint16_t f2h(float f) { __half h = __float2half_rn(f); return __half_raw(h).x; }
static float b2f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); return __nv_bfloat16_raw(h).x; }
static float frand() { return float(rand()) / float(RAND_MAX) * 2.f - 1.f; }

static void run_mode(int mode) {
  const int KA = (mode == 1 || mode == 4) ? 128 : 64;
  const int N = (mode == 1) ? 64 : 128;
  std::vector<uint16_t> a(128 * KA), img(16384, 0);
  std::vector<float> af(128 * KA);
  // logical W[n][k] (n < 128, k < 64) for modes 0,1,2,3,5; Act[m][k] (m,k < 128) for mode 4
  std::vector<float> W(128 * 128, 0.f);
  for (int i = 0; i < 128 * KA; ++i) {
    const float v = frand();
    if (mode == 5) { a[i] = f2b(v); af[i] = b2f(a[i]); } else { a[i] = f2h(v); af[i] = h2f(a[i]); }
  }
  if (mode == 4) {
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 128; ++k) {
      const uint16_t h = f2h(frand()); W[m * 128 + k] = h2f(h);
      img[((k / 8) * 2048 + m * 16 + (k % 8) * 2) / 2] = h;
    }
  } else {
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 64; ++k) {
      const uint16_t h = f2h(frand()); W[n * 64 + k] = h2f(h);
      img[((k / 8) * 2048 + n * 16 + (k % 8) * 2) / 2] = h;
    }
  }
  std::vector<double> ref(128 * N, 0.0);
  for (int m = 0; m < 128; ++m) for (int c = 0; c < N; ++c) {
    double s = 0;
    if (mode == 1) { for (int n = 0; n < 128; ++n) s += double(af[m * 128 + n]) * W[n * 64 + c]; }
    else if (mode == 4) { for (int t = 0; t < 128; ++t) s += double(af[t * 128 + m]) * W[t * 128 + c]; }   // D[n=m][k=c] = sum_t dZ[t][n] Act[t][k]
    else { for (int k = 0; k < 64; ++k) s += double(af[m * 64 + k]) * W[c * 64 + k]; if (mode == 3) s += 1.0; }
    ref[m * N + c] = s;
  }
  uint16_t *da, *db; float* dd;
  CK(cudaMalloc(&da, a.size() * 2)); CK(cudaMalloc(&db, 32768)); CK(cudaMalloc(&dd, 128 * 128 * 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, img.data(), 32768, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, 128 * 128 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
  Args g{da, db, dd, mode};
  probe_kernel<<<1, 128, 65536 + 1024>>>(g);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); exit(1); }
  std::vector<float> d(128 * N);
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < 128 * N; ++i) { maxerr = fmax(maxerr, fabs(d[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  const char* names[] = {"T1 TS  A=tmem fp16 pairs, B K-major", "T2 TS  same image as MN-major B (contract n)", "T3 SS  A K-major tile", "T4 TS  scale-input-d 2^-11 (D0 = 2048)",
                         "T5 SS  A,B MN-major tiles (wgrad form)", "T6 TS  A=bf16 tmem, B=fp16 smem"};
  printf("%-48s max|err| %.3e  (max|ref| %.3f)  d[0..3] = %.5f %.5f %.5f %.5f  ref %.5f %.5f %.5f %.5f  %s\n", names[mode], maxerr, maxref, d[0], d[1], d[2], d[3],
         ref[0], ref[1], ref[2], ref[3], maxerr < 1e-4 * (maxref + 1) ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(dd);
}

// ---------------------------------------------------------------------------------------------- T7: rates
// form 0: TS B K-major; 1: TS B MN-major; 2: SS both MN-major; 3: SS K-major
template <int FORM, int N, int NMMA>
__global__ void __launch_bounds__(128) rate_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
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
  uint32_t phase = 0;
  long long best = 1ll << 60;
  for (int rep = 0; rep < 6; ++rep) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
      const uint32_t abase = smem_u32(sm), bbase = smem_u32(sm) + 65536;
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          const int ks = i & 7;
          if constexpr (FORM == 0) mma_ts(tmem, tmem + 256 + ks * 8, make_desc(bbase + ks * 2 * N * 16, N * 16, 128), make_idesc(0, 0, 0, 0, 128, N), i > 0);
          else if constexpr (FORM == 1) mma_ts(tmem, tmem + 256 + ks * 8, make_desc(bbase + ks * 256, 128, 2048), make_idesc(0, 0, 0, 1, 128, N), i > 0);
          else if constexpr (FORM == 2) mma_ss(tmem, make_desc(abase + ks * 256, 128, 2048), make_desc(bbase + ks * 256, 128, 2048), make_idesc(0, 0, 1, 1, 128, N), i > 0);
          else mma_ss(tmem, make_desc(abase + (ks & 3) * 2 * 2048, 2048, 128), make_desc(bbase + (ks & 3) * 2 * N * 16, N * 16, 128), make_idesc(0, 0, 0, 0, 128, N), i > 0);
        }
        tc_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
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
template <int FORM, int N>
void rate() {
  long long* d; CK(cudaMalloc(&d, 16));
  long long h8 = 0, h72 = 0;
  const int smem = 131072 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<FORM, N, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate_kernel<FORM, N, 72>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rate_kernel<FORM, N, 8><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h8, d, 8, cudaMemcpyDeviceToHost));
  rate_kernel<FORM, N, 72><<<1, 128, smem>>>(d); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h72, d, 8, cudaMemcpyDeviceToHost));
  const char* forms[] = {"TS B K-major ", "TS B MN-major", "SS MN x MN   ", "SS K x K     "};
  printf("T7 kind::f16 %s N=%3d: 8 MMAs %5lld clk, 72 MMAs %5lld clk => %.1f clk per MMA (K = 16)\n", forms[FORM], N, h8, h72, double(h72 - h8) / 64.0);
  cudaFree(d);
}

int main() {
  srand(1);
  run_mode(0); run_mode(1); run_mode(2);
#ifdef WITH_SCALE_D
  run_mode(3);
#endif
  run_mode(4); run_mode(5);
  rate<0, 128>(); rate<0, 64>(); rate<0, 32>(); rate<0, 16>();
  rate<1, 128>(); rate<1, 32>(); rate<1, 16>();
  rate<2, 128>(); rate<2, 32>(); rate<2, 16>();
  rate<3, 128>(); rate<3, 16>();
  return 0;
}
