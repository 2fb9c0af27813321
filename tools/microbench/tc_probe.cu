// Descriptor-driven tcgen05 probe: the host builds the shared-memory images of the two operands and the
// descriptor parameters, the kernel issues the 3xTF32 MMA sequence and dumps all 128 TMEM lanes.  Used to
// pin down (on the device, there are no public docs in this sandbox) the operand layout the fused rollout
// uses: the no-swizzle "interleaved" tile  X[c/4][r][c%4]  read
//   * K-major  (rows r = M/N index, columns c = K index)     by the forward and data-gradient GEMMs and
//   * MN-major (rows r = K index,  columns c = M/N index)    by the weight-gradient GEMM (contraction over
//     the trajectories) -- the SAME bytes, two descriptors.
// Also measures the issue/commit/wait round trip and the MMA rate for the small layer shapes.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct OpDesc {            // one operand: descriptor fields + how the start address advances per k-step
  uint32_t lbo, sbo;       // bytes
  uint32_t kstep;          // bytes added to the start address per MMA k-step (K = 8 tf32)
  uint32_t layout;         // 0 none, 1 128B_base32B, 2 128B, 4 64B, 6 32B
  uint32_t major;          // 0 K-major, 1 MN-major
};
struct Probe {
  OpDesc a, b;
  uint32_t M, N, nks;
  uint32_t a_floats, b_floats;     // size of one image (hi or lo)
  uint32_t reps;                   // timing repetitions of the whole sequence (1 = functional)
  uint32_t three;                  // 1: 3xTF32 (needs lo images), 0: single pass on the hi image
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout & 7) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_major, uint32_t b_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// images: [a_hi | a_lo | b_hi | b_lo] in global, copied verbatim to shared memory (1024-aligned each)
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ img, float* __restrict__ Dg, const Probe pr, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = (pr.a_floats * 4 + 1023) & ~1023u, b_bytes = (pr.b_floats * 4 + 1023) & ~1023u;
  float* a_hi = reinterpret_cast<float*>(sm);
  float* a_lo = reinterpret_cast<float*>(sm + a_bytes);
  float* b_hi = reinterpret_cast<float*>(sm + 2 * a_bytes);
  float* b_lo = reinterpret_cast<float*>(sm + 2 * a_bytes + b_bytes);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < pr.a_floats; i += 128) { a_hi[i] = img[i]; a_lo[i] = img[pr.a_floats + i]; }
  for (uint32_t i = tid; i < pr.b_floats; i += 128) { b_hi[i] = img[2 * pr.a_floats + i]; b_lo[i] = img[2 * pr.a_floats + pr.b_floats + i]; }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // poison the accumulator columns so that "not written" is visible (0x7fc00000 = NaN)
  {
    for (uint32_t c0 = 0; c0 < 256; c0 += 8) {
      const uint32_t q = 0x7fc00000u;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(tmem + (uint32_t(warp * 32) << 16) + c0), "r"(q));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  const uint32_t idesc = make_idesc(pr.M, pr.N, pr.a.major, pr.b.major);
  uint32_t phase = 0;
  long long t0 = clock64();
  for (uint32_t rep = 0; rep < pr.reps; ++rep) {
    if (tid == 0) {
      if (pr.three) {
        for (uint32_t ks = 0; ks < pr.nks; ++ks) {
          const uint64_t ah = make_desc(smem_u32(a_hi) + ks * pr.a.kstep, pr.a.lbo, pr.a.sbo, pr.a.layout);
          const uint64_t al = make_desc(smem_u32(a_lo) + ks * pr.a.kstep, pr.a.lbo, pr.a.sbo, pr.a.layout);
          const uint64_t bh = make_desc(smem_u32(b_hi) + ks * pr.b.kstep, pr.b.lbo, pr.b.sbo, pr.b.layout);
          const uint64_t bl = make_desc(smem_u32(b_lo) + ks * pr.b.kstep, pr.b.lbo, pr.b.sbo, pr.b.layout);
          mma_tf32(tmem, al, bh, idesc, ks > 0);           // small corrections first: the accumulator is
          mma_tf32(tmem, ah, bl, idesc, 1);                // still small, so its truncation costs nothing
        }
      }
      for (uint32_t ks = 0; ks < pr.nks; ++ks) {
        const uint64_t ah = make_desc(smem_u32(a_hi) + ks * pr.a.kstep, pr.a.lbo, pr.a.sbo, pr.a.layout);
        const uint64_t bh = make_desc(smem_u32(b_hi) + ks * pr.b.kstep, pr.b.lbo, pr.b.sbo, pr.b.layout);
        mma_tf32(tmem, ah, bh, idesc, (pr.three || ks > 0) ? 1u : 0u);
      }
      tc_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0 && cycles) *cycles = t1 - t0;
  for (uint32_t c0 = 0; c0 < pr.N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < pr.N) Dg[tid * pr.N + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem));
}

// ---------------------------------------------------------------- host side
static float tf32_rna(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xFFFFE000u;
  float r; memcpy(&r, &u, 4); return r;
}
// interleaved tile X[c/4][r][c%4], R rows
static inline int il(int r, int c, int R) { return (c >> 2) * (R * 4) + r * 4 + (c & 3); }
// K-major SW128 tile (blocks of 32 columns, 128-byte rows, 16-byte chunk ^ (row & 7))
static inline int sw128(int r, int c, int R) { const int blk = c >> 5, cc = c & 31; return blk * (R * 32) + r * 32 + ((((cc >> 2) ^ (r & 7)) << 2) | (cc & 3)); }

struct Mat { int R, C; std::vector<float> v; float at(int r, int c) const { return v[r * C + c]; } };
static Mat rnd(int R, int C, unsigned seed) {
  Mat m{R, C, std::vector<float>(size_t(R) * C)};
  srand(seed);
  for (auto& x : m.v) x = (rand() / float(RAND_MAX)) * 2.f - 1.f;
  return m;
}
// build hi/lo images of `m` in layout `fn` (image size floats, rows padded to Rp)
template <class F>
static void image(const Mat& m, int Rp, int floats, F fn, std::vector<float>& hi, std::vector<float>& lo) {
  hi.assign(floats, 0.f); lo.assign(floats, 0.f);
  for (int r = 0; r < m.R; ++r)
    for (int c = 0; c < m.C; ++c) {
      const float v = m.at(r, c), h = tf32_rna(v);
      hi[fn(r, c, Rp)] = h; lo[fn(r, c, Rp)] = v - h;
    }
}

struct Result { std::vector<float> D; long long cyc; };
static Result launch(const Probe& pr, const std::vector<float>& ahi, const std::vector<float>& alo, const std::vector<float>& bhi, const std::vector<float>& blo) {
  std::vector<float> img;
  img.insert(img.end(), ahi.begin(), ahi.end()); img.insert(img.end(), alo.begin(), alo.end());
  img.insert(img.end(), bhi.begin(), bhi.end()); img.insert(img.end(), blo.begin(), blo.end());
  float *dimg, *dD; long long* dc;
  CK(cudaMalloc(&dimg, img.size() * 4)); CK(cudaMalloc(&dD, 128 * pr.N * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dimg, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * pr.N * 4));
  const size_t smem = 2 * ((pr.a_floats * 4 + 1023) & ~1023u) + 2 * ((pr.b_floats * 4 + 1023) & ~1023u) + 1024 + 65536;  // slack: M'=128 reads past small tiles
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dimg, dD, pr, dc);
  cudaError_t e = cudaDeviceSynchronize();
  Result r; r.D.resize(128 * pr.N); r.cyc = 0;
  if (e != cudaSuccess) { printf("   kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
  CK(cudaMemcpy(r.D.data(), dD, 128 * pr.N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&r.cyc, dc, 8, cudaMemcpyDeviceToHost));
  cudaFree(dimg); cudaFree(dD); cudaFree(dc);
  return r;
}

// compare D (lane-major [128][N]) with ref[rows][cols]; lane_of(row) gives the TMEM lane of a row
template <class L>
static int check(const char* name, const Result& r, int N, int rows, int cols, const std::vector<double>& ref, L lane_of, double tol) {
  double maxerr = 0, maxref = 0; int nan = 0;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      const float d = r.D[lane_of(i) * N + j];
      if (d != d) { ++nan; continue; }
      maxerr = fmax(maxerr, fabs(ref[i * cols + j] - d)); maxref = fmax(maxref, fabs(ref[i * cols + j]));
    }
  const bool ok = nan == 0 && maxerr / maxref < tol;
  printf("%-58s rel %.3e  nan %d  %s\n", name, maxerr / maxref, nan, ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  // ---------------- forward form  D[m][n] = sum_k A[m][k] W[n][k], both K-major
  for (int variant = 0; variant < 2; ++variant) {          // 0 interleaved/no swizzle, 1 SW128
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int K = cfg == 0 ? 8 : cfg == 1 ? 40 : 64, N = cfg == 0 ? 32 : cfg == 1 ? 32 : 64;
      if (variant == 1 && K % 32) continue;
      Mat A = rnd(128, K, 1), W = rnd(N, K, 2);
      std::vector<float> ah, al, bh, bl;
      Probe pr{}; pr.M = 128; pr.N = N; pr.nks = K / 8; pr.reps = 1; pr.three = 1;
      pr.a_floats = 128 * K; pr.b_floats = N * K;
      if (variant == 0) {
        image(A, 128, pr.a_floats, il, ah, al); image(W, N, pr.b_floats, il, bh, bl);
        pr.a = {128 * 16, 128, 2 * 128 * 16, 0, 0};
        pr.b = {uint32_t(N) * 16, 128, 2 * uint32_t(N) * 16, 0, 0};
      } else {
        image(A, 128, pr.a_floats, sw128, ah, al); image(W, N, pr.b_floats, sw128, bh, bl);
        pr.a = {16, 1024, 32, 2, 0}; pr.b = {16, 1024, 32, 2, 0};      // NB k-steps 4.. cross a 32-column block: only K<=32 valid here
        if (K > 32) continue;
      }
      std::vector<double> ref(128 * N);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += double(A.at(m, k)) * W.at(n, k); ref[m * N + n] = s; }
      char nm[128]; snprintf(nm, 128, "fwd  K-major x K-major  %s K=%d N=%d", variant ? "SW128" : "interleaved", K, N);
      bad += check(nm, launch(pr, ah, al, bh, bl), N, 128, N, ref, [](int i) { return i; }, 2e-6);
    }
  }
  // ---------------- data-gradient form  D[m][k] = sum_n dZ[m][n] W[n][k]:  A-op dZ K-major, B-op W read MN-major
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int Nn = cfg == 0 ? 32 : 64, K = cfg == 0 ? 48 : 64;      // contraction over n (Nn), output width K
    Mat dZ = rnd(128, Nn, 3), W = rnd(Nn, K, 4);
    std::vector<float> ah, al, bh, bl;
    Probe pr{}; pr.M = 128; pr.N = K; pr.nks = Nn / 8; pr.reps = 1; pr.three = 1;
    pr.a_floats = 128 * Nn; pr.b_floats = Nn * K;
    image(dZ, 128, pr.a_floats, il, ah, al); image(W, Nn, pr.b_floats, il, bh, bl);
    pr.a = {128 * 16, 128, 2 * 128 * 16, 0, 0};
    // W tile X[k/4][n][k%4] (rows n): MN-major view: contraction rows n at 16 B, 8-row groups at 128 B (LBO), MN chunks of 4 at Nn*16 B (SBO)
    pr.b = {128, uint32_t(Nn) * 16, 128, 0, 1};
    std::vector<double> ref(128 * K);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) { double s = 0; for (int n = 0; n < Nn; ++n) s += double(dZ.at(m, n)) * W.at(n, k); ref[m * K + k] = s; }
    char nm[128]; snprintf(nm, 128, "dgrad K-major x MN-major interleaved n=%d k=%d", Nn, K);
    bad += check(nm, launch(pr, ah, al, bh, bl), K, 128, K, ref, [](int i) { return i; }, 2e-6);
    // same with LBO/SBO swapped (diagnostic)
    pr.b = {uint32_t(Nn) * 16, 128, 128, 0, 1};
    snprintf(nm, 128, "dgrad  (diagnostic: B LBO/SBO swapped) n=%d k=%d", Nn, K);
    check(nm, launch(pr, ah, al, bh, bl), K, 128, K, ref, [](int i) { return i; }, 2e-6);
  }
  // ---------------- weight-gradient form  D'[n][k] = sum_m dZ[m][n] A[m][k]: both operands read MN-major
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int Nn = cfg == 0 ? 32 : cfg == 1 ? 64 : 128, K = cfg == 0 ? 48 : cfg == 1 ? 64 : 16;
    Mat dZ = rnd(128, Nn, 5), A = rnd(128, K, 6);
    std::vector<float> ah, al, bh, bl;
    std::vector<double> ref(Nn * K);
    for (int n = 0; n < Nn; ++n) for (int k = 0; k < K; ++k) { double s = 0; for (int m = 0; m < 128; ++m) s += double(dZ.at(m, n)) * A.at(m, k); ref[n * K + k] = s; }
    for (int Mp = 128; Mp >= 64; Mp -= 64) {
      if (Mp < Nn) continue;
      Probe pr{}; pr.M = Mp; pr.N = K; pr.nks = 16; pr.reps = 1; pr.three = 1;
      pr.a_floats = 128 * Nn; pr.b_floats = 128 * K;
      image(dZ, 128, pr.a_floats, il, ah, al); image(A, 128, pr.b_floats, il, bh, bl);
      pr.a = {128, 128 * 16, 128, 0, 1}; pr.b = {128, 128 * 16, 128, 0, 1};
      Result r = launch(pr, ah, al, bh, bl);
      char nm[128]; snprintf(nm, 128, "wgrad MN-major x MN-major interleaved n=%d k=%d M'=%d", Nn, K, Mp);
      if (Mp == 128) bad += check(nm, r, K, Nn, K, ref, [](int i) { return i; }, 2e-6);
      else {
        // where did the 64 rows land?  match every lane against every reference row
        printf("%s lane map:", nm);
        for (int n = 0; n < Nn; n += 8) {
          int found = -1;
          for (int lane = 0; lane < 128 && found < 0; ++lane) {
            bool ok = true;
            for (int k = 0; k < K && ok; ++k) { const float d = r.D[lane * K + k]; ok = d == d && fabs(d - ref[n * K + k]) < 1e-4; }
            if (ok) found = lane;
          }
          printf(" %d->%d", n, found);
        }
        printf("\n");
      }
    }
  }
  // ---------------- timing: round trip and throughput, forward shapes
  {
    struct T { int K, N; } ts[] = {{8, 32}, {40, 32}, {64, 64}, {136, 128}, {64, 16}};
    for (auto t : ts) {
      Mat A = rnd(128, t.K, 1), W = rnd(t.N, t.K, 2);
      std::vector<float> ah, al, bh, bl;
      Probe pr{}; pr.M = 128; pr.N = t.N; pr.nks = t.K / 8; pr.three = 1;
      pr.a_floats = 128 * t.K; pr.b_floats = t.N * t.K;
      image(A, 128, pr.a_floats, il, ah, al); image(W, t.N, pr.b_floats, il, bh, bl);
      pr.a = {128 * 16, 128, 2 * 128 * 16, 0, 0};
      pr.b = {uint32_t(t.N) * 16, 128, 2 * uint32_t(t.N) * 16, 0, 0};
      pr.reps = 1; long long c1 = launch(pr, ah, al, bh, bl).cyc;
      pr.reps = 201; long long c2 = launch(pr, ah, al, bh, bl).cyc;
      const double per = double(c2 - c1) / 200.0, mmas = 3.0 * pr.nks;
      printf("timing fwd K=%3d N=%3d: %7.1f clk per issue+commit+wait round trip (%d MMAs, %.1f clk/MMA, %.0f MAC/clk)\n", t.K, t.N, per, int(mmas), per / mmas,
             128.0 * t.N * t.K * 3 / per);
    }
    // wgrad shape: 48 MMAs of M'=128 N'=K
    for (int K : {16, 48, 64, 136}) {
      Mat dZ = rnd(128, 64, 5), A = rnd(128, K, 6);
      std::vector<float> ah, al, bh, bl;
      Probe pr{}; pr.M = 128; pr.N = K; pr.nks = 16; pr.three = 1;
      pr.a_floats = 128 * 64; pr.b_floats = 128 * K;
      image(dZ, 128, pr.a_floats, il, ah, al); image(A, 128, pr.b_floats, il, bh, bl);
      pr.a = {128, 128 * 16, 128, 0, 1}; pr.b = {128, 128 * 16, 128, 0, 1};
      pr.reps = 1; long long c1 = launch(pr, ah, al, bh, bl).cyc;
      pr.reps = 201; long long c2 = launch(pr, ah, al, bh, bl).cyc;
      const double per = double(c2 - c1) / 200.0;
      printf("timing wgrad M'=128 N'=%3d (48 MMAs): %7.1f clk per round trip, %.1f clk/MMA\n", K, per, per / 48.0);
    }
  }
  printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
  return 0;
}
