// exp_bulk.cu -- standalone tuning harness for the D3Q19 BGK bulk kernel (not part of the
// product).  Times even / odd AA steps of kernel variants on an all-fluid box with a ghost
// ring, prints ms and algorithmic GB/s (152 B per fluid node).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I sailfish_b200/csrc -o gpurun_out/exp_bulk tools/exp_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "slf_kernels.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using L = D3Q19;

template <int HINT> __device__ __forceinline__ float ldx(const float *p) {
  if (HINT == 1) return __ldcs(p);
  if (HINT == 2) return __ldcg(p);
  if (HINT == 3) { float v; asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
  if (HINT == 4) { float v; asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
  return *p;
}
template <int HINT> __device__ __forceinline__ void stx(float *p, float v) {
  if (HINT == 1) __stcs(p, v);
  else if (HINT == 2) __stwt(p, v);
  else if (HINT == 3) asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v));
  else if (HINT == 4) asm volatile("st.global.L1::evict_first.f32 [%0], %1;" :: "l"(p), "f"(v));
  else *p = v;
}

// variant A: one node per thread (the product kernel's structure), templated on block size / min blocks / hints
template <int MODE, int LH, int SH, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB)
kA(const DevParams P, const uint32_t *__restrict__ fmask, float *dist, long long gi0, long long gi1) {
  const long long gi = gi0 + (long long)blockIdx.x * BS + threadIdx.x;
  if (gi >= gi1) return;
  if (!((fmask[gi >> 5] >> (gi & 31)) & 1u)) return;
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (MODE == MODE_AA_ODD) f[i] = ldx<LH>(dist + (gi - noff<L>(P, i) + P.DS * L::opp(i)));
    else f[i] = ldx<LH>(dist + (gi + P.DS * i));
  }
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  collide_bgk<L>(P, rho, v, f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (MODE == MODE_AA_EVEN) stx<SH>(dist + (gi + P.DS * L::opp(i)), f[i]);
    else stx<SH>(dist + (gi + noff<L>(P, i) + P.DS * i), f[i]);
  }
}

// variant B: two nodes per thread, 32 apart (same warp covers 64 consecutive nodes): 2x loads in flight
template <int MODE, int LH, int SH, int BS, int MINB, int NPT = 2>
__global__ void __launch_bounds__(BS, MINB)
kB(const DevParams P, const uint32_t *__restrict__ fmask, float *dist, long long gi0, long long gi1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = gi0 + ((long long)blockIdx.x * (BS / 32) + warp) * (32 * NPT) + lane;
  float f[NPT][L::Q];
  bool act[NPT];
#pragma unroll
  for (int n = 0; n < NPT; n++) {
    const long long gi = base + 32 * n;
    act[n] = gi < gi1 && ((fmask[gi >> 5] >> (gi & 31)) & 1u);
    if (act[n]) {
#pragma unroll
      for (int i = 0; i < L::Q; i++) {
        if (MODE == MODE_AA_ODD) f[n][i] = ldx<LH>(dist + (gi - noff<L>(P, i) + P.DS * L::opp(i)));
        else f[n][i] = ldx<LH>(dist + (gi + P.DS * i));
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NPT; n++) {
    const long long gi = base + 32 * n;
    if (!act[n]) continue;
    float rho, v[3];
    moments<L>(f[n], rho, v);
    const float inv = 1.0f / rho;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
    collide_bgk<L>(P, rho, v, f[n]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (MODE == MODE_AA_EVEN) stx<SH>(dist + (gi + P.DS * L::opp(i)), f[n][i]);
      else stx<SH>(dist + (gi + noff<L>(P, i) + P.DS * i), f[n][i]);
    }
  }
}

// variant C: persistent grid-stride (grid = k * SMs), one node per thread per iteration
template <int MODE, int LH, int SH, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB)
kC(const DevParams P, const uint32_t *__restrict__ fmask, float *dist, long long gi0, long long gi1) {
  for (long long gi = gi0 + (long long)blockIdx.x * BS + threadIdx.x; gi < gi1; gi += (long long)gridDim.x * BS) {
    if (!((fmask[gi >> 5] >> (gi & 31)) & 1u)) continue;
    float f[L::Q];
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (MODE == MODE_AA_ODD) f[i] = ldx<LH>(dist + (gi - noff<L>(P, i) + P.DS * L::opp(i)));
      else f[i] = ldx<LH>(dist + (gi + P.DS * i));
    }
    float rho, v[3];
    moments<L>(f, rho, v);
    const float inv = 1.0f / rho;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
    collide_bgk<L>(P, rho, v, f);
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (MODE == MODE_AA_EVEN) stx<SH>(dist + (gi + P.DS * L::opp(i)), f[i]);
      else stx<SH>(dist + (gi + noff<L>(P, i) + P.DS * i), f[i]);
    }
  }
}

struct Ctx {
  DevParams P; uint32_t *fmask, *wmask; float *dist; long long gi0, gi1; long long fluid;
};

template <class K>
static void timeit(const char *name, const Ctx &c, K launch_even, K launch_odd) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms[2] = {0, 0};
  for (int par = 0; par < 2; par++) {
    for (int w = 0; w < 2; w++) { if (par) launch_odd(c); else launch_even(c); }
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) { if (par) launch_odd(c); else launch_even(c); }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    ms[par] = t / reps;
  }
  const double gb = 152.0 * c.fluid * 1e-9;
  printf("%-34s even %.3f ms %.0f GB/s | odd %.3f ms %.0f GB/s | avg %.3f ms %.0f GB/s  %.1f GLUPS\n", name, ms[0],
         gb / ms[0] * 1e3, ms[1], gb / ms[1] * 1e3, 0.5 * (ms[0] + ms[1]), gb / (0.5 * (ms[0] + ms[1])) * 1e3,
         c.fluid / (0.5 * (ms[0] + ms[1])) * 1e-6);
  fflush(stdout);
}

typedef void (*Launch)(const Ctx &);

#define VA(LH, SH, BS, MINB)                                                                             \
  timeit<Launch>("A 1npt lh" #LH " sh" #SH " bs" #BS " minb" #MINB, c,                                   \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; kA<MODE_AA_EVEN, LH, SH, BS, MINB><<<(unsigned)((n + BS - 1) / BS), BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); }, \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; kA<MODE_AA_ODD, LH, SH, BS, MINB><<<(unsigned)((n + BS - 1) / BS), BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); })
#define VB(LH, SH, BS, MINB, NPT)                                                                        \
  timeit<Launch>("B " #NPT "npt lh" #LH " sh" #SH " bs" #BS " minb" #MINB, c,                            \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; kB<MODE_AA_EVEN, LH, SH, BS, MINB, NPT><<<(unsigned)((n + NPT * BS - 1) / (NPT * BS)), BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); }, \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; kB<MODE_AA_ODD, LH, SH, BS, MINB, NPT><<<(unsigned)((n + NPT * BS - 1) / (NPT * BS)), BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); })
#define VP(HINT)                                                                                         \
  timeit<Launch>("P product kernel hint" #HINT, c,                                                       \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; collide_bulk<L, SLF_BGK, MODE_AA_EVEN, false, true, HINT><<<(unsigned)((n + 127) / 128), 128>>>(c.P, nullptr, c.fmask, c.wmask, c.dist, c.dist, nullptr, nullptr, nullptr, nullptr, c.gi0, c.gi1); }, \
      [](const Ctx &c) { long long n = c.gi1 - c.gi0; collide_bulk<L, SLF_BGK, MODE_AA_ODD, false, true, HINT><<<(unsigned)((n + 127) / 128), 128>>>(c.P, nullptr, c.fmask, c.wmask, c.dist, c.dist, nullptr, nullptr, nullptr, nullptr, c.gi0, c.gi1); })
#define VC(LH, SH, BS, MINB, GRIDMUL)                                                                    \
  timeit<Launch>("C persist lh" #LH " sh" #SH " bs" #BS " minb" #MINB " x" #GRIDMUL, c,                  \
      [](const Ctx &c) { kC<MODE_AA_EVEN, LH, SH, BS, MINB><<<148 * GRIDMUL, BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); }, \
      [](const Ctx &c) { kC<MODE_AA_ODD, LH, SH, BS, MINB><<<148 * GRIDMUL, BS>>>(c.P, c.fmask, c.dist, c.gi0, c.gi1); })

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 512;
  Ctx c;
  DevParams &P = c.P;
  memset(&P, 0, sizeof(P));
  P.nx = P.ny = P.nz = S + 2;
  P.ax = ((P.nx + 31) / 32) * 32; P.ay = P.ny; P.az = P.nz;
  P.sy = P.ax; P.sz = (long long)P.ax * P.ay; P.DS = P.sz * P.az;
  P.omega = 1.0f / 0.6f; P.tau = 0.6f; P.relax = 1; P.prop = 1;
  const long long words = (P.DS + 31) / 32;
  std::vector<uint32_t> fm(words, 0u);
  c.fluid = 0;
  for (int z = 1; z <= S; z++) for (int y = 1; y <= S; y++) for (int x = 1; x <= S; x++) {
    long long gi = x + P.sy * y + P.sz * z; fm[gi >> 5] |= 1u << (gi & 31); c.fluid++;
  }
  CK(cudaMalloc(&c.fmask, words * 4));
  CK(cudaMemcpy(c.fmask, fm.data(), words * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c.wmask, words * 4));
  CK(cudaMemset(c.wmask, 0, words * 4));
  CK(cudaMalloc(&c.dist, (size_t)P.DS * 19 * 4));
  {
    std::vector<float> h((size_t)P.DS);
    for (int i = 0; i < 19; i++) {
      for (long long k = 0; k < P.DS; k++) h[k] = L::w(i) * (1.0f + 1e-3f * (float)((k * 7 + i) % 13));
      CK(cudaMemcpy(c.dist + (size_t)P.DS * i, h.data(), (size_t)P.DS * 4, cudaMemcpyHostToDevice));
    }
  }
  c.gi0 = 0; c.gi1 = P.DS;
  printf("size %d^3, fluid %lld nodes, lattice %.2f GB\n", S, c.fluid, P.DS * 76.0 / 1e9);
  VP(0); VP(1);
  VA(0, 0, 128, 10); VA(1, 1, 128, 10); VA(0, 0, 128, 12); VA(1, 1, 128, 12); VA(0, 0, 128, 9); VA(0, 0, 64, 20); VA(0, 0, 256, 5);
  VB(0, 0, 128, 1, 2); VB(0, 0, 128, 4, 2); VB(0, 0, 128, 5, 2); VB(0, 0, 128, 6, 2); VB(0, 1, 128, 1, 2); VB(1, 0, 128, 1, 2);
  VB(0, 0, 64, 1, 2); VB(0, 0, 256, 1, 2);
  VB(0, 0, 128, 1, 3); VB(0, 0, 128, 1, 4); VB(0, 0, 64, 1, 4); VB(0, 0, 128, 3, 4); VB(0, 0, 128, 4, 3);
  return 0;
}
