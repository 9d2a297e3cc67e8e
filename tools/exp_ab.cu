// exp_ab.cu -- standalone experiment harness for the AB (two-lattice) D3Q19 BGK bulk kernel (not part of
// the product).  Background (profiles/r1d_variant_kernels.txt): in the AB pattern the stores of the 10
// directions with e_x != 0 land one float off the 128-byte line grid in a lattice that was not just read,
// so the two partially written sectors per warp and direction are filled from DRAM: +20 % DRAM reads,
// 5.15 TB/s instead of the 6.3 TB/s of the AA kernel.  Variants:
//   V0  direct: dout[gi + noff(i) + DS * i] = f_i            (the product's collide_bulk structure)
//   V1  address tables (constant-bank base pointers, 32-bit thread offset) -- addressing cost only
//   V2  V1 + x-shift through shared memory: a thread stores what lands on ITS x, so every direction is
//       written as whole aligned lines; only the value crossing the block edge is stored on its own
//       (2 partial sectors per block and direction instead of 2 per warp)
// Every variant is checked bit for bit against V0 before it is timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I sailfish_b200/csrc -o tools/bin/exp_ab tools/exp_ab.cu     &&   tools/bin/exp_ab 256
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "slf_kernels.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using L = D3Q19;
struct Tab { const float *ld[L::Q]; float *st[L::Q]; };

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k0(const DevParams P, const uint32_t *__restrict__ fmask, const float *__restrict__ din, float *__restrict__ dout,
   long long gi0, long long gi1) {
  const long long gi = gi0 + (long long)blockIdx.x * 128 + threadIdx.x;
  if (gi >= gi1) return;
  if (!((fmask[gi >> 5] >> (gi & 31)) & 1u)) return;
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f[i] = din[gi + P.DS * i];
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  collide_bgk<L>(P, rho, v, f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) dout[gi + noff<L>(P, i) + P.DS * i] = f[i];
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k1(const DevParams P, const Tab T, const uint32_t *__restrict__ fmask, long long gi0, unsigned n) {
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= n) return;
  const long long gi = gi0 + t;
  if (!((fmask[gi >> 5] >> (gi & 31)) & 1u)) return;
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f[i] = T.ld[i][t];
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  collide_bgk<L>(P, rho, v, f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) T.st[i][t] = f[i];
}

// V1 with cache operators: LD = 0 default / 1 ld.cs (evict first) / 2 ld.cg (L2 only); ST = 0 default / 1 st.cs / 2 st.wt
template <int MINB, int LD, int ST>
__global__ void __launch_bounds__(128, MINB)
k1h(const DevParams P, const Tab T, const uint32_t *__restrict__ fmask, long long gi0, unsigned n) {
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= n) return;
  const long long gi = gi0 + t;
  if (!((fmask[gi >> 5] >> (gi & 31)) & 1u)) return;
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f[i] = LD == 1 ? __ldcs(T.ld[i] + t) : (LD == 2 ? __ldcg(T.ld[i] + t) : T.ld[i][t]);
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  collide_bgk<L>(P, rho, v, f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (ST == 1) __stcs(T.st[i] + t, f[i]); else if (ST == 2) __stwt(T.st[i] + t, f[i]); else T.st[i][t] = f[i];
  }
}

// number of directions with e_x != 0 and their position among them
__host__ __device__ constexpr int xdir_index(int i) {
  int k = 0;
  for (int j = 0; j < i; j++) k += L::ex(j) != 0;
  return k;
}
constexpr int NXD = xdir_index(L::Q);

// V2.  Block = 128 consecutive nodes starting at a multiple of 128 (rows are padded to 32, so a block may
// span two rows; the row change is invisible here because the x shift is a shift of the linear index).
// st[i] for e_x != 0 is the table WITHOUT the x part of the offset: thread t stores the value computed by
// thread t - e_x to st[i][t] (an aligned, coalesced line); the value that leaves the block (from thread
// 127 for e_x = +1, thread 0 for e_x = -1) is stored by its producer to st[i][t + e_x].
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k2(const DevParams P, const Tab T, const uint32_t *__restrict__ fmask, long long gi0, unsigned n) {
  __shared__ float sh[NXD][128];
  __shared__ uint32_t live[4];
  const int tid = threadIdx.x;
  const unsigned t = blockIdx.x * 128u + tid;
  const long long gi = gi0 + t;
  const bool on = t < n && ((fmask[gi >> 5] >> (gi & 31)) & 1u);
  float f[L::Q];
  if (on) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) f[i] = T.ld[i][t];
    float rho, v[3];
    moments<L>(f, rho, v);
    const float inv = 1.0f / rho;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
    collide_bgk<L>(P, rho, v, f);
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (L::ex(i) == 0) T.st[i][t] = f[i];
      else sh[xdir_index(i)][tid] = f[i];
    }
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, on);
  if ((tid & 31) == 0) live[tid >> 5] = ballot;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (L::ex(i) == 0) continue;
    const int src = tid - L::ex(i);                       // producer of what lands on this thread's x
    if (src >= 0 && src < 128) {
      if ((live[src >> 5] >> (src & 31)) & 1u) T.st[i][t] = sh[xdir_index(i)][src];
    }
    // the value leaving the block
    if (on && ((L::ex(i) == 1 && tid == 127) || (L::ex(i) == -1 && tid == 0))) T.st[i][(long long)t + L::ex(i)] = f[i];
  }
}

struct Ctx {
  DevParams P;
  uint32_t *fmask;
  float *din, *dout, *ref;
  long long gi0, gi1, fluid;
};

static Tab make_tab(const Ctx &c, bool x_in_table) {
  Tab T;
  for (int i = 0; i < L::Q; i++) {
    const long long off = (x_in_table ? L::ex(i) : 0) + c.P.sy * L::ey(i) + c.P.sz * L::ez(i);
    T.ld[i] = c.din + c.P.DS * i + c.gi0;
    T.st[i] = c.dout + c.P.DS * i + c.gi0 + off;
  }
  return T;
}

template <class K>
static void run(const char *name, Ctx &c, K launch, bool is_ref) {
  const size_t bytes = (size_t)c.P.DS * L::Q * 4;
  CK(cudaMemset(c.dout, 0, bytes));
  launch(c);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  if (is_ref) {
    CK(cudaMemcpy(c.ref, c.dout, bytes, cudaMemcpyDeviceToDevice));
  } else {
    std::vector<float> a((size_t)c.P.DS), b((size_t)c.P.DS);
    long long bad = 0;
    for (int i = 0; i < L::Q; i++) {
      CK(cudaMemcpy(a.data(), c.dout + (size_t)c.P.DS * i, (size_t)c.P.DS * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b.data(), c.ref + (size_t)c.P.DS * i, (size_t)c.P.DS * 4, cudaMemcpyDeviceToHost));
      bad += memcmp(a.data(), b.data(), (size_t)c.P.DS * 4) != 0;
    }
    if (bad) { printf("%-28s MISMATCH against V0 in %lld of %d slots -- not timed\n", name, bad, L::Q); return; }
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 10;
  launch(c);
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) launch(c);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double gb = 152.0 * c.fluid * 1e-9;
  printf("%-28s %.3f ms  %.0f GB/s algorithmic  %.1f GLUPS\n", name, ms, gb / ms * 1e3, c.fluid / ms * 1e-6);
  fflush(stdout);
}

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 256;
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
  const size_t bytes = (size_t)P.DS * L::Q * 4;
  CK(cudaMalloc(&c.din, bytes)); CK(cudaMalloc(&c.dout, bytes)); CK(cudaMalloc(&c.ref, bytes));
  {
    std::vector<float> h((size_t)P.DS);
    for (int i = 0; i < L::Q; i++) {
      for (long long k = 0; k < P.DS; k++) h[k] = L::w(i) * (1.0f + 1e-3f * (float)((k * 7 + i) % 13));
      CK(cudaMemcpy(c.din + (size_t)P.DS * i, h.data(), (size_t)P.DS * 4, cudaMemcpyHostToDevice));
    }
  }
  // whole z layers 1..S: block starts stay multiples of 128 because sz is a multiple of 32 -- round down
  c.gi0 = (P.sz / 128) * 128; c.gi1 = P.DS - P.sz;
  const long long n = c.gi1 - c.gi0;
  const unsigned nb = (unsigned)((n + 127) / 128);
  printf("AB D3Q19 BGK %d^3, fluid %lld nodes, 2 x %.2f GB\n", S, c.fluid, bytes / 1e9);
  run("V0 direct minb10", c, [=](Ctx &c) { k0<10><<<nb, 128>>>(c.P, c.fmask, c.din, c.dout, c.gi0, c.gi1); }, true);
  run("V1 tables minb10", c, [=](Ctx &c) { k1<10><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V2 tables+xshift minb10", c, [=](Ctx &c) { k2<10><<<nb, 128>>>(c.P, make_tab(c, false), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 ld.cs st.cs minb10", c, [=](Ctx &c) { k1h<10, 1, 1><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 ld.cg st.default minb10", c, [=](Ctx &c) { k1h<10, 2, 0><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 ld.default st.cs minb10", c, [=](Ctx &c) { k1h<10, 0, 1><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 ld.cs st.default minb10", c, [=](Ctx &c) { k1h<10, 1, 0><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 ld.cg st.wt minb10", c, [=](Ctx &c) { k1h<10, 2, 2><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 tables minb8", c, [=](Ctx &c) { k1<8><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 tables minb12", c, [=](Ctx &c) { k1<12><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V1 tables minb6", c, [=](Ctx &c) { k1<6><<<nb, 128>>>(c.P, make_tab(c, true), c.fmask, c.gi0, (unsigned)n); }, false);
  run("V2 tables+xshift minb8", c, [=](Ctx &c) { k2<8><<<nb, 128>>>(c.P, make_tab(c, false), c.fmask, c.gi0, (unsigned)n); }, false);
  return 0;
}
