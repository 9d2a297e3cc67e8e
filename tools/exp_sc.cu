// exp_sc.cu -- standalone experiment harness for the fused binary Shan-Chen collide kernel (not part of
// the product).  Hypothesis: the psi stencil reads 9 rows per field and node row; with
// 128 x 1 blocks those rows are shared between blocks on different SMs (L1 hit rate 25-37 %).  Tile-shaped
// blocks keep every population access a full 128-byte line and cut the gather lines per node row to 3-4.
// Variants differ only in the thread -> node map (TX x TY x TZ tile, TX = 32 keeps rows coalesced) and in
// the register cap; every variant is checked bit for bit against the 128 x 1 x 1 one before it is timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I sailfish_b200/csrc -o tools/bin/exp_sc tools/exp_sc.cu     &&   tools/bin/exp_sc 256
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "slf_shan_chen.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using L = D3Q19;

// body of sc_collide_fused (slf_shan_chen.cuh) for a fluid node at thread offset t
__device__ __forceinline__ void fused_node(const DevParams &P, const ScParams &S, const ScTables<L> &T, unsigned t) {
  float f0[L::Q], f1[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f0[i] = T.ld1[i][t];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f1[i] = T.ld2[i][t];
  const float g0 = T.g0[0][t], g1 = T.g1[0][t];
  float grad0[3] = {0.f, 0.f, 0.f}, grad1[3] = {0.f, 0.f, 0.f};
  if (S.g[0] != 0.0f || S.g[2] != 0.0f) sc_psi_gradient_tab<L, false>(T.g0, t, grad0);
  if (S.g[1] != 0.0f || S.g[3] != 0.0f) sc_psi_gradient_tab<L, false>(T.g1, t, grad1);
  float r0, r1, j0[3], j1[3];
  moments<L>(f0, r0, j0);
  moments<L>(f1, r1, j1);
  const float total = S.tau0_inv * r0 + S.tau1_inv * r1;
  const float v[3] = {(S.tau0_inv * j0[0] + S.tau1_inv * j1[0]) / total, (S.tau0_inv * j0[1] + S.tau1_inv * j1[1]) / total,
                      (S.tau0_inv * j0[2] + S.tau1_inv * j1[2]) / total};
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[0] != 0.0f) sc_add_force(S.potential, g0, S.g[0], grad0, a);
    if (S.g[1] != 0.0f) sc_add_force(S.potential, g0, S.g[1], grad1, a);
    const float ig = 1.0f / g0;
    a[0] *= ig; a[1] *= ig; a[2] *= ig;
    sc_relax<L>(f0, g0, a, v, S.omega[0], S.tau[0]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st1[i][t] = f0[i];
  }
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[2] != 0.0f) sc_add_force(S.potential, g1, S.g[2], grad0, a);
    if (S.g[3] != 0.0f) sc_add_force(S.potential, g1, S.g[3], grad1, a);
    const float ig = 1.0f / g1;
    a[0] *= ig; a[1] *= ig; a[2] *= ig;
    sc_relax<L>(f1, g1, a, v, S.omega[1], S.tau[1]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st2[i][t] = f1[i];
  }
}

// TX x TY x TZ tile of real nodes per block; grid = (ceil(S / TX), ceil(S / TY), ceil(S / TZ))
template <int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__(TX * TY * TZ, MINB)
k_tile(const DevParams P, const ScParams S, const ScTables<L> T, int s) {
  const int lx = threadIdx.x % TX, ly = (threadIdx.x / TX) % TY, lz = threadIdx.x / (TX * TY);
  const int x = blockIdx.x * TX + lx, y = blockIdx.y * TY + ly, z = blockIdx.z * TZ + lz;
  if (x >= s || y >= s || z >= s) return;
  // tables are based at node (1, 1, 1): offset of real node (x, y, z)
  fused_node(P, S, T, (unsigned)(x + P.sy * y + P.sz * z));
}

// smem-staged psi stencil: the block's (TX+2) x (TY+2) x (TZ+2) halo tile of rho and phi is loaded once
// (coalesced rows) and the 2 x 18 neighbour values of every node come from shared memory.
template <int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__(TX * TY * TZ, MINB)
k_smem(const DevParams P, const ScParams S, const ScTables<L> T, const float *__restrict__ rho,
       const float *__restrict__ phi, int s) {
  constexpr int HX = TX + 2, HY = TY + 2, HZ = TZ + 2, NT = TX * TY * TZ;
  __shared__ float sm[2][HZ][HY][HX];
  const int tid = threadIdx.x;
  const int lx = tid % TX, ly = (tid / TX) % TY, lz = tid / (TX * TY);
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = blockIdx.z * TZ;      // real-node coordinates (0-based)
  // node (x, y, z) real -> array index (x + 1, y + 1, z + 1); the halo starts one node before the tile
  for (int idx = tid; idx < HX * HY * HZ; idx += NT) {
    const int hx = idx % HX, hy = (idx / HX) % HY, hz = idx / (HX * HY);
    const int ax = x0 + hx, ay = y0 + hy, az = z0 + hz;                              // array coordinates
    float a = 0.f, b = 0.f;
    if (ax < P.ax && ay < P.ay && az < P.az) {
      const long long gi = ax + P.sy * ay + P.sz * az;
      a = rho[gi]; b = phi[gi];
    }
    sm[0][hz][hy][hx] = a; sm[1][hz][hy][hx] = b;
  }
  __syncthreads();
  const int x = x0 + lx, y = y0 + ly, z = z0 + lz;
  if (x >= s || y >= s || z >= s) return;
  const unsigned t = (unsigned)(x + P.sy * y + P.sz * z);
  float f0[L::Q], f1[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f0[i] = T.ld1[i][t];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f1[i] = T.ld2[i][t];
  const float g0 = sm[0][lz + 1][ly + 1][lx + 1], g1 = sm[1][lz + 1][ly + 1][lx + 1];
  float grad0[3] = {0.f, 0.f, 0.f}, grad1[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; k++) {
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int i = 1; i < L::Q; i++) {
      const float psi = sm[k][lz + 1 + L::ez(i)][ly + 1 + L::ey(i)][lx + 1 + L::ex(i)];
      if (L::ex(i)) fx = fmaf(psi, (float)L::ex(i) * L::w(i), fx);
      if (L::ey(i)) fy = fmaf(psi, (float)L::ey(i) * L::w(i), fy);
      if (L::ez(i)) fz = fmaf(psi, (float)L::ez(i) * L::w(i), fz);
    }
    if (k == 0) { grad0[0] = fx; grad0[1] = fy; grad0[2] = fz; } else { grad1[0] = fx; grad1[1] = fy; grad1[2] = fz; }
  }
  float r0, r1, j0[3], j1[3];
  moments<L>(f0, r0, j0);
  moments<L>(f1, r1, j1);
  const float total = S.tau0_inv * r0 + S.tau1_inv * r1;
  const float v[3] = {(S.tau0_inv * j0[0] + S.tau1_inv * j1[0]) / total, (S.tau0_inv * j0[1] + S.tau1_inv * j1[1]) / total,
                      (S.tau0_inv * j0[2] + S.tau1_inv * j1[2]) / total};
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[0] != 0.0f) sc_add_force(S.potential, g0, S.g[0], grad0, a);
    if (S.g[1] != 0.0f) sc_add_force(S.potential, g0, S.g[1], grad1, a);
    const float ig = 1.0f / g0;
    a[0] *= ig; a[1] *= ig; a[2] *= ig;
    sc_relax<L>(f0, g0, a, v, S.omega[0], S.tau[0]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st1[i][t] = f0[i];
  }
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[2] != 0.0f) sc_add_force(S.potential, g1, S.g[2], grad0, a);
    if (S.g[3] != 0.0f) sc_add_force(S.potential, g1, S.g[3], grad1, a);
    const float ig = 1.0f / g1;
    a[0] *= ig; a[1] *= ig; a[2] *= ig;
    sc_relax<L>(f1, g1, a, v, S.omega[1], S.tau[1]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st2[i][t] = f1[i];
  }
}

struct Ctx {
  DevParams P;
  ScParams S;
  int s;
  float *d1, *d2, *rho, *phi, *init1, *init2, *ref1, *ref2;
  long long fluid;
  size_t lat_bytes;
};

static ScTables<L> tables(const Ctx &c, bool odd) {
  const DevParams &P = c.P;
  const long long base = 1 + P.sy + P.sz;
  ScTables<L> T;
  for (int i = 0; i < L::Q; i++) {
    const long long off = (long long)L::ex(i) + P.sy * L::ey(i) + P.sz * L::ez(i);
    const long long ld = odd ? P.DS * L::opp(i) + base - off : P.DS * i + base;
    const long long st = odd ? P.DS * i + base + off : P.DS * L::opp(i) + base;
    T.ld1[i] = c.d1 + ld; T.ld2[i] = c.d2 + ld;
    T.st1[i] = c.d1 + st; T.st2[i] = c.d2 + st;
    T.g0[i] = c.rho + base + off; T.g1[i] = c.phi + base + off;
  }
  return T;
}

template <class K>
static void run(const char *name, Ctx &c, K launch, bool is_ref) {
  // correctness: one even + one odd step from the pristine lattices
  CK(cudaMemcpy(c.d1, c.init1, c.lat_bytes, cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(c.d2, c.init2, c.lat_bytes, cudaMemcpyDeviceToDevice));
  launch(c, false); launch(c, true);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  if (is_ref) {
    CK(cudaMemcpy(c.ref1, c.d1, c.lat_bytes, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(c.ref2, c.d2, c.lat_bytes, cudaMemcpyDeviceToDevice));
  } else {
    std::vector<float> a(c.lat_bytes / 4), b(c.lat_bytes / 4);
    int bad = 0;
    CK(cudaMemcpy(a.data(), c.d1, c.lat_bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), c.ref1, c.lat_bytes, cudaMemcpyDeviceToHost));
    bad += memcmp(a.data(), b.data(), c.lat_bytes) != 0;
    CK(cudaMemcpy(a.data(), c.d2, c.lat_bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), c.ref2, c.lat_bytes, cudaMemcpyDeviceToHost));
    bad += memcmp(a.data(), b.data(), c.lat_bytes) != 0;
    if (bad) { printf("%-26s MISMATCH against the 128x1x1 variant -- not timed\n", name); return; }
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 6;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) { launch(c, false); launch(c, true); }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 2 * reps;
  printf("%-26s %.3f ms per launch  %.0f GB/s at 312 B/LUP  (%.2f GLUPS for this pass)\n", name, ms,
         312.0 * c.fluid * 1e-9 / ms * 1e3, c.fluid / ms * 1e-6);
  fflush(stdout);
}

#define VARIANT(TX, TY, TZ, MINB, REF)                                                                    \
  run(#TX "x" #TY "x" #TZ " minb" #MINB, c, [](Ctx &c, bool odd) {                                        \
        dim3 grid((c.s + TX - 1) / TX, (c.s + TY - 1) / TY, (c.s + TZ - 1) / TZ);                          \
        k_tile<TX, TY, TZ, MINB><<<grid, TX * TY * TZ>>>(c.P, c.S, tables(c, odd), c.s); }, REF)

#define SVARIANT(TX, TY, TZ, MINB)                                                                        \
  run("smem " #TX "x" #TY "x" #TZ " minb" #MINB, c, [](Ctx &c, bool odd) {                                \
        dim3 grid((c.s + TX - 1) / TX, (c.s + TY - 1) / TY, (c.s + TZ - 1) / TZ);                          \
        k_smem<TX, TY, TZ, MINB><<<grid, TX * TY * TZ>>>(c.P, c.S, tables(c, odd), c.rho, c.phi, c.s); }, false)

int main(int argc, char **argv) {
  Ctx c;
  c.s = argc > 1 ? atoi(argv[1]) : 256;
  DevParams &P = c.P;
  memset(&P, 0, sizeof(P));
  P.nx = P.ny = P.nz = c.s + 2;
  P.ax = ((P.nx + 31) / 32) * 32; P.ay = P.ny; P.az = P.nz;
  P.sy = P.ax; P.sz = (long long)P.ax * P.ay; P.DS = P.sz * P.az;
  P.relax = 1; P.prop = 1;
  memset(&c.S, 0, sizeof(c.S));
  c.S.tau[0] = 1.0f; c.S.tau[1] = 1.0f; c.S.tau0_inv = c.S.tau1_inv = 1.0f; c.S.omega[0] = c.S.omega[1] = 1.0f;
  c.S.g[1] = c.S.g[2] = 1.2f;                      // G12 = G21 = 1.2 (examples/binary_fluid/sc_separation_3d.py)
  c.fluid = (long long)c.s * c.s * c.s;
  c.lat_bytes = (size_t)P.DS * L::Q * 4;
  float **lats[] = {&c.d1, &c.d2, &c.init1, &c.init2, &c.ref1, &c.ref2};
  for (float **p : lats) CK(cudaMalloc(p, c.lat_bytes));
  CK(cudaMalloc(&c.rho, (size_t)P.DS * 4)); CK(cudaMalloc(&c.phi, (size_t)P.DS * 4));
  {
    std::vector<float> h((size_t)P.DS);
    for (int g = 0; g < 2; g++)
      for (int i = 0; i < L::Q; i++) {
        for (long long k = 0; k < P.DS; k++) h[k] = L::w(i) * (1.0f + 1e-3f * (float)((k * 7 + i + 5 * g) % 13));
        CK(cudaMemcpy((g ? c.init2 : c.init1) + (size_t)P.DS * i, h.data(), (size_t)P.DS * 4, cudaMemcpyHostToDevice));
      }
    for (long long k = 0; k < P.DS; k++) h[k] = 1.0f + 1e-3f * (float)(k % 11);
    CK(cudaMemcpy(c.rho, h.data(), (size_t)P.DS * 4, cudaMemcpyHostToDevice));
    for (long long k = 0; k < P.DS; k++) h[k] = 1.0f + 1e-3f * (float)(k % 7);
    CK(cudaMemcpy(c.phi, h.data(), (size_t)P.DS * 4, cudaMemcpyHostToDevice));
  }
  printf("binary Shan-Chen D3Q19 %d^3, 2 x %.2f GB\n", c.s, c.lat_bytes / 1e9);
  VARIANT(128, 1, 1, 5, true);
  VARIANT(128, 1, 1, 6, false);
  VARIANT(32, 4, 1, 5, false);
  VARIANT(32, 2, 2, 5, false);
  VARIANT(32, 4, 1, 6, false);
  VARIANT(32, 2, 2, 6, false);
  VARIANT(32, 8, 1, 2, false);
  VARIANT(32, 4, 2, 2, false);
  VARIANT(64, 2, 1, 5, false);
  VARIANT(64, 2, 2, 2, false);
  SVARIANT(32, 2, 2, 6);
  SVARIANT(32, 4, 2, 3);
  SVARIANT(32, 4, 2, 2);
  SVARIANT(32, 4, 4, 1);
  SVARIANT(32, 8, 2, 1);
  SVARIANT(32, 2, 4, 3);
  SVARIANT(32, 4, 1, 6);
  SVARIANT(64, 2, 2, 3);
  return 0;
}
