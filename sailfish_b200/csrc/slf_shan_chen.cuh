// slf_shan_chen.cuh -- binary Shan-Chen mixture (BASELINE config 5).
//
// Reference: templates/models/binary_shan_chen.mako:19-141 (ShanChenPrepareMacroFields,
// ShanChenCollideAndPropagate{0,1}), templates/shan_chen.mako:9-84 (force stencil),
// templates/models/lb_binary_fluid.mako:88-127 (SetInitialConditions),
// relaxation.mako:99-181 + relaxation_common.mako:56-149 (BGK_relaxate{0,1} with Guo forcing),
// kernel_utils.mako:410-474 (ApplyMacroPeriodicBoundaryConditions).
//
// Same three-pass structure as the reference (the density of the next step depends on the
// neighbours' post-collision populations, so the macro pass cannot be folded into the collide
// pass): per lattice update 2 x Q x 4 B are read by the macro pass and 2 x (2 Q x 4 B) moved by
// the two collide passes, plus the 4 + 4 + 4 D bytes of rho, phi, v written once and read twice
// (the 18-point psi stencil hits L1/L2): 516 B/LUP for D3Q19.
#pragma once
#include "slf_kernels.cuh"

struct ScParams {
  float tau0_inv, tau1_inv;   // 1.0f / tau, 1.0f / tau_phi   (lb_binary_fluid.mako:36-45)
  float omega[2];             // (float)(1 / tau_g)
  float tau[2];
  float g[4];                 // G11, G12, G21, G22
  int potential;
  float body[2][3];           // constant body acceleration of each component (add_body_force(grid=k))
};

__device__ __forceinline__ float sc_ppot(int pot, float x) {
  return pot == SLF_SC_CLASSIC ? 1.0f - expf(-x) : x;       // sym.py:896-908
}

template <class L>
__device__ __forceinline__ bool sc_is_wet(const DevParams &P, uint32_t type) {
  return type == P.tid[SLF_NT_FLUID];
}

// sum_i w_i e_i psi(field(x + e_i))   (shan_chen.mako:45-84)
template <class L>
__device__ __forceinline__ void sc_psi_gradient(const DevParams &P, int pot, const float *__restrict__ field,
                                                long long gi, float (&grad)[3]) {
  float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const float psi = sc_ppot(pot, field[gi + noff<L>(P, i)]);
    if (L::ex(i)) fx = fmaf(psi, (float)L::ex(i) * L::w(i), fx);
    if (L::ey(i)) fy = fmaf(psi, (float)L::ey(i) * L::w(i), fy);
    if (L::D == 3 && L::ez(i)) fz = fmaf(psi, (float)L::ez(i) * L::w(i), fz);
  }
  grad[0] = fx; grad[1] = fy; grad[2] = fz;
}

// F += -G psi(own density) grad
__device__ __forceinline__ void sc_add_force(int pot, float gm, float G, const float (&grad)[3], float (&a)[3]) {
  const float c = -sc_ppot(pot, gm) * G;
  a[0] = fmaf(grad[0], c, a[0]); a[1] = fmaf(grad[1], c, a[1]); a[2] = fmaf(grad[2], c, a[2]);
}

// BGK relaxation of one component towards the equilibrium at the force-shifted common velocity,
// with the Guo source term (relaxation.mako:99-181, relaxation_common.mako:56-149)
template <class L>
__device__ __forceinline__ void sc_relax(float (&f)[L::Q], float gm, const float (&a)[3], const float (&v)[3], float omega,
                                         float tau) {
  // Opposite directions share everything that is even in e_i: the pair (i, opp(i)) is relaxed together.
  //   f_i^eq = w gm (base + 4.5 ev^2 +- 3 ev),   Guo term = pref w (3 ev ea - av +- ea)
  const float v0[3] = {fmaf(0.5f, a[0], v[0]), fmaf(0.5f, a[1], v[1]), L::D == 3 ? fmaf(0.5f, a[2], v[2]) : 0.f};
  const float vv = v0[0] * v0[0] + v0[1] * v0[1] + (L::D == 3 ? v0[2] * v0[2] : 0.f);
  const float base = fmaf(-1.5f, vv, 1.0f);
  const float pref = gm * 3.0f * (1.0f - 1.0f / (2.0f * tau));
  const float av = a[0] * v0[0] + a[1] * v0[1] + (L::D == 3 ? a[2] * v0[2] : 0.f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    const int o = L::opp(i);
    if (o < i) continue;
    const float wg = L::w(i) * gm, pw = pref * L::w(i);
    if (o == i) {                                   // rest population
      f[i] = fmaf(omega, wg * base - f[i], f[i]) - pw * av;
      continue;
    }
    const float ev = edotv<L>(i, v0), ea = edotv<L>(i, a);
    const float ev3 = 3.0f * ev;
    const float sym = wg * fmaf(1.5f * ev3, ev, base), asym = wg * ev3;
    const float fs = pw * fmaf(ev3, ea, -av), fa = pw * ea;
    f[i] = fmaf(omega, (sym + asym) - f[i], f[i]) + (fs + fa);
    f[o] = fmaf(omega, (sym - asym) - f[o], f[o]) + (fs - fa);
  }
}

template <class L, int MODE>
__device__ __forceinline__ void sc_store(const DevParams &P, float *dout, long long gi, const float (&f)[L::Q]) {
  if (!P.prop) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) dout[gi + P.DS * i] = f[i];
  } else if (MODE == MODE_AA_EVEN) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) dout[gi + P.DS * L::opp(i)] = f[i];
  } else {
#pragma unroll
    for (int i = 0; i < L::Q; i++) dout[gi + noff<L>(P, i) + P.DS * i] = f[i];
  }
}

template <class L>
__device__ __forceinline__ void sc_wall(const DevParams &P, float *dout, long long gi, const float (&f)[L::Q]) {
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (!P.prop) dout[gi + P.DS * i] = f[L::opp(i)];
    else dout[gi + noff<L>(P, i) + P.DS * i] = f[L::opp(i)];   // AB; walls next to the array edge are rejected at bind time
  }
}

// ShanChenPrepareMacroFields
template <class L, int MODE, bool WITH_V>
__global__ void __launch_bounds__(128)
sc_prepare_macro(const DevParams P, const ScParams S, const uint32_t *__restrict__ map, const float *d1,
                 const float *d2, float *__restrict__ orho, float *__restrict__ ophi,
                 float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, long long gi0,
                 long long gi1) {
  const long long gi = gi0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= gi1) return;
  if ((map[gi] & P.tmask) != P.tid[SLF_NT_FLUID]) return;      // wet nodes only
  float f[L::Q], rho0, rho1, j0[3], j1[3];
  load_node<L, MODE>(P, d1, gi, f);
  moments<L>(f, rho0, j0);
  load_node<L, MODE>(P, d2, gi, f);
  moments<L>(f, rho1, j1);
  orho[gi] = rho0; ophi[gi] = rho1;
  if (!WITH_V) return;     // the fused collide kernel recomputes the common velocity from the populations
  const float total = S.tau0_inv * rho0 + S.tau1_inv * rho1;
  ovx[gi] = (S.tau0_inv * j0[0] + S.tau1_inv * j1[0]) / total;
  ovy[gi] = (S.tau0_inv * j0[1] + S.tau1_inv * j1[1]) / total;
  if (L::D == 3) ovz[gi] = (S.tau0_inv * j0[2] + S.tau1_inv * j1[2]) / total;
}

// ShanChenCollideAndPropagate<GRID>
template <class L, int MODE, int GRID>
__global__ void __launch_bounds__(128)
sc_collide(const DevParams P, const ScParams S, const uint32_t *__restrict__ map, const float *din, float *dout,
           const float *__restrict__ rho_f, const float *__restrict__ phi_f, const float *__restrict__ ovx,
           const float *__restrict__ ovy, const float *__restrict__ ovz, long long gi0, long long gi1) {
  const long long gi = gi0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= gi1) return;
  const uint32_t type = map[gi] & P.tmask;
  const bool fluid = type == P.tid[SLF_NT_FLUID];
  const bool wall = type == P.tid[SLF_NT_FULLBB_WALL];
  if (!fluid && !wall) return;
  // full-way bounce-back is a no-op in the AA pattern (see collide_bulk_aa)
  if (wall && MODE != MODE_AB && P.prop) return;

  float f[L::Q];
  load_node<L, MODE>(P, din, gi, f);
  if (wall) { sc_wall<L>(P, dout, gi, f); return; }
  const float *own = GRID == 0 ? rho_f : phi_f;
  const float gm = own[gi];
  float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const float G = S.g[2 * GRID + k];
    if (G == 0.0f) continue;
    float grad[3];
    sc_psi_gradient<L>(P, S.potential, k == 0 ? rho_f : phi_f, gi, grad);
    sc_add_force(S.potential, gm, G, grad, a);
  }
  const float ig = 1.0f / gm;
  a[0] = fmaf(a[0], ig, S.body[GRID][0]); a[1] = fmaf(a[1], ig, S.body[GRID][1]); a[2] = fmaf(a[2], ig, S.body[GRID][2]);
  const float v[3] = {ovx[gi], ovy[gi], L::D == 3 ? ovz[gi] : 0.f};
  if (P.relax) sc_relax<L>(f, gm, a, v, S.omega[GRID], S.tau[GRID]);
  sc_store<L, MODE>(P, dout, gi, f);
}

// Both components in one pass (native runner): the common velocity is recomputed from the two sets of
// populations the thread holds anyway, each psi-gradient is evaluated once instead of twice, and the
// v fields are neither written by the macro pass nor read here: 2 Q x 4 B in, 2 Q x 4 B out, + rho, phi
// = 312 B/LUP (D3Q19) against 2 x 172 for the two separate launches.  Same arithmetic, same order.
// Measured alternatives that lost (profiles/r01_sweep_sc_fused.txt): re-reading component 2 instead of
// holding it (64 registers, 8 blocks/SM: 8.9 GLUPS vs 9.2), and one half-block per component with the
// moments / gradients exchanged through shared memory (48-64 registers: 7.0-7.6 GLUPS).
// All addresses come from per-slot base pointers in the kernel parameters (see collide_bulk_aa_tab):
// the access pattern (AB / AA even / AA odd / propagation off) lives in the tables, not in the code.
#ifndef SLF_MINB_SC_FUSED
#define SLF_MINB_SC_FUSED 5
#endif
template <class L> struct ScTables {
  const float *ld1[L::Q], *ld2[L::Q];   // f_i of component 1 / 2 is read from ld[i][t]
  float *st1[L::Q], *st2[L::Q];         // ... and written to st[i][t]
  const float *g0[L::Q], *g1[L::Q];     // rho / phi at x + e_i
};

template <class L, bool CLASSIC>
__device__ __forceinline__ void sc_psi_gradient_tab(const float *const (&g)[L::Q], unsigned t, float (&grad)[3]) {
  float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const float x = g[i][t];
    const float psi = CLASSIC ? 1.0f - expf(-x) : x;
    if (L::ex(i)) fx = fmaf(psi, (float)L::ex(i) * L::w(i), fx);
    if (L::ey(i)) fy = fmaf(psi, (float)L::ey(i) * L::w(i), fy);
    if (L::D == 3 && L::ez(i)) fz = fmaf(psi, (float)L::ez(i) * L::w(i), fz);
  }
  grad[0] = fx; grad[1] = fy; grad[2] = fz;
}

template <class L, bool SAVE>
__device__ __forceinline__ void
sc_fused_node(const DevParams &P, const ScParams &S, const ScTables<L> &T, const uint32_t *__restrict__ map,
              float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, const long long gi,
              const unsigned t, const int bounce_walls) {
  const uint32_t type = map[gi] & P.tmask;
  const bool fluid = type == P.tid[SLF_NT_FLUID];
  const bool wall = type == P.tid[SLF_NT_FULLBB_WALL];
  if (!fluid && !(wall && bounce_walls)) return;    // full-way bounce-back is a no-op in the AA pattern

  float f0[L::Q], f1[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f0[i] = T.ld1[i][t];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f1[i] = T.ld2[i][t];
  if (wall) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) { T.st1[i][t] = f0[L::opp(i)]; T.st2[i][t] = f1[L::opp(i)]; }
    return;
  }
  const float g0 = T.g0[0][t], g1 = T.g1[0][t];
  float grad0[3] = {0.f, 0.f, 0.f}, grad1[3] = {0.f, 0.f, 0.f};
  if (S.potential == SLF_SC_CLASSIC) {
    if (S.g[0] != 0.0f || S.g[2] != 0.0f) sc_psi_gradient_tab<L, true>(T.g0, t, grad0);
    if (S.g[1] != 0.0f || S.g[3] != 0.0f) sc_psi_gradient_tab<L, true>(T.g1, t, grad1);
  } else {
    if (S.g[0] != 0.0f || S.g[2] != 0.0f) sc_psi_gradient_tab<L, false>(T.g0, t, grad0);
    if (S.g[1] != 0.0f || S.g[3] != 0.0f) sc_psi_gradient_tab<L, false>(T.g1, t, grad1);
  }

  float r0, r1, j0[3], j1[3];
  moments<L>(f0, r0, j0);
  moments<L>(f1, r1, j1);
  const float total = S.tau0_inv * r0 + S.tau1_inv * r1;
  const float v[3] = {(S.tau0_inv * j0[0] + S.tau1_inv * j1[0]) / total, (S.tau0_inv * j0[1] + S.tau1_inv * j1[1]) / total,
                      L::D == 3 ? (S.tau0_inv * j0[2] + S.tau1_inv * j1[2]) / total : 0.f};
  if (SAVE) {
    ovx[gi] = v[0]; ovy[gi] = v[1];
    if (L::D == 3) ovz[gi] = v[2];
  }
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[0] != 0.0f) sc_add_force(S.potential, g0, S.g[0], grad0, a);
    if (S.g[1] != 0.0f) sc_add_force(S.potential, g0, S.g[1], grad1, a);
    const float ig = 1.0f / g0;
    a[0] = fmaf(a[0], ig, S.body[0][0]); a[1] = fmaf(a[1], ig, S.body[0][1]); a[2] = fmaf(a[2], ig, S.body[0][2]);
    if (P.relax) sc_relax<L>(f0, g0, a, v, S.omega[0], S.tau[0]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st1[i][t] = f0[i];
  }
  {
    float a[3] = {0.f, 0.f, 0.f};
    if (S.g[2] != 0.0f) sc_add_force(S.potential, g1, S.g[2], grad0, a);
    if (S.g[3] != 0.0f) sc_add_force(S.potential, g1, S.g[3], grad1, a);
    const float ig = 1.0f / g1;
    a[0] = fmaf(a[0], ig, S.body[1][0]); a[1] = fmaf(a[1], ig, S.body[1][1]); a[2] = fmaf(a[2], ig, S.body[1][2]);
    if (P.relax) sc_relax<L>(f1, g1, a, v, S.omega[1], S.tau[1]);
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st2[i][t] = f1[i];
  }
}

template <class L, bool SAVE>
__global__ void __launch_bounds__(128, SLF_MINB_SC_FUSED)
sc_collide_fused(const DevParams P, const ScParams S, const ScTables<L> T, const uint32_t *__restrict__ map,
                 float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, long long gi0,
                 unsigned n, int bounce_walls) {
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= n) return;
  sc_fused_node<L, SAVE>(P, S, T, map, ovx, ovy, ovz, gi0 + t, t, bounce_walls);
}

// The same node body on tile-shaped blocks: 32 x 2 x 2 nodes (3-D) keep every population access a full line and
// cut the rows the two psi stencils touch per block from 9 + 9 per node row to 16 + 16 per four node rows.
// Measured on B200 (tools/exp_sc.cu, profiles/r02_exp_sc.txt, 256^3): 1.093 ms per launch against 1.207 ms for
// 128 x 1 x 1 blocks at 5 blocks/SM and 1.146 ms at 6; staging the stencil in shared memory on top of that LOSES
// (1.216 ms: the gathers already hit L1, what limits the kernel is the latency of its 38 population loads at
// 35 % occupancy) and so do larger tiles (fewer blocks per SM).
#ifndef SLF_MINB_SC_TILE
#define SLF_MINB_SC_TILE 6
#endif
template <class L, bool SAVE>
__global__ void __launch_bounds__(128, SLF_MINB_SC_TILE)
sc_collide_fused_tile(const DevParams P, const ScParams S, const ScTables<L> T, const uint32_t *__restrict__ map,
                      float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz, long long gi0,
                      int z0, int z1, int bounce_walls) {
  const int lx = threadIdx.x & 31, ly = (threadIdx.x >> 5) & 1, lz = threadIdx.x >> 6;
  const int x = blockIdx.x * 32 + lx, y = blockIdx.y * 2 + ly, z = z0 + blockIdx.z * 2 + lz;
  if (x >= P.ax || y >= P.ay || z >= z1) return;
  const long long gi = (long long)x + P.sy * y + P.sz * z;
  sc_fused_node<L, SAVE>(P, S, T, map, ovx, ovy, ovz, gi, (unsigned)(gi - gi0), bounce_walls);
}

// mode: MODE_AB / MODE_AA_EVEN / MODE_AA_ODD as seen by load_node / sc_store
template <class L>
static inline ScTables<L> make_sc_tables(const DevParams &P, int mode, const float *d1in, float *d1out,
                                         const float *d2in, float *d2out, const float *rho, const float *phi,
                                         long long gi0) {
  ScTables<L> T;
  for (int i = 0; i < L::Q; i++) {
    const long long off = (long long)L::ex(i) + P.sy * L::ey(i) + P.sz * L::ez(i);
    const long long ld = mode == MODE_AA_ODD ? P.DS * L::opp(i) + gi0 - off : P.DS * i + gi0;
    long long st;
    if (!P.prop) st = P.DS * i + gi0;
    else if (mode == MODE_AA_EVEN) st = P.DS * L::opp(i) + gi0;
    else st = P.DS * i + gi0 + off;
    T.ld1[i] = d1in + ld; T.ld2[i] = d2in + ld;
    T.st1[i] = d1out + st; T.st2[i] = d2out + st;
    T.g0[i] = rho + gi0 + off; T.g1[i] = phi + gi0 + off;
  }
  return T;
}

// SetInitialConditions of the binary model
template <class L>
__global__ void __launch_bounds__(256)
sc_set_initial_conditions(const DevParams P, const uint32_t *__restrict__ map, float *__restrict__ d1,
                          float *__restrict__ d2, const float *__restrict__ ivx, const float *__restrict__ ivy,
                          const float *__restrict__ ivz, const float *__restrict__ irho,
                          const float *__restrict__ iphi) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= P.DS) return;
  const int x = (int)(gi % P.ax);
  const long long t = gi / P.ax;
  const int y = (int)(t % P.ay), z = (int)(t / P.ay);
  if (x >= P.nx || y >= P.ny || z >= P.nz) return;
  const uint32_t type = map[gi] & P.tmask;
  const int cls = node_class<L>(P, type);
  const bool wet = cls == SLF_NT_FLUID || cls == SLF_NT_HALFBB_WALL || cls == SLF_NT_EQUILIBRIUM_VELOCITY ||
                   cls == SLF_NT_EQUILIBRIUM_DENSITY || cls == SLF_NT_REGULARIZED_VELOCITY ||
                   cls == SLF_NT_REGULARIZED_DENSITY || cls == SLF_NT_DO_NOTHING || cls == SLF_NT_ZOUHE_VELOCITY ||
                   cls == SLF_NT_ZOUHE_DENSITY || cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW ||
                   cls == SLF_NT_WALL_TMS;
  if (!wet && cls != SLF_NT_FULLBB_WALL) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) { d1[gi + P.DS * i] = INFINITY; d2[gi + P.DS * i] = INFINITY; }
    return;
  }
  float v[3] = {ivx[gi], ivy[gi], L::D == 3 ? ivz[gi] : 0.f};
  float feq[L::Q];
  equilibrium<L>(irho[gi], v, feq);
#pragma unroll
  for (int i = 0; i < L::Q; i++) d1[gi + P.DS * i] = feq[i];
  equilibrium<L>(iphi[gi], v, feq);
#pragma unroll
  for (int i = 0; i < L::Q; i++) d2[gi + P.DS * i] = feq[i];
}

// ApplyMacroPeriodicBoundaryConditions: ghost <- opposite real layer, both sides, if finite
__global__ void __launch_bounds__(128)
macro_pbc(const DevParams P, float *field, int axis, int dim, int link_lo, int link_hi) {
  const int n1 = axis == 0 ? P.ny : P.nx;
  const int n2 = dim == 3 ? (axis == 2 ? P.ny : P.nz) : 1;
  const int i1 = blockIdx.x * blockDim.x + threadIdx.x, i2 = blockIdx.y;
  if (i1 >= n1 || i2 >= n2) return;
  int x = 0, y = 0, z = 0;
  if (axis == 0) { y = i1; z = i2; } else if (axis == 1) { x = i1; z = i2; } else { x = i1; y = i2; }
  // ghost layers across a split face are delivered by the neighbour (slf_halo_push_field)
  const int io = dim == 3 ? z : y, no = dim == 3 ? P.nz : P.ny;
  if (axis != dim - 1 && ((link_lo && io == 0) || (link_hi && io == no - 1))) return;
  const long long stride = axis == 0 ? 1 : (axis == 1 ? P.sy : P.sz);
  const int n = axis == 0 ? P.nx : (axis == 1 ? P.ny : P.nz);
  const long long base = (long long)x + P.sy * y + P.sz * z;   // coordinate along `axis` is 0 here
  const float hi = field[base + stride * (n - 2)];
  if (isfinite(hi)) field[base] = hi;
  const float lo = field[base + stride];
  if (isfinite(lo)) field[base + stride * (n - 1)] = lo;
}
