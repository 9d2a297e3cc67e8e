// slf_device.cuh -- device-side building blocks of the B200 lattice-Boltzmann kernels.
//
// One thread updates one lattice node.  Populations live in registers (all loops over the
// Q slots are fully unrolled over constexpr lattice tables, so every slot index is a
// compile-time constant); HBM traffic is one coalesced read and one coalesced write of
// every population per node and step (AA pattern: the same addresses, in place).
//
// Reference semantics being reproduced (file:line of sailfish/templates):
//   getDist                          geo_helpers.mako:225-276
//   getMacro                         boundary.mako:465-507
//   precollisionBoundaryConditions   boundary.mako:784-878
//   BGK_relaxate0 / MS_relaxate      relaxation.mako:99-181 / relaxation_mrt.mako:31-97
//   Guo forcing                      relaxation_common.mako:56-64,110-149; sym_force.py:121-160
//   half-way bounce-back             boundary.mako:653-683
//   save_macro_fields                kernel_common.mako:213-240
//   propagate                        propagation.mako:384-421, bounds :36-41
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slf_b200.h"

#define SLF_HD __host__ __device__ __forceinline__
#include "lattice_tables.cuh"

// Access modes of one CollideAndPropagate launch.
//   MODE_AB      : read natural slots at the node, write natural slots at the neighbours (other array)
//   MODE_AA_EVEN : read natural slots at the node, write opposite slots at the node (in place)
//   MODE_AA_ODD  : read opposite slots at the upstream neighbours, write natural slots downstream
enum { MODE_AB = 0, MODE_AA_EVEN = 1, MODE_AA_ODD = 2 };

struct DevParams {
  int nx, ny, nz;            // ghost-inclusive lattice
  int ax, ay, az;            // padded array
  long long DS;              // dist slot stride = ax*ay*az
  long long sy, sz;          // strides of y and z in nodes
  uint32_t tmask, pmask;
  int pshift, oshift;
  uint32_t tid[SLF_NT_COUNT];
  float omega;               // (float)(1/tau)            relaxation.mako:133
  float tau;
  float inv_tau_mrt;         // 1/(0.5 + 3 visc) in fp32  relaxation_mrt.mako:44-46
  float guo_pref;            // 3 (1 - 1/(2 tau))
  float a[3];                // body acceleration
  int forced, relax, prop, use_tags;
  unsigned int unused_tags;  // link-tag bits without missing-link code (geo_encoder.py:283-289)
  const float *node_params;
};

template <class L>
struct Node {
  float f[L::Q];
};

// ---- moments ---------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ void moments(const float (&f)[L::Q], float &rho, float (&j)[3]) {
  float r = 0.f, jx = 0.f, jy = 0.f, jz = 0.f;
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    r += f[i];
    if (L::ex(i) == 1) jx += f[i]; else if (L::ex(i) == -1) jx -= f[i];
    if (L::ey(i) == 1) jy += f[i]; else if (L::ey(i) == -1) jy -= f[i];
    if (L::D == 3) { if (L::ez(i) == 1) jz += f[i]; else if (L::ez(i) == -1) jz -= f[i]; }
  }
  rho = r; j[0] = jx; j[1] = jy; j[2] = jz;
}

template <class L>
__device__ __forceinline__ float edotv(int i, const float (&v)[3]) {
  float s = 0.f;
  if (L::ex(i) == 1) s += v[0]; else if (L::ex(i) == -1) s -= v[0];
  if (L::ey(i) == 1) s += v[1]; else if (L::ey(i) == -1) s -= v[1];
  if (L::D == 3) { if (L::ez(i) == 1) s += v[2]; else if (L::ez(i) == -1) s -= v[2]; }
  return s;
}

// f^eq_i = w_i (rho + rho (3 e.v + 4.5 (e.v)^2 - 1.5 v.v))   sym_equilibrium.py:90-120
template <class L>
__device__ __forceinline__ void equilibrium(float rho, const float (&v)[3], float (&feq)[L::Q]) {
  const float vv = v[0] * v[0] + v[1] * v[1] + (L::D == 3 ? v[2] * v[2] : 0.f);
  const float base = fmaf(-1.5f, vv, 1.0f);
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    const float ev = edotv<L>(i, v);
    const float h = fmaf(ev, fmaf(4.5f, ev, 3.0f), base);
    feq[i] = L::w(i) * rho * h;
  }
}

// ---- collision operators ---------------------------------------------------------------
// FORCED: 0 / 1 fix the presence of a body force at compile time, -1 reads P.forced.
template <class L, int FORCED = -1>
__device__ __forceinline__ void collide_bgk(const DevParams &P, float rho, float (&v)[3], float (&f)[L::Q]) {
  const bool forced = FORCED < 0 ? (P.forced != 0) : (FORCED != 0);
  float v0[3] = {v[0], v[1], v[2]};
  if (forced) {
#pragma unroll
    for (int d = 0; d < L::D; d++) v0[d] = fmaf(0.5f, P.a[d], v[d]);
  }
  const float vv = v0[0] * v0[0] + v0[1] * v0[1] + (L::D == 3 ? v0[2] * v0[2] : 0.f);
  const float base = fmaf(-1.5f, vv, 1.0f);
  const float omega = P.omega;
  if (!forced) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      const float ev = edotv<L>(i, v0);
      const float feq = L::w(i) * rho * fmaf(ev, fmaf(4.5f, ev, 3.0f), base);
      f[i] = fmaf(omega, feq - f[i], f[i]);
    }
  } else {
    const float pref = rho * P.guo_pref;
    const float av = P.a[0] * v0[0] + P.a[1] * v0[1] + (L::D == 3 ? P.a[2] * v0[2] : 0.f);
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      const float ev = edotv<L>(i, v0);
      const float feq = L::w(i) * rho * fmaf(ev, fmaf(4.5f, ev, 3.0f), base);
      const float ea = edotv<L>(i, P.a);
      // (e_i - v + 3 (e_i.v) e_i) . a
      const float s = ea - av + 3.0f * ev * ea;
      f[i] = fmaf(omega, feq - f[i], f[i]) + pref * L::w(i) * s;
    }
#pragma unroll
    for (int d = 0; d < L::D; d++) v[d] = fmaf(0.5f, P.a[d], v[d]);
  }
}

// MS_relaxate (relaxation_mrt.mako:31-97): m = M f; non-conserved moments relax towards their
// equilibria with rate s_k; f = M^-1 m.  The rows of M are mutually orthogonal (Gram-Schmidt,
// sym.py:918-942, :1044), hence M^-1 = M^T diag(1/|row|^2).  Evaluated here in the
// algebraically identical "delta" form  f += M^-1 [ -s_k (m_k - m_k^eq) ]  : only the small
// change of each moment goes through the back-transform, so its round-off scales with the
// change rather than with the populations (closer to the exact result than transforming the
// full moment vector; the parity tests measure both against the fp64 oracle).
template <class L, bool EQ_NODE, int FORCED = -1>
__device__ __forceinline__ void collide_mrt(const DevParams &P, float (&v)[3], float (&f)[L::Q]) {
  if constexpr (L::HAS_MRT) {
    const bool forced = FORCED < 0 ? (P.forced != 0) : (FORCED != 0);
    float m[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; k++) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < L::Q; i++) {
        const int c = L::m(k, i);
        if (c == 1) s += f[i];
        else if (c == -1) s -= f[i];
        else if (c != 0) s = fmaf((float)c, f[i], s);
      }
      m[k] = s;
    }
    constexpr int IZ = L::I_MZ < 0 ? 0 : L::I_MZ;
    if (forced) {                                 // fluid_momentum, relaxation_mrt.mako:11-28
      m[L::I_MX] = fmaf(0.5f, P.a[0], m[L::I_MX]);
      m[L::I_MY] = fmaf(0.5f, P.a[1], m[L::I_MY]);
      if (L::D == 3) m[IZ] = fmaf(0.5f, P.a[2], m[IZ]);
    }
    const float rho = m[0], mx = m[L::I_MX], my = m[L::I_MY];
    const float mz = L::D == 3 ? m[IZ] : 0.f;
    float meq[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; k++) meq[k] = 0.f;
    if (L::ID == 0) {            // D2Q9, sym.py:122-149
      const float m2 = mx * mx + my * my;
      meq[1] = fmaf(3.0f, m2, -2.0f * rho);
      meq[2] = fmaf(-3.0f, m2, rho);
      meq[4] = -mx; meq[6] = -my;
      meq[7] = mx * mx - my * my; meq[8] = mx * my;
    } else {                     // D3Q19, sym.py:384-400
      const float ir = 1.0f / rho, m2 = mx * mx + my * my + mz * mz;
      meq[1] = fmaf(19.0f * ir, m2, -11.0f * rho);
      meq[2] = (float)(-475.0 / 63.0) * ir * m2;
      meq[4] = (float)(-2.0 / 3.0) * mx; meq[6] = (float)(-2.0 / 3.0) * my; meq[8] = (float)(-2.0 / 3.0) * mz;
      meq[9] = ir * (2.0f * mx * mx - (my * my + mz * mz));
      meq[11] = ir * (my * my - mz * mz);
      meq[13] = ir * mx * my; meq[14] = ir * my * mz; meq[15] = ir * mx * mz;
    }
    // change of each moment, pre-divided by |row|^2
    float dm[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; k++) {
      if (L::rate(k) != 0.0f) {
        const float s = L::rate(k) < 0.f ? P.inv_tau_mrt : L::rate(k);
        // equilibrium nodes first overwrite the moment with its equilibrium (:59-77), i.e. s = 1
        dm[k] = (EQ_NODE ? 1.0f : s) * (meq[k] - m[k]) * L::mnorm_inv(k);
      } else {
        dm[k] = 0.f;
      }
    }
    if (forced) {
      // both half-steps of fluid_momentum add a/2 to the (conserved) momentum moments
      dm[L::I_MX] = P.a[0] * L::mnorm_inv(L::I_MX);
      dm[L::I_MY] = P.a[1] * L::mnorm_inv(L::I_MY);
      if (L::D == 3) dm[IZ] = P.a[2] * L::mnorm_inv(IZ);
#pragma unroll
      for (int d = 0; d < L::D; d++) v[d] = fmaf(0.5f, P.a[d], v[d]);   // fluid_output_velocity
    }
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < L::Q; k++) {
        const int c = L::m(k, i);
        if (L::rate(k) == 0.0f && !(k == L::I_MX || k == L::I_MY || k == IZ)) continue;
        if (c == 1) s += dm[k];
        else if (c == -1) s -= dm[k];
        else if (c != 0) s = fmaf((float)c, dm[k], s);
      }
      f[i] += s;
    }
  }
}

// ---- addressing ------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ long long noff(const DevParams &P, int i) {
  return (long long)L::ex(i) + (long long)L::ey(i) * P.sy + (long long)L::ez(i) * P.sz;
}

// HINT selects the cache operator of the population loads/stores of the bulk kernel:
// 0 default (ld.global / st.global), 1 streaming (.cs: evict-first in L1/L2 -- every value is
// used exactly once per step), 2 L2-only loads (.cg) + write-through stores (.wt).
template <int HINT>
__device__ __forceinline__ float ldp(const float *p) {
  if (HINT == 1) return __ldcs(p);
  if (HINT == 2) return __ldcg(p);
  return *p;
}
template <int HINT>
__device__ __forceinline__ void stp(float *p, float v) {
  if (HINT == 1) __stcs(p, v);
  else if (HINT == 2) __stwt(p, v);
  else *p = v;
}

template <class L, int MODE, int HINT = 0>
__device__ __forceinline__ void load_node(const DevParams &P, const float *din, long long gi,
                                          float (&f)[L::Q]) {
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    if (MODE == MODE_AA_ODD) f[i] = ldp<HINT>(din + (gi - noff<L>(P, i) + P.DS * L::opp(i)));
    else f[i] = ldp<HINT>(din + (gi + P.DS * i));
  }
}
