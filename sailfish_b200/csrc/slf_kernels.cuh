// slf_kernels.cuh -- the kernels behind the reference's kernel names.
//
//   collide_bulk     : CollideAndPropagate for plain fluid nodes (the 97+% case).  No
//                      branches on node type beyond "is it fluid"; this is the kernel the
//                      roofline is quoted on.
//   collide_boundary : CollideAndPropagate for every other active node (walls, velocity /
//                      density / do-nothing nodes).  Same arithmetic as the reference's
//                      generic path, orientation handled with run-time normals.
//   The two kernels touch disjoint memory (each node owns the Q addresses it reads and
//   writes in the AA pattern; in AB the target slots of different nodes never coincide),
//   so they may run in either order or concurrently.
#pragma once
#include "slf_device.cuh"

template <class L>
__device__ __forceinline__ int node_class(const DevParams &P, uint32_t type) {
#pragma unroll
  for (int k = 0; k < SLF_NT_COUNT; k++)
    if (P.tid[k] == type) return k;
  return -1;
}

// ----------------------------------------------------------------------------------------
// Bulk kernel.  gi range [gi0, gi1) is a contiguous run of nodes (whole z-slabs).
// ----------------------------------------------------------------------------------------
// MASKED = true: node classes come from two 1-bit-per-node masks built once per node map
// (slf_plan_bind_map): bit set in `fluid` -> plain fluid node; bit set in `wall` -> full-way
// bounce-back wall.  A warp covers 32 consecutive nodes = one mask word, so the 4 B/node
// map read of the reference layout becomes 2 x 4 B per warp.  Walls are handled here too
// (bounce-back is a pure slot permutation) so that wall columns on x faces share their
// sectors with the neighbouring fluid nodes instead of being fetched one sector per value.
template <class L, int MODEL, int MODE, bool SAVE, bool MASKED, int HINT = 0>
__global__ void __launch_bounds__(128)
collide_bulk(const DevParams P, const uint32_t *__restrict__ map, const uint32_t *__restrict__ fluid_mask,
             const uint32_t *__restrict__ wall_mask, const float *din, float *dout,
             float *__restrict__ orho, float *__restrict__ ovx, float *__restrict__ ovy,
             float *__restrict__ ovz, long long gi0, long long gi1) {
  const long long gi = gi0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= gi1) return;
  bool is_wall = false;
  if (MASKED) {
    const uint32_t bit = 1u << (gi & 31);
    const bool is_fluid = (fluid_mask[gi >> 5] & bit) != 0;
    is_wall = (wall_mask[gi >> 5] & bit) != 0;
    if (!is_fluid && !is_wall) return;
  } else {
    const uint32_t code = map[gi];
    if ((code & P.tmask) != P.tid[SLF_NT_FLUID]) return;
  }

  // NTFullBBWall (boundary.mako:791-795: reflect, no relaxation, no macroscopic output).  In
  // the AA pattern a node reads and writes the same Q addresses, and reflecting then storing
  // "opposite slot, same node" (even) or "natural slot, downstream node" (odd) puts every
  // value back where it was read from: a full-way bounce-back node is a no-op.
  if (MASKED && is_wall && MODE != MODE_AB && P.prop) return;

  float f[L::Q];
  load_node<L, MODE, HINT>(P, din, gi, f);

  if (MASKED && is_wall) {
    if (!P.prop) {
#pragma unroll
      for (int i = 0; i < L::Q; i++) dout[gi + P.DS * i] = f[L::opp(i)];
    } else {
      // AB: the mask only admits walls whose every neighbour lies inside the array
#pragma unroll
      for (int i = 0; i < L::Q; i++) dout[gi + noff<L>(P, i) + P.DS * i] = f[L::opp(i)];
    }
    return;
  }

  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;

  if (P.relax) {
    if (MODEL == SLF_MRT) collide_mrt<L, false>(P, v, f);
    else collide_bgk<L>(P, rho, v, f);
  }
  if (SAVE) {
    orho[gi] = rho; ovx[gi] = v[0]; ovy[gi] = v[1];
    if (L::D == 3) ovz[gi] = v[2];
  }
  if (!P.prop) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) stp<HINT>(dout + (gi + P.DS * i), f[i]);
  } else if (MODE == MODE_AA_EVEN) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) stp<HINT>(dout + (gi + P.DS * L::opp(i)), f[i]);
  } else {
    // A fluid node is never on the outermost layer (that layer is always ghost,
    // subdomain.py:952-977), so every target is inside the array: the bounds tests of
    // propagation.mako:36-41 are vacuous here and live in the boundary kernel only.
#pragma unroll
    for (int i = 0; i < L::Q; i++) stp<HINT>(dout + (gi + noff<L>(P, i) + P.DS * i), f[i]);
  }
}

// ----------------------------------------------------------------------------------------
// AA-pattern bulk kernel (the hot kernel; roofline is quoted on this one).
//
// In the AA pattern a node reads and writes the same Q addresses (even step: slot i <-> slot
// opp(i) at the node; odd step: slot opp(i) at x - e_i is both the source of f_i and the
// target of the post-collision f_opp(i)).  The kernel therefore takes ONE lattice pointer, so
// that every address is computed once and shared by the load and the store, and needs only
// the 1-bit fluid mask: full-way bounce-back walls are no-ops in this pattern (see
// collide_bulk) and every other node type goes to the boundary kernel.
// Launch shape (measured on B200, profiles/r01_exp_bulk_variants.txt, r01_sweep_minb_mrt_q27.txt):
// 128-thread blocks, register budget for 10 resident blocks per SM (48 registers; D3Q27: 8 blocks,
// 64 registers), default cache policy: 6.2 TB/s of algorithmic traffic on D3Q19 BGK 512^3 = 94 % of
// the measured copy bandwidth.
#ifndef SLF_MINB_D3Q19_MRT
#define SLF_MINB_D3Q19_MRT 10
#endif
#ifndef SLF_MINB_D3Q27
#define SLF_MINB_D3Q27 8
#endif
template <class L, int MODEL> struct BulkTuning { static constexpr int MINB = 1; };
#ifndef SLF_MINB_D3Q19_BGK
#define SLF_MINB_D3Q19_BGK 10
#endif
template <> struct BulkTuning<D3Q19, SLF_BGK> { static constexpr int MINB = SLF_MINB_D3Q19_BGK; };
template <> struct BulkTuning<D3Q19, SLF_MRT> { static constexpr int MINB = SLF_MINB_D3Q19_MRT; };
template <> struct BulkTuning<D3Q27, SLF_BGK> { static constexpr int MINB = SLF_MINB_D3Q27; };
template <> struct BulkTuning<D2Q9, SLF_BGK> { static constexpr int MINB = 10; };
template <> struct BulkTuning<D2Q9, SLF_MRT> { static constexpr int MINB = 10; };

// The Q addresses come from a table of per-slot base pointers (kernel parameters, i.e. the
// constant bank): address = table[i] + 4 * t with t the 32-bit offset of the thread in the
// launch.  That is one LDC.64 + one IMAD.WIDE.U32 per access instead of the
// 64-bit index arithmetic of `dist[gi +- noff(i) + DS * slot]` (about 60 % of the dynamic
// instructions of the BGK kernel), and it folds the even/odd difference into the table:
//   even: table[i] = dist + DS * i + gi0;   odd: table[i] = dist + DS * opp(i) + gi0 - noff(i);
// f_i is loaded from table[i] and stored to table[opp(i)] in both cases.
template <class L> struct SlotTable { float *p[L::Q]; };

// one plain fluid node: f_i from T.p[i][t], relax, store to T.p[opp(i)][t]
template <class L, int MODEL, bool SAVE, int HINT = 0>
__device__ __forceinline__ void
bulk_aa_node(const DevParams &P, const SlotTable<L> &T, float *__restrict__ orho, float *__restrict__ ovx,
             float *__restrict__ ovy, float *__restrict__ ovz, const long long gi, const unsigned t) {
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f[i] = ldp<HINT>(T.p[i] + t);
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  if (MODEL == SLF_MRT) collide_mrt<L, false>(P, v, f);
  else collide_bgk<L>(P, rho, v, f);
  if (SAVE) {
    orho[gi] = rho; ovx[gi] = v[0]; ovy[gi] = v[1];
    if (L::D == 3) ovz[gi] = v[2];
  }
#pragma unroll
  for (int i = 0; i < L::Q; i++) T.p[L::opp(i)][t] = f[i];
}

template <class L, int MODEL, bool SAVE>
__global__ void __launch_bounds__(128, BulkTuning<L, MODEL>::MINB)
collide_bulk_aa_tab(const DevParams P, const SlotTable<L> T, const uint32_t *__restrict__ fluid_mask,
                    float *__restrict__ orho, float *__restrict__ ovx, float *__restrict__ ovy,
                    float *__restrict__ ovz, long long gi0, unsigned n) {
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= n) return;
  const long long gi = gi0 + t;
  if (!((fluid_mask[gi >> 5] >> (gi & 31)) & 1u)) return;
  bulk_aa_node<L, MODEL, SAVE>(P, T, orho, ovx, ovy, ovz, gi, t);
}

// Sparse geometries (a pipe in a box: BASELINE config 3): the same node body driven by a list of the 32-node
// segments that contain fluid at all (built by slf_plan_bind_map), one warp per segment.  40 % of the threads of the
// dense launch are padding or solid there and leave at once; the list removes them (ncu, profiles/r02_c3_ab_kernels.txt:
// 49 % occupancy, issue-bound).  Segments are aligned mask words, so the accesses stay whole lines.
template <class L, int MODEL, bool SAVE>
__global__ void __launch_bounds__(128, BulkTuning<L, MODEL>::MINB)
collide_bulk_aa_seg(const DevParams P, const SlotTable<L> T, const uint32_t *__restrict__ fluid_mask,
                    const uint32_t *__restrict__ seg, unsigned nseg, float *__restrict__ orho,
                    float *__restrict__ ovx, float *__restrict__ ovy, float *__restrict__ ovz) {
  const unsigned wi = blockIdx.x * 4u + (threadIdx.x >> 5);
  if (wi >= nseg) return;
  const unsigned w = seg[wi];
  const unsigned lane = threadIdx.x & 31u;
  if (!((fluid_mask[w] >> lane) & 1u)) return;
  const unsigned t = w * 32u + lane;               // tables are based at node 0
  bulk_aa_node<L, MODEL, SAVE>(P, T, orho, ovx, ovy, ovz, (long long)t, t);
}

// AB (two-lattice) counterpart: loads from `din` at the node, stores to `dout` at the downstream neighbours, both
// through per-slot base pointers.  Measured (tools/exp_ab.cu on B200, profiles/r02_exp_ab.txt): the tables are worth
// +1 % over 64-bit index arithmetic on this kernel, and a shared-memory x shift that turns every store into whole
// aligned lines changes nothing (33.6 vs 33.4 GLUPS) -- the AB pattern is bound by its 38 concurrent DRAM streams
// (19 read in one array, 19 written in the other), not by partially written sectors.  Full-way bounce-back walls are
// applied here (reflect, then stream), as in collide_bulk.
template <class L> struct SlotTableAB { const float *ld[L::Q]; float *st[L::Q]; };

template <class L, int MODEL, bool SAVE>
__global__ void __launch_bounds__(128, BulkTuning<L, MODEL>::MINB)
collide_bulk_ab_tab(const DevParams P, const SlotTableAB<L> T, const uint32_t *__restrict__ fluid_mask,
                    const uint32_t *__restrict__ wall_mask, float *__restrict__ orho, float *__restrict__ ovx,
                    float *__restrict__ ovy, float *__restrict__ ovz, long long gi0, unsigned n) {
  const unsigned t = blockIdx.x * 128u + threadIdx.x;
  if (t >= n) return;
  const long long gi = gi0 + t;
  const uint32_t bit = 1u << (gi & 31);
  const bool is_fluid = (fluid_mask[gi >> 5] & bit) != 0;
  const bool is_wall = (wall_mask[gi >> 5] & bit) != 0;
  if (!is_fluid && !is_wall) return;
  float f[L::Q];
#pragma unroll
  for (int i = 0; i < L::Q; i++) f[i] = T.ld[i][t];
  if (is_wall) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) T.st[i][t] = f[L::opp(i)];
    return;
  }
  float rho, v[3];
  moments<L>(f, rho, v);
  const float inv = 1.0f / rho;
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
  if (MODEL == SLF_MRT) collide_mrt<L, false>(P, v, f);
  else collide_bgk<L>(P, rho, v, f);
  if (SAVE) {
    orho[gi] = rho; ovx[gi] = v[0]; ovy[gi] = v[1];
    if (L::D == 3) ovz[gi] = v[2];
  }
#pragma unroll
  for (int i = 0; i < L::Q; i++) T.st[i][t] = f[i];
}

template <class L>
static inline SlotTableAB<L> make_ab_table(const DevParams &P, const float *din, float *dout, long long gi0) {
  SlotTableAB<L> T;
  for (int i = 0; i < L::Q; i++) {
    const long long off = (long long)L::ex(i) + P.sy * L::ey(i) + P.sz * L::ez(i);
    T.ld[i] = din + P.DS * i + gi0;
    T.st[i] = dout + P.DS * i + gi0 + off;
  }
  return T;
}

template <class L>
static inline SlotTable<L> make_aa_table(const DevParams &P, float *dist, long long gi0, bool odd) {
  SlotTable<L> T;
  for (int i = 0; i < L::Q; i++) {
    const long long off = (long long)L::ex(i) + P.sy * L::ey(i) + P.sz * L::ez(i);
    T.p[i] = odd ? dist + P.DS * L::opp(i) + gi0 - off : dist + P.DS * i + gi0;
  }
  return T;
}

// ----------------------------------------------------------------------------------------
// Boundary kernel: all active non-fluid nodes.
// ----------------------------------------------------------------------------------------
template <class L>
SLF_HD constexpr int axis_comp(int i, int a) { return a == 0 ? L::ex(i) : (a == 1 ? L::ey(i) : L::ez(i)); }

template <class L>
__device__ __forceinline__ int edotn(int i, int n0, int n1, int n2) {
  return L::ex(i) * n0 + L::ey(i) * n1 + L::ez(i) * n2;
}

// LISTED = true: `list[k]`, k in [gi0, gi1) are the indices of the nodes to process (built by
// slf_plan_bind_map: every active node that the masked bulk kernel does not handle, sorted by
// gi).  LISTED = false: scan the node range [gi0, gi1) and pick the non-fluid nodes.
// zouhe_bb (boundary.mako:347-381): non-equilibrium bounce-back of the unknown populations (sym.py:750-766), then the
// momentum excess rho v - sum e f tangential to the wall is spread evenly over the unknown populations that have a
// component along it (zouhe_fixup, sym.py:770-815).  (n0, n1, n2): inward normal of the node.
template <class L>
__device__ __forceinline__ void zouhe_bb(int n0, int n1, int n2, float rho, const float (&v)[3], float (&f)[L::Q]) {
  float feq[L::Q];
  equilibrium<L>(rho, v, feq);
#pragma unroll
  for (int i = 0; i < L::Q; i++)
    if (edotn<L>(i, n0, n1, n2) > 0) f[i] = f[L::opp(i)] - feq[L::opp(i)] + feq[i];
  float j[3], r;
  moments<L>(f, r, j);
  const int nn[3] = {n0, n1, n2};
  float nv[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    nv[d] = 0.f;
    if (d < L::D && nn[d] == 0) {
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < L::Q; i++) {
        const int sp = edotn<L>(i, n0, n1, n2);
        const int e2 = L::ex(i) * L::ex(i) + L::ey(i) * L::ey(i) + L::ez(i) * L::ez(i);
        const int ed = d == 0 ? L::ex(i) : (d == 1 ? L::ey(i) : L::ez(i));
        if (sp > 0 && e2 > 1 && ed != 0) cnt++;          // unknown, not the normal direction itself
      }
      nv[d] = cnt ? (rho * v[d] - j[d]) / (float)cnt : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    const int e2 = L::ex(i) * L::ex(i) + L::ey(i) * L::ey(i) + L::ez(i) * L::ez(i);
    if (edotn<L>(i, n0, n1, n2) > 0 && e2 > 1) {
      float val = 0.f;
      if (L::ex(i) != 0 && n0 == 0) val += nv[0] * (float)L::ex(i);
      if (L::ey(i) != 0 && n1 == 0) val += nv[1] * (float)L::ey(i);
      if (L::D == 3 && L::ez(i) != 0 && n2 == 0) val += nv[2] * (float)L::ez(i);
      f[i] += val;
    }
  }
}

// One node of the generic path (every node class, run-time normals); `skip_fluid`: the caller scans a node
// range in which plain fluid nodes belong to the bulk kernel.
// WITH_TMS = false compiles the NTWallTMS branches out (the multi-step kernel, whose launch is refused for lattices
// with such nodes, stays within its register budget).
template <class L, int MODEL, int MODE, bool SAVE, int HINT = 0, bool WITH_TMS = true>
__device__ __forceinline__ void
boundary_node(const DevParams &P, const uint32_t *__restrict__ map, const float *din, float *dout,
              float *__restrict__ orho, float *__restrict__ ovx, float *__restrict__ ovy,
              float *__restrict__ ovz, const long long gi, const bool skip_fluid) {
  const uint32_t code = map[gi];
  const uint32_t type = code & P.tmask;
  if (skip_fluid && type == P.tid[SLF_NT_FLUID]) return;
  const int cls = node_class<L>(P, type);
  if (cls < 0 || cls == SLF_NT_GHOST || cls == SLF_NT_UNUSED || cls == SLF_NT_PROPAGATION_ONLY) return;
  const int x = (int)(gi % P.ax);
  const long long t = gi / P.ax;
  const int y = (int)(t % P.ay), z = (int)(t / P.ay);
  if (x >= P.nx) return;
  const int orient = (int)(code >> P.oshift);
  constexpr bool ODD = MODE == MODE_AA_ODD;
  constexpr bool AA = MODE != MODE_AB;

  float f[L::Q];
  load_node<L, MODE, HINT>(P, din, gi, f);

  // inward normal of the node (primary orientation); zero when unknown
  int n0 = 0, n1 = 0, n2 = 0;
  const bool is_tms = WITH_TMS && cls == SLF_NT_WALL_TMS;
  const bool link_bounce = cls == SLF_NT_HALFBB_WALL || is_tms;      // boundary.mako:653-668
  const bool tagged = link_bounce && P.use_tags;
  if (orient > 0 && orient <= 2 * L::D && !tagged) {
#pragma unroll
    for (int d = 1; d <= 2 * L::D; d++)
      if (orient == d) { n0 = L::ex(L::dir2vec(d)); n1 = L::ey(L::dir2vec(d)); n2 = L::ez(L::dir2vec(d)); }
  }
  // fixMissingDistributions (boundary.mako:574-605): copy / extrapolation outflow nodes take the populations that
  // would arrive from outside from dist_in of their inward neighbour(s).  AB only (boundary.mako:626-628; plan
  // creation refuses these classes with the AA pattern).
  if (!AA && (cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW) && (n0 | n1 | n2) != 0) {
    const long long nb = (long long)n0 + (long long)P.ax * n1 + (long long)P.ax * P.ay * n2;
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (edotn<L>(i, n0, n1, n2) > 0) {
        const float a = din[gi + nb + P.DS * i];
        f[i] = cls == SLF_NT_COPY ? a : 2.0f * a - din[gi + 2 * nb + P.DS * i];
      }
    }
  }
  // NTWallTMS (boundary.mako:631-648): target moments from the populations as loaded -- the slots of the links into
  // the wall hold what was bounced back at the end of the previous step -- then those populations are replaced by
  // the equilibrium at the target moments.
  float tg_rho = 0.f, tg_v[3] = {0.f, 0.f, 0.f};
  if (is_tms) {
    moments<L>(f, tg_rho, tg_v);
    tg_v[0] /= tg_rho; tg_v[1] /= tg_rho; tg_v[2] /= tg_rho;
    float feq[L::Q];
    equilibrium<L>(tg_rho, tg_v, feq);
#pragma unroll
    for (int i = 1; i < L::Q; i++) {
      // link i -> (missing, name) as in the store loop below; only `missing` matters here
      if (P.use_tags) {
        if ((orient & (1 << (i - 1))) == 0 && (P.unused_tags & (1u << (i - 1))) == 0) f[L::opp(i)] = feq[L::opp(i)];
      } else if (orient != 0 && edotn<L>(i, n0, n1, n2) > 0) {
        f[i] = feq[i];
      }
    }
  }
  const bool vel_bc = cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_REGULARIZED_VELOCITY ||
                      cls == SLF_NT_ZOUHE_VELOCITY;
  const bool den_bc = cls == SLF_NT_EQUILIBRIUM_DENSITY || cls == SLF_NT_REGULARIZED_DENSITY ||
                      cls == SLF_NT_ZOUHE_DENSITY;
  const bool standard = !(vel_bc || den_bc);

  float rho, v[3];
  if (standard || orient == 0) {
    moments<L>(f, rho, v);
    v[0] /= rho; v[1] /= rho; v[2] /= rho;
  } else {
    const int pidx = (int)((code >> P.pshift) & P.pmask);
#pragma unroll
    for (int i = 0; i < L::Q; i++)            // fill missing with opposite, boundary.mako:418-422
      if (edotn<L>(i, n0, n1, n2) > 0) f[i] = f[L::opp(i)];
    float j[3];
    moments<L>(f, rho, j);
    if (vel_bc) {
      v[0] = P.node_params[pidx]; v[1] = P.node_params[pidx + 1];
      v[2] = L::D == 3 ? P.node_params[pidx + 2] : 0.f;
      const float nv = (float)n0 * v[0] + (float)n1 * v[1] + (float)n2 * v[2];
      rho = rho / (1.0f - nv);                // sym.py:627
    } else {
      const float par_rho = P.node_params[pidx];
      const float s = (rho - par_rho) / par_rho;   // sym.py:666-680
      v[0] = (float)(-n0) * s; v[1] = (float)(-n1) * s; v[2] = (float)(-n2) * s;
      if (cls == SLF_NT_ZOUHE_DENSITY) {          // boundary.mako:487-492: fix up, recompute the moments, impose rho
        zouhe_bb<L>(n0, n1, n2, par_rho, v, f);
        float r2, j2[3];
        moments<L>(f, r2, j2);
        v[0] = j2[0] / r2; v[1] = j2[1] / r2; v[2] = j2[2] / r2;
      }
      rho = par_rho;
    }
  }

  // pre-collision boundary conditions
  if (cls == SLF_NT_FULLBB_WALL) {
#pragma unroll
    for (int i = 0; i < L::Q; i++)
      if (i < L::opp(i)) { const float tmp = f[i]; f[i] = f[L::opp(i)]; f[L::opp(i)] = tmp; }
  } else if (cls == SLF_NT_SLIP) {                // boundary.mako:837-856: specular reflection (sym.py:481-497)
#pragma unroll
    for (int a = 0; a < L::D; a++) {
      const int na = a == 0 ? n0 : (a == 1 ? n1 : n2);
      if (na != 0) {
#pragma unroll
        for (int i = 0; i < L::Q; i++) {
          // populations that point into the fluid change places with their mirror images in the wall plane
          if (axis_comp<L>(i, a) != 0 && axis_comp<L>(i, a) * na > 0) {
            const float tmp = f[i]; f[i] = f[L::mirror(a, i)]; f[L::mirror(a, i)] = tmp;
          }
        }
      }
    }
  } else if (cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_EQUILIBRIUM_DENSITY) {
    equilibrium<L>(rho, v, f);
  } else if (cls == SLF_NT_ZOUHE_VELOCITY) {      // boundary.mako:812-816
    if (orient == 0) {
#pragma unroll
      for (int i = 0; i < L::Q; i++)
        if (i < L::opp(i)) { const float tmp = f[i]; f[i] = f[L::opp(i)]; f[L::opp(i)] = tmp; }
    } else {
      zouhe_bb<L>(n0, n1, n2, rho, v, f);
    }
  } else if (cls == SLF_NT_REGULARIZED_VELOCITY || cls == SLF_NT_REGULARIZED_DENSITY) {
    if (orient == 0) {
#pragma unroll
      for (int i = 0; i < L::Q; i++)
        if (i < L::opp(i)) { const float tmp = f[i]; f[i] = f[L::opp(i)]; f[L::opp(i)] = tmp; }
    } else {
      float feq[L::Q];
      equilibrium<L>(rho, v, feq);
#pragma unroll
      for (int i = 0; i < L::Q; i++)          // non-equilibrium bounce-back, sym.py:750-766
        if (edotn<L>(i, n0, n1, n2) > 0) f[i] = f[L::opp(i)] - feq[L::opp(i)] + feq[i];
      // Pi^neq = sum e e f - (rho v v + rho cs^2 I), boundary.mako:298-304
      float pxx = 0.f, pyy = 0.f, pzz = 0.f, pxy = 0.f, pxz = 0.f, pyz = 0.f;
#pragma unroll
      for (int i = 0; i < L::Q; i++) {
        if (L::ex(i) * L::ex(i)) pxx += f[i];
        if (L::ey(i) * L::ey(i)) pyy += f[i];
        if (L::ez(i) * L::ez(i)) pzz += f[i];
        if (L::ex(i) * L::ey(i) == 1) pxy += f[i]; else if (L::ex(i) * L::ey(i) == -1) pxy -= f[i];
        if (L::ex(i) * L::ez(i) == 1) pxz += f[i]; else if (L::ex(i) * L::ez(i) == -1) pxz -= f[i];
        if (L::ey(i) * L::ez(i) == 1) pyz += f[i]; else if (L::ey(i) * L::ez(i) == -1) pyz -= f[i];
      }
      const float third = (float)(1.0 / 3.0);
      pxx -= rho * (v[0] * v[0] + third); pyy -= rho * (v[1] * v[1] + third);
      pzz -= L::D == 3 ? rho * (v[2] * v[2] + third) : 0.f;
      pxy -= rho * v[0] * v[1]; pxz -= rho * v[0] * v[2]; pyz -= rho * v[1] * v[2];
#pragma unroll
      for (int i = 0; i < L::Q; i++) {        // regularisation, sym.py:882-891
        float s = pxx * ((float)(L::ex(i) * L::ex(i)) - third) + pyy * ((float)(L::ey(i) * L::ey(i)) - third);
        if (L::D == 3) s += pzz * ((float)(L::ez(i) * L::ez(i)) - third);
        s += 2.0f * pxy * (float)(L::ex(i) * L::ey(i));
        if (L::D == 3) s += 2.0f * (pxz * (float)(L::ex(i) * L::ez(i)) + pyz * (float)(L::ey(i) * L::ez(i)));
        f[i] = fmaxf(1e-7f, feq[i] + L::w(i) * 4.5f * s);
      }
    }
  } else if (cls == SLF_NT_DO_NOTHING && AA && orient != 0) {   // boundary.mako:862-877
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      if (edotn<L>(i, n0, n1, n2) > 0) {
        if (ODD) dout[gi + P.DS * i] = f[i];
        else dout[gi + noff<L>(P, L::opp(i)) + P.DS * L::opp(i)] = f[i];
      }
    }
  }

  const bool wet = cls == SLF_NT_FLUID || link_bounce || vel_bc || den_bc || cls == SLF_NT_DO_NOTHING ||
                   cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW;
  if (wet && P.relax) {
    if (MODEL == SLF_MRT) {
      if (cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_EQUILIBRIUM_DENSITY) collide_mrt<L, true>(P, v, f);
      else collide_mrt<L, false>(P, v, f);
    } else {
      collide_bgk<L>(P, rho, v, f);
    }
  }

  if (is_tms) {                               // boundary.mako:696-722: + f^eq(target) - f^eq(instantaneous)
    float feq[L::Q];
    equilibrium<L>(tg_rho, tg_v, feq);
#pragma unroll
    for (int i = 0; i < L::Q; i++) f[i] += feq[i];
    equilibrium<L>(rho, v, feq);
#pragma unroll
    for (int i = 0; i < L::Q; i++) f[i] -= feq[i];
  }

  if (link_bounce) {                          // boundary.mako:653-683
#pragma unroll
    for (int i = 1; i < L::Q; i++) {
      bool hit; int missing, name;
      if (P.use_tags) { hit = (orient & (1 << (i - 1))) == 0 && (P.unused_tags & (1u << (i - 1))) == 0; name = i; missing = L::opp(i); }
      else { hit = orient != 0 && edotn<L>(i, n0, n1, n2) > 0; missing = i; name = L::opp(i); }
      if (hit) {
        // `name`/`missing` are compile-time per branch: resolve through a tiny switch
        float val = 0.f;
#pragma unroll
        for (int k = 0; k < L::Q; k++) if (k == name) val = f[k];
        if (!AA || ODD) dout[gi + P.DS * missing] = val;
        else {
          long long off = 0; int slot = 0;
#pragma unroll
          for (int k = 0; k < L::Q; k++) if (k == missing) { off = noff<L>(P, L::opp(k)); slot = L::opp(k); }
          dout[gi + off + P.DS * slot] = val;
        }
      }
    }
  }

  if (SAVE && wet) {
    orho[gi] = rho; ovx[gi] = v[0]; ovy[gi] = v[1];
    if (L::D == 3) ovz[gi] = v[2];
  }

  if (!P.prop) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) dout[gi + P.DS * i] = f[i];
  } else if (MODE == MODE_AA_EVEN) {
#pragma unroll
    for (int i = 0; i < L::Q; i++) dout[gi + P.DS * L::opp(i)] = f[i];
  } else {
#pragma unroll
    for (int i = 0; i < L::Q; i++) {
      bool ok = true;                          // prop_bnd, propagation.mako:36-41
      if (L::ex(i) > 0) ok = ok && x < P.nx - 1; else if (L::ex(i) < 0) ok = ok && x > 0;
      if (L::ey(i) > 0) ok = ok && y < P.ny - 1; else if (L::ey(i) < 0) ok = ok && y > 0;
      if (L::D == 3) { if (L::ez(i) > 0) ok = ok && z < P.nz - 1; else if (L::ez(i) < 0) ok = ok && z > 0; }
      if (ok) dout[gi + noff<L>(P, i) + P.DS * i] = f[i];
    }
  }
}

template <class L, int MODEL, int MODE, bool SAVE, bool LISTED>
__global__ void __launch_bounds__(128)
collide_boundary(const DevParams P, const uint32_t *__restrict__ map, const uint32_t *__restrict__ list,
                 const float *din, float *dout,
                 float *__restrict__ orho, float *__restrict__ ovx, float *__restrict__ ovy,
                 float *__restrict__ ovz, long long gi0, long long gi1) {
  long long gi = gi0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= gi1) return;
  if (LISTED) gi = list[gi];
  boundary_node<L, MODEL, MODE, SAVE>(P, map, din, dout, orho, ovx, ovy, ovz, gi, !LISTED);
}

// ----------------------------------------------------------------------------------------
// Many steps in ONE launch, for lattices that fit a single co-resident grid (ldc_2d 256^2: 74 k nodes,
// 2.7 MB -- it lives in L2).  At that size a step is a few hundred nanoseconds of memory time and the
// cost of a time step is the launch: two kernels per step, ~6 us even when replayed from a CUDA graph
// (profiles/r01_bench_configs.txt).  Here a persistent cooperative grid keeps ONE node per thread for the
// whole run and the steps are separated by point-to-point synchronisation instead of kernel boundaries
// (a grid-wide barrier -- one serialised atomic per block -- measured 3.7 us per step,
// profiles/r02_bench_configs.txt): a block covers B consecutive nodes, so everything its nodes read or write
// lies within R = sz + sy + 1 nodes of them, i.e. in the few blocks around it; after step s a block
// publishes s + 1 (release store) and before step s + 1 it polls the counters of those blocks only (acquire
// loads, one poller thread per neighbour).  In the AA pattern that is enough in both directions: neighbours
// are never more than one step apart, and a node's Q addresses of step s + 1 are touched in step s only by
// its neighbours.  Population loads bypass L1 (ld.global.cg): other blocks rewrite those lines every step and
// L1 is not coherent.  Boundary nodes (the plan's node list, here as a third 1-bit mask) are handled by the
// thread that covers them.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

template <class L, int MODEL> struct MultiStepTuning { static constexpr int BLOCK = 512; };
template <int MODEL> struct MultiStepTuning<D2Q9, MODEL> { static constexpr int BLOCK = 1024; };

template <class L, int MODEL>
__global__ void __launch_bounds__(MultiStepTuning<L, MODEL>::BLOCK)
collide_multistep(const DevParams P, const SlotTable<L> Te, const SlotTable<L> To,
                  const uint32_t *__restrict__ fluid_mask, const uint32_t *__restrict__ other_mask,
                  const uint32_t *__restrict__ map, float *dist,
                  float *__restrict__ orho, float *__restrict__ ovx, float *__restrict__ ovy,
                  float *__restrict__ ovz, unsigned n, unsigned it0, int nsteps, int save_last,
                  unsigned int *flags, int reach) {
  constexpr unsigned B = MultiStepTuning<L, MODEL>::BLOCK;
  const unsigned t = blockIdx.x * B + threadIdx.x;
  const long long gi = t;
  bool fluid = false, other = false;
  if (t < n) {
    fluid = (fluid_mask[gi >> 5] >> (gi & 31)) & 1u;
    other = (other_mask[gi >> 5] >> (gi & 31)) & 1u;
  }
  // the blocks this one exchanges data with: blockIdx.x - reach .. + reach
  const int dep = (int)blockIdx.x - reach + (int)threadIdx.x + ((int)threadIdx.x >= reach ? 1 : 0);
  const bool poller = threadIdx.x < 2u * (unsigned)reach && dep >= 0 && dep < (int)gridDim.x;
  for (int s = 0; s < nsteps; s++) {
    if (s > 0) {
      if (poller) while (ld_acquire_gpu_u32(flags + dep) < (unsigned)s) { }
      __syncthreads();
    }
    const bool odd = ((it0 + (unsigned)s) & 1u) != 0;
    const bool save = save_last && s == nsteps - 1;
    if (fluid) {
      if (save) {
        if (odd) bulk_aa_node<L, MODEL, true, 2>(P, To, orho, ovx, ovy, ovz, gi, t);
        else bulk_aa_node<L, MODEL, true, 2>(P, Te, orho, ovx, ovy, ovz, gi, t);
      } else {
        if (odd) bulk_aa_node<L, MODEL, false, 2>(P, To, orho, ovx, ovy, ovz, gi, t);
        else bulk_aa_node<L, MODEL, false, 2>(P, Te, orho, ovx, ovy, ovz, gi, t);
      }
    } else if (other) {
      if (save) {
        if (odd) boundary_node<L, MODEL, MODE_AA_ODD, true, 2, false>(P, map, dist, dist, orho, ovx, ovy, ovz, gi, false);
        else boundary_node<L, MODEL, MODE_AA_EVEN, true, 2, false>(P, map, dist, dist, orho, ovx, ovy, ovz, gi, false);
      } else {
        if (odd) boundary_node<L, MODEL, MODE_AA_ODD, false, 2, false>(P, map, dist, dist, orho, ovx, ovy, ovz, gi, false);
        else boundary_node<L, MODEL, MODE_AA_EVEN, false, 2, false>(P, map, dist, dist, orho, ovx, ovy, ovz, gi, false);
      }
    }
    if (s + 1 < nsteps) {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        st_release_gpu_u32(flags + blockIdx.x, (unsigned)(s + 1));
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// SetInitialConditions (lb_single_fluid.mako:101-127): every node with gx < lat_nx.
// ----------------------------------------------------------------------------------------
template <class L>
__global__ void __launch_bounds__(256)
set_initial_conditions(const DevParams P, float *__restrict__ dist, const float *__restrict__ ivx,
                       const float *__restrict__ ivy, const float *__restrict__ ivz,
                       const float *__restrict__ irho) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= P.DS) return;
  const int x = (int)(gi % P.ax);
  const long long t = gi / P.ax;
  const int y = (int)(t % P.ay), z = (int)(t / P.ay);
  if (x >= P.nx || y >= P.ny || z >= P.nz) return;
  float v[3] = {ivx[gi], ivy[gi], L::D == 3 ? ivz[gi] : 0.f};
  float feq[L::Q];
  equilibrium<L>(irho[gi], v, feq);
#pragma unroll
  for (int i = 0; i < L::Q; i++) dist[gi + P.DS * i] = feq[i];
}

// PrepareMacroFields (lb_single_fluid.mako:129-159): density only.
template <class L, int MODE>
__global__ void __launch_bounds__(256)
prepare_macro_fields(const DevParams P, const uint32_t *__restrict__ map, const float *__restrict__ din,
                     float *__restrict__ orho) {
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= P.DS) return;
  const int cls = node_class<L>(P, map[gi] & P.tmask);
  if (cls < 0 || cls == SLF_NT_GHOST || cls == SLF_NT_UNUSED || cls == SLF_NT_PROPAGATION_ONLY) return;
  if ((int)(gi % P.ax) >= P.nx) return;
  float f[L::Q];
  load_node<L, MODE>(P, din, gi, f);
  float rho = 0.f;
#pragma unroll
  for (int i = 0; i < L::Q; i++) rho += f[i];
  orho[gi] = rho;
}

// ----------------------------------------------------------------------------------------
// ApplyPeriodicBoundaryConditions[WithSwap] (kernel_utils.mako:19-384).  One thread per
// face position; the per-population conditions the template prints at render time are
// evaluated here from (basis, periodicity, lattice size).
// ----------------------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ long long gidx3(const DevParams &P, int c0, int c1, int c2) {
  return (long long)c0 + P.sy * c1 + P.sz * c2;
}

template <class L, bool SWAP>
__device__ __forceinline__ bool pbc_main_cond(int axis, int i, const int (&per)[3], const int (&oth)[2],
                                              const int (&idx)[2], const int (&maxd)[2], int link_lo, int link_hi) {
  const int e[3] = {L::ex(i), L::ey(i), L::ez(i)};
  if (SWAP) {
    bool ok = idx[0] >= 1 && idx[0] <= maxd[0];
    if (L::D == 3) ok = ok && idx[1] >= 1 && idx[1] <= maxd[1];
    return ok;
  }
  if (L::D == 2) {
    // The reference's 2-D rule stops at the last real row.  With a neighbour across a y face (plane push, axis == 0
    // only) the ghost row on that side holds this subdomain's outgoing populations; the ones that also crossed the
    // periodic x face are wrapped inside it before the row leaves, as the 3-D rule does for z-splits.
    const int c = e[1 - axis];
    const bool split_other = axis == 0;
    if (c == 1) return idx[0] > 1 && (idx[0] <= maxd[0] || (split_other && link_hi && idx[0] == maxd[0] + 1));
    if (c == -1) return idx[0] < maxd[0] && (idx[0] >= 1 || (split_other && link_lo && idx[0] == 0));
    return true;
  }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int c = e[oth[k]];
    const bool both = per[axis] && per[oth[k]];
    if (c == 1) { ok = ok && idx[k] > 1; if (both) ok = ok && idx[k] <= maxd[k]; }
    else if (c == -1) { ok = ok && idx[k] < maxd[k]; if (both) ok = ok && idx[k] >= 1; }
  }
  return ok;
}

template <class L, bool SWAP>
__device__ __forceinline__ bool pbc_corner(int axis, int i, const int (&per)[3], const int (&lim)[3],
                                           const int (&oth)[2], const int (&idx)[2], int axis_target,
                                           int (&tgt)[3]) {
  const int e[3] = {L::ex(i), L::ey(i), L::ez(i)};
  constexpr int nother = L::D - 1;
  {
    int ncond = 0; bool ok = true; int t[3] = {0, 0, 0};
    t[axis] = axis_target;
#pragma unroll
    for (int k = 0; k < nother; k++) {
      const int oa = oth[k], c = e[oa];
      if (c == 0 || !per[oa]) { t[oa] = idx[k]; continue; }
      if (c == 1) { const int shift = SWAP ? 2 : 1; ncond++; ok = ok && idx[k] == lim[oa] - shift; t[oa] = SWAP ? 0 : 1; }
      else { const int shift = SWAP ? 1 : 0; ncond++; ok = ok && idx[k] == shift; t[oa] = lim[oa] - (SWAP ? 1 : 2); }
    }
    if (ncond > 0 && ok) { tgt[0] = t[0]; tgt[1] = t[1]; tgt[2] = t[2]; return true; }
  }
  if (L::D == 3 && e[0] * e[0] + e[1] * e[1] + e[2] * e[2] == 3) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int oa = oth[k], ob = oth[1 - k], id_ = idx[k], id2 = idx[1 - k];
      bool ok = true; int t[3] = {0, 0, 0};
      t[axis] = axis_target;
      if (e[oa] == 1) { const int shift = SWAP ? 2 : 1, start = SWAP ? 1 : 2; ok = ok && id_ >= start && id_ < lim[oa] - shift; }
      else { const int shift = SWAP ? 1 : 2; ok = ok && id_ < lim[oa] - shift && id_ >= 1; }
      t[oa] = id_;
      if (e[ob] == 1) { const int shift = SWAP ? 2 : 1; ok = ok && id2 == lim[ob] - shift; t[ob] = SWAP ? 0 : 1; }
      else { const int shift = SWAP ? 1 : 0; ok = ok && id2 == shift; t[ob] = lim[ob] - (SWAP ? 1 : 2); }
      if (ok) { tgt[0] = t[0]; tgt[1] = t[1]; tgt[2] = t[2]; return true; }
    }
  }
  return false;
}

// The reference's edge entries for vectors with three non-zero components wrap along the second other axis whether or
// not it is periodic (kernel_utils.mako:84-118).  Alone, the stray copy lands in a ghost layer nobody reads; with a
// neighbour across that face the layer is, after in-place steps, the neighbour's landing zone and is left alone.
template <bool SWAP>
__device__ __forceinline__ bool pbc_landing_zone(int t_split, int n_split, int link_lo, int link_hi) {
  return SWAP && ((link_lo && t_split == 0) || (link_hi && t_split == n_split - 1));
}

__device__ __forceinline__ void set3(int a, int v, int &x, int &y, int &z) {
  if (a == 0) x = v; else if (a == 1) y = v; else z = v;
}

template <class L, bool SWAP>
__global__ void __launch_bounds__(128)
apply_pbc(const DevParams P, float *dist, int axis, int per0, int per1, int per2, int link_lo, int link_hi) {
  const int lim[3] = {P.nx, P.ny, P.nz};
  const int per[3] = {per0, per1, per2};
  int oth[2];
  if (L::D == 2) { oth[0] = 1 - axis; oth[1] = 2; }
  else { oth[0] = axis == 0 ? 1 : 0; oth[1] = axis == 2 ? 1 : 2; }
  const int n1 = oth[0] == 0 ? P.nx : P.ny, n2 = L::D == 3 ? (oth[1] == 1 ? P.ny : P.nz) : 1;
  const int i1 = blockIdx.x * blockDim.x + threadIdx.x, i2 = blockIdx.y;
  if (i1 >= n1 || i2 >= n2) return;
  const int idx[2] = {i1, i2};
  const int maxd[2] = {n1 - 2, L::D == 3 ? n2 - 2 : 0};
  if (SWAP && L::D == 3) { if (i1 >= n1 - 1 || i2 >= n2 - 1 || i1 < 1 || i2 < 1) return; }
  const long long stride = axis == 0 ? 1 : (axis == 1 ? P.sy : P.sz);
  const long long offset = SWAP ? -stride : stride;
  // (explicit selects instead of run-time indexed local arrays: the latter were miscompiled
  //  for the 2-D instantiation -- gi_low came out equal to gi_high)
  int lx = 0, ly = 0, lz = 0, hx = 0, hy = 0, hz = 0;
  set3(oth[0], i1, lx, ly, lz); set3(oth[0], i1, hx, hy, hz);
  if (L::D == 3) { set3(oth[1], i2, lx, ly, lz); set3(oth[1], i2, hx, hy, hz); }
  set3(axis, SWAP ? 1 : 0, lx, ly, lz);
  set3(axis, SWAP ? lim[axis] - 1 : lim[axis] - 2, hx, hy, hz);
  const long long gi_low = gidx3<L>(P, lx, ly, lz), gi_high = gidx3<L>(P, hx, hy, hz);
  // Subdomain split along the outermost axis: the neighbours deliver whole, PBC-completed planes
  // into the ghost layer (after in-place steps) or into the crossing slots of the first/last real
  // layer (after streaming steps), possibly while this kernel runs -- leave exactly those alone.
  const int so = L::D - 1;                       // split axis
  const int io = (oth[0] == so) ? i1 : ((L::D == 3 && oth[1] == so) ? i2 : -1);
  const int no = so == 2 ? P.nz : P.ny;
  // (after streaming steps the ghost layer holds this subdomain's own outgoing populations and must be
  //  completed here before it is pushed; after in-place steps it is the neighbour's landing zone)
  if (SWAP && io >= 0 && ((link_lo && io == 0) || (link_hi && io == no - 1))) return;
  const int skip_e = (SWAP || io < 0) ? 0 : ((link_lo && io == 1) ? 1 : ((link_hi && io == no - 2) ? -1 : 0));
  float val[L::Q];
  // low -> high
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const int ea = axis == 0 ? L::ex(i) : (axis == 1 ? L::ey(i) : L::ez(i));
    if (ea == -1) val[i] = dist[gi_low + P.DS * (SWAP ? L::opp(i) : i)];
  }
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const int ea = axis == 0 ? L::ex(i) : (axis == 1 ? L::ey(i) : L::ez(i));
    if (ea != -1) continue;
    if (skip_e != 0 && (so == 2 ? L::ez(i) : L::ey(i)) == skip_e) continue;
    const int j = SWAP ? L::opp(i) : i;
    const float fv = val[i];
    if (!isfinite(fv)) continue;
    const int n2e = L::ex(i) * L::ex(i) + L::ey(i) * L::ey(i) + L::ez(i) * L::ez(i);
    if (n2e > 1) {
      const bool main_ok = pbc_main_cond<L, SWAP>(axis, i, per, oth, idx, maxd, link_lo, link_hi);
      if (main_ok) dist[gi_high + P.DS * j] = fv;
      if (SWAP || !main_ok) {
        int tgt[3];
        if (pbc_corner<L, SWAP>(axis, i, per, lim, oth, idx, SWAP ? lim[axis] - 1 : lim[axis] - 2, tgt) &&
            !pbc_landing_zone<SWAP>(tgt[so], no, link_lo, link_hi))
          dist[gidx3<L>(P, tgt[0], tgt[1], tgt[2]) + P.DS * j] = fv;
      }
    } else {
      dist[gi_high + P.DS * j] = fv;
    }
  }
  // high -> low
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const int ea = axis == 0 ? L::ex(i) : (axis == 1 ? L::ey(i) : L::ez(i));
    if (ea == 1) val[i] = dist[gi_high + offset + P.DS * (SWAP ? L::opp(i) : i)];
  }
#pragma unroll
  for (int i = 1; i < L::Q; i++) {
    const int ea = axis == 0 ? L::ex(i) : (axis == 1 ? L::ey(i) : L::ez(i));
    if (ea != 1) continue;
    if (skip_e != 0 && (so == 2 ? L::ez(i) : L::ey(i)) == skip_e) continue;
    const int j = SWAP ? L::opp(i) : i;
    const float fv = val[i];
    if (!isfinite(fv)) continue;
    const int n2e = L::ex(i) * L::ex(i) + L::ey(i) * L::ey(i) + L::ez(i) * L::ez(i);
    if (n2e > 1) {
      const bool main_ok = pbc_main_cond<L, SWAP>(axis, i, per, oth, idx, maxd, link_lo, link_hi);
      if (main_ok) dist[gi_low + offset + P.DS * j] = fv;
      if (SWAP || !main_ok) {
        int tgt[3];
        if (pbc_corner<L, SWAP>(axis, i, per, lim, oth, idx, SWAP ? 0 : 1, tgt) &&
            !pbc_landing_zone<SWAP>(tgt[so], no, link_lo, link_hi))
          dist[gidx3<L>(P, tgt[0], tgt[1], tgt[2]) + P.DS * j] = fv;
      }
    } else {
      dist[gi_low + offset + P.DS * j] = fv;
    }
  }
}

// ----------------------------------------------------------------------------------------
// Collect/DistributeContinuousData[WithSwap] (kernel_utils.mako:476-794); buffer layout
// [dist][other][x].  grid = (ceil(max_lx/128), max_other, ndists).
// ----------------------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ int face_slot(int face, int k) {
  // k-th population with e[axis] * sign >= 1  (sym.get_interblock_dists, sym.py:829-845)
  const int axis = face >> 1, sgn = (face & 1) ? 1 : -1;
  int n = 0, out = -1;
#pragma unroll
  for (int i = 0; i < L::Q; i++) {
    const int ea = axis == 0 ? L::ex(i) : (axis == 1 ? L::ey(i) : L::ez(i));
    if (ea * sgn >= 1) { if (n == k) out = i; n++; }
  }
  return out;
}

template <class L>
__device__ __forceinline__ int opp_rt(int i) {
  int o = 0;
#pragma unroll
  for (int k = 0; k < L::Q; k++) if (k == i) o = L::opp(k);
  return o;
}

template <class L, bool DISTRIBUTE>
__global__ void __launch_bounds__(128)
xfer_continuous(const DevParams P, float *dist, int face, int base_gx, int base_other, int max_lx,
                int max_other, float *buffer, int with_swap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y, k = blockIdx.z;
  if (x >= max_lx || o >= max_other) return;
  const int lim[3] = {P.nx, P.ny, P.nz};
  const int axis = face >> 1, hi = face & 1, n = lim[axis];
  int layer;
  if (!DISTRIBUTE) layer = with_swap ? (hi ? n - 2 : 1) : (hi ? n - 1 : 0);    // lat_linear_macro / lat_linear
  else layer = with_swap ? (hi ? 0 : n - 1) : (hi ? 1 : n - 2);               // lat_linear_with_swap / _dist
  int c0 = base_gx + x, c1, c2 = 0;
  if (L::D == 2) c1 = layer;
  else if (axis == 1) { c1 = layer; c2 = base_other + o; }
  else { c1 = base_other + o; c2 = layer; }
  int slot = face_slot<L>(face, k);
  if (with_swap) slot = opp_rt<L>(slot);
  const long long di = gidx3<L>(P, c0, c1, c2) + P.DS * slot;
  const long long bi = ((long long)k * max_other + o) * max_lx + x;
  if (DISTRIBUTE) dist[di] = buffer[bi]; else buffer[bi] = dist[di];
}

template <bool DISTRIBUTE>
__global__ void xfer_sparse(const uint32_t *__restrict__ idx_array, float *dist, float *buffer, int max_idx,
                            unsigned long long limit) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_idx) return;
  const uint32_t gi = idx_array[idx];
  if (gi == 0xffffffffu || gi >= limit) return;
  if (DISTRIBUTE) dist[gi] = buffer[idx]; else buffer[idx] = dist[gi];
}

// Collect/DistributeContinuousMacroData (kernel_utils.mako:838-953): one scalar field, faces >= 2.
// Collect reads the first/last real layer (lat_linear_macro), distribute writes the ghost layer
// (lat_linear).  2-D: the row is passed explicitly (the reference's `gy` argument) in base_other.
// Buffer rows are dense ([other][x], row length max_lx = the shape the reference allocates,
// subdomain_connection.py:530-534).
template <bool DISTRIBUTE>
static __global__ void __launch_bounds__(128)
xfer_macro(const DevParams P, float *field, int dim, int face, int base_gx, int base_other, int max_lx,
           int max_other, float *buffer) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (x >= max_lx || o >= max_other) return;
  long long gi;
  if (dim == 2) {
    gi = (long long)(base_gx + x) + P.sy * base_other;
  } else {
    const int lim[3] = {P.nx, P.ny, P.nz};
    const int axis = face >> 1, hi = face & 1, n = lim[axis];
    const int layer = DISTRIBUTE ? (hi ? n - 1 : 0) : (hi ? n - 2 : 1);
    const int c1 = axis == 1 ? layer : base_other + o, c2 = axis == 1 ? base_other + o : layer;
    gi = (long long)(base_gx + x) + P.sy * c1 + P.sz * c2;
  }
  const long long bi = (long long)o * max_lx + x;
  if (DISTRIBUTE) field[gi] = buffer[bi]; else buffer[bi] = field[gi];
}

// ----------------------------------------------------------------------------------------
// Output fast path: the device side of _fields_to_host + mask_nonfluid_nodes
// (subdomain_runner.py:1141-1151, io.py:53-59).  Packs the real (non-ghost, non-padding) nodes
// of one field into a dense [nz-2][ny-2][nx-2] array and writes NaN where the node is not wet,
// so that the host only has to hand the buffer to the file writer.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_wet_type(const DevParams &P, uint32_t type) {
  return type == P.tid[SLF_NT_FLUID] || type == P.tid[SLF_NT_HALFBB_WALL] ||
         type == P.tid[SLF_NT_EQUILIBRIUM_VELOCITY] || type == P.tid[SLF_NT_EQUILIBRIUM_DENSITY] ||
         type == P.tid[SLF_NT_REGULARIZED_VELOCITY] || type == P.tid[SLF_NT_REGULARIZED_DENSITY] ||
         type == P.tid[SLF_NT_DO_NOTHING] || type == P.tid[SLF_NT_ZOUHE_VELOCITY] ||
         type == P.tid[SLF_NT_ZOUHE_DENSITY] || type == P.tid[SLF_NT_COPY] || type == P.tid[SLF_NT_YU_OUTFLOW] ||
         type == P.tid[SLF_NT_WALL_TMS];
}

static __global__ void __launch_bounds__(256)
pack_field(const DevParams P, const uint32_t *__restrict__ map, const float *__restrict__ field,
           float *__restrict__ out, int dim) {
  const int rx = P.nx - 2, ry = P.ny - 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= rx) return;
  const int y = blockIdx.y, z = blockIdx.z;                 // z = 0 in 2-D
  const long long gi = (long long)(x + 1) + P.sy * (y + 1) + (dim == 3 ? P.sz * (z + 1) : 0);
  const bool wet = is_wet_type(P, map[gi] & P.tmask);
  out[(long long)x + (long long)rx * (y + (long long)ry * z)] = wet ? field[gi] : __int_as_float(0x7fc00000);
}
