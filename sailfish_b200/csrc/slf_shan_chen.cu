// slf_shan_chen.cu -- C entry points of the binary Shan-Chen path (BASELINE config 5).
#include <cuda_runtime.h>

#include <cmath>
#include <string>

#include <cstdlib>

#include "slf_plan.cuh"
#include "slf_shan_chen.cuh"

extern "C" {

// ---- binary Shan-Chen ------------------------------------------------------------------------
static ScParams make_sc(const SlfPlan *pl) {
  ScParams S;
  const double tau = (6.0 * pl->hp.visc + 1.0) / 2.0;
  const float tauf = (float)tau, tphi = (float)pl->hp.tau_phi;
  S.tau0_inv = 1.0f / tauf; S.tau1_inv = 1.0f / tphi;
  S.omega[0] = (float)(1.0 / tau); S.omega[1] = (float)(1.0 / pl->hp.tau_phi);
  S.tau[0] = tauf; S.tau[1] = tphi;
  for (int k = 0; k < 4; k++) S.g[k] = (float)pl->hp.sc_g[k];
  S.potential = pl->hp.sc_potential;
  for (int d = 0; d < 3; d++) {
    S.body[0][d] = pl->hp.force_enabled ? (float)pl->hp.accel[d] : 0.f;
    S.body[1][d] = (float)pl->hp.accel1[d];
  }
  return S;
}

static int sc_check(SlfPlan *pl) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (pl->hp.simtype != SLF_SIM_SHAN_CHEN) return fail(SLF_EINVAL, "plan was not created for the Shan-Chen model");
  if (pl->hp.model != SLF_BGK) return fail(SLF_EUNSUPPORTED, "Shan-Chen uses BGK (lb_binary.py:131-135)");
  return 0;
}

static int slab_range(SlfPlan *pl, int z0, int z1, long long *gi0, long long *gi1) {
  const int outer = pl->D == 3 ? pl->dp.nz : pl->dp.ny;
  if (z1 < 0) z1 = outer;
  if (z0 < 0 || z1 > outer || z0 > z1) return fail(SLF_EINVAL, "bad slab range");
  const long long slab = pl->D == 3 ? pl->dp.sz : pl->dp.sy;
  *gi0 = slab * z0; *gi1 = slab * z1;
  return 0;
}

int slf_sc_set_initial_conditions(SlfPlan *pl, uint64_t map, uint64_t dist1, uint64_t dist2, uint64_t vx, uint64_t vy,
                                  uint64_t vz, uint64_t rho, uint64_t phi, uint64_t stream) {
  if (int rc = sc_check(pl)) return rc;
  const int bs = 256;
  const unsigned nb = (unsigned)((pl->dp.DS + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
  SLF_LATTICE_SWITCH(pl, (sc_set_initial_conditions<L><<<nb, bs, 0, s>>>(
                             pl->dp, (const uint32_t *)map, (float *)dist1, (float *)dist2, (const float *)vx,
                             (const float *)vy, (const float *)vz, (const float *)rho, (const float *)phi)));
  LAUNCH_OK("sc_set_initial_conditions");
  return 0;
}

int slf_sc_prepare_macro_fields(SlfPlan *pl, uint64_t map, uint64_t dist1, uint64_t dist2, uint64_t rho, uint64_t phi,
                                uint64_t vx, uint64_t vy, uint64_t vz, int options, uint32_t iteration, int z0, int z1,
                                uint64_t stream) {
  if (int rc = sc_check(pl)) return rc;
  long long gi0, gi1;
  if (int rc = slab_range(pl, z0, z1, &gi0, &gi1)) return rc;
  if (gi1 <= gi0) return 0;
  const ScParams S = make_sc(pl);
  const int bs = 128;
  const unsigned nb = (unsigned)((gi1 - gi0 + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
  const bool odd = pl->hp.access_pattern == SLF_AA && (iteration & 1);
  const bool with_v = !(options & SLF_OPTION_DENSITIES_ONLY);
#define SC_MACRO(MODE, WITH_V)                                                                                   \
  SLF_LATTICE_SWITCH(pl, (sc_prepare_macro<L, MODE, WITH_V><<<nb, bs, 0, s>>>(                                    \
                             pl->dp, S, (const uint32_t *)map, (const float *)dist1, (const float *)dist2,        \
                             (float *)rho, (float *)phi, (float *)vx, (float *)vy, (float *)vz, gi0, gi1)))
  if (with_v) {
    if (odd) { SC_MACRO(MODE_AA_ODD, true); } else { SC_MACRO(MODE_AB, true); }
  } else {
    if (odd) { SC_MACRO(MODE_AA_ODD, false); } else { SC_MACRO(MODE_AB, false); }
  }
#undef SC_MACRO
  LAUNCH_OK("sc_prepare_macro");
  return 0;
}

int slf_sc_collide_and_propagate(SlfPlan *pl, int grid_idx, uint64_t map, uint64_t dist_in, uint64_t dist_out,
                                 uint64_t rho, uint64_t phi, uint64_t vx, uint64_t vy, uint64_t vz, int options,
                                 uint32_t iteration, int z0, int z1, uint64_t stream) {
  (void)options;
  if (int rc = sc_check(pl)) return rc;
  if (grid_idx != 0 && grid_idx != 1) return fail(SLF_EINVAL, "grid_idx must be 0 or 1");
  long long gi0, gi1;
  if (int rc = slab_range(pl, z0, z1, &gi0, &gi1)) return rc;
  if (gi1 <= gi0) return 0;
  int mode;
  if (pl->hp.access_pattern == SLF_AB) {
    mode = MODE_AB;
    if (dist_in == dist_out) return fail(SLF_EINVAL, "AB access pattern needs two distinct lattices");
  } else {
    mode = (iteration & 1) ? MODE_AA_ODD : MODE_AA_EVEN;
    if (dist_in != dist_out) return fail(SLF_EINVAL, "AA access pattern works in place");
  }
  const ScParams S = make_sc(pl);
  const int bs = 128;
  const unsigned nb = (unsigned)((gi1 - gi0 + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
#define SC_COLL(MODE, GRID)                                                                                       \
  SLF_LATTICE_SWITCH(pl, (sc_collide<L, MODE, GRID><<<nb, bs, 0, s>>>(                                             \
                             pl->dp, S, (const uint32_t *)map, (const float *)dist_in, (float *)dist_out,          \
                             (const float *)rho, (const float *)phi, (const float *)vx, (const float *)vy,         \
                             (const float *)vz, gi0, gi1)))
  if (grid_idx == 0) {
    if (mode == MODE_AB) { SC_COLL(MODE_AB, 0); } else if (mode == MODE_AA_EVEN) { SC_COLL(MODE_AA_EVEN, 0); } else { SC_COLL(MODE_AA_ODD, 0); }
  } else {
    if (mode == MODE_AB) { SC_COLL(MODE_AB, 1); } else if (mode == MODE_AA_EVEN) { SC_COLL(MODE_AA_EVEN, 1); } else { SC_COLL(MODE_AA_ODD, 1); }
  }
#undef SC_COLL
  LAUNCH_OK("sc_collide");
  return 0;
}

int slf_sc_collide_and_propagate_fused(SlfPlan *pl, uint64_t map, uint64_t dist1_in, uint64_t dist1_out,
                                       uint64_t dist2_in, uint64_t dist2_out, uint64_t rho, uint64_t phi,
                                       uint64_t vx, uint64_t vy, uint64_t vz, int options, uint32_t iteration,
                                       int z0, int z1, uint64_t stream) {
  if (int rc = sc_check(pl)) return rc;
  long long gi0, gi1;
  if (int rc = slab_range(pl, z0, z1, &gi0, &gi1)) return rc;
  if (gi1 <= gi0) return 0;
  int mode;
  if (pl->hp.access_pattern == SLF_AB) {
    mode = MODE_AB;
    if (dist1_in == dist1_out || dist2_in == dist2_out) return fail(SLF_EINVAL, "AB access pattern needs two distinct lattices");
  } else {
    mode = (iteration & 1) ? MODE_AA_ODD : MODE_AA_EVEN;
    if (dist1_in != dist1_out || dist2_in != dist2_out) return fail(SLF_EINVAL, "AA access pattern works in place");
  }
  const bool save = options & SLF_OPTION_SAVE_MACRO_FIELDS;
  const ScParams S = make_sc(pl);
  const int bs = 128;
  const unsigned nb = (unsigned)((gi1 - gi0 + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
  // walls: bounce-back copies in AB / with propagation off, nothing to do in the AA pattern
  const int bounce_walls = (mode == MODE_AB || !pl->dp.prop) ? 1 : 0;
  const DevParams &d = pl->dp;
  const int outer_n = pl->D == 3 ? d.nz : d.ny;
  const int zb = z0, ze = z1 < 0 ? outer_n : z1;
  // tile-shaped blocks for thick 3-D ranges (single boundary layers keep the linear map)
  const bool tiled = pl->D == 3 && ze - zb >= 2 && !getenv("SLF_SC_LINEAR");
  const dim3 tgrid((d.ax + 31) / 32, (d.ay + 1) / 2, (ze - zb + 1) / 2);
#define SC_FUSED(SAVE)                                                                                            \
  if (tiled) {                                                                                                    \
    SLF_LATTICE_SWITCH(pl, (sc_collide_fused_tile<L, SAVE><<<tgrid, 128, 0, s>>>(                                  \
                               pl->dp, S,                                                                          \
                               make_sc_tables<L>(pl->dp, mode, (const float *)dist1_in, (float *)dist1_out,        \
                                                 (const float *)dist2_in, (float *)dist2_out, (const float *)rho,  \
                                                 (const float *)phi, gi0),                                         \
                               (const uint32_t *)map, (float *)vx, (float *)vy, (float *)vz, gi0, zb, ze,          \
                               bounce_walls)));                                                                    \
  } else {                                                                                                        \
    SLF_LATTICE_SWITCH(pl, (sc_collide_fused<L, SAVE><<<nb, bs, 0, s>>>(                                           \
                               pl->dp, S,                                                                          \
                               make_sc_tables<L>(pl->dp, mode, (const float *)dist1_in, (float *)dist1_out,        \
                                                 (const float *)dist2_in, (float *)dist2_out, (const float *)rho,  \
                                                 (const float *)phi, gi0),                                         \
                               (const uint32_t *)map, (float *)vx, (float *)vy, (float *)vz, gi0,                  \
                               (unsigned)(gi1 - gi0), bounce_walls)));                                             \
  }
  if (save) { SC_FUSED(true); } else { SC_FUSED(false); }
#undef SC_FUSED
  LAUNCH_OK("sc_collide_fused");
  return 0;
}

int slf_apply_macro_periodic_boundary_conditions(SlfPlan *pl, uint64_t field, int axis, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (axis == 2 && pl->D == 2) return 0;
  if (axis < 0 || axis >= pl->D) return fail(SLF_EINVAL, "bad axis");
  const DevParams &d = pl->dp;
  const int n1 = axis == 0 ? d.ny : d.nx;
  const int n2 = pl->D == 3 ? (axis == 2 ? d.ny : d.nz) : 1;
  dim3 grid((n1 + 127) / 128, n2);
  macro_pbc<<<grid, 128, 0, (cudaStream_t)stream>>>(d, (float *)field, axis, pl->D, pl->link_lo, pl->link_hi);
  LAUNCH_OK("macro_pbc");
  return 0;
}

}  // extern "C"
