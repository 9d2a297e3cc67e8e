// slf_collide.cuh -- launch + template dispatch of the collide kernels (included by slf_collide_*.cu).
#pragma once
#include <cstdlib>

#include "slf_plan.cuh"

// ---- dispatch helpers --------------------------------------------------------------------
template <class L, int MODEL, int MODE, bool SAVE>
static int launch_collide(SlfPlan *pl, const uint32_t *map, const float *din, float *dout, float *rho,
                          float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s) {
  const long long slab = pl->D == 3 ? pl->dp.sz : pl->dp.sy;
  const long long gi0 = slab * z0, gi1 = slab * z1;
  const long long n = gi1 - gi0;
  if (n <= 0) return 0;
  const bool bound = pl->bound_map != 0 && pl->bound_map == (uint64_t)map;
  {
    const int bs = 128;
    const long long nb = (n + bs - 1) / bs;
    const bool plain = bound && pl->dp.relax && pl->dp.prop;
    if constexpr (MODE != MODE_AB) {
      if (plain) {
        float *dist = dout;      // AA: dist_in == dist_out (checked by the caller)
        if (pl->d_seg) {
          const long long s0 = pl->seg_off[z0], s1 = pl->seg_off[z1];
          if (s1 > s0) {
            collide_bulk_aa_seg<L, MODEL, SAVE><<<(unsigned)((s1 - s0 + 3) / 4), bs, 0, s>>>(
                pl->dp, make_aa_table<L>(pl->dp, dist, 0, MODE == MODE_AA_ODD), pl->d_fluid_mask, pl->d_seg + s0,
                (unsigned)(s1 - s0), rho, vx, vy, vz);
            LAUNCH_OK("collide_bulk_aa_seg");
          }
        } else {
          collide_bulk_aa_tab<L, MODEL, SAVE><<<(unsigned)nb, bs, 0, s>>>(
              pl->dp, make_aa_table<L>(pl->dp, dist, gi0, MODE == MODE_AA_ODD), pl->d_fluid_mask, rho, vx, vy, vz,
              gi0, (unsigned)n);
          LAUNCH_OK("collide_bulk_aa");
        }
      }
    }
    if (MODE == MODE_AB && plain) {
      collide_bulk_ab_tab<L, MODEL, SAVE><<<(unsigned)nb, bs, 0, s>>>(
          pl->dp, make_ab_table<L>(pl->dp, din, dout, gi0), pl->d_fluid_mask, pl->d_wall_mask, rho, vx, vy, vz, gi0,
          (unsigned)n);
      LAUNCH_OK("collide_bulk_ab");
    } else if (MODE == MODE_AB || !plain) {
      if (bound)
        collide_bulk<L, MODEL, MODE, SAVE, true><<<(unsigned)nb, bs, 0, s>>>(
            pl->dp, map, pl->d_fluid_mask, pl->d_wall_mask, din, dout, rho, vx, vy, vz, gi0, gi1);
      else
        collide_bulk<L, MODEL, MODE, SAVE, false><<<(unsigned)nb, bs, 0, s>>>(
            pl->dp, map, nullptr, nullptr, din, dout, rho, vx, vy, vz, gi0, gi1);
      LAUNCH_OK("collide_bulk");
    }
  }
  {
    const int bs = 128;
    if (bound) {
      const long long k0 = pl->layer_off[z0], k1 = pl->layer_off[z1];
      if (k1 > k0) {
        const long long nb = (k1 - k0 + bs - 1) / bs;
        collide_boundary<L, MODEL, MODE, SAVE, true><<<(unsigned)nb, bs, 0, s>>>(
            pl->dp, map, pl->d_list, din, dout, rho, vx, vy, vz, k0, k1);
        LAUNCH_OK("collide_boundary");
      }
    } else {
      const long long nb = (n + bs - 1) / bs;
      collide_boundary<L, MODEL, MODE, SAVE, false><<<(unsigned)nb, bs, 0, s>>>(
          pl->dp, map, nullptr, din, dout, rho, vx, vy, vz, gi0, gi1);
      LAUNCH_OK("collide_boundary");
    }
  }
  return 0;
}

template <class L, int MODEL>
static int dispatch_mode(SlfPlan *pl, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                         float *rho, float *vx, float *vy, float *vz, int gi0, int gi1,
                         cudaStream_t s) {
#define SLF_CASE(M)                                                                                     \
  case M:                                                                                               \
    return save ? launch_collide<L, MODEL, M, true>(pl, map, din, dout, rho, vx, vy, vz, gi0, gi1, s)   \
                : launch_collide<L, MODEL, M, false>(pl, map, din, dout, rho, vx, vy, vz, gi0, gi1, s);
  switch (mode) {
    SLF_CASE(MODE_AB)
    SLF_CASE(MODE_AA_EVEN)
    SLF_CASE(MODE_AA_ODD)
  }
#undef SLF_CASE
  return fail(SLF_EINVAL, "bad mode");
}


// ---- many steps in one cooperative launch (small lattices) ---------------------------------------------
// Returns 1 when the plan / lattice does not qualify (the caller then steps kernel by kernel), 0 when the
// steps were issued, an error code otherwise.
template <class L, int MODEL>
static int launch_multistep(SlfPlan *pl, const uint32_t *map, float *dist, float *rho, float *vx, float *vy,
                            float *vz, uint32_t it0, int nsteps, int save_last, cudaStream_t s) {
  constexpr int B = MultiStepTuning<L, MODEL>::BLOCK;
  const DevParams &d = pl->dp;
  if (pl->bound_map == 0 || pl->bound_map != (uint64_t)map || !d.relax || !d.prop || !pl->d_other_mask) return 1;
  const long long n = d.DS;
  static int per_sm = -1, sms = 0;
  if (per_sm < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, collide_multistep<L, MODEL>, B, 0) != cudaSuccess) per_sm = 0;
  }
  const long long resident = (long long)per_sm * sms;
  const long long nb = (n + B - 1) / B;
  if (resident <= 0 || nb > resident || nb > SLF_MULTISTEP_MAX_BLOCKS) return 1;     // one node per thread, all co-resident
  // everything a node touches lies within R nodes of it: the blocks within `reach` of its own
  const long long R = d.sz * (pl->D == 3 ? 1 : 0) + d.sy + 1;
  const int reach = (int)((R + B - 1) / B);
  if (2 * reach > 32) return 1;
  if (!pl->d_flags) CUDA_OK(cudaMalloc((void **)&pl->d_flags, SLF_MULTISTEP_MAX_BLOCKS * sizeof(unsigned int)));
  CUDA_OK(cudaMemsetAsync(pl->d_flags, 0, (size_t)nb * sizeof(unsigned int), s));
  DevParams P = d;
  SlotTable<L> Te = make_aa_table<L>(d, dist, 0, false), To = make_aa_table<L>(d, dist, 0, true);
  const uint32_t *fm = pl->d_fluid_mask, *om = pl->d_other_mask;
  unsigned un = (unsigned)n;
  unsigned int *flags = pl->d_flags;
  int reach_arg = reach;
  void *args[] = {&P, &Te, &To, &fm, &om, &map, &dist, &rho, &vx, &vy, &vz, &un, &it0, &nsteps, &save_last, &flags,
                  &reach_arg};
  CUDA_OK(cudaLaunchCooperativeKernel((void *)collide_multistep<L, MODEL>, dim3((unsigned)nb), dim3(B), args, 0, s));
  LAUNCH_OK("collide_multistep");
  return 0;
}
