#include "slf_collide.cuh"
int slf_collide_d2q9(SlfPlan *pl, int model, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                     float *rho, float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s) {
  return model == SLF_MRT ? dispatch_mode<D2Q9, SLF_MRT>(pl, mode, save, map, din, dout, rho, vx, vy, vz, z0, z1, s)
                          : dispatch_mode<D2Q9, SLF_BGK>(pl, mode, save, map, din, dout, rho, vx, vy, vz, z0, z1, s);
}
int slf_multistep_d2q9(SlfPlan *pl, int model, const uint32_t *map, float *dist, float *rho, float *vx, float *vy,
                       float *vz, uint32_t it0, int nsteps, int save_last, cudaStream_t s) {
  return model == SLF_MRT ? launch_multistep<D2Q9, SLF_MRT>(pl, map, dist, rho, vx, vy, vz, it0, nsteps, save_last, s)
                          : launch_multistep<D2Q9, SLF_BGK>(pl, map, dist, rho, vx, vy, vz, it0, nsteps, save_last, s);
}
