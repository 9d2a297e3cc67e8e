// slf_plan.cuh -- host-side plan object and error plumbing shared by the translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "slf_kernels.cuh"

int slf_fail(int code, const std::string &msg);
void slf_count_launch();
#define fail slf_fail
#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return slf_fail((int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
  } while (0)
#define LAUNCH_OK(name)                                                                       \
  do {                                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) return slf_fail((int)e_, std::string(name) + ": " + cudaGetErrorString(e_)); \
    slf_count_launch();                                                                       \
  } while (0)

// step-pair graphs of slf_run_steps, owned by the plan (guarded by SlfPlan::graph_mutex)
struct StepGraphKey { uint64_t map, a, b, rho; int parity; };
struct StepGraph { StepGraphKey key; cudaGraphExec_t exec; uint64_t kernels; };

#define SLF_MULTISTEP_MAX_BLOCKS 4096

struct SlfPlan {
  SlfParams hp;
  DevParams dp;
  float *d_node_params = nullptr;
  int Q = 0, D = 0;
  // derived from a bound node map (slf_plan_bind_map)
  uint64_t bound_map = 0;
  uint32_t *d_fluid_mask = nullptr, *d_wall_mask = nullptr, *d_list = nullptr;
  uint32_t *d_seg = nullptr;           // indices of the mask words that contain fluid (sparse geometries), or null
  std::vector<long long> seg_off;      // segment-list offsets per outermost-axis layer (size n + 1)
  uint32_t *d_other_mask = nullptr;    // 1 bit per node: member of d_list (multi-step kernel)
  unsigned int *d_flags = nullptr;     // per-block step counters of the multi-step kernel
  std::vector<long long> layer_off;    // list offsets per outermost-axis layer (size n + 1)
  int link_lo = 0, link_hi = 0;        // neighbours across the low / high face of the split axis
  int face_links = 0x3f;               // bit f: face f (2 * axis + side) has a neighbour (slf_plan_set_face_links);
                                       // default "all": the windows of slf_halo_push_face are then the plain ones
  std::vector<StepGraph> graphs;
  std::mutex graph_mutex;
  unsigned int *d_push_done = nullptr; // block counters of halo_push_kernel (last block signals), [2]
};

// one per lattice, each in its own translation unit (slf_collide_*.cu) so they build in parallel
int slf_collide_d2q9(SlfPlan *pl, int model, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                     float *rho, float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s);
int slf_collide_d3q19_bgk(SlfPlan *pl, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                          float *rho, float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s);
int slf_collide_d3q19_mrt(SlfPlan *pl, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                          float *rho, float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s);
int slf_collide_d3q27(SlfPlan *pl, int mode, bool save, const uint32_t *map, const float *din, float *dout,
                      float *rho, float *vx, float *vy, float *vz, int z0, int z1, cudaStream_t s);

// many steps in one cooperative launch (small lattices); 1 = does not qualify
int slf_multistep_d2q9(SlfPlan *pl, int model, const uint32_t *map, float *dist, float *rho, float *vx, float *vy,
                       float *vz, uint32_t it0, int nsteps, int save_last, cudaStream_t s);
int slf_multistep_d3q19_bgk(SlfPlan *pl, const uint32_t *map, float *dist, float *rho, float *vx, float *vy, float *vz,
                            uint32_t it0, int nsteps, int save_last, cudaStream_t s);
int slf_multistep_d3q19_mrt(SlfPlan *pl, const uint32_t *map, float *dist, float *rho, float *vx, float *vy, float *vz,
                            uint32_t it0, int nsteps, int save_last, cudaStream_t s);
int slf_multistep_d3q27(SlfPlan *pl, const uint32_t *map, float *dist, float *rho, float *vx, float *vy, float *vz,
                        uint32_t it0, int nsteps, int save_last, cudaStream_t s);

#define SLF_LATTICE_SWITCH(pl, CALL)            \
  switch ((pl)->hp.lattice) {                    \
    case SLF_D2Q9: { using L = D2Q9; CALL; } break;   \
    case SLF_D3Q19: { using L = D3Q19; CALL; } break; \
    case SLF_D3Q27: { using L = D3Q27; CALL; } break; \
  }

