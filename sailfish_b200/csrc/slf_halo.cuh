// slf_halo.cuh -- ghost-layer exchange for subdomains split along the outermost axis
// (conn_axis = z in 3-D, y in 2-D), device to device over NVLink peer mappings.
//
// Replaces the reference's collect kernel -> D2H -> ZeroMQ -> H2D -> distribute kernel chain
// (subdomain_runner.py:1064-1139, kernel_utils.mako:526-794, connector.py:73-117).  With the
// SoA layout a face message is |dists| whole planes that are already contiguous in the
// lattice array, so "pack" and "unpack" disappear: one kernel copies the planes straight
// from this GPU's lattice into the neighbour's lattice through the peer pointer.
// Layer/slot selection follows the reference exactly:
//   after an AB or AA-odd step : ghost layer, natural slots  -> neighbour's first real layer
//   after an AA-even step      : first real layer, opposite slots -> neighbour's ghost layer
// (lat_linear / lat_linear_macro / lat_linear_dist / lat_linear_with_swap,
//  subdomain_runner.py:488-510).
#pragma once

struct HaloArgs {
  const float *src[9];
  float *dst[9];
  int n;                 // number of planes
  long long plane;       // floats per plane (multiple of 4)
};

__global__ void __launch_bounds__(256) halo_push_kernel(const HaloArgs a) {
  const long long n4 = a.plane >> 2;
  const int k = blockIdx.y;
  const float4 *__restrict__ s = reinterpret_cast<const float4 *>(a.src[k]);
  float4 *__restrict__ d = reinterpret_cast<float4 *>(a.dst[k]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    d[i] = s[i];
  __threadfence_system();
}

__global__ void sem_signal_kernel(unsigned int *sem) {
  __threadfence_system();
  atomicAdd_system(sem, 1u);
}

__global__ void sem_wait_kernel(volatile unsigned int *sem, unsigned int *consumed) {
  const unsigned int want = *consumed + 1u;
  while (*sem < want) { __nanosleep(200); }
  __threadfence_system();
  *consumed = want;
}

extern "C" {

int slf_halo_push(SlfPlan *pl, uint64_t dist, int face, uint64_t peer_dist, int peer_lat_n, int peer_arr_n,
                  int with_swap, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  const DevParams &d = pl->dp;
  const int axis = pl->D - 1;                   // outermost axis
  if ((face >> 1) != axis) return fail(SLF_EUNSUPPORTED, "halo push is implemented for the outermost axis only");
  const int hi = face & 1;
  const int n = axis == 2 ? d.nz : d.ny;
  const long long plane = axis == 2 ? d.sz : d.sy;           // nodes per layer
  if (plane % 4) return fail(SLF_EINVAL, "plane size must be a multiple of 4 floats");
  const long long peer_DS = plane * peer_arr_n;
  const int src_layer = with_swap ? (hi ? n - 2 : 1) : (hi ? n - 1 : 0);
  const int dst_layer = with_swap ? (hi ? 0 : peer_lat_n - 1) : (hi ? 1 : peer_lat_n - 2);
  HaloArgs a;
  a.n = 0; a.plane = plane;
  const int sgn = hi ? 1 : -1;
  const int Q = pl->Q;
  for (int i = 0; i < Q; i++) {
    int e, opp;
    switch (pl->hp.lattice) {
      case SLF_D2Q9: e = D2Q9::ey(i); opp = D2Q9::opp(i); break;
      case SLF_D3Q19: e = D3Q19::ez(i); opp = D3Q19::opp(i); break;
      default: e = D3Q27::ez(i); opp = D3Q27::opp(i); break;
    }
    if (e * sgn < 1) continue;
    const int slot = with_swap ? opp : i;
    a.src[a.n] = (const float *)dist + d.DS * slot + plane * src_layer;
    a.dst[a.n] = (float *)peer_dist + peer_DS * slot + plane * dst_layer;
    a.n++;
  }
  const long long n4 = plane / 4;
  int bx = (int)((n4 + 255) / 256);
  if (bx > 592) bx = 592;                       // 4 x 148 SMs, grid-stride beyond
  dim3 grid(bx, a.n);
  halo_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  LAUNCH_OK("halo_push");
  return 0;
}

int slf_halo_push_field(SlfPlan *pl, uint64_t field, int face, uint64_t peer_field, int peer_lat_n, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  const DevParams &d = pl->dp;
  const int axis = pl->D - 1;
  if ((face >> 1) != axis) return fail(SLF_EUNSUPPORTED, "halo push is implemented for the outermost axis only");
  const int hi = face & 1;
  const int n = axis == 2 ? d.nz : d.ny;
  const long long plane = axis == 2 ? d.sz : d.sy;
  if (plane % 4) return fail(SLF_EINVAL, "plane size must be a multiple of 4 floats");
  HaloArgs a;
  a.n = 1; a.plane = plane;
  a.src[0] = (const float *)field + plane * (hi ? n - 2 : 1);             // first / last real layer
  a.dst[0] = (float *)peer_field + plane * (hi ? 0 : peer_lat_n - 1);     // neighbour's ghost layer
  const long long n4 = plane / 4;
  int bx = (int)((n4 + 255) / 256);
  if (bx > 592) bx = 592;
  halo_push_kernel<<<dim3(bx, 1), 256, 0, (cudaStream_t)stream>>>(a);
  LAUNCH_OK("halo_push_field");
  return 0;
}

int slf_sem_signal(uint64_t sem, uint64_t stream) {
  sem_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned int *)sem);
  LAUNCH_OK("sem_signal");
  return 0;
}

int slf_sem_wait(uint64_t sem, uint64_t consumed, uint64_t stream) {
  sem_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((volatile unsigned int *)sem, (unsigned int *)consumed);
  LAUNCH_OK("sem_wait");
  return 0;
}

}  // extern "C"
