// slf_api.cu -- C ABI of libslf_b200.so (see include/slf_b200.h for the contract and the
// reference interface each entry point replaces).
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "slf_plan.cuh"

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static int g_device = -1;

int slf_fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
void slf_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }



extern "C" {

int slf_abi_version(void) { return SLF_ABI_VERSION; }
const char *slf_last_error(void) { return g_err.c_str(); }
uint64_t slf_launch_count(void) { return g_launches.load(); }

int slf_device_count(int *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(SLF_ENODEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
  *count = n;
  return 0;
}

int slf_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SLF_ENODEVICE, "no CUDA device available; libslf_b200 has no CPU fallback");
  if (device < 0 || device >= n) return fail(SLF_EINVAL, "device index out of range");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(SLF_EUNSUPPORTED, "libslf_b200 is built for sm_100a (B200) only");
  CUDA_OK(cudaFree(0));
  g_device = device;
  return 0;
}

int slf_device_info(char *name, int name_len, uint64_t *total_mem, int *sm_count) {
  cudaDeviceProp prop;
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
  if (total_mem) *total_mem = prop.totalGlobalMem;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return 0;
}

int slf_malloc(uint64_t bytes, uint64_t *dptr) {
  void *p = nullptr;
  CUDA_OK(cudaMalloc(&p, bytes ? bytes : 4));
  *dptr = (uint64_t)p;
  return 0;
}
int slf_free(uint64_t dptr) { CUDA_OK(cudaFree((void *)dptr)); return 0; }
int slf_memset(uint64_t dptr, int value, uint64_t bytes, uint64_t stream) {
  CUDA_OK(cudaMemsetAsync((void *)dptr, value, bytes, (cudaStream_t)stream));
  return 0;
}
int slf_host_alloc(uint64_t bytes, void **hptr) {
  CUDA_OK(cudaHostAlloc(hptr, bytes ? bytes : 4, cudaHostAllocDefault));
  memset(*hptr, 0, bytes);
  return 0;
}
int slf_host_free(void *hptr) { CUDA_OK(cudaFreeHost(hptr)); return 0; }
int slf_memcpy_h2d(uint64_t dptr, const void *hptr, uint64_t bytes) {
  CUDA_OK(cudaMemcpy((void *)dptr, hptr, bytes, cudaMemcpyHostToDevice));
  return 0;
}
int slf_memcpy_d2h(void *hptr, uint64_t dptr, uint64_t bytes) {
  CUDA_OK(cudaMemcpy(hptr, (const void *)dptr, bytes, cudaMemcpyDeviceToHost));
  return 0;
}
int slf_memcpy_h2d_async(uint64_t dptr, const void *hptr, uint64_t bytes, uint64_t stream) {
  CUDA_OK(cudaMemcpyAsync((void *)dptr, hptr, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
int slf_memcpy_d2h_async(void *hptr, uint64_t dptr, uint64_t bytes, uint64_t stream) {
  CUDA_OK(cudaMemcpyAsync(hptr, (const void *)dptr, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
int slf_memcpy_d2d_async(uint64_t dst, uint64_t src, uint64_t bytes, uint64_t stream) {
  CUDA_OK(cudaMemcpyAsync((void *)dst, (const void *)src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int slf_sync(void) { CUDA_OK(cudaDeviceSynchronize()); return 0; }
int slf_stream_create(uint64_t *stream, int high_priority) {
  cudaStream_t s;
  int lo = 0, hi = 0;
  CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_OK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
  *stream = (uint64_t)s;
  return 0;
}
int slf_stream_destroy(uint64_t stream) { CUDA_OK(cudaStreamDestroy((cudaStream_t)stream)); return 0; }
int slf_stream_sync(uint64_t stream) { CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }
int slf_stream_wait_event(uint64_t stream, uint64_t event) {
  CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return 0;
}
int slf_event_create(uint64_t *event, int timing) {
  cudaEvent_t e;
  CUDA_OK(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
  *event = (uint64_t)e;
  return 0;
}
int slf_event_destroy(uint64_t event) { CUDA_OK(cudaEventDestroy((cudaEvent_t)event)); return 0; }
int slf_event_record(uint64_t event, uint64_t stream) {
  CUDA_OK(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return 0;
}
int slf_event_sync(uint64_t event) { CUDA_OK(cudaEventSynchronize((cudaEvent_t)event)); return 0; }
int slf_event_elapsed_ms(uint64_t start, uint64_t end, float *ms) {
  CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end));
  return 0;
}

// ---- plan ------------------------------------------------------------------------------
int slf_plan_create(const SlfParams *p, SlfPlan **out) {
  if (!p || !out) return fail(SLF_EINVAL, "null argument");
  if (p->struct_bytes != sizeof(SlfParams))
    return fail(SLF_EINVAL, "SlfParams size mismatch (ABI): got " + std::to_string(p->struct_bytes) +
                                ", expected " + std::to_string(sizeof(SlfParams)));
  if (p->lattice < SLF_D2Q9 || p->lattice > SLF_D3Q27) return fail(SLF_EINVAL, "unknown lattice");
  if (p->model != SLF_BGK && p->model != SLF_MRT) return fail(SLF_EINVAL, "unknown model");
  if (p->model == SLF_MRT && p->lattice == SLF_D3Q27)
    return fail(SLF_EUNSUPPORTED, "D3Q27 has no MRT basis in the reference (sym.py:409-428)");
  if (p->access_pattern != SLF_AB && p->access_pattern != SLF_AA) return fail(SLF_EINVAL, "unknown access pattern");
  if (p->access_pattern == SLF_AA &&
      (p->type_id[SLF_NT_COPY] != SLF_NT_ABSENT || p->type_id[SLF_NT_YU_OUTFLOW] != SLF_NT_ABSENT))
    return fail(SLF_EUNSUPPORTED, "NTCopy / NTYuOutflow nodes need the AB access pattern (boundary.mako:626-628)");
  if (p->lat_nx < 3 || p->lat_ny < 3 || p->arr_nx < p->lat_nx || p->arr_ny < p->lat_ny)
    return fail(SLF_EINVAL, "bad lattice/array size");
  if (p->bits_type < 1 || p->bits_type + p->bits_param + p->bits_scratch > 31)
    return fail(SLF_EINVAL, "bad node-code bit layout");
  if (p->type_id[SLF_NT_FLUID] == SLF_NT_ABSENT) return fail(SLF_EINVAL, "fluid type id missing");
  if (!(p->visc > 0.0)) return fail(SLF_EINVAL, "viscosity must be positive");
  SlfPlan *pl = new SlfPlan();
  pl->hp = *p;
  pl->hp.node_params = nullptr;
  const int D = p->lattice == SLF_D2Q9 ? 2 : 3;
  pl->D = D;
  pl->Q = p->lattice == SLF_D2Q9 ? 9 : (p->lattice == SLF_D3Q19 ? 19 : 27);
  DevParams &d = pl->dp;
  memset(&d, 0, sizeof(d));
  d.nx = p->lat_nx; d.ny = p->lat_ny; d.nz = D == 3 ? p->lat_nz : 1;
  d.ax = p->arr_nx; d.ay = p->arr_ny; d.az = D == 3 ? p->arr_nz : 1;
  if (D == 3 && (d.nz < 3 || d.az < d.nz)) { delete pl; return fail(SLF_EINVAL, "bad z size"); }
  d.DS = (long long)d.ax * d.ay * d.az;
  d.sy = d.ax; d.sz = (long long)d.ax * d.ay;
  d.tmask = (1u << p->bits_type) - 1u;
  d.pmask = (1u << p->bits_param) - 1u;
  d.pshift = p->bits_type;
  d.oshift = p->bits_type + p->bits_param + p->bits_scratch;
  for (int k = 0; k < SLF_NT_COUNT; k++) d.tid[k] = p->type_id[k];
  const double tau = (6.0 * p->visc + 1.0) / 2.0;            // sym.py:847-848
  d.tau = (float)tau;
  d.omega = (float)(1.0 / tau);
  const float viscf = (float)p->visc;
  d.inv_tau_mrt = 1.0f / (0.5f + 3.0f * viscf);
  d.guo_pref = 3.0f * (1.0f - 1.0f / (2.0f * d.tau));
  d.forced = p->force_enabled ? 1 : 0;
  for (int k = 0; k < 3; k++) d.a[k] = p->force_enabled ? (float)p->accel[k] : 0.f;
  d.relax = p->relaxation_enabled; d.prop = p->propagation_enabled; d.use_tags = p->use_link_tags;
  d.unused_tags = (unsigned int)p->unused_tag_bits;
  const int np = p->num_node_params > 0 ? p->num_node_params : 0;
  std::vector<float> h(np + 4, 0.f);
  for (int k = 0; k < np; k++) h[k] = (float)p->node_params[k];
  cudaError_t e = cudaMalloc((void **)&pl->d_node_params, h.size() * sizeof(float));
  if (e != cudaSuccess) { delete pl; return fail((int)e, std::string("cudaMalloc(node_params): ") + cudaGetErrorString(e)); }
  e = cudaMemcpy(pl->d_node_params, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(pl->d_node_params); delete pl; return fail((int)e, cudaGetErrorString(e)); }
  d.node_params = pl->d_node_params;
  e = cudaMalloc((void **)&pl->d_push_done, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(pl->d_push_done, 0, 4 * sizeof(unsigned int));
  if (e != cudaSuccess) { cudaFree(pl->d_node_params); delete pl; return fail((int)e, cudaGetErrorString(e)); }
  *out = pl;
  return 0;
}

static void purge_graphs(SlfPlan *plan);

int slf_plan_set_links(SlfPlan *plan, int link_low, int link_high) {
  if (!plan) return fail(SLF_EINVAL, "null plan");
  purge_graphs(plan);
  plan->link_lo = link_low != 0; plan->link_hi = link_high != 0;
  return 0;
}

// Step graphs capture kernel arguments by value (mask pointers, link flags): drop them whenever those change.
static void purge_graphs(SlfPlan *plan) {
  std::lock_guard<std::mutex> lock(plan->graph_mutex);
  for (auto &g : plan->graphs) cudaGraphExecDestroy(g.exec);
  plan->graphs.clear();
}

int slf_plan_destroy(SlfPlan *plan) {
  if (!plan) return 0;
  purge_graphs(plan);
  if (plan->d_node_params) cudaFree(plan->d_node_params);
  if (plan->d_fluid_mask) cudaFree(plan->d_fluid_mask);
  if (plan->d_wall_mask) cudaFree(plan->d_wall_mask);
  if (plan->d_list) cudaFree(plan->d_list);
  if (plan->d_push_done) cudaFree(plan->d_push_done);
  if (plan->d_other_mask) cudaFree(plan->d_other_mask);
  if (plan->d_seg) cudaFree(plan->d_seg);
  if (plan->d_flags) cudaFree(plan->d_flags);
  delete plan;
  return 0;
}


// Derives the per-map acceleration structures: 1-bit masks of plain fluid nodes and of
// full-way bounce-back walls (handled by the bulk kernel without touching the uint32 map)
// and the sorted list of all other active nodes (handled by the boundary kernel).  Call
// again after the map changes (Subdomain.update_node); unbound maps still work through the
// generic scanning kernels.
int slf_plan_bind_map(SlfPlan *pl, uint64_t map) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  purge_graphs(pl);
  const DevParams &d = pl->dp;
  if (d.DS >= (1ll << 32)) return fail(SLF_EUNSUPPORTED, "node lists use 32-bit indices");
  pl->bound_map = 0;
  if (!map) return 0;
  std::vector<uint32_t> h((size_t)d.DS);
  CUDA_OK(cudaMemcpy(h.data(), (const void *)map, (size_t)d.DS * 4, cudaMemcpyDeviceToHost));
  const long long words = (d.DS + 31) / 32;
  std::vector<uint32_t> fm((size_t)words, 0u), wm((size_t)words, 0u), om((size_t)words, 0u), list;
  const int outer = pl->D == 3 ? d.nz : d.ny;
  const long long slab = pl->D == 3 ? d.sz : d.sy;
  pl->layer_off.assign(outer + 1, 0);
  const uint32_t t_fluid = d.tid[SLF_NT_FLUID], t_wall = d.tid[SLF_NT_FULLBB_WALL];
  const uint32_t t_ghost = d.tid[SLF_NT_GHOST], t_unused = d.tid[SLF_NT_UNUSED], t_prop = d.tid[SLF_NT_PROPAGATION_ONLY];
  for (int o = 0; o < outer; o++) {
    pl->layer_off[o] = (long long)list.size();
    for (long long gi = slab * o; gi < slab * (o + 1); gi++) {
      const uint32_t type = h[gi] & d.tmask;
      if (type == t_ghost || type == t_unused || type == t_prop) continue;
      const int x = (int)(gi % d.ax);
      const long long t = gi / d.ax;
      const int y = (int)(t % d.ay), z = (int)(t / d.ay);
      if (x >= d.nx || y >= d.ny || z >= d.nz) continue;
      const bool interior = x > 0 && x < d.nx - 1 && y > 0 && y < d.ny - 1 && (pl->D == 2 || (z > 0 && z < d.nz - 1));
      if (type == t_fluid && interior) fm[gi >> 5] |= 1u << (gi & 31);
      else if (type == t_wall && interior) wm[gi >> 5] |= 1u << (gi & 31);
      else { list.push_back((uint32_t)gi); om[gi >> 5] |= 1u << (gi & 31); }
    }
  }
  pl->layer_off[outer] = (long long)list.size();
  if (pl->d_fluid_mask) { cudaFree(pl->d_fluid_mask); pl->d_fluid_mask = nullptr; }
  if (pl->d_wall_mask) { cudaFree(pl->d_wall_mask); pl->d_wall_mask = nullptr; }
  if (pl->d_list) { cudaFree(pl->d_list); pl->d_list = nullptr; }
  if (pl->d_other_mask) { cudaFree(pl->d_other_mask); pl->d_other_mask = nullptr; }
  CUDA_OK(cudaMalloc((void **)&pl->d_other_mask, (size_t)words * 4));
  CUDA_OK(cudaMemcpy(pl->d_other_mask, om.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc((void **)&pl->d_fluid_mask, (size_t)words * 4));
  CUDA_OK(cudaMalloc((void **)&pl->d_wall_mask, (size_t)words * 4));
  CUDA_OK(cudaMalloc((void **)&pl->d_list, (list.size() + 1) * 4));
  CUDA_OK(cudaMemcpy(pl->d_fluid_mask, fm.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl->d_wall_mask, wm.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
  if (!list.empty()) CUDA_OK(cudaMemcpy(pl->d_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice));
  // segment list: worth it when a good part of the mask words of the real layers holds no fluid at all
  if (pl->d_seg) { cudaFree(pl->d_seg); pl->d_seg = nullptr; }
  pl->seg_off.clear();
  if (slab % 32 == 0 && d.DS < (1ll << 32) && !getenv("SLF_NO_SEGMENTS")) {
    std::vector<uint32_t> seg;
    std::vector<long long> off(outer + 1, 0);
    const long long wps = slab / 32;
    for (int o = 0; o < outer; o++) {
      off[o] = (long long)seg.size();
      for (long long w = wps * o; w < wps * (o + 1); w++) if (fm[(size_t)w]) seg.push_back((uint32_t)w);
    }
    off[outer] = (long long)seg.size();
    const long long real_words = wps * (outer > 2 ? outer - 2 : outer);
    if (!seg.empty() && (double)seg.size() < 0.85 * (double)real_words) {
      CUDA_OK(cudaMalloc((void **)&pl->d_seg, seg.size() * 4));
      CUDA_OK(cudaMemcpy(pl->d_seg, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice));
      pl->seg_off = off;
    }
  }
  pl->bound_map = map;
  return 0;
}

}  // extern "C"

extern "C" {

int slf_collide_and_propagate(SlfPlan *pl, uint64_t map, uint64_t dist_in, uint64_t dist_out, uint64_t rho,
                              uint64_t vx, uint64_t vy, uint64_t vz, int options, uint32_t iteration,
                              int z_begin, int z_end, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  const DevParams &d = pl->dp;
  const int outer = pl->D == 3 ? d.nz : d.ny;
  if (z_end < 0) z_end = outer;
  if (z_begin < 0 || z_end > outer || z_begin > z_end) return fail(SLF_EINVAL, "bad slab range");
  const int gi0 = z_begin, gi1 = z_end;   // slab range; converted to node / list ranges at launch
  int mode;
  if (pl->hp.access_pattern == SLF_AB) {
    mode = MODE_AB;
    if (dist_in == dist_out) return fail(SLF_EINVAL, "AB access pattern needs two distinct lattices");
  } else {
    mode = (iteration & 1) ? MODE_AA_ODD : MODE_AA_EVEN;
    if (dist_in != dist_out) return fail(SLF_EINVAL, "AA access pattern works in place (dist_in == dist_out)");
  }
  const bool save = (options & SLF_OPTION_SAVE_MACRO_FIELDS) != 0;
  if (save && (!rho || !vx || !vy || (pl->D == 3 && !vz))) return fail(SLF_EINVAL, "macro fields requested but null");
  const uint32_t *m = (const uint32_t *)map;
  const float *din = (const float *)dist_in;
  float *dout = (float *)dist_out, *r = (float *)rho, *x = (float *)vx, *y = (float *)vy, *z = (float *)vz;
  cudaStream_t s = (cudaStream_t)stream;
  const int model = pl->hp.model;
  switch (pl->hp.lattice) {
    case SLF_D2Q9: return slf_collide_d2q9(pl, model, mode, save, m, din, dout, r, x, y, z, gi0, gi1, s);
    case SLF_D3Q19:
      return model == SLF_MRT ? slf_collide_d3q19_mrt(pl, mode, save, m, din, dout, r, x, y, z, gi0, gi1, s)
                              : slf_collide_d3q19_bgk(pl, mode, save, m, din, dout, r, x, y, z, gi0, gi1, s);
    case SLF_D3Q27: return slf_collide_d3q27(pl, mode, save, m, din, dout, r, x, y, z, gi0, gi1, s);
  }
  return fail(SLF_EINVAL, "bad lattice");
}


int slf_set_initial_conditions(SlfPlan *pl, uint64_t dist, uint64_t vx, uint64_t vy, uint64_t vz, uint64_t rho,
                               uint64_t map, uint64_t stream) {
  (void)map;
  if (!pl) return fail(SLF_EINVAL, "null plan");
  const int bs = 256;
  const unsigned nb = (unsigned)((pl->dp.DS + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
  SLF_LATTICE_SWITCH(pl, (set_initial_conditions<L><<<nb, bs, 0, s>>>(pl->dp, (float *)dist, (const float *)vx,
                                                                      (const float *)vy, (const float *)vz,
                                                                      (const float *)rho)));
  LAUNCH_OK("set_initial_conditions");
  return 0;
}

int slf_prepare_macro_fields(SlfPlan *pl, uint64_t map, uint64_t dist, uint64_t rho, int options,
                             uint32_t iteration, uint64_t stream) {
  (void)options;
  if (!pl) return fail(SLF_EINVAL, "null plan");
  const int bs = 256;
  const unsigned nb = (unsigned)((pl->dp.DS + bs - 1) / bs);
  cudaStream_t s = (cudaStream_t)stream;
  const bool odd = pl->hp.access_pattern == SLF_AA && (iteration & 1);
  if (odd) {
    SLF_LATTICE_SWITCH(pl, (prepare_macro_fields<L, MODE_AA_ODD><<<nb, bs, 0, s>>>(
                               pl->dp, (const uint32_t *)map, (const float *)dist, (float *)rho)));
  } else {
    SLF_LATTICE_SWITCH(pl, (prepare_macro_fields<L, MODE_AB><<<nb, bs, 0, s>>>(
                               pl->dp, (const uint32_t *)map, (const float *)dist, (float *)rho)));
  }
  LAUNCH_OK("prepare_macro_fields");
  return 0;
}

int slf_apply_periodic_boundary_conditions(SlfPlan *pl, uint64_t dist, int axis, int with_swap, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (axis == 2 && pl->D == 2) return 0;   // the reference prepares a z kernel in 2-D and never runs it
  if (axis < 0 || axis >= pl->D) return fail(SLF_EINVAL, "bad axis");
  const DevParams &d = pl->dp;
  const int lim[3] = {d.nx, d.ny, d.nz};
  int o0, o1;
  if (pl->D == 2) { o0 = 1 - axis; o1 = 2; } else { o0 = axis == 0 ? 1 : 0; o1 = axis == 2 ? 1 : 2; }
  const int n1 = lim[o0], n2 = pl->D == 3 ? lim[o1] : 1;
  dim3 grid((n1 + 127) / 128, n2);
  cudaStream_t s = (cudaStream_t)stream;
  const int *per = pl->hp.periodic;
  if (with_swap) {
    SLF_LATTICE_SWITCH(pl, (apply_pbc<L, true><<<grid, 128, 0, s>>>(d, (float *)dist, axis, per[0], per[1], per[2], pl->link_lo, pl->link_hi)));
  } else {
    SLF_LATTICE_SWITCH(pl, (apply_pbc<L, false><<<grid, 128, 0, s>>>(d, (float *)dist, axis, per[0], per[1], per[2], pl->link_lo, pl->link_hi)));
  }
  LAUNCH_OK("apply_pbc");
  return 0;
}

static int face_ndists(const SlfPlan *pl, int face) {
  // |{i: e_i . n >= 1}|
  if (pl->hp.lattice == SLF_D2Q9) return 3;
  if (pl->hp.lattice == SLF_D3Q19) return 5;
  (void)face;
  return 9;
}

static int xfer(SlfPlan *pl, uint64_t dist, int face, int base_gx, int base_other, int max_lx, int max_other,
                uint64_t buffer, int with_swap, bool distribute, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (face < 2 || face >= 2 * pl->D) return fail(SLF_EINVAL, "continuous transfers are for y/z faces");
  if (pl->D == 2) max_other = 1;
  if (max_lx <= 0 || max_other <= 0) return 0;
  dim3 grid((max_lx + 127) / 128, max_other, face_ndists(pl, face));
  cudaStream_t s = (cudaStream_t)stream;
  if (distribute) {
    SLF_LATTICE_SWITCH(pl, (xfer_continuous<L, true><<<grid, 128, 0, s>>>(pl->dp, (float *)dist, face, base_gx, base_other,
                                                                       max_lx, max_other, (float *)buffer, with_swap)));
  } else {
    SLF_LATTICE_SWITCH(pl, (xfer_continuous<L, false><<<grid, 128, 0, s>>>(pl->dp, (float *)dist, face, base_gx, base_other,
                                                                        max_lx, max_other, (float *)buffer, with_swap)));
  }
  LAUNCH_OK("xfer_continuous");
  return 0;
}

int slf_collect_continuous_data(SlfPlan *pl, uint64_t dist, int face, int base_gx, int base_other, int max_lx,
                                int max_other, uint64_t buffer, int with_swap, uint64_t stream) {
  return xfer(pl, dist, face, base_gx, base_other, max_lx, max_other, buffer, with_swap, false, stream);
}
int slf_distribute_continuous_data(SlfPlan *pl, uint64_t dist, int face, int base_gx, int base_other, int max_lx,
                                   int max_other, uint64_t buffer, int with_swap, uint64_t stream) {
  return xfer(pl, dist, face, base_gx, base_other, max_lx, max_other, buffer, with_swap, true, stream);
}

int slf_collect_sparse_data(SlfPlan *pl, uint64_t idx_array, uint64_t dist, uint64_t buffer, int max_idx,
                            uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (max_idx <= 0) return 0;
  xfer_sparse<false><<<(max_idx + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      (const uint32_t *)idx_array, (float *)dist, (float *)buffer, max_idx,
      (unsigned long long)pl->dp.DS * pl->Q);
  LAUNCH_OK("collect_sparse");
  return 0;
}
int slf_distribute_sparse_data(SlfPlan *pl, uint64_t idx_array, uint64_t dist, uint64_t buffer, int max_idx,
                               uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (max_idx <= 0) return 0;
  xfer_sparse<true><<<(max_idx + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      (const uint32_t *)idx_array, (float *)dist, (float *)buffer, max_idx,
      (unsigned long long)pl->dp.DS * pl->Q);
  LAUNCH_OK("distribute_sparse");
  return 0;
}



// ---- Collect/DistributeContinuousMacroData ---------------------------------------------------
static int xfer_macro_launch(SlfPlan *pl, uint64_t field, int face, int base_gx, int base_other, int max_lx,
                             int max_other, uint64_t buffer, bool distribute, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (!field || !buffer) return fail(SLF_EINVAL, "null buffer");
  const DevParams &d = pl->dp;
  if (pl->D == 2) {
    max_other = 1;
    if (base_other < 0 || base_other >= d.ny) return fail(SLF_EINVAL, "row outside the lattice");
  } else {
    if (face < 2 || face > 5) return fail(SLF_EINVAL, "macro data moves through y and z faces (x faces use the sparse kernels)");
    const int lim_o = (face >> 1) == 1 ? d.nz : d.ny;
    if (base_other < 0 || max_other < 0 || base_other + max_other > lim_o) return fail(SLF_EINVAL, "rows outside the lattice");
  }
  if (base_gx < 0 || max_lx < 0 || base_gx + max_lx > d.nx) return fail(SLF_EINVAL, "columns outside the lattice");
  if (max_lx == 0 || max_other == 0) return 0;
  dim3 grid((max_lx + 127) / 128, max_other);
  cudaStream_t s = (cudaStream_t)stream;
  if (distribute)
    xfer_macro<true><<<grid, 128, 0, s>>>(d, (float *)field, pl->D, face, base_gx, base_other, max_lx, max_other, (float *)buffer);
  else
    xfer_macro<false><<<grid, 128, 0, s>>>(d, (float *)field, pl->D, face, base_gx, base_other, max_lx, max_other, (float *)buffer);
  LAUNCH_OK("xfer_macro");
  return 0;
}

int slf_collect_continuous_macro_data(SlfPlan *pl, uint64_t field, int face, int base_gx, int base_other, int max_lx,
                                      int max_other, uint64_t buffer, uint64_t stream) {
  return xfer_macro_launch(pl, field, face, base_gx, base_other, max_lx, max_other, buffer, false, stream);
}

int slf_distribute_continuous_macro_data(SlfPlan *pl, uint64_t field, int face, int base_gx, int base_other, int max_lx,
                                         int max_other, uint64_t buffer, uint64_t stream) {
  return xfer_macro_launch(pl, field, face, base_gx, base_other, max_lx, max_other, buffer, true, stream);
}

// ---- output fast path ---------------------------------------------------------------------
int slf_pack_field(SlfPlan *pl, uint64_t map, uint64_t field, uint64_t out, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (!map || !field || !out) return fail(SLF_EINVAL, "null buffer");
  const DevParams &d = pl->dp;
  const int rx = d.nx - 2, ry = d.ny - 2, rz = pl->D == 3 ? d.nz - 2 : 1;
  if (rx <= 0 || ry <= 0 || rz <= 0) return 0;
  dim3 grid((rx + 255) / 256, ry, rz);
  pack_field<<<grid, 256, 0, (cudaStream_t)stream>>>(d, (const uint32_t *)map, (const float *)field, (float *)out,
                                                      pl->D);
  LAUNCH_OK("pack_field");
  return 0;
}

// ---- multi-step driver -------------------------------------------------------------------
// The reference pays ~10 Python calls + 1-2 pycuda launches per time step
// (subdomain_runner.py:1673-1748); on small lattices (ldc_2d 256^2: ~66 k nodes, a few
// microseconds of device work) that host overhead is the whole cost.
// slf_run_steps issues whole steps from C and replays pairs of steps (one even + one odd: all
// kernel arguments depend on the iteration only through its parity) as a CUDA graph.
static int one_step(SlfPlan *pl, uint64_t map, uint64_t dist_a, uint64_t dist_b, uint64_t rho, uint64_t vx,
                    uint64_t vy, uint64_t vz, uint32_t it, int save, uint64_t stream) {
  const bool aa = pl->hp.access_pattern == SLF_AA;
  uint64_t din = dist_a, dout = dist_a;
  if (!aa) { din = (it & 1) ? dist_b : dist_a; dout = (it & 1) ? dist_a : dist_b; }
  int rc = slf_collide_and_propagate(pl, map, din, dout, rho, vx, vy, vz, save ? SLF_OPTION_SAVE_MACRO_FIELDS : 0,
                                     it, 0, -1, stream);
  if (rc) return rc;
  const int swap = aa && (it & 1) == 0;
  for (int axis = 0; axis < pl->D; axis++)
    if (pl->hp.periodic[axis]) {
      rc = slf_apply_periodic_boundary_conditions(pl, dout, axis, swap, stream);
      if (rc) return rc;
    }
  return 0;
}

int slf_run_steps(SlfPlan *pl, uint64_t map, uint64_t dist_a, uint64_t dist_b, uint64_t rho, uint64_t vx,
                  uint64_t vy, uint64_t vz, uint32_t iteration, int nsteps, int save_last, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (nsteps <= 0) return 0;
  if (pl->hp.access_pattern == SLF_AB && !dist_b) return fail(SLF_EINVAL, "AB needs two lattices");
  cudaStream_t s = (cudaStream_t)stream;
  // Small lattices without periodic axes: all steps in one cooperative launch (grid barrier between steps).
  if (pl->hp.access_pattern == SLF_AA && nsteps >= 2 && s != 0 && !pl->hp.periodic[0] && !pl->hp.periodic[1] &&
      !pl->hp.periodic[2] && pl->hp.type_id[SLF_NT_WALL_TMS] == SLF_NT_ABSENT && !getenv("SLF_NO_MULTISTEP")) {
    const uint32_t *m = (const uint32_t *)map;
    float *d = (float *)dist_a, *r = (float *)rho, *x = (float *)vx, *y = (float *)vy, *z = (float *)vz;
    int left = nsteps;
    uint32_t it = iteration;
    int rc = 1;
    while (left > 0) {
      const int chunk = left > (1 << 30) ? (1 << 30) : left;
      const int sl = save_last && chunk == left;
      switch (pl->hp.lattice) {
        case SLF_D2Q9: rc = slf_multistep_d2q9(pl, pl->hp.model, m, d, r, x, y, z, it, chunk, sl, s); break;
        case SLF_D3Q19:
          rc = pl->hp.model == SLF_MRT ? slf_multistep_d3q19_mrt(pl, m, d, r, x, y, z, it, chunk, sl, s)
                                       : slf_multistep_d3q19_bgk(pl, m, d, r, x, y, z, it, chunk, sl, s);
          break;
        default: rc = slf_multistep_d3q27(pl, m, d, r, x, y, z, it, chunk, sl, s); break;
      }
      if (rc != 0) break;
      left -= chunk; it += chunk;
    }
    if (left == 0) return 0;
    if (rc != 1) cudaGetLastError();                      // cooperative launch refused: step kernel by kernel
    iteration = it; nsteps = left;                        // (chunks already issued stay issued)
  }
  int done = 0;
  const int plain = nsteps - (save_last ? 1 : 0);       // steps that may be replayed from a graph
  if (plain >= 4 && s != 0) {
    const int parity = iteration & 1;
    std::lock_guard<std::mutex> lock(pl->graph_mutex);
    StepGraph *g = nullptr;
    for (auto &e : pl->graphs)
      if (e.key.map == map && e.key.a == dist_a && e.key.b == dist_b && e.key.rho == rho && e.key.parity == parity)
        g = &e;
    if (!g) {
      const uint64_t before = slf_launch_count();
      cudaGraph_t graph;
      CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rc = one_step(pl, map, dist_a, dist_b, rho, vx, vy, vz, iteration, 0, stream);
      if (!rc) rc = one_step(pl, map, dist_a, dist_b, rho, vx, vy, vz, iteration + 1, 0, stream);
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc) { if (ce == cudaSuccess) cudaGraphDestroy(graph); return rc; }
      CUDA_OK(ce);
      StepGraph ng;
      ng.key = {map, dist_a, dist_b, rho, parity};
      ng.kernels = slf_launch_count() - before;          // kernels captured (counted, not yet run)
      CUDA_OK(cudaGraphInstantiate(&ng.exec, graph, 0));
      cudaGraphDestroy(graph);
      pl->graphs.push_back(ng);
      g = &pl->graphs.back();
      for (uint64_t k = 0; k < ng.kernels; k++) g_launches.fetch_sub(1, std::memory_order_relaxed);
    }
    const int pairs = plain / 2;
    for (int k = 0; k < pairs; k++) {
      CUDA_OK(cudaGraphLaunch(g->exec, s));
      g_launches.fetch_add(g->kernels, std::memory_order_relaxed);
    }
    done = 2 * pairs;
  }
  for (int k = done; k < nsteps; k++) {
    int rc = one_step(pl, map, dist_a, dist_b, rho, vx, vy, vz, iteration + k, save_last && k == nsteps - 1, stream);
    if (rc) return rc;
  }
  return 0;
}

// ---- generic stream capture: the host-side runner records whatever its step() issues (collide slabs,
// semaphore waits, plane pushes, fork/join events across its two streams) once per step parity and
// replays it; every kernel argument depends on the iteration only through its parity, and the
// semaphore kernels read their targets from device memory. ------------------------------------
struct SlfGraph { cudaGraphExec_t exec; uint64_t kernels; };
static thread_local uint64_t t_capture_base = 0;

int slf_graph_begin_capture(uint64_t stream) {
  t_capture_base = slf_launch_count();
  CUDA_OK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return 0;
}

int slf_graph_end_capture(uint64_t stream, uint64_t *graph_exec) {
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture((cudaStream_t)stream, &graph);
  const uint64_t kernels = slf_launch_count() - t_capture_base;
  for (uint64_t k = 0; k < kernels; k++) g_launches.fetch_sub(1, std::memory_order_relaxed);   // recorded, not run
  if (ce != cudaSuccess) { cudaGetLastError(); return fail((int)ce, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); }
  SlfGraph *g = new SlfGraph();
  g->kernels = kernels;
  // The recorded step forks onto a high-priority stream (slabs + pushes) beside the interior kernel.  Without this
  // flag every node runs at the priority of the stream the graph is launched into, the slab kernels' blocks queue
  // behind the interior kernel's pending blocks, and the chain wait -> slabs -> pushes is only dispatched at the end
  // of the step instead of beside it.
  const unsigned long long flags = getenv("SLF_GRAPH_NO_NODE_PRIORITY") ? 0ull : (unsigned long long)cudaGraphInstantiateFlagUseNodePriority;
  ce = cudaGraphInstantiateWithFlags(&g->exec, graph, flags);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { delete g; return fail((int)ce, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
  *graph_exec = (uint64_t)g;
  return 0;
}

int slf_graph_launch(uint64_t graph_exec, uint64_t stream) {
  SlfGraph *g = (SlfGraph *)graph_exec;
  if (!g) return fail(SLF_EINVAL, "null graph");
  CUDA_OK(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  g_launches.fetch_add(g->kernels, std::memory_order_relaxed);
  return 0;
}

int slf_graph_destroy(uint64_t graph_exec) {
  SlfGraph *g = (SlfGraph *)graph_exec;
  if (!g) return 0;
  cudaGraphExecDestroy(g->exec);
  delete g;
  return 0;
}

// ---- peer access ---------------------------------------------------------------------------
int slf_ipc_get_handle(uint64_t dptr, void *handle64) {
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, (void *)dptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int slf_ipc_open_handle(const void *handle64, uint64_t *dptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr = (uint64_t)p;
  return 0;
}
int slf_ipc_close_handle(uint64_t dptr) { CUDA_OK(cudaIpcCloseMemHandle((void *)dptr)); return 0; }
int slf_enable_peer_access(int peer_device) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  CUDA_OK(e);
  return 0;
}

}  // extern "C"

#include "slf_halo.cuh"
