// slf_halo.cuh -- ghost-layer exchange for subdomains split along the outermost axis
// (conn_axis = z in 3-D, y in 2-D), device to device over NVLink peer mappings.
//
// Replaces the reference's collect kernel -> D2H -> ZeroMQ -> H2D -> distribute kernel chain
// (subdomain_runner.py:1064-1139, kernel_utils.mako:526-794, connector.py:73-117).  With the
// SoA layout a face message is |dists| whole planes that are already contiguous in the
// lattice array, so "pack" and "unpack" disappear: one kernel copies the planes straight
// from this GPU's lattice into the neighbour's lattice through the peer pointer and, when its
// last block is done, raises the neighbour's `ready` counter (copy + signal are ONE launch).
// Layer/slot selection follows the reference exactly:
//   after an AB or AA-odd step : ghost layer, natural slots  -> neighbour's first real layer
//   after an AA-even step      : first real layer, opposite slots -> neighbour's ghost layer
// (lat_linear / lat_linear_macro / lat_linear_dist / lat_linear_with_swap,
//  subdomain_runner.py:488-510).
//
// Ordering protocol (one semaphore block of SLF_SEM_WORDS uint32 per subdomain, zeroed once):
//   ready[ch][side]  incremented by the neighbour on that side after each of its pushes on
//                    channel ch (0: populations, 1: macroscopic fields);
//   seq[ch]          number of steps this subdomain has waited for on channel ch.
// Step s of a subdomain needs the neighbours' population pushes of their steps 0..s-1, i.e.
// ready >= seq (fields, exchanged in mid-step: those of step s too, ready >= seq + 1 = `ahead`);
// the wait kernel spins on that (acquire loads at system scope), then ++seq.  Both counters only
// ever grow and are compared as differences, so a run may be re-initialised, restored from a
// checkpoint or continued any number of times without resetting anything -- all that is needed
// around a re-initialisation is a host-side barrier (every rank idle) before and after it.
// A wait that exceeds the time limit (a dead or diverged peer) or sees the abort word set
// records an error in status[0] and RETURNS: the stream drains and the host reports the error
// at its next synchronisation instead of hanging in an uncancellable kernel.
#pragma once
#include <cstdlib>

struct HaloArgs {
  const float *src[9];
  float *dst[9];
  int n;                      // number of planes
  long long plane;            // floats per plane (multiple of 4)
  unsigned int *signal;       // neighbour's ready word (peer pointer) or null
  unsigned int *done;         // local block counter for the last-block-signals pattern
};

__global__ void __launch_bounds__(256) halo_push_kernel(const HaloArgs a) {
  const long long n4 = a.plane >> 2;
  const int k = blockIdx.y;
  const float4 *__restrict__ s = reinterpret_cast<const float4 *>(a.src[k]);
  float4 *__restrict__ d = reinterpret_cast<float4 *>(a.dst[k]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    d[i] = s[i];
  __threadfence_system();
  if (a.signal) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int total = gridDim.x * gridDim.y;
      const unsigned int prev = atomicAdd(a.done, 1u);
      if (prev == total - 1u) {
        *a.done = 0u;                     // ready for the next launch (launches on one stream are ordered)
        __threadfence_system();
        atomicAdd_system(a.signal, 1u);
      }
    }
  }
}

__global__ void sem_signal_kernel(unsigned int *sem) {
  __threadfence_system();
  atomicAdd_system(sem, 1u);
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// status[0]: error (0 ok, 1 timeout, 2 aborted), status[1]: abort request (host-settable)
__global__ void sem_wait_kernel(const unsigned int *ready_a, const unsigned int *ready_b, unsigned int *seq,
                                unsigned int *status, unsigned int ahead, unsigned long long timeout_ns) {
  const unsigned int want = *seq + ahead;
  const unsigned int *r[2] = {ready_a, ready_b};
  const unsigned long long t0 = global_timer_ns();
  unsigned int err = 0;
  for (int k = 0; k < 2 && !err; k++) {
    if (!r[k]) continue;
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(r[k]) - want) < 0) {
      __nanosleep(64);
      if ((++spins & 63u) == 0) {
        if (ld_acquire_sys(status + 1)) { err = 2; break; }
        if (global_timer_ns() - t0 > timeout_ns) { err = 1; break; }
      }
    }
  }
  if (err) atomicCAS(status, 0u, err);
  __threadfence_system();
  *seq = want + 1u - ahead;
}

// ---- any face (subdomains split along x and / or y as well: regtest/subdomains/3d_ldc.py block layouts) ------
// The reference connects every pair of subdomains that share a face, an edge or a corner and ships each
// population set separately (subdomain_connection.py:348-407, sparse index lists for x faces, edges and
// corners).  Here the exchange is dimension ordered -- all x faces, then all y faces, then z -- and edge / corner
// data reach their diagonal destination in two or three hops, riding in the ghost ring of the face messages.
// Per face and crossing population i the copied window is chosen such that (a) everything a real node of the
// receiver will read and every entry that must be forwarded further is included, and (b) no two messages -- of this
// step or of neighbouring steps, from different senders, which nothing orders against each other -- ever write the
// same location:
//   after streaming steps (data sit where they were streamed to; the receiver's node reads its own slots):
//     along an EARLIER axis (already exchanged in this step): the real nodes 1 .. n - 2 (entries forwarded by the
//       earlier exchange sit in the first / last real column);
//     along a LATER axis b: the real nodes shifted by e_i[b] (the entries real nodes of this subdomain streamed
//       into, ghost-row entries -- to be forwarded by the later exchange -- included);
//   after in-place steps (with_swap: data sit un-streamed at their node; receiver node p reads q = p - e_i):
//     along an EARLIER axis b: q_b in 1 - e_i[b] .. n - 2 - e_i[b] (what some real node of the receiver reads; the
//       ghost-column entries delivered by the earlier exchange included);
//     along a LATER axis: the real nodes 1 .. n - 2 (entries to be forwarded by the later exchange sit in real rows).
// With the looser "whole extent along earlier axes" a late in-place message of one neighbour could overwrite an early
// streamed message of another in a ghost row (found on hardware with one rank per block, tools/gpu_round2_f.sh).
struct FaceArgs {
  const float *src[9];
  float *dst[9];
  int lo1[9], n1[9], lo2[9], n2[9];     // window along the two other axes (other1 < other2)
  long long s1, s2, d1, d2;             // strides of those axes in the source / destination array
  int n;
  unsigned int *signal, *done;
};

__global__ void __launch_bounds__(128) halo_push_face_kernel(const FaceArgs a) {
  const int k = blockIdx.z;
  const int i1 = blockIdx.x * blockDim.x + threadIdx.x, i2 = blockIdx.y;
  if (i1 < a.n1[k] && i2 < a.n2[k]) {
    const long long so = (long long)(a.lo1[k] + i1) * a.s1 + (long long)(a.lo2[k] + i2) * a.s2;
    const long long dn = (long long)(a.lo1[k] + i1) * a.d1 + (long long)(a.lo2[k] + i2) * a.d2;
    a.dst[k][dn] = a.src[k][so];
  }
  __threadfence_system();
  if (a.signal) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
      const unsigned int prev = atomicAdd(a.done, 1u);
      if (prev == total - 1u) {
        *a.done = 0u;
        __threadfence_system();
        atomicAdd_system(a.signal, 1u);
      }
    }
  }
}

static unsigned long long sem_timeout_ns() {
  static unsigned long long v = 0;
  if (!v) {
    const char *e = getenv("SLF_SEM_TIMEOUT_MS");
    const double ms = e ? atof(e) : 30000.0;
    v = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
  }
  return v;
}

extern "C" {

int slf_halo_push(SlfPlan *pl, uint64_t dist, int face, uint64_t peer_dist, int peer_lat_n, int peer_arr_n,
                  int with_swap, uint64_t signal_sem, uint64_t stream) {
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
  a.signal = (unsigned int *)signal_sem;
  a.done = nullptr;
  if (signal_sem) a.done = pl->d_push_done + hi;
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

// Window of the positions along the other axes that population e carries across a face of `axis` (dimension-ordered
// exchange: the axes before `axis` have been exchanged in this step, the ones after it have not):
//   after streaming steps: earlier axes 1 .. n-2,        later axes 1+e .. n-2+e;
//   after in-place steps : earlier axes 1-e .. n-2-e,    later axes 1 .. n-2.
// Along an EARLIER axis the end of the window on the side the population comes from holds data only if a neighbour
// delivered it there; on a side without a neighbour (domain boundary; bit 2*ax+side of `links` clear) that position is
// dropped -- its content never left a node (streaming) or belongs to the receiving subdomain's own boundary nodes
// (in-place: do-nothing and link-bounce nodes store into their ghost neighbours).
static inline void face_window(int D, const int *lat, const int *e, int axis, int with_swap, int links, int *lo, int *cnt) {
  int b = 0;
  for (int ax = 0; ax < D; ax++) {
    if (ax == axis) continue;
    int l, c = lat[ax] - 2;
    if (ax < axis) {
      l = with_swap ? 1 - e[ax] : 1;
      const int from_side = e[ax] > 0 ? 0 : 1;               // the side of `ax` the population entered through
      if (e[ax] != 0 && !((links >> (2 * ax + from_side)) & 1)) {
        c -= 1;
        if (with_swap) { if (e[ax] > 0) l = 1; else l = 2; }
        else if (e[ax] > 0) l = 2;
      }
    } else {
      l = with_swap ? 1 : 1 + e[ax];
    }
    lo[b] = l; cnt[b] = c < 0 ? 0 : c;
    b++;
  }
}

// The same windows for the tests (pure host code, no device needed): 1 if population i of the lattice crosses `face`.
int slf_halo_face_window(int lattice, const int *lat, int face, int with_swap, int links, int i, int *lo, int *cnt) {
  const int D = lattice == SLF_D2Q9 ? 2 : 3;
  const int axis = face >> 1, sgn = (face & 1) ? 1 : -1;
  if (!lat || !lo || !cnt || face < 0 || axis >= D) return fail(SLF_EINVAL, "bad arguments");
  int e[3];
  switch (lattice) {
    case SLF_D2Q9: if (i < 0 || i >= 9) return 0; e[0] = D2Q9::ex(i); e[1] = D2Q9::ey(i); e[2] = 0; break;
    case SLF_D3Q19: if (i < 0 || i >= 19) return 0; e[0] = D3Q19::ex(i); e[1] = D3Q19::ey(i); e[2] = D3Q19::ez(i); break;
    case SLF_D3Q27: if (i < 0 || i >= 27) return 0; e[0] = D3Q27::ex(i); e[1] = D3Q27::ey(i); e[2] = D3Q27::ez(i); break;
    default: return fail(SLF_EINVAL, "unknown lattice");
  }
  if (e[axis] * sgn < 1) return 0;
  lo[0] = lo[1] = 0; cnt[0] = cnt[1] = 1;
  face_window(D, lat, e, axis, with_swap, links, lo, cnt);
  return 1;
}

int slf_plan_set_face_links(SlfPlan *pl, int links) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  pl->face_links = links & 0x3f;
  return 0;
}

int slf_halo_push_face(SlfPlan *pl, uint64_t dist, int face, uint64_t peer_dist, const int *peer_lat, const int *peer_arr,
                       int with_swap, uint64_t signal_sem, uint64_t stream) {
  if (!pl) return fail(SLF_EINVAL, "null plan");
  if (!peer_lat || !peer_arr) return fail(SLF_EINVAL, "null peer geometry");
  const DevParams &d = pl->dp;
  const int D = pl->D, axis = face >> 1, hi = face & 1;
  if (face < 0 || axis >= D) return fail(SLF_EINVAL, "bad face");
  const int lat[3] = {d.nx, d.ny, d.nz};
  const long long str[3] = {1, d.sy, d.sz};
  const long long pstr[3] = {1, (long long)peer_arr[0], (long long)peer_arr[0] * peer_arr[1]};
  const long long peer_DS = (long long)peer_arr[0] * peer_arr[1] * (D == 3 ? peer_arr[2] : 1);
  int o1, o2;                                    // the other axes, o1 < o2 (2-D: o2 unused)
  if (D == 2) { o1 = 1 - axis; o2 = 2; } else { o1 = axis == 0 ? 1 : 0; o2 = axis == 2 ? 1 : 2; }
  if (lat[o1] != peer_lat[o1] || (D == 3 && lat[o2] != peer_lat[o2]))
    return fail(SLF_EUNSUPPORTED, "neighbouring subdomains must share the whole face (equal cross-sections)");
  const int n = lat[axis], pn = peer_lat[axis];
  const int src_layer = with_swap ? (hi ? n - 2 : 1) : (hi ? n - 1 : 0);
  const int dst_layer = with_swap ? (hi ? 0 : pn - 1) : (hi ? 1 : pn - 2);
  FaceArgs a;
  a.n = 0; a.s1 = str[o1]; a.d1 = pstr[o1];
  a.s2 = D == 3 ? str[o2] : 0; a.d2 = D == 3 ? pstr[o2] : 0;
  a.signal = (unsigned int *)signal_sem;
  a.done = signal_sem ? pl->d_push_done + hi : nullptr;
  const int sgn = hi ? 1 : -1;
  int max1 = 0, max2 = 1;
  for (int i = 0; i < pl->Q; i++) {
    int e[3], opp;
    switch (pl->hp.lattice) {
      case SLF_D2Q9: e[0] = D2Q9::ex(i); e[1] = D2Q9::ey(i); e[2] = 0; opp = D2Q9::opp(i); break;
      case SLF_D3Q19: e[0] = D3Q19::ex(i); e[1] = D3Q19::ey(i); e[2] = D3Q19::ez(i); opp = D3Q19::opp(i); break;
      default: e[0] = D3Q27::ex(i); e[1] = D3Q27::ey(i); e[2] = D3Q27::ez(i); opp = D3Q27::opp(i); break;
    }
    if (e[axis] * sgn < 1) continue;
    const int slot = with_swap ? opp : i;
    int lo[2] = {0, 0}, cnt[2] = {1, 1};
    face_window(D, lat, e, axis, with_swap, pl->face_links, lo, cnt);
    a.src[a.n] = (const float *)dist + d.DS * slot + str[axis] * src_layer;
    a.dst[a.n] = (float *)peer_dist + peer_DS * slot + pstr[axis] * dst_layer;
    a.lo1[a.n] = lo[0]; a.n1[a.n] = cnt[0]; a.lo2[a.n] = lo[1]; a.n2[a.n] = cnt[1];
    if (cnt[0] > max1) max1 = cnt[0];
    if (cnt[1] > max2) max2 = cnt[1];
    a.n++;
  }
  if (a.n == 0 || max1 <= 0) return 0;
  dim3 grid((max1 + 127) / 128, max2, a.n);
  halo_push_face_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  LAUNCH_OK("halo_push_face");
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
  a.n = 1; a.plane = plane; a.signal = nullptr; a.done = nullptr;
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

int slf_sem_wait(uint64_t ready_a, uint64_t ready_b, uint64_t seq, uint64_t status, int ahead, uint64_t stream) {
  if (!seq || !status) return fail(SLF_EINVAL, "null semaphore block");
  if (!ready_a && !ready_b) return 0;
  sem_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned int *)ready_a, (const unsigned int *)ready_b,
                                                     (unsigned int *)seq, (unsigned int *)status, ahead ? 1u : 0u, sem_timeout_ns());
  LAUNCH_OK("sem_wait");
  return 0;
}

// Blocking 4-byte reads/writes of the status words on a private stream (the compute streams may
// be busy -- or stuck in a wait -- when the host looks).
static cudaStream_t sem_side_stream() {
  static cudaStream_t s = nullptr;
  if (!s) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  return s;
}

int slf_sem_status(uint64_t status, uint32_t *err) {
  if (!status || !err) return fail(SLF_EINVAL, "null argument");
  cudaStream_t s = sem_side_stream();
  CUDA_OK(cudaMemcpyAsync(err, (const void *)status, 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int slf_sem_abort(uint64_t status) {
  if (!status) return fail(SLF_EINVAL, "null argument");
  static const uint32_t one = 1;
  cudaStream_t s = sem_side_stream();
  CUDA_OK(cudaMemcpyAsync((void *)(status + 4), &one, 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"
