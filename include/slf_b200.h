/*
 * slf_b200.h -- C ABI of libslf_b200.so, the B200 (sm_100a) implementation of the
 * sailfish lattice-Boltzmann hot path.
 *
 * The reference (sailfish-team/sailfish) has no FFI: its device interface is the
 * "backend" duck type (sailfish/backend_cuda.py:34-332, pycuda driver calls) plus
 * the kernel-name/argument contract of the Mako-generated module
 * (sailfish/templates/models/lb_single_fluid.mako, sailfish/templates/kernel_utils.mako).
 * Every entry point below names the reference interface it replaces.  The Python
 * binding is sailfish_b200/backend_b200.py (ctypes); INTEGRATION.md shows how the
 * reference selects it with --backends=b200.
 *
 * Conventions
 *   - plain C types only; device pointers travel as uint64_t, streams/events as
 *     opaque uint64_t handles (0 = default stream);
 *   - every function returns 0 on success, a negative SLF_E* code or a positive
 *     cudaError_t otherwise; slf_last_error() gives the message (thread local);
 *   - all launches are asynchronous on the given stream;
 *   - no CPU fallback exists: without a CUDA device slf_init() fails.
 *
 * Memory layout (identical to the reference, subdomain_runner.py:359-379,
 * templates/kernel_common.mako:459-461, geo_helpers.mako:163-176):
 *   gi        = gx + arr_nx * (gy + arr_ny * gz)      ghost-inclusive, x padded
 *   dist[i]   = base[i * DIST_SIZE + gi],  DIST_SIZE = arr_nx*arr_ny*arr_nz
 *   fields    = one float per gi;  node map = one uint32 per gi
 *   node code = orientation | scratch | param | type  (geo_encoder.py:365-382)
 */
#ifndef SLF_B200_H_
#define SLF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLF_ABI_VERSION 6

/* error codes (negative; positive values are cudaError_t) */
#define SLF_EINVAL      (-1)   /* bad argument / unsupported configuration          */
#define SLF_ENODEVICE   (-2)   /* no CUDA device: there is no CPU fallback           */
#define SLF_EUNSUPPORTED (-3)  /* feature of the reference that is out of scope      */
#define SLF_ENCCL       (-4)

/* lattices: sailfish/sym.py:61 (D2Q9), :312 (D3Q19), :409 (D3Q27) */
enum { SLF_D2Q9 = 0, SLF_D3Q19 = 1, SLF_D3Q27 = 2 };
/* collision models: templates/relaxation.mako:99 (BGK), relaxation_mrt.mako:31 (MRT) */
enum { SLF_BGK = 0, SLF_MRT = 1 };
/* --access_pattern, lb_base.py:61-65 */
enum { SLF_AB = 0, SLF_AA = 1 };
/* ctx['simtype']: single fluid (lb_single.py:60) or binary Shan-Chen (lb_binary.py:422-427) */
enum { SLF_SIM_LBM = 0, SLF_SIM_SHAN_CHEN = 1 };
/* --sc_potential, sym.py:896-908 */
enum { SLF_SC_LINEAR = 0, SLF_SC_CLASSIC = 1 };

/* Node classes understood by the kernels (sailfish/node_type.py).  type_id[] in
 * SlfParams holds the per-run remapped id (geo_encoder.py:83-91) of each class,
 * or SLF_NT_ABSENT when the class does not occur in the subdomain. */
enum {
  SLF_NT_FLUID = 0,             /* node_type.py:87  _NTFluid              */
  SLF_NT_GHOST,                 /* :94  _NTGhost (excluded)               */
  SLF_NT_UNUSED,                /* :99  _NTUnused (excluded)              */
  SLF_NT_PROPAGATION_ONLY,      /* :104 _NTPropagationOnly                */
  SLF_NT_FULLBB_WALL,           /* :148 NTFullBBWall                      */
  SLF_NT_HALFBB_WALL,           /* :117 NTHalfBBWall                      */
  SLF_NT_EQUILIBRIUM_VELOCITY,  /* :246 NTEquilibriumVelocity             */
  SLF_NT_EQUILIBRIUM_DENSITY,   /* :200 NTEquilibriumDensity              */
  SLF_NT_REGULARIZED_VELOCITY,  /* :269 NTRegularizedVelocity             */
  SLF_NT_REGULARIZED_DENSITY,   /* :211 NTRegularizedDensity              */
  SLF_NT_DO_NOTHING,            /* :296 NTDoNothing                       */
  SLF_NT_ZOUHE_VELOCITY,        /* :256 NTZouHeVelocity                   */
  SLF_NT_ZOUHE_DENSITY,         /* :229 NTZouHeDensity                    */
  SLF_NT_COPY,                  /* :310 NTCopy (AB access pattern only)   */
  SLF_NT_YU_OUTFLOW,            /* :336 NTYuOutflow (AB only)             */
  SLF_NT_WALL_TMS,              /* :171 NTWallTMS                         */
  SLF_NT_SLIP,                  /* :402 NTSlip (orientation from the user) */
  SLF_NT_COUNT
};
#define SLF_NT_ABSENT 0xfffffffeu   /* geo_encoder.py:44-50 */

/* CollideAndPropagate `options` bits, templates/kernel_common.mako:507-508 */
#define SLF_OPTION_SAVE_MACRO_FIELDS 1
#define SLF_OPTION_BULK              2
#define SLF_OPTION_DENSITIES_ONLY    4   /* library extension: slf_sc_prepare_macro_fields skips the velocity */

/* Faces, sailfish/subdomain.py SubdomainSpec.{X_LOW..Z_HIGH} */
enum { SLF_X_LOW = 0, SLF_X_HIGH = 1, SLF_Y_LOW = 2, SLF_Y_HIGH = 3, SLF_Z_LOW = 4, SLF_Z_HIGH = 5 };

/*
 * Everything the reference bakes into the generated source as literals
 * (codegen.py:195-204 _build_context -> lb_base.py:120-137, lb_single.py:52-70,
 * subdomain_runner.py:182-248, geo_encoder.py:341-363), as one POD.
 */
typedef struct SlfParams {
  uint32_t struct_bytes;        /* sizeof(SlfParams): ABI check                         */
  int32_t  lattice;             /* SLF_D2Q9 | SLF_D3Q19 | SLF_D3Q27   (ctx['grid'])      */
  int32_t  model;               /* SLF_BGK | SLF_MRT                  (ctx['model'])     */
  int32_t  access_pattern;      /* SLF_AB | SLF_AA                                      */
  int32_t  lat_nx, lat_ny, lat_nz;   /* ghost-inclusive lattice (lat_nz = 1 in 2-D)     */
  int32_t  arr_nx, arr_ny, arr_nz;   /* in-memory size, x padded to --mem_alignment     */
  int32_t  periodic[3];         /* ctx['block_periodicity']                             */
  int32_t  bits_type;           /* nt_misc_shift                                        */
  int32_t  bits_param;          /* nt_param_shift                                       */
  int32_t  bits_scratch;        /* nt_scratch_shift                                     */
  uint32_t type_id[SLF_NT_COUNT];
  int32_t  use_link_tags;       /* --nouse_link_tags clears it (lb_base.py:91-97)       */
  int32_t  relaxation_enabled;  /* config.relaxation_enabled (regtest/subdomains)     */
  int32_t  propagation_enabled;
  int32_t  force_enabled;       /* LBForcedSim.add_body_force, Guo scheme only          */
  int32_t  unused_tag_bits;     /* link-tag bits set in every tagged node: no missing-link code exists for them
                                 * (geo_encoder.py:283-289, boundary.mako:393-401)                           */
  double   visc;                /* --visc; tau = (6 visc + 1) / 2  (sym.py:847)          */
  double   accel[3];            /* body acceleration (sym_force.py:72-119, accel=True)  */
  int32_t  num_node_params;     /* len(ctx['node_params'])                              */
  int32_t  simtype;             /* SLF_SIM_LBM | SLF_SIM_SHAN_CHEN   (ctx['simtype'])   */
  const double *node_params;    /* host pointer, copied by slf_plan_create              */
  /* binary Shan-Chen mixture (lb_binary.py:375-420): second lattice relaxation time,
   * coupling constants G11, G12, G21 (= G12), G22 and the pseudopotential */
  double   tau_phi;             /* --tau_phi                                            */
  double   sc_g[4];             /* G11, G12, G21, G22  (constants(), lb_binary.py:387-393) */
  int32_t  sc_potential;        /* SLF_SC_LINEAR | SLF_SC_CLASSIC (sym.py:896-908)       */
  int32_t  reserved1;
  /* binary models: `accel` above acts on the first component, accel1 on the second
   * (add_body_force(..., grid=1), lb_base.py:331-370; relaxation_common.mako:24-33: the body
   * acceleration is added to the Shan-Chen acceleration of that component) */
  double   accel1[3];
} SlfParams;

typedef struct SlfPlan SlfPlan;

/* ---- device / memory: backend_cuda.py:52-191 ------------------------------------ */
int  slf_abi_version(void);
const char *slf_last_error(void);
int  slf_device_count(int *count);
int  slf_init(int device);                                   /* CUDABackend.__init__ :52-107 */
int  slf_device_info(char *name, int name_len, uint64_t *total_mem, int *sm_count);
int  slf_malloc(uint64_t bytes, uint64_t *dptr);             /* alloc_buf :132-154           */
int  slf_free(uint64_t dptr);
int  slf_memset(uint64_t dptr, int value, uint64_t bytes, uint64_t stream);
int  slf_host_alloc(uint64_t bytes, void **hptr);            /* alloc_async_host_buf :156    */
int  slf_host_free(void *hptr);
int  slf_memcpy_h2d(uint64_t dptr, const void *hptr, uint64_t bytes);                 /* to_buf :161     */
int  slf_memcpy_d2h(void *hptr, uint64_t dptr, uint64_t bytes);                       /* from_buf :174   */
int  slf_memcpy_h2d_async(uint64_t dptr, const void *hptr, uint64_t bytes, uint64_t stream); /* :187 */
int  slf_memcpy_d2h_async(void *hptr, uint64_t dptr, uint64_t bytes, uint64_t stream);       /* :190 */
int  slf_memcpy_d2d_async(uint64_t dst, uint64_t src, uint64_t bytes, uint64_t stream);
int  slf_sync(void);                                         /* sync :294                    */
int  slf_stream_create(uint64_t *stream, int high_priority); /* make_stream :297             */
int  slf_stream_destroy(uint64_t stream);
int  slf_stream_sync(uint64_t stream);
int  slf_stream_wait_event(uint64_t stream, uint64_t event);
int  slf_event_create(uint64_t *event, int timing);          /* make_event :300-308          */
int  slf_event_destroy(uint64_t event);
int  slf_event_record(uint64_t event, uint64_t stream);
int  slf_event_sync(uint64_t event);
int  slf_event_elapsed_ms(uint64_t start, uint64_t end, float *ms);

/* ---- plan: replaces BlockCodeGenerator.get_code + backend.build (codegen.py:104,
 *      backend_cuda.py:193-218): no source is generated, the context becomes data ---- */
int  slf_plan_create(const SlfParams *params, SlfPlan **plan);
int  slf_plan_destroy(SlfPlan *plan);
/* Declares which faces of the split (outermost) axis have a neighbouring subdomain
 * (SubdomainSpec.has_face_conn, subdomain.py:242-304).  The PBC kernels of the other axes then leave
 * alone what the neighbour delivers through slf_halo_push*: the ghost layer of that face and the
 * crossing populations of the adjacent real layer. */
int  slf_plan_set_links(SlfPlan *plan, int link_low, int link_high);
/* Binds a node map (device pointer to the uint32 array the kernels receive as `map`) to the
 * plan and derives, once, what the reference re-reads every step: 1-bit masks of the plain
 * fluid and full-bounce-back nodes for the bulk kernel and the sorted list of the remaining
 * active nodes for the boundary kernel.  Launches with a different / unbound map pointer fall
 * back to kernels that decode the map per node.  Call again after the map is modified
 * (Subdomain.update_node -> backend.to_buf, subdomain.py:564-591); map = 0 unbinds. */
int  slf_plan_bind_map(SlfPlan *plan, uint64_t map);
/* number of kernels this library has launched since slf_init (bench `gpu_launches`) */
uint64_t slf_launch_count(void);

/* ---- kernels (names are the reference's, kernel-name contract in SURVEY.md 8b) ---- */

/* SetInitialConditions(dist, ivx, ivy[, ivz], irho, map)  lb_single_fluid.mako:101-127 */
int  slf_set_initial_conditions(SlfPlan *plan, uint64_t dist, uint64_t vx, uint64_t vy,
                                uint64_t vz, uint64_t rho, uint64_t map, uint64_t stream);

/* CollideAndPropagate(map, dist_in, dist_out, rho, vx, vy[, vz], options, iteration)
 * lb_single_fluid.mako:161-229.  z_begin/z_end select a slab of gz planes
 * [z_begin, z_end) (2-D: gy rows); pass 0, -1 for the whole subdomain.  This replaces
 * the reference's bulk/boundary block lists (kernel_common.mako:245-457). */
int  slf_collide_and_propagate(SlfPlan *plan, uint64_t map, uint64_t dist_in, uint64_t dist_out,
                               uint64_t rho, uint64_t vx, uint64_t vy, uint64_t vz,
                               int options, uint32_t iteration,
                               int z_begin, int z_end, uint64_t stream);

/* Runs `nsteps` whole time steps from C: CollideAndPropagate over the subdomain followed by the
 * PBC kernels of the locally periodic axes, with the AA in-place / AB ping-pong lattice selection
 * of lb_single.py:96-185 -- the body of SubdomainRunner.main()/step() (subdomain_runner.py:960-974,
 * :1673-1748) for stretches without output or halo exchange.  Pairs of steps are replayed as a CUDA
 * graph, which removes the per-step host cost that dominates small lattices (ldc_2d 256^2).
 * `iteration` is the number of the first step; save_last stores rho, v in the last one. */
int  slf_run_steps(SlfPlan *plan, uint64_t map, uint64_t dist_a, uint64_t dist_b, uint64_t rho,
                   uint64_t vx, uint64_t vy, uint64_t vz, uint32_t iteration, int nsteps,
                   int save_last, uint64_t stream);

/* Output fast path: device side of SubdomainRunner._fields_to_host + LBOutput.mask_nonfluid_nodes
 * (subdomain_runner.py:1141-1151, io.py:53-59).  Packs the real nodes of `field` (ghost layer and x
 * padding stripped) into the dense float array `out` of (lat_nz-2)*(lat_ny-2)*(lat_nx-2) elements and
 * stores NaN for nodes that are not wet; the host then only copies and writes. */
int  slf_pack_field(SlfPlan *plan, uint64_t map, uint64_t field, uint64_t out, uint64_t stream);

/* PrepareMacroFields(map, dist, rho, options, iteration)  lb_single_fluid.mako:129-159 */
int  slf_prepare_macro_fields(SlfPlan *plan, uint64_t map, uint64_t dist, uint64_t rho,
                              int options, uint32_t iteration, uint64_t stream);

/* ---- binary Shan-Chen mixture (BASELINE config 5; lb_binary.py:375-517) ---------------------
 * SetInitialConditions(map, dist1, dist2, ivx, ivy[, ivz], irho, iphi)  lb_binary_fluid.mako:88-127 */
int  slf_sc_set_initial_conditions(SlfPlan *plan, uint64_t map, uint64_t dist1, uint64_t dist2,
                                   uint64_t vx, uint64_t vy, uint64_t vz, uint64_t rho, uint64_t phi,
                                   uint64_t stream);
/* ShanChenPrepareMacroFields(map, dist1, dist2, orho0, orho1, ovx, ovy[, ovz], options, iteration)
 * binary_shan_chen.mako:19-87: both densities and the tau-weighted common velocity. */
int  slf_sc_prepare_macro_fields(SlfPlan *plan, uint64_t map, uint64_t dist1, uint64_t dist2,
                                 uint64_t rho, uint64_t phi, uint64_t vx, uint64_t vy, uint64_t vz,
                                 int options, uint32_t iteration, int z_begin, int z_end, uint64_t stream);
/* ShanChenCollideAndPropagate{0,1}(map, dist_in, dist_out, gg0m0, gg1m0, ovx, ovy[, ovz], options,
 * iteration)  binary_shan_chen.mako:89-141 + shan_chen.mako:9-84: grid_idx selects the component. */
int  slf_sc_collide_and_propagate(SlfPlan *plan, int grid_idx, uint64_t map, uint64_t dist_in,
                                  uint64_t dist_out, uint64_t rho, uint64_t phi, uint64_t vx, uint64_t vy,
                                  uint64_t vz, int options, uint32_t iteration, int z_begin, int z_end,
                                  uint64_t stream);
/* Both components in one launch (native runner; the reference issues ShanChenCollideAndPropagate0 and 1
 * back to back, lb_binary.py:494-515, "These kernels can actually run in parallel").  The common
 * velocity is recomputed from the populations, so the preceding slf_sc_prepare_macro_fields may run
 * with SLF_OPTION_DENSITIES_ONLY; with SLF_OPTION_SAVE_MACRO_FIELDS the velocity the macro pass
 * would have stored is written to vx, vy, vz. */
int  slf_sc_collide_and_propagate_fused(SlfPlan *plan, uint64_t map, uint64_t dist1_in, uint64_t dist1_out,
                                        uint64_t dist2_in, uint64_t dist2_out, uint64_t rho, uint64_t phi,
                                        uint64_t vx, uint64_t vy, uint64_t vz, int options, uint32_t iteration,
                                        int z_begin, int z_end, uint64_t stream);
/* ApplyMacroPeriodicBoundaryConditions(field, axis)  kernel_utils.mako:410-474 */
int  slf_apply_macro_periodic_boundary_conditions(SlfPlan *plan, uint64_t field, int axis, uint64_t stream);
/* Ghost-layer exchange of a scalar field across the split axis (replaces
 * Collect/DistributeContinuousMacroData + connectors, kernel_utils.mako:839-953,
 * subdomain_runner.py:1900-2100): first/last real layer -> neighbour's ghost layer. */
int  slf_halo_push_field(SlfPlan *plan, uint64_t field, int face, uint64_t peer_field, int peer_lat_n,
                         uint64_t stream);

/* ApplyPeriodicBoundaryConditions[WithSwap](dist, axis)  kernel_utils.mako:295-384 */
int  slf_apply_periodic_boundary_conditions(SlfPlan *plan, uint64_t dist, int axis,
                                            int with_swap, uint64_t stream);

/* Collect/DistributeContinuousData[WithSwap](dist, face, base_gx, base_other, max_lx,
 * max_other, buffer)  kernel_utils.mako:526-794.  2-D: base_other is ignored and
 * max_other must be 1 (the reference's 2-D signature has no such arguments; its
 * max_lx = ndists * nx is passed here as max_lx = nx). */
int  slf_collect_continuous_data(SlfPlan *plan, uint64_t dist, int face, int base_gx,
                                 int base_other, int max_lx, int max_other, uint64_t buffer,
                                 int with_swap, uint64_t stream);
int  slf_distribute_continuous_data(SlfPlan *plan, uint64_t dist, int face, int base_gx,
                                    int base_other, int max_lx, int max_other, uint64_t buffer,
                                    int with_swap, uint64_t stream);
/* Collect/DistributeSparseData(idx_array, dist, buffer, max_idx)  kernel_utils.mako:796-836 */
int  slf_collect_sparse_data(SlfPlan *plan, uint64_t idx_array, uint64_t dist, uint64_t buffer,
                             int max_idx, uint64_t stream);
int  slf_distribute_sparse_data(SlfPlan *plan, uint64_t idx_array, uint64_t dist, uint64_t buffer,
                                int max_idx, uint64_t stream);

/* Collect/DistributeContinuousMacroData(field, face, base_gx, base_other, max_lx, max_other, buffer)
 * kernel_utils.mako:838-953 (call sites subdomain_runner.py:1984-2058): ghost-layer exchange of a
 * scalar field through y and z faces; collect reads the first/last real layer, distribute writes the
 * ghost layer.  2-D (reference signature field, base_gx, max_lx, gy, buffer): pass gy as base_other,
 * face is ignored, max_other = 1.  Buffer rows are dense, [max_other][max_lx]. */
int  slf_collect_continuous_macro_data(SlfPlan *plan, uint64_t field, int face, int base_gx, int base_other,
                                       int max_lx, int max_other, uint64_t buffer, uint64_t stream);
int  slf_distribute_continuous_macro_data(SlfPlan *plan, uint64_t field, int face, int base_gx, int base_other,
                                          int max_lx, int max_other, uint64_t buffer, uint64_t stream);

/* ---- halo exchange between the GPUs of one box (subdomains split along the outermost
 * axis: conn_axis = z in 3-D, y in 2-D) -------------------------------------------------
 * Replaces SubdomainRunner._send_dists/_recv_dists + the collect/distribute kernels + the
 * ZeroMQ connectors (subdomain_runner.py:1064-1139, kernel_utils.mako:526-794,
 * connector.py:73-117): the |dists| planes of a face message are contiguous in the SoA
 * lattice, so one kernel copies them from this GPU's lattice straight into the
 * neighbour's lattice through a peer-mapped pointer (no pack, no host staging).
 *   face        local face the populations leave through (SLF_Z_LOW/HIGH; 2-D: SLF_Y_*)
 *   peer_dist   neighbour's dist base, mapped here (slf_ipc_open_handle / peer access)
 *   peer_lat_n, peer_arr_n   neighbour's lattice / array extent along the split axis
 *   with_swap   1 after an AA even (in-place) step: first real layer, opposite slots ->
 *               neighbour's ghost layer; 0 after AB / AA-odd steps: ghost layer, natural
 *               slots -> neighbour's first real layer (the reference's WithSwap variants).
 *   signal_sem  0, or the neighbour's `ready` word (peer pointer): the last block of the copy
 *               kernel increments it after a system-scope fence, so copy + signal are one launch.
 */
int  slf_halo_push(SlfPlan *plan, uint64_t dist, int face, uint64_t peer_dist, int peer_lat_n,
                   int peer_arr_n, int with_swap, uint64_t signal_sem, uint64_t stream);
/* The same for ANY face of a subdomain that is split along x and / or y as well (block decompositions,
 * regtest/subdomains/3d_ldc.py).  Replaces, together with a dimension-ordered schedule of the caller (all x faces,
 * then y, then z; the caller waits for axis k before it pushes axis k + 1), the reference's per-pair face, edge and
 * corner connections with their sparse index lists (subdomain_connection.py:348-407, Collect/DistributeSparseData,
 * subdomain_runner.py:624-738): edge and corner data travel in the ghost ring of the face messages and reach their
 * diagonal destination in two or three hops.  peer_lat / peer_arr: the neighbour's lattice / array extents (x, y, z);
 * cross-sections must be equal.  The copied windows are disjoint between messages (see csrc/slf_halo.cuh). */
int  slf_halo_push_face(SlfPlan *plan, uint64_t dist, int face, uint64_t peer_dist, const int *peer_lat,
                        const int *peer_arr, int with_swap, uint64_t signal_sem, uint64_t stream);
/* Which faces of this subdomain have a neighbour (bit 2 * axis + side; default: all).  Across a face WITHOUT one -- the
 * domain boundary -- nothing was delivered into the ghost layer, so slf_halo_push_face leaves the positions that
 * would forward such data out of its windows.  The reference gets the same effect from the partial maps of its
 * connections (subdomain_connection.py:96-180: populations whose source lies outside the domain are not sent); it
 * matters to nodes that keep what they find in the slots nobody streams into (do-nothing outflow, link bounce-back). */
int  slf_plan_set_face_links(SlfPlan *plan, int links);
/* The windows slf_halo_push_face uses, as a pure host function (tests): returns 1 if population i of `lattice` crosses
 * `face`, and then lo[0..1] / cnt[0..1] = first position and count along the other axes (ascending axis order). */
int  slf_halo_face_window(int lattice, const int *lat, int face, int with_swap, int links, int i, int *lo, int *cnt);
/* Stream-ordered counting semaphores in device memory for cross-GPU ordering of the pushes
 * (they stand where the reference blocks in zmq recv, connector.py:73-117).  A subdomain owns a
 * zero-initialised block of uint32 words: ready[channel][side] (incremented by the neighbour on
 * that side after each push), seq[channel] (steps waited for so far) and two status words.
 *   signal: system-scope fence + atomic increment of `sem` (a peer pointer);
 *   wait  : spin (acquire loads, system scope) until ready_a and ready_b (either may be 0) have
 *           reached *seq + ahead (ahead = 0: pushes of the previous steps, the population planes;
 *           ahead = 1: of this step too, the mid-step field exchange), then ++*seq.  Nothing is ever reset: re-initialising or restoring a run
 *           only needs a host barrier.  A wait that lasts longer than SLF_SEM_TIMEOUT_MS (default
 *           30000) or sees status[1] != 0 stores 1 (timeout) / 2 (aborted) in status[0] and returns;
 *   status: blocking read of status[0] on a private stream;  abort: sets status[1]. */
int  slf_sem_signal(uint64_t sem, uint64_t stream);
int  slf_sem_wait(uint64_t ready_a, uint64_t ready_b, uint64_t seq, uint64_t status, int ahead, uint64_t stream);
int  slf_sem_status(uint64_t status, uint32_t *err);
int  slf_sem_abort(uint64_t status);

/* Stream capture for step replay of linked subdomains (SubdomainRunner.main() pays ~10 Python calls
 * per step, subdomain_runner.py:1673-1748): everything issued on `stream` (and on streams forked from
 * it with events) between begin and end is recorded instead of run; launch replays it. */
int  slf_graph_begin_capture(uint64_t stream);
int  slf_graph_end_capture(uint64_t stream, uint64_t *graph_exec);
int  slf_graph_launch(uint64_t graph_exec, uint64_t stream);
int  slf_graph_destroy(uint64_t graph_exec);

/* CUDA IPC handles for the above (backend_cuda.py:122-126 wraps them but never calls them) */
int  slf_ipc_get_handle(uint64_t dptr, void *handle64);      /* 64 bytes */
int  slf_ipc_open_handle(const void *handle64, uint64_t *dptr);
int  slf_ipc_close_handle(uint64_t dptr);
int  slf_enable_peer_access(int peer_device);

#ifdef __cplusplus
}
#endif
#endif  /* SLF_B200_H_ */
