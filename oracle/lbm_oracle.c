/*
 * lbm_oracle.c -- CPU restatement of the reference's collide-and-propagate path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (sailfish_b200/, libslf_b200.so) never does and has no CPU fallback.
 *
 * Why a restatement: the reference has no compiled kernel source -- its kernels
 * are Mako templates rendered and nvcc-JITed at run time through pycuda
 * (sailfish/backend_cuda.py:193-218); neither Mako, pycuda nor a GPU exist in the
 * build container, and `--backends=dummy` is a no-op (backend_dummy.py:31-38).
 * The control flow below follows the templates line by line; the closed-form
 * expressions are restated from sailfish/sym*.py.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - every formula is checked against tests/golden/formulas.npz, produced by
 *     lambdifying the reference's own SymPy expressions (tools/gen_golden.py);
 *   - whole steps are checked against tests/golden/steps_*.npz, produced by
 *     tools/ref_stepper.py whose arithmetic is the reference's SymPy objects and
 *     whose geometry is the reference's own encoded node map;
 *   - streaming/halo/PBC are checked against the literal known-answer vectors of
 *     regtest/subdomains/{2d,3d}_propagation.py.
 *
 * Built twice: -DREAL=float (parity target, the reference computes in fp32) and
 * -DREAL=double (error budgeting).  OpenMP over the outermost axis is used for the
 * CPU baseline timing only.
 *
 * Reference map (execution order inside CollideAndPropagate,
 * templates/models/lb_single_fluid.mako:161-229):
 *   local_indices / load_node_type   kernel_common.mako:83-115, :191-201
 *   getDist                          geo_helpers.mako:225-276
 *   getMacro                         boundary.mako:465-507 (+ :418-459, :267-319)
 *   precollisionBoundaryConditions   boundary.mako:784-878
 *   relaxate -> BGK_relaxate0        relaxation.mako:99-181, relaxation_common.mako
 *            -> MS_relaxate          relaxation_mrt.mako:31-97
 *   postcollisionBoundaryConditions  boundary.mako:670-683 (HalfBB :653-666)
 *   save_macro_fields                kernel_common.mako:213-240
 *   propagate                        propagation.mako:384-421 (+ :12-60 bounds)
 */
#define _GNU_SOURCE
#include <math.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/slf_b200.h"

#ifndef REAL
#define REAL float
#endif

#define MAXQ 27

typedef struct Lat {
  int Q, D;
  int e[MAXQ][3];
  double w[MAXQ];
  int opp[MAXQ];
  int dir2vec[7];      /* direction number 1..2D -> slot */
  int has_mrt;
  int M[MAXQ][MAXQ];   /* orthogonal integer moment basis (sym.py:1044) */
  double Mnorm[MAXQ];
  double rates[MAXQ];  /* -1 -> 1/tau */
  int i_mx, i_my, i_mz;
} Lat;

static const int E_D2Q9[9][3] = {{0,0,0},{1,0,0},{0,1,0},{-1,0,0},{0,-1,0},{1,1,0},{-1,1,0},{-1,-1,0},{1,-1,0}};
static const int E_D3Q27[27][3] = {
  {0,0,0},{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1},
  {1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1},
  {1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},
  {1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{-1,1,1},{-1,1,-1},{-1,-1,1},{-1,-1,-1}};

static Lat g_lat[3];
static int g_lat_ready = 0;

/* Gram-Schmidt with integer rescaling, sym.py:918-942. Uses long double rationals
 * via brute force: the bases are tiny so exact fractions are done with int64. */
static long long gcdll(long long a, long long b) { if (a < 0) a = -a; if (b < 0) b = -b; while (b) { long long t = a % b; a = b; b = t; } return a; }

static void orthogonalize(int Q, long long rows[MAXQ][MAXQ], int out[MAXQ][MAXQ]) {
  /* keep vectors as integer numerators with a common positive denominator */
  static long long num[MAXQ][MAXQ];
  for (int k = 0; k < Q; k++) {
    long long v[MAXQ], den = 1;
    for (int i = 0; i < Q; i++) v[i] = rows[k][i];
    for (int j = 0; j < k; j++) {
      long long uu = 0, uv = 0;
      for (int i = 0; i < Q; i++) { uu += num[j][i] * num[j][i]; uv += num[j][i] * v[i]; }
      /* v <- v - (uv/(den*uu)) u  with v = v/den : v*uu - uv*u over den*uu */
      for (int i = 0; i < Q; i++) v[i] = v[i] * uu - uv * num[j][i];
      den *= uu;
      long long g = 0;
      for (int i = 0; i < Q; i++) g = gcdll(g, v[i]);
      g = gcdll(g, den);
      if (g > 1) { for (int i = 0; i < Q; i++) v[i] /= g; den /= g; }
    }
    long long g = 0;
    for (int i = 0; i < Q; i++) g = gcdll(g, v[i]);
    if (g > 1) for (int i = 0; i < Q; i++) v[i] /= g;
    for (int i = 0; i < Q; i++) { num[k][i] = v[i]; out[k][i] = (int)v[i]; }
  }
}

static void lat_finish(Lat *L) {
  for (int i = 0; i < L->Q; i++)
    for (int j = 0; j < L->Q; j++)
      if (L->e[i][0] == -L->e[j][0] && L->e[i][1] == -L->e[j][1] && L->e[i][2] == -L->e[j][2]) L->opp[i] = j;
  int d = 1;
  for (int i = 0; i < L->Q; i++) {
    int n2 = L->e[i][0]*L->e[i][0] + L->e[i][1]*L->e[i][1] + L->e[i][2]*L->e[i][2];
    if (n2 == 1) L->dir2vec[d++] = i;
  }
}

static void lat_init(void) {
  if (g_lat_ready) return;
  /* D2Q9, sym.py:61-95 */
  Lat *L = &g_lat[SLF_D2Q9];
  memset(L, 0, sizeof(*L));
  L->Q = 9; L->D = 2;
  for (int i = 0; i < 9; i++) for (int d = 0; d < 3; d++) L->e[i][d] = E_D2Q9[i][d];
  L->w[0] = 4.0/9; for (int i = 1; i < 5; i++) L->w[i] = 1.0/9; for (int i = 5; i < 9; i++) L->w[i] = 1.0/36;
  lat_finish(L);
  {
    long long rows[MAXQ][MAXQ];
    for (int i = 0; i < 9; i++) {
      long long x = L->e[i][0], y = L->e[i][1], sq = x*x + y*y;
      rows[0][i] = 1; rows[1][i] = sq; rows[2][i] = sq*sq; rows[3][i] = x; rows[4][i] = x*sq;
      rows[5][i] = y; rows[6][i] = y*sq; rows[7][i] = x*x - y*y; rows[8][i] = x*y;
    }
    orthogonalize(9, rows, L->M);
    const double r[9] = {0, 1.63, 1.14, 0, 1.9, 0, 1.9, -1, -1};   /* sym.py:83 */
    for (int i = 0; i < 9; i++) L->rates[i] = r[i];
    L->has_mrt = 1; L->i_mx = 3; L->i_my = 5; L->i_mz = -1;
  }
  /* D3Q19 (first 19 vectors of the D3Q27 list), sym.py:312-358 */
  L = &g_lat[SLF_D3Q19];
  memset(L, 0, sizeof(*L));
  L->Q = 19; L->D = 3;
  for (int i = 0; i < 19; i++) for (int d = 0; d < 3; d++) L->e[i][d] = E_D3Q27[i][d];
  L->w[0] = 1.0/3; for (int i = 1; i < 7; i++) L->w[i] = 1.0/18; for (int i = 7; i < 19; i++) L->w[i] = 1.0/36;
  lat_finish(L);
  {
    long long rows[MAXQ][MAXQ];
    for (int i = 0; i < 19; i++) {
      long long x = L->e[i][0], y = L->e[i][1], z = L->e[i][2], sq = x*x + y*y + z*z;
      rows[0][i] = 1; rows[1][i] = sq; rows[2][i] = sq*sq;
      rows[3][i] = x; rows[4][i] = x*sq; rows[5][i] = y; rows[6][i] = y*sq; rows[7][i] = z; rows[8][i] = z*sq;
      rows[9][i] = 3*x*x - sq; rows[10][i] = (3*sq - 5)*(3*x*x - sq);
      rows[11][i] = y*y - z*z; rows[12][i] = (3*sq - 5)*(y*y - z*z);
      rows[13][i] = x*y; rows[14][i] = y*z; rows[15][i] = x*z;
      rows[16][i] = (y*y - z*z)*x; rows[17][i] = (z*z - x*x)*y; rows[18][i] = (x*x - y*y)*z;
    }
    orthogonalize(19, rows, L->M);
    const double r[19] = {0.0, 1.19, 1.4, 0.0, 1.2, 0.0, 1.2, 0.0, 1.2, -1, 1.4, -1, 1.4, -1, -1, -1, 1.98, 1.98, 1.98};
    for (int i = 0; i < 19; i++) L->rates[i] = r[i];   /* sym.py:334-335 */
    L->has_mrt = 1; L->i_mx = 3; L->i_my = 5; L->i_mz = 7;
  }
  /* D3Q27, sym.py:409-428 (BGK only) */
  L = &g_lat[SLF_D3Q27];
  memset(L, 0, sizeof(*L));
  L->Q = 27; L->D = 3;
  for (int i = 0; i < 27; i++) for (int d = 0; d < 3; d++) L->e[i][d] = E_D3Q27[i][d];
  L->w[0] = 8.0/27; for (int i = 1; i < 7; i++) L->w[i] = 2.0/27; for (int i = 7; i < 19; i++) L->w[i] = 1.0/54;
  for (int i = 19; i < 27; i++) L->w[i] = 1.0/216;
  lat_finish(L);
  for (int l = 0; l < 3; l++)
    for (int k = 0; k < g_lat[l].Q; k++) {
      double s = 0; for (int i = 0; i < g_lat[l].Q; i++) s += (double)g_lat[l].M[k][i] * g_lat[l].M[k][i];
      g_lat[l].Mnorm[k] = s;
    }
  g_lat_ready = 1;
}

/* introspection for tests */
int orc_lattice_info(int lattice, int *Q, int *D, int *e, double *w, int *opp, int *M) {
  lat_init();
  if (lattice < 0 || lattice > 2) return SLF_EINVAL;
  const Lat *L = &g_lat[lattice];
  *Q = L->Q; *D = L->D;
  for (int i = 0; i < L->Q; i++) {
    for (int d = 0; d < 3; d++) e[3*i + d] = L->e[i][d];
    w[i] = L->w[i]; opp[i] = L->opp[i];
    for (int j = 0; j < L->Q; j++) M[i*L->Q + j] = L->M[i][j];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ */
typedef struct Ctx {
  const SlfParams *P;
  const Lat *L;
  int nx, ny, nz, ax, ay, az;
  long long DS;
  long long stride[3];
  uint32_t tmask; int oshift, pshift; uint32_t pmask;
  REAL omega;        /* (float)(1/tau), relaxation.mako:133 */
  REAL tau;
  REAL visc;
  REAL a[3];
  int forced;
  REAL rates[MAXQ];
} Ctx;

static int ctx_init(Ctx *C, const SlfParams *P) {
  lat_init();
  if (!P || P->struct_bytes != sizeof(SlfParams)) return SLF_EINVAL;
  if (P->lattice < 0 || P->lattice > 2) return SLF_EINVAL;
  C->P = P; C->L = &g_lat[P->lattice];
  if (P->model == SLF_MRT && !C->L->has_mrt) return SLF_EUNSUPPORTED;
  if (P->access_pattern == SLF_AA &&                 /* boundary.mako:626-628 */
      (P->type_id[SLF_NT_COPY] != SLF_NT_ABSENT || P->type_id[SLF_NT_YU_OUTFLOW] != SLF_NT_ABSENT))
    return SLF_EUNSUPPORTED;
  C->nx = P->lat_nx; C->ny = P->lat_ny; C->nz = C->L->D == 3 ? P->lat_nz : 1;
  C->ax = P->arr_nx; C->ay = P->arr_ny; C->az = C->L->D == 3 ? P->arr_nz : 1;
  C->DS = (long long)C->ax * C->ay * C->az;
  C->stride[0] = 1; C->stride[1] = C->ax; C->stride[2] = (long long)C->ax * C->ay;
  C->tmask = (1u << P->bits_type) - 1u;
  C->pshift = P->bits_type; C->pmask = (1u << P->bits_param) - 1u;
  C->oshift = P->bits_type + P->bits_param + P->bits_scratch;
  double tau = (6.0 * P->visc + 1.0) / 2.0;     /* sym.py:847-848 */
  C->tau = (REAL)tau;
  C->omega = (REAL)(1.0 / tau);
  C->visc = (REAL)P->visc;
  C->forced = P->force_enabled;
  for (int d = 0; d < 3; d++) C->a[d] = (REAL)P->accel[d];
  /* relaxation_mrt.mako:44-46 evaluates inv_tau = 1/(0.5 + 3 visc) in REAL */
  REAL inv_tau = (REAL)1.0 / ((REAL)0.5 + (REAL)3.0 * C->visc);
  for (int k = 0; k < C->L->Q; k++) C->rates[k] = C->L->rates[k] < 0 ? inv_tau : (REAL)C->L->rates[k];
  return 0;
}

static inline int nclass(const Ctx *C, uint32_t type) {
  for (int k = 0; k < SLF_NT_COUNT; k++) if (C->P->type_id[k] == type) return k;
  return -1;
}

/* f^eq, sym_equilibrium.py:90-120: w_i (rho + rho0 * horner(3 e.v + 4.5 (e.v)^2 - 1.5 v.v)), rho0 = rho */
static inline void feq_all(const Lat *L, REAL rho, const REAL *v, REAL *out) {
  REAL vv = 0;
  for (int d = 0; d < L->D; d++) vv += v[d] * v[d];
  for (int i = 0; i < L->Q; i++) {
    REAL ev = 0;
    for (int d = 0; d < L->D; d++) ev += (REAL)L->e[i][d] * v[d];
    REAL h = (REAL)3.0 * ev + (REAL)4.5 * ev * ev - (REAL)1.5 * vv;
    out[i] = (REAL)L->w[i] * (rho + rho * h);
  }
}

static inline void fill_missing_with_opposite(const Lat *L, int orient, REAL *f) {
  /* boundary.mako:418-422 with sym.get_missing_dists (sym.py:534-543) */
  const int *n = L->e[L->dir2vec[orient]];
  for (int i = 0; i < L->Q; i++) {
    int sp = 0; for (int d = 0; d < L->D; d++) sp += L->e[i][d] * n[d];
    if (sp > 0) f[i] = f[L->opp[i]];
  }
}

/* for_orientation_dist(link_tags=True), boundary.mako:391-416, for link i = 1..Q-1 of a node with orientation /
 * tag word `orient`: is a population of this link undefined, and which (missing) is replaced from which (name)? */
static inline int missing_link(const Ctx *C, int orient, int i, int *missing, int *name) {
  const Lat *L = C->L; const SlfParams *P = C->P;
  if (P->use_link_tags) {
    if (orient & (1 << (i - 1))) return 0;               /* link i points to a fluid node */
    if (P->unused_tag_bits & (1 << (i - 1))) return 0;   /* no code is generated for this bit, boundary.mako:396 */
    *name = i; *missing = L->opp[i];
    return 1;
  }
  if (orient == 0) return 0;
  const int *n = L->e[L->dir2vec[orient]];
  int sp = 0; for (int d = 0; d < L->D; d++) sp += L->e[i][d] * n[d];
  if (sp <= 0) return 0;
  *missing = i; *name = L->opp[i];
  return 1;
}

static inline void bounce_back(const Lat *L, REAL *f) {   /* boundary.mako:255-264 */
  for (int i = 0; i < L->Q; i++) { int j = L->opp[i]; if (i < j) { REAL t = f[i]; f[i] = f[j]; f[j] = t; } }
}

/* MS_relaxate, relaxation_mrt.mako:31-97 */
static inline void mrt_relax(const Ctx *C, int cls, REAL *f, REAL *v) {
  const Lat *L = C->L; const int Q = L->Q, D = L->D;
  REAL m[MAXQ], meq[MAXQ];
  for (int k = 0; k < Q; k++) { REAL s = 0; for (int i = 0; i < Q; i++) if (L->M[k][i]) s += (REAL)L->M[k][i] * f[i]; m[k] = s; }
  const int im[3] = {L->i_mx, L->i_my, L->i_mz};
  if (C->forced) for (int d = 0; d < D; d++) m[im[d]] += (REAL)0.5 * C->a[d];   /* fluid_momentum :11-28 */
  REAL rho = m[0], mx = m[im[0]], my = m[im[1]], mz = D == 3 ? m[im[2]] : 0;
  for (int k = 0; k < Q; k++) meq[k] = 0;
  if (C->P->lattice == SLF_D2Q9) {               /* sym.py:122-149 */
    meq[1] = (REAL)-2.0 * rho + (REAL)3.0 * (mx*mx + my*my);
    meq[2] = rho - (REAL)3.0 * (mx*mx + my*my);
    meq[4] = -mx; meq[6] = -my;
    meq[7] = mx*mx - my*my; meq[8] = mx*my;
  } else {                                        /* D3Q19, sym.py:384-400 */
    REAL ir = (REAL)1.0 / rho, m2 = mx*mx + my*my + mz*mz;
    meq[1] = (REAL)-11.0 * rho + (REAL)19.0 * ir * m2;
    meq[2] = (REAL)(-475.0/63.0) * ir * m2;
    meq[4] = (REAL)(-2.0/3.0) * mx; meq[6] = (REAL)(-2.0/3.0) * my; meq[8] = (REAL)(-2.0/3.0) * mz;
    meq[9] = ir * ((REAL)2.0*mx*mx - (my*my + mz*mz));
    meq[11] = ir * (my*my - mz*mz);
    meq[13] = ir * mx*my; meq[14] = ir * my*mz; meq[15] = ir * mx*mz;
  }
  if (cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_EQUILIBRIUM_DENSITY)   /* :59-77 */
    for (int k = 0; k < Q; k++) if (L->rates[k] != 0) m[k] = meq[k];
  for (int k = 0; k < Q; k++) if (L->rates[k] != 0) m[k] -= C->rates[k] * (m[k] - meq[k]);
  if (C->forced) for (int d = 0; d < D; d++) m[im[d]] += (REAL)0.5 * C->a[d];
  for (int k = 0; k < Q; k++) m[k] = m[k] / (REAL)L->Mnorm[k];
  for (int i = 0; i < Q; i++) { REAL s = 0; for (int k = 0; k < Q; k++) if (L->M[k][i]) s += (REAL)L->M[k][i] * m[k]; f[i] = s; }
  if (C->forced) for (int d = 0; d < D; d++) v[d] += (REAL)0.5 * C->a[d];       /* fluid_output_velocity */
}

/* BGK_relaxate0, relaxation.mako:99-181 + relaxation_common.mako:56-64,110-149,208-237 */
static inline void bgk_relax(const Ctx *C, REAL rho, REAL *f, REAL *v) {
  const Lat *L = C->L; const int Q = L->Q, D = L->D;
  REAL v0[3] = {0, 0, 0}, feq[MAXQ];
  for (int d = 0; d < D; d++) v0[d] = C->forced ? v[d] + (REAL)0.5 * C->a[d] : v[d];
  feq_all(L, rho, v0, feq);
  for (int i = 0; i < Q; i++) f[i] += C->omega * (feq[i] - f[i]);
  if (C->forced) {
    /* Guo, sym_force.py:121-160: pref = rho/cs^2 (1 - 1/(2 tau)) */
    REAL pref = rho * (REAL)3.0 * ((REAL)1.0 - (REAL)1.0 / ((REAL)2.0 * C->tau));
    for (int i = 0; i < Q; i++) {
      REAL ev = 0; for (int d = 0; d < D; d++) ev += (REAL)L->e[i][d] * v0[d];
      REAL s = 0;
      for (int d = 0; d < D; d++) s += ((REAL)L->e[i][d] - v0[d] + (REAL)3.0 * ev * (REAL)L->e[i][d]) * C->a[d];
      f[i] += pref * (REAL)L->w[i] * s;
    }
    for (int d = 0; d < D; d++) v[d] += (REAL)0.5 * C->a[d];
  }
}

/* zouhe_bb, boundary.mako:347-381 (orientation 1..2D): bounce-back of the non-equilibrium parts of the unknown
 * populations (sym.py:750-766), then the momentum excess rho v - sum e f along the directions tangential to the wall
 * is spread evenly over the unknown populations that have a component along them (zouhe_fixup, sym.py:770-815). */
static inline void zouhe_bb(const Lat *L, int orient, REAL rho, const REAL *v, REAL *f) {
  const int Q = L->Q, D = L->D;
  const int *n = L->e[L->dir2vec[orient]];
  REAL feq[MAXQ];
  feq_all(L, rho, v, feq);
  int unknown[MAXQ];
  for (int i = 0; i < Q; i++) {
    int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * n[d];
    unknown[i] = sp > 0;
  }
  for (int i = 0; i < Q; i++)
    if (unknown[i]) f[i] = f[L->opp[i]] - feq[L->opp[i]] + feq[i];
  REAL nv[3] = {0, 0, 0};
  for (int d = 0; d < D; d++) {
    REAL s = 0; for (int i = 0; i < Q; i++) s += (REAL)L->e[i][d] * f[i];
    nv[d] = rho * v[d] - s;
  }
  const int normal_slot = L->dir2vec[orient];
  for (int d = 0; d < D; d++) {
    if (n[d] != 0) { nv[d] = 0; continue; }
    int cnt = 0;
    for (int i = 0; i < Q; i++) if (unknown[i] && i != normal_slot && L->e[i][d] != 0) cnt++;
    nv[d] = cnt ? nv[d] / (REAL)cnt : (REAL)0;
  }
  for (int i = 0; i < Q; i++) {
    if (!unknown[i] || i == normal_slot) continue;
    REAL val = 0;
    for (int d = 0; d < D; d++) if (L->e[i][d] != 0 && n[d] == 0) val += nv[d] * (REAL)L->e[i][d];
    f[i] = f[i] + val;
  }
}

static inline int in_bounds_target(const Ctx *C, const int *e, int x, int y, int z) {
  /* prop_bnd, propagation.mako:36-41: positive component needs coord < lim-1, negative coord > 0 */
  const int c[3] = {x, y, z}, lim[3] = {C->nx, C->ny, C->nz};
  for (int d = 0; d < C->L->D; d++) {
    if (e[d] > 0 && !(c[d] < lim[d] - 1)) return 0;
    if (e[d] < 0 && !(c[d] > 0)) return 0;
  }
  return 1;
}

/* Plain fluid node, BGK, no body force: the same operations as the generic path below
 * (getDist -> standard getMacro -> BGK_relaxate0 -> propagate) with the lattice tables as
 * compile-time constants so that gcc unrolls and vectorises.  Used for the CPU baseline; the
 * golden tests run through it as well (it is bit-identical to the generic path by construction:
 * same expressions, same order). */
static const REAL W_D2Q9[9] = {(REAL)(4.0/9), (REAL)(1.0/9), (REAL)(1.0/9), (REAL)(1.0/9), (REAL)(1.0/9),
                               (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36)};
static const REAL W_D3Q19[19] = {(REAL)(1.0/3), (REAL)(1.0/18), (REAL)(1.0/18), (REAL)(1.0/18), (REAL)(1.0/18), (REAL)(1.0/18), (REAL)(1.0/18),
                                 (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36),
                                 (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36), (REAL)(1.0/36)};
static const int OPP_D2Q9[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
static const int OPP_D3Q19[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

static inline __attribute__((always_inline)) void fluid_bgk_node(
    const int Q, const int D, const int (*E)[3], const REAL *W, const int *OPP, const long long *off,
    const long long DS, const long long gi, const REAL *din, REAL *dout, const int odd, const int aa,
    const REAL omega, const int save, REAL *orho, REAL **ov) {
  REAL f[MAXQ];
#pragma GCC unroll 27
  for (int i = 0; i < Q; i++) f[i] = odd ? din[gi - off[i] + DS * OPP[i]] : din[gi + DS * i];
  REAL rho = 0;
#pragma GCC unroll 27
  for (int i = 0; i < Q; i++) rho += f[i];
  REAL v[3] = {0, 0, 0};
  for (int d = 0; d < D; d++) {
    REAL s = 0;
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) s += (REAL)E[i][d] * f[i];
    v[d] = s / rho;
  }
  REAL vv = 0;
  for (int d = 0; d < D; d++) vv += v[d] * v[d];
#pragma GCC unroll 27
  for (int i = 0; i < Q; i++) {
    REAL ev = 0;
    for (int d = 0; d < D; d++) ev += (REAL)E[i][d] * v[d];
    REAL h = (REAL)3.0 * ev + (REAL)4.5 * ev * ev - (REAL)1.5 * vv;
    REAL feq = W[i] * (rho + rho * h);
    f[i] += omega * (feq - f[i]);
  }
  if (save) { orho[gi] = rho; for (int d = 0; d < D; d++) ov[d][gi] = v[d]; }
  if (aa && !odd) {
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) dout[gi + DS * OPP[i]] = f[i];
  } else {
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) dout[gi + off[i] + DS * i] = f[i];
  }
}

/* The same arithmetic for a run of `n` consecutive plain fluid nodes of one row, VW nodes at a time: every node
 * sees exactly the operation sequence of fluid_bgk_node (so the results are bit-identical to it), but the loops over
 * the VW lanes are what gcc vectorises.  CPU-baseline plumbing: the scalar path reaches ~10 MLUPS per core. */
#define VW 16
static inline __attribute__((always_inline)) void fluid_bgk_span(
    const int Q, const int D, const int (*E)[3], const REAL *W, const int *OPP, const long long *off,
    const long long DS, const long long gi0, const int n, const REAL *din, REAL *dout, const int odd, const int aa,
    const REAL omega, const int save, REAL *orho, REAL **ov) {
  for (int b = 0; b < n; b += VW) {
    const int m = n - b < VW ? n - b : VW;
    const long long gi = gi0 + b;
    REAL f[MAXQ][VW] __attribute__((aligned(64))), rho[VW] __attribute__((aligned(64)));
    REAL v[3][VW] __attribute__((aligned(64))), vv[VW] __attribute__((aligned(64)));
    /* lanes beyond the run compute on ones and are never stored */
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) {
      const REAL *src = odd ? din + (gi - off[i] + DS * OPP[i]) : din + (gi + DS * i);
      if (m == VW) { for (int l = 0; l < VW; l++) f[i][l] = src[l]; }
      else { for (int l = 0; l < VW; l++) f[i][l] = l < m ? src[l] : (REAL)1; }
    }
    for (int l = 0; l < VW; l++) rho[l] = 0;
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) for (int l = 0; l < VW; l++) rho[l] += f[i][l];
    for (int d = 0; d < 3; d++) for (int l = 0; l < VW; l++) v[d][l] = 0;
    for (int d = 0; d < D; d++) {
      REAL s[VW] __attribute__((aligned(64)));
      for (int l = 0; l < VW; l++) s[l] = 0;
#pragma GCC unroll 27
      for (int i = 0; i < Q; i++) { const REAL e = (REAL)E[i][d]; for (int l = 0; l < VW; l++) s[l] += e * f[i][l]; }
      for (int l = 0; l < VW; l++) v[d][l] = s[l] / rho[l];
    }
    for (int l = 0; l < VW; l++) vv[l] = 0;
    for (int d = 0; d < D; d++) for (int l = 0; l < VW; l++) vv[l] += v[d][l] * v[d][l];
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) {
      REAL ev[VW] __attribute__((aligned(64)));
      for (int l = 0; l < VW; l++) ev[l] = 0;
      for (int d = 0; d < D; d++) { const REAL e = (REAL)E[i][d]; for (int l = 0; l < VW; l++) ev[l] += e * v[d][l]; }
      for (int l = 0; l < VW; l++) {
        const REAL h = (REAL)3.0 * ev[l] + (REAL)4.5 * ev[l] * ev[l] - (REAL)1.5 * vv[l];
        const REAL feq = W[i] * (rho[l] + rho[l] * h);
        f[i][l] += omega * (feq - f[i][l]);
      }
    }
    if (save) {
      for (int l = 0; l < m; l++) orho[gi + l] = rho[l];
      for (int d = 0; d < D; d++) for (int l = 0; l < m; l++) ov[d][gi + l] = v[d][l];
    }
#pragma GCC unroll 27
    for (int i = 0; i < Q; i++) {
      REAL *dst = (aa && !odd) ? dout + (gi + DS * OPP[i]) : dout + (gi + off[i] + DS * i);
      if (m == VW) { for (int l = 0; l < VW; l++) dst[l] = f[i][l]; }
      else { for (int l = 0; l < m; l++) dst[l] = f[i][l]; }
    }
  }
}

int orc_collide_and_propagate(const SlfParams *P, const uint32_t *map, const REAL *din, REAL *dout,
                              REAL *orho, REAL *ovx, REAL *ovy, REAL *ovz,
                              int options, uint32_t iteration, int z_begin, int z_end) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L; const int Q = L->Q, D = L->D;
  const long long DS = C->DS;
  REAL *ov[3] = {ovx, ovy, ovz};
  const int aa = P->access_pattern == SLF_AA;
  const int odd = aa && (iteration & 1);
  const int outer_n = D == 3 ? C->nz : C->ny;
  if (z_end < 0) z_end = outer_n;
  long long off[MAXQ];
  for (int i = 0; i < Q; i++) off[i] = L->e[i][0] + C->stride[1] * L->e[i][1] + C->stride[2] * L->e[i][2];
  const int fast_ok = P->model == SLF_BGK && !C->forced && P->relaxation_enabled && P->propagation_enabled &&
                      (P->lattice == SLF_D3Q19 || P->lattice == SLF_D2Q9);

  /* rows (z, y) are independent work items: enough of them to balance any core count */
  const int rows_per_o = D == 3 ? C->ny : 1;
  const long long row0 = (long long)z_begin * rows_per_o, row1 = (long long)z_end * rows_per_o;
#pragma omp parallel for schedule(static)
  for (long long row = row0; row < row1; row++) {
    const int o = (int)(row / rows_per_o);
    const int z = D == 3 ? o : 0;
    const int y = D == 3 ? (int)(row % rows_per_o) : o;
    const int row_interior = fast_ok && y > 0 && y < C->ny - 1 && (D == 2 || (z > 0 && z < C->nz - 1));
    for (int x = 0; x < C->nx; x++) {
      const long long gi = x + C->stride[1] * y + C->stride[2] * z;
      const uint32_t code = map[gi];
      const uint32_t type = code & C->tmask;
      const int cls = nclass(C, type);
      if (cls == SLF_NT_GHOST || cls == SLF_NT_UNUSED) continue;      /* isExcludedNode */
      if (cls == SLF_NT_PROPAGATION_ONLY) continue;                   /* nothing but barriers */
      if (cls < 0) continue;
      if (cls == SLF_NT_FLUID && row_interior && x > 0 && x < C->nx - 1) {
        /* run of plain fluid nodes: vectorised span (in the AB pattern and in AA even steps a node only writes
         * what no other node of the run reads; AA odd steps read and write the same addresses per node) */
        int n = 1;
        while (x + n < C->nx - 1 && (map[gi + n] & C->tmask) == type) n++;
        const int save = (options & SLF_OPTION_SAVE_MACRO_FIELDS) != 0;
        if (P->lattice == SLF_D3Q19)
          fluid_bgk_span(19, 3, E_D3Q27, W_D3Q19, OPP_D3Q19, off, DS, gi, n, din, dout, odd, aa, C->omega, save, orho, ov);
        else
          fluid_bgk_span(9, 2, E_D2Q9, W_D2Q9, OPP_D2Q9, off, DS, gi, n, din, dout, odd, aa, C->omega, save, orho, ov);
        x += n - 1;
        continue;
      }
      const int orient = (int)(code >> C->oshift);
      REAL f[MAXQ], rho, v[3] = {0, 0, 0};
      /* getDist */
      if (!odd) for (int i = 0; i < Q; i++) f[i] = din[gi + DS * i];
      else      for (int i = 0; i < Q; i++) f[i] = din[gi - off[i] + DS * L->opp[i]];
      /* fixMissingDistributions, boundary.mako:574-605: copy / extrapolation outflow takes the populations that
       * would arrive from outside from dist_in of the inward neighbour(s); AB pattern only (:626-628) */
      if ((cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW) && orient != 0) {
        const int *n = L->e[L->dir2vec[orient]];
        const long long noff = n[0] + C->stride[1] * n[1] + C->stride[2] * n[2];
        for (int i = 0; i < Q; i++) {
          int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * n[d];
          if (sp <= 0) continue;
          if (cls == SLF_NT_COPY) f[i] = din[gi + noff + DS * i];
          else f[i] = (REAL)2.0 * din[gi + noff + DS * i] - din[gi + 2 * noff + DS * i];
        }
      }
      /* NTWallTMS, boundary.mako:631-648: target moments from the populations as loaded (the slots of the links into
       * the wall hold the values bounced back at the end of the previous step); those populations are then replaced
       * by the equilibrium at the target moments */
      REAL tg_rho = 0, tg_v[3] = {0, 0, 0};
      if (cls == SLF_NT_WALL_TMS) {
        for (int i = 0; i < Q; i++) tg_rho += f[i];
        for (int d = 0; d < D; d++) { REAL s = 0; for (int i = 0; i < Q; i++) s += (REAL)L->e[i][d] * f[i]; tg_v[d] = s / tg_rho; }
        REAL feq[MAXQ]; feq_all(L, tg_rho, tg_v, feq);
        for (int i = 1; i < Q; i++) {
          int missing, name;
          if (!missing_link(C, orient, i, &missing, &name)) continue;
          f[missing] = feq[missing];
        }
      }
      /* getMacro */
      const int standard = (cls == SLF_NT_FLUID || cls == SLF_NT_FULLBB_WALL || cls == SLF_NT_HALFBB_WALL ||
                            cls == SLF_NT_DO_NOTHING || cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW ||
                            cls == SLF_NT_WALL_TMS || cls == SLF_NT_SLIP);
      const int vel_bc = (cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_REGULARIZED_VELOCITY ||
                          cls == SLF_NT_ZOUHE_VELOCITY);
      const int den_bc = (cls == SLF_NT_EQUILIBRIUM_DENSITY || cls == SLF_NT_REGULARIZED_DENSITY ||
                          cls == SLF_NT_ZOUHE_DENSITY);
      if (standard || orient == 0) {
        rho = 0; for (int i = 0; i < Q; i++) rho += f[i];
        for (int d = 0; d < D; d++) { REAL s = 0; for (int i = 0; i < Q; i++) s += (REAL)L->e[i][d] * f[i]; v[d] = s / rho; }
      } else if (vel_bc) {
        const int pidx = (int)((code >> C->pshift) & C->pmask);
        fill_missing_with_opposite(L, orient, f);
        rho = 0; for (int i = 0; i < Q; i++) rho += f[i];
        const int *n = L->e[L->dir2vec[orient]];
        REAL nv = 0;
        for (int d = 0; d < D; d++) { v[d] = (REAL)P->node_params[pidx + d]; nv += (REAL)n[d] * v[d]; }
        rho = rho / ((REAL)1.0 - nv);                                  /* sym.py:627 */
      } else { /* den_bc */
        const int pidx = (int)((code >> C->pshift) & C->pmask);
        fill_missing_with_opposite(L, orient, f);
        rho = 0; for (int i = 0; i < Q; i++) rho += f[i];
        const REAL par_rho = (REAL)P->node_params[pidx];
        const int *n = L->e[L->dir2vec[orient]];
        for (int d = 0; d < D; d++) v[d] = (REAL)(-n[d]) * (rho - par_rho) / par_rho;   /* sym.py:666-680 */
        if (cls == SLF_NT_ZOUHE_DENSITY) {                              /* boundary.mako:487-492 */
          zouhe_bb(L, orient, par_rho, v, f);
          REAL r = 0; for (int i = 0; i < Q; i++) r += f[i];
          for (int d = 0; d < D; d++) { REAL q = 0; for (int i = 0; i < Q; i++) q += (REAL)L->e[i][d] * f[i]; v[d] = q / r; }
        }
        rho = par_rho;
      }
      /* precollisionBoundaryConditions */
      if (cls == SLF_NT_FULLBB_WALL) {
        bounce_back(L, f);
      } else if (cls == SLF_NT_SLIP) {                                  /* boundary.mako:837-856, sym.py:481-497 */
        if (orient != 0) {
          const int *n = L->e[L->dir2vec[orient]];
          for (int i = 0; i < Q; i++) {
            int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * n[d];
            if (sp <= 0) continue;
            int r[3] = {0, 0, 0};                                       /* e_i with its normal components reversed */
            for (int d = 0; d < D; d++) r[d] = n[d] != 0 ? -L->e[i][d] : L->e[i][d];
            for (int j = 0; j < Q; j++)
              if (L->e[j][0] == r[0] && L->e[j][1] == r[1] && L->e[j][2] == r[2]) {
                REAL t = f[i]; f[i] = f[j]; f[j] = t;
                break;
              }
          }
        }
      } else if (cls == SLF_NT_ZOUHE_VELOCITY) {                        /* boundary.mako:812-816 */
        if (orient == 0) bounce_back(L, f); else zouhe_bb(L, orient, rho, v, f);
      } else if (cls == SLF_NT_EQUILIBRIUM_VELOCITY || cls == SLF_NT_EQUILIBRIUM_DENSITY) {
        feq_all(L, rho, v, f);
      } else if (cls == SLF_NT_REGULARIZED_VELOCITY || cls == SLF_NT_REGULARIZED_DENSITY) {
        if (orient == 0) {
          bounce_back(L, f);                                           /* boundary.mako:336-338 */
        } else {
          REAL feq[MAXQ]; feq_all(L, rho, v, feq);
          const int *n = L->e[L->dir2vec[orient]];
          for (int i = 0; i < Q; i++) {                                 /* noneq_bb, sym.py:750-766 */
            int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * n[d];
            if (sp > 0) f[i] = f[L->opp[i]] - feq[L->opp[i]] + feq[i];
          }
          REAL flux[3][3];
          for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) {     /* boundary.mako:298-304 */
            REAL s = 0; for (int i = 0; i < Q; i++) s += (REAL)(L->e[i][a] * L->e[i][b]) * f[i];
            REAL eq = a == b ? rho * (v[a]*v[a] + (REAL)(1.0/3.0)) : rho * v[a] * v[b];
            flux[a][b] = s - eq;
          }
          for (int i = 0; i < Q; i++) {                                 /* sym.py:882-891 */
            REAL s = 0;
            for (int a = 0; a < D; a++) for (int b = 0; b < D; b++)
              s += flux[a][b] * ((REAL)(L->e[i][a] * L->e[i][b]) - (a == b ? (REAL)(1.0/3.0) : (REAL)0));
            REAL t = feq[i] + (REAL)L->w[i] * (REAL)4.5 * s;
            f[i] = t > (REAL)1e-7 ? t : (REAL)1e-7;
          }
        }
      } else if (cls == SLF_NT_DO_NOTHING && aa) {                      /* boundary.mako:862-877 */
        if (orient != 0) {
          const int *n = L->e[L->dir2vec[orient]];
          for (int i = 0; i < Q; i++) {
            int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * n[d];
            if (sp <= 0) continue;
            if (odd) dout[gi + DS * i] = f[i];
            else     dout[gi + off[L->opp[i]] + DS * L->opp[i]] = f[i];
          }
        }
      }
      /* relaxate */
      const int wet = (cls == SLF_NT_FLUID || cls == SLF_NT_HALFBB_WALL || vel_bc || den_bc || cls == SLF_NT_DO_NOTHING ||
                       cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW || cls == SLF_NT_WALL_TMS);
      if (wet && P->relaxation_enabled) {
        if (P->model == SLF_MRT) mrt_relax(C, cls, f, v);
        else bgk_relax(C, rho, f, v);
      }
      /* postcollisionBoundaryConditions: TMS adds f^eq(target) - f^eq(instantaneous), boundary.mako:696-722 */
      if (cls == SLF_NT_WALL_TMS) {
        REAL feq[MAXQ];
        feq_all(L, tg_rho, tg_v, feq);
        for (int i = 0; i < Q; i++) f[i] += feq[i];
        feq_all(L, rho, v, feq);
        for (int i = 0; i < Q; i++) f[i] -= feq[i];
      }
      /* half-way bounce-back and TMS: the populations of the links into the wall are stored where the next step
       * reads the ones coming out of it, boundary.mako:653-668 */
      if (cls == SLF_NT_HALFBB_WALL || cls == SLF_NT_WALL_TMS) {
        for (int i = 1; i < Q; i++) {
          int missing;   /* slot whose data is undefined; f[name] is written into it */
          int name;
          if (!missing_link(C, orient, i, &missing, &name)) continue;
          if (!aa || odd) dout[gi + DS * missing] = f[name];
          else dout[gi + off[L->opp[missing]] + DS * L->opp[missing]] = f[name];
        }
      }
      /* save_macro_fields */
      if ((options & SLF_OPTION_SAVE_MACRO_FIELDS) && wet) {
        orho[gi] = rho;
        for (int d = 0; d < D; d++) ov[d][gi] = v[d];
      }
      /* propagate */
      if (!P->propagation_enabled) {
        for (int i = 0; i < Q; i++) dout[gi + DS * i] = f[i];
      } else if (aa && !odd) {
        for (int i = 0; i < Q; i++) dout[gi + DS * L->opp[i]] = f[i];
      } else {
        for (int i = 0; i < Q; i++)
          if (in_bounds_target(C, L->e[i], x, y, z)) dout[gi + off[i] + DS * i] = f[i];
      }
    }
  }
  return 0;
}

/* SetInitialConditions, lb_single_fluid.mako:101-127: every node with gx < lat_nx, no type test */
int orc_set_initial_conditions(const SlfParams *P, REAL *dist, const REAL *ivx, const REAL *ivy,
                               const REAL *ivz, const REAL *irho, const uint32_t *map) {
  (void)map;
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L;
  /* same row decomposition and static schedule as the step: on a multi-socket host every thread
   * first-touches the rows it will later stream through */
  const long long rows = (long long)C->nz * C->ny;
#pragma omp parallel for schedule(static)
  for (long long row = 0; row < rows; row++) {
    const int z = (int)(row / C->ny), y = (int)(row % C->ny);
    for (int x = 0; x < C->nx; x++) {
      long long gi = x + C->stride[1] * y + C->stride[2] * z;
      REAL feq[MAXQ], v[3] = {ivx[gi], ivy[gi], L->D == 3 ? ivz[gi] : (REAL)0};
      feq_all(L, irho[gi], v, feq);
      for (int i = 0; i < L->Q; i++) dist[gi + C->DS * i] = feq[i];
    }
  }
  return 0;
}

/* PrepareMacroFields, lb_single_fluid.mako:129-159 (density only) */
int orc_prepare_macro_fields(const SlfParams *P, const uint32_t *map, const REAL *din, REAL *orho,
                             int options, uint32_t iteration) {
  (void)options;
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L; const int Q = L->Q;
  const int odd = P->access_pattern == SLF_AA && (iteration & 1);
  for (int z = 0; z < C->nz; z++) for (int y = 0; y < C->ny; y++) for (int x = 0; x < C->nx; x++) {
    long long gi = x + C->stride[1] * y + C->stride[2] * z;
    int cls = nclass(C, map[gi] & C->tmask);
    if (cls < 0 || cls == SLF_NT_GHOST || cls == SLF_NT_UNUSED || cls == SLF_NT_PROPAGATION_ONLY) continue;
    REAL rho = 0;
    for (int i = 0; i < Q; i++) {
      long long off = L->e[i][0] + C->stride[1] * L->e[i][1] + C->stride[2] * L->e[i][2];
      rho += odd ? din[gi - off + C->DS * L->opp[i]] : din[gi + C->DS * i];
    }
    orho[gi] = rho;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * ApplyPeriodicBoundaryConditions[WithSwap], kernel_utils.mako:19-384.
 * The template generates one condition chain per population at render time; here the
 * same rules are evaluated at run time from (basis, block_periodicity, bnd_limits).
 * ------------------------------------------------------------------------------------ */
static inline long long gidx(const Ctx *C, const int *c) { return c[0] + C->stride[1] * c[1] + C->stride[2] * c[2]; }

/* make_cond(i), kernel_utils.mako:120-171 */
static int pbc_make_cond(const Ctx *C, int axis, int opposite, const int *e, const int *oth, int nother,
                         const int *idx, const int *maxd, int link_lo, int link_hi) {
  const int *per = C->P->periodic;
  if (opposite) {
    for (int k = 0; k < nother; k++) if (!(idx[k] >= 1 && idx[k] <= maxd[k])) return 0;
    return 1;
  }
  if (C->L->D == 2) {
    /* The reference's 2-D rule stops at the last real row.  With a neighbour across a y face (native plane push,
     * axis == 0 only) the ghost row on that side holds this subdomain's outgoing populations, and the ones that also
     * crossed the periodic x face must be wrapped inside it before the row leaves -- what the 3-D rule (no bound on a
     * non-periodic other axis) does for z-splits. */
    int c = e[1 - axis];
    const int split_other = axis == 0;
    if (c == 1) return idx[0] > 1 && (idx[0] <= maxd[0] || (split_other && link_hi && idx[0] == maxd[0] + 1));
    if (c == -1) return idx[0] < maxd[0] && (idx[0] >= 1 || (split_other && link_lo && idx[0] == 0));
    return 1;
  }
  for (int k = 0; k < 2; k++) {
    int c = e[oth[k]];
    int both = per[axis] && per[oth[k]];
    if (c == 1) { if (!(idx[k] > 1)) return 0; if (both && !(idx[k] <= maxd[k])) return 0; }
    else if (c == -1) { if (!(idx[k] < maxd[k])) return 0; if (both && !(idx[k] >= 1)) return 0; }
  }
  return 1;
}

/* handle_corners, kernel_utils.mako:32-118.  Returns 1 and fills `tgt` with the node the
 * value must be copied to when one of the generated conditions matches (first match wins). */
static int pbc_corner_target(const Ctx *C, int axis, int opposite, const int *e, const int *oth, int nother,
                             const int *idx, int axis_target, int *tgt) {
  const int *per = C->P->periodic; const int lim[3] = {C->nx, C->ny, C->nz};
  /* entry 0: corner nodes */
  {
    int ncond = 0, ok = 1, t[3] = {0, 0, 0};
    t[axis] = axis_target;
    for (int k = 0; k < nother; k++) {
      int oa = oth[k], c = e[oa];
      if (c == 0 || !per[oa]) { t[oa] = idx[k]; continue; }
      if (c == 1) { int shift = opposite ? 2 : 1; ncond++; if (idx[k] != lim[oa] - shift) ok = 0; t[oa] = opposite ? 0 : 1; }
      else { int shift = opposite ? 1 : 0; ncond++; if (idx[k] != shift) ok = 0; t[oa] = lim[oa] - (opposite ? 1 : 2); }
    }
    if (ncond > 0 && ok) { tgt[0] = t[0]; tgt[1] = t[1]; tgt[2] = t[2]; return 1; }
  }
  /* edge entries: only for vectors with three non-zero components (D3Q27) */
  if (abs(e[0]) + abs(e[1]) + abs(e[2]) == 3) {
    for (int k = 0; k < 2; k++) {
      int oa = oth[k], ob = oth[1 - k], id_ = idx[k], id2 = idx[1 - k];
      int ok = 1, t[3] = {0, 0, 0};
      t[axis] = axis_target;
      if (e[oa] == 1) { int shift = opposite ? 2 : 1, start = opposite ? 1 : 2; if (!(id_ >= start && id_ < lim[oa] - shift)) ok = 0; }
      else { int shift = opposite ? 1 : 2; if (!(id_ < lim[oa] - shift && id_ >= 1)) ok = 0; }
      t[oa] = id_;
      if (e[ob] == 1) { int shift = opposite ? 2 : 1; if (id2 != lim[ob] - shift) ok = 0; t[ob] = opposite ? 0 : 1; }
      else { int shift = opposite ? 1 : 0; if (id2 != shift) ok = 0; t[ob] = lim[ob] - (opposite ? 1 : 2); }
      if (ok) { tgt[0] = t[0]; tgt[1] = t[1]; tgt[2] = t[2]; return 1; }
    }
  }
  return 0;
}

/* The reference's edge entries for vectors with three non-zero components (handle_corners, kernel_utils.mako:84-118)
 * wrap along the second other axis whether or not that axis is periodic.  In a single subdomain the stray copy lands
 * in a ghost layer nobody reads; with a neighbour across that face the ghost layer is, after in-place steps, where the
 * neighbour's planes land -- it is left alone like the rest of that layer. */
static inline int pbc_target_is_landing_zone(int with_swap, int t_split, int n_split, int link_lo, int link_hi) {
  return with_swap && ((link_lo && t_split == 0) || (link_hi && t_split == n_split - 1));
}

/* link_lo / link_hi: this subdomain has a neighbour across the low / high face of the outermost (split) axis
 * that delivers whole, PBC-completed planes (slf_halo_push): after in-place steps (with_swap) into the ghost
 * layer, after streaming steps into the crossing slots of the first / last real layer.  Those cells belong to
 * the neighbour and are left alone (same rule as apply_pbc in slf_kernels.cuh). */
int orc_apply_periodic_boundary_conditions_linked(const SlfParams *P, REAL *dist, int axis, int with_swap,
                                                  int link_lo, int link_hi) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L; const int D = L->D, Q = L->Q;
  if (axis < 0 || axis >= D) return D == 2 && axis == 2 ? 0 : SLF_EINVAL;
  const int lim[3] = {C->nx, C->ny, C->nz};
  int oth[2], nother = D - 1;
  if (D == 2) { oth[0] = 1 - axis; oth[1] = 2; }
  else { const int oa[3][2] = {{1, 2}, {0, 2}, {0, 1}}; oth[0] = oa[axis][0]; oth[1] = oa[axis][1]; }
  const int maxd[2] = {lim[oth[0]] - 2, D == 3 ? lim[oth[1]] - 2 : 0};
  const long long offset = with_swap ? -C->stride[axis] : C->stride[axis];
  const int n1 = lim[oth[0]], n2 = D == 3 ? lim[oth[1]] : 1;
  for (int i2 = 0; i2 < n2; i2++) for (int i1 = 0; i1 < n1; i1++) {
    const int idx[2] = {i1, i2};
    if (with_swap && D == 3) { if (i1 >= lim[oth[0]] - 1 || i2 >= lim[oth[1]] - 1 || i1 < 1 || i2 < 1) continue; }
    int cl[3] = {0, 0, 0}, ch[3] = {0, 0, 0};
    cl[oth[0]] = ch[oth[0]] = i1; if (D == 3) cl[oth[1]] = ch[oth[1]] = i2;
    cl[axis] = with_swap ? 1 : 0;
    ch[axis] = with_swap ? lim[axis] - 1 : lim[axis] - 2;
    const long long gi_low = gidx(C, cl), gi_high = gidx(C, ch);
    const int so = D - 1;                                   /* split axis */
    const int io = oth[0] == so ? i1 : ((D == 3 && oth[1] == so) ? i2 : -1), no = lim[so];
    if (with_swap && io >= 0 && ((link_lo && io == 0) || (link_hi && io == no - 1))) continue;
    const int skip_e = (with_swap || io < 0) ? 0 : ((link_lo && io == 1) ? 1 : ((link_hi && io == no - 2) ? -1 : 0));
    /* low -> high: populations with e[axis] == -1 */
    REAL val[MAXQ];
    for (int i = 1; i < Q; i++) if (L->e[i][axis] == -1) { int j = with_swap ? L->opp[i] : i; val[i] = dist[gi_low + C->DS * j]; }
    for (int i = 1; i < Q; i++) if (L->e[i][axis] == -1) {
      if (skip_e != 0 && L->e[i][so] == skip_e) continue;
      const int j = with_swap ? L->opp[i] : i; const REAL fv = val[i];
      const int n2e = L->e[i][0]*L->e[i][0] + L->e[i][1]*L->e[i][1] + L->e[i][2]*L->e[i][2];
      if (!isfinite(fv)) continue;
      if (n2e > 1) {
        int main_ok = pbc_make_cond(C, axis, with_swap, L->e[i], oth, nother, idx, maxd, link_lo, link_hi);
        if (main_ok) dist[gi_high + C->DS * j] = fv;
        if (with_swap || !main_ok) {
          int tgt[3];
          if (pbc_corner_target(C, axis, with_swap, L->e[i], oth, nother, idx, with_swap ? lim[axis] - 1 : lim[axis] - 2, tgt) &&
              !pbc_target_is_landing_zone(with_swap, tgt[so], no, link_lo, link_hi))
            dist[gidx(C, tgt) + C->DS * j] = fv;
        }
      } else {
        dist[gi_high + C->DS * j] = fv;
      }
    }
    /* high -> low: populations with e[axis] == +1 */
    for (int i = 1; i < Q; i++) if (L->e[i][axis] == 1) { int j = with_swap ? L->opp[i] : i; val[i] = dist[gi_high + offset + C->DS * j]; }
    for (int i = 1; i < Q; i++) if (L->e[i][axis] == 1) {
      if (skip_e != 0 && L->e[i][so] == skip_e) continue;
      const int j = with_swap ? L->opp[i] : i; const REAL fv = val[i];
      const int n2e = L->e[i][0]*L->e[i][0] + L->e[i][1]*L->e[i][1] + L->e[i][2]*L->e[i][2];
      if (!isfinite(fv)) continue;
      if (n2e > 1) {
        int main_ok = pbc_make_cond(C, axis, with_swap, L->e[i], oth, nother, idx, maxd, link_lo, link_hi);
        if (main_ok) dist[gi_low + offset + C->DS * j] = fv;
        if (with_swap || !main_ok) {
          int tgt[3];
          if (pbc_corner_target(C, axis, with_swap, L->e[i], oth, nother, idx, with_swap ? 0 : 1, tgt) &&
              !pbc_target_is_landing_zone(with_swap, tgt[so], no, link_lo, link_hi))
            dist[gidx(C, tgt) + C->DS * j] = fv;
        }
      } else {
        dist[gi_low + offset + C->DS * j] = fv;
      }
    }
  }
  return 0;
}

int orc_apply_periodic_boundary_conditions(const SlfParams *P, REAL *dist, int axis, int with_swap) {
  return orc_apply_periodic_boundary_conditions_linked(P, dist, axis, with_swap, 0, 0);
}

/* ------------------------------------------------------------------------------------
 * Collect/DistributeContinuousData[WithSwap], kernel_utils.mako:476-794.
 * Buffer layout [dist][other][x]; `face` >= 2 (y/z faces).  In 2-D pass max_other = 1.
 * ------------------------------------------------------------------------------------ */
static int face_dists(const Lat *L, int face, int *out) {
  /* sym.get_interblock_dists(grid, face_to_normal(face)), sym.py:829-845 */
  int axis = face / 2, sgn = (face & 1) ? 1 : -1, n = 0;
  for (int i = 0; i < L->Q; i++) if (L->e[i][axis] * sgn >= 1) out[n++] = i;
  return n;
}

static int face_layer(const Ctx *C, int face, int which) {
  /* which: 0 lat_linear, 1 lat_linear_macro, 2 lat_linear_dist, 3 lat_linear_with_swap
   * subdomain_runner.py:488-510 (envelope 1) */
  const int lim[3] = {C->nx, C->ny, C->nz}; int axis = face / 2, hi = face & 1, n = lim[axis];
  switch (which) {
    case 0: return hi ? n - 1 : 0;
    case 1: return hi ? n - 2 : 1;
    case 2: return hi ? 1 : n - 2;
    default: return hi ? 0 : n - 1;
  }
}

static int xfer(const SlfParams *P, REAL *dist, int face, int base_gx, int base_other, int max_lx,
                int max_other, REAL *buffer, int with_swap, int distribute) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L;
  if (face < 2 || face >= 2 * L->D) return SLF_EINVAL;
  int dl[MAXQ], nd = face_dists(L, face, dl);
  int layer = face_layer(C, face, distribute ? (with_swap ? 3 : 2) : (with_swap ? 1 : 0));
  int axis = face / 2;
  if (L->D == 2) max_other = 1;
  for (int k = 0; k < nd; k++) for (int o = 0; o < max_other; o++) for (int x = 0; x < max_lx; x++) {
    int c[3] = {base_gx + x, 0, 0};
    if (L->D == 2) c[1] = layer;
    else if (axis == 1) { c[1] = layer; c[2] = base_other + o; }
    else { c[1] = base_other + o; c[2] = layer; }
    int slot = with_swap ? L->opp[dl[k]] : dl[k];
    long long di = gidx(C, c) + C->DS * slot;
    long long bi = ((long long)k * max_other + o) * max_lx + x;
    if (distribute) dist[di] = buffer[bi]; else buffer[bi] = dist[di];
  }
  return 0;
}

int orc_collect_continuous_data(const SlfParams *P, const REAL *dist, int face, int base_gx, int base_other,
                                int max_lx, int max_other, REAL *buffer, int with_swap) {
  return xfer(P, (REAL *)dist, face, base_gx, base_other, max_lx, max_other, buffer, with_swap, 0);
}
int orc_distribute_continuous_data(const SlfParams *P, REAL *dist, int face, int base_gx, int base_other,
                                   int max_lx, int max_other, const REAL *buffer, int with_swap) {
  return xfer(P, dist, face, base_gx, base_other, max_lx, max_other, (REAL *)buffer, with_swap, 1);
}

/* Collect/DistributeContinuousMacroData, kernel_utils.mako:838-953: collect reads layer
 * lat_linear_macro[face] (first/last real layer), distribute writes lat_linear[face] (ghost layer);
 * 2-D kernels take the row explicitly (passed in base_other).  Buffer [other][x], dense rows. */
static int xfer_macro(const SlfParams *P, REAL *field, int face, int base_gx, int base_other, int max_lx,
                      int max_other, REAL *buffer, int distribute) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L;
  if (L->D == 2) max_other = 1;
  else if (face < 2 || face > 5) return SLF_EINVAL;
  for (int o = 0; o < max_other; o++) for (int x = 0; x < max_lx; x++) {
    int c[3] = {base_gx + x, 0, 0};
    if (L->D == 2) c[1] = base_other;
    else {
      int layer = face_layer(C, face, distribute ? 0 : 1);
      if (face / 2 == 1) { c[1] = layer; c[2] = base_other + o; } else { c[1] = base_other + o; c[2] = layer; }
    }
    long long gi = gidx(C, c), bi = (long long)o * max_lx + x;
    if (distribute) field[gi] = buffer[bi]; else buffer[bi] = field[gi];
  }
  return 0;
}
int orc_collect_continuous_macro_data(const SlfParams *P, const REAL *field, int face, int base_gx, int base_other,
                                      int max_lx, int max_other, REAL *buffer) {
  return xfer_macro(P, (REAL *)field, face, base_gx, base_other, max_lx, max_other, buffer, 0);
}
int orc_distribute_continuous_macro_data(const SlfParams *P, REAL *field, int face, int base_gx, int base_other,
                                         int max_lx, int max_other, const REAL *buffer) {
  return xfer_macro(P, field, face, base_gx, base_other, max_lx, max_other, (REAL *)buffer, 1);
}

int orc_sizeof_real(void) { return (int)sizeof(REAL); }

/* CPU-baseline plumbing (bench.py): the OpenMP runtime may have been initialised by another library
 * with OMP_NUM_THREADS=1 (torchrun exports that), so the thread count is set -- and reported -- here,
 * and each thread is pinned to one of the CPUs the process may use (first touch in
 * orc_set_initial_conditions then places every thread's rows on its own NUMA node). */
#ifdef _OPENMP
#include <omp.h>
int orc_omp_config(int nthreads, const int *cpus, int ncpu, int pin) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  int used = 1;
#pragma omp parallel
  {
#pragma omp single
    used = omp_get_num_threads();
    if (cpus && ncpu > 0) {
      cpu_set_t set;
      CPU_ZERO(&set);
      if (pin) CPU_SET(cpus[omp_get_thread_num() % ncpu], &set);        /* thread k on the k-th CPU of the list */
      else for (int c = 0; c < ncpu; c++) CPU_SET(cpus[c], &set);         /* free to move within the list */
      sched_setaffinity(0, sizeof(set), &set);
    }
  }
  return used;
}
#else
int orc_omp_config(int nthreads, const int *cpus, int ncpu, int pin) { (void)nthreads; (void)cpus; (void)ncpu; (void)pin; return 1; }
#endif

/* ------------------------------------------------------------------------------------
 * Formula-level hooks for tests/test_oracle_formulas.py (golden: tests/golden/formulas.npz,
 * produced by lambdifying the reference's SymPy expressions).  n independent samples;
 * arrays are [component][sample].
 * ------------------------------------------------------------------------------------ */
int orc_eval_feq(int lattice, int n, const REAL *rho, const REAL *v, REAL *out) {
  lat_init();
  if (lattice < 0 || lattice > 2) return SLF_EINVAL;
  const Lat *L = &g_lat[lattice];
  for (int s = 0; s < n; s++) {
    REAL vv[3] = {0, 0, 0}, f[MAXQ];
    for (int d = 0; d < L->D; d++) vv[d] = v[d * n + s];
    feq_all(L, rho[s], vv, f);
    for (int i = 0; i < L->Q; i++) out[i * n + s] = f[i];
  }
  return 0;
}

/* Guo term added to f_i (sym_force.py:121-160) for velocity v (already shifted by a/2),
 * acceleration a and relaxation time tau. */
int orc_eval_guo(int lattice, int n, const REAL *rho, const REAL *v, const REAL *a, REAL tau, REAL *out) {
  lat_init();
  if (lattice < 0 || lattice > 2) return SLF_EINVAL;
  const Lat *L = &g_lat[lattice];
  for (int s = 0; s < n; s++) {
    REAL pref = rho[s] * (REAL)3.0 * ((REAL)1.0 - (REAL)1.0 / ((REAL)2.0 * tau));
    for (int i = 0; i < L->Q; i++) {
      REAL ev = 0, t = 0;
      for (int d = 0; d < L->D; d++) ev += (REAL)L->e[i][d] * v[d * n + s];
      for (int d = 0; d < L->D; d++) t += ((REAL)L->e[i][d] - v[d * n + s] + (REAL)3.0 * ev * (REAL)L->e[i][d]) * a[d * n + s];
      out[i * n + s] = pref * (REAL)L->w[i] * t;
    }
  }
  return 0;
}

/* Moments m = M f, their equilibria and the relaxation rates (relaxation_mrt.mako:31-57). */
int orc_eval_mrt(int lattice, int n, const REAL *f, double visc, REAL *m_out, REAL *meq_out, REAL *rates_out) {
  lat_init();
  if (lattice != SLF_D2Q9 && lattice != SLF_D3Q19) return SLF_EUNSUPPORTED;
  const Lat *L = &g_lat[lattice]; const int Q = L->Q;
  REAL inv_tau = (REAL)1.0 / ((REAL)0.5 + (REAL)3.0 * (REAL)visc);
  for (int k = 0; k < Q; k++) rates_out[k] = L->rates[k] < 0 ? inv_tau : (REAL)L->rates[k];
  for (int s = 0; s < n; s++) {
    REAL m[MAXQ], meq[MAXQ];
    for (int k = 0; k < Q; k++) { REAL t = 0; for (int i = 0; i < Q; i++) t += (REAL)L->M[k][i] * f[i * n + s]; m[k] = t; meq[k] = 0; }
    REAL rho = m[0], mx = m[L->i_mx], my = m[L->i_my], mz = L->D == 3 ? m[L->i_mz] : 0;
    if (lattice == SLF_D2Q9) {
      meq[1] = (REAL)-2.0 * rho + (REAL)3.0 * (mx*mx + my*my); meq[2] = rho - (REAL)3.0 * (mx*mx + my*my);
      meq[4] = -mx; meq[6] = -my; meq[7] = mx*mx - my*my; meq[8] = mx*my;
    } else {
      REAL ir = (REAL)1.0 / rho, m2 = mx*mx + my*my + mz*mz;
      meq[1] = (REAL)-11.0 * rho + (REAL)19.0 * ir * m2; meq[2] = (REAL)(-475.0/63.0) * ir * m2;
      meq[4] = (REAL)(-2.0/3.0) * mx; meq[6] = (REAL)(-2.0/3.0) * my; meq[8] = (REAL)(-2.0/3.0) * mz;
      meq[9] = ir * ((REAL)2.0*mx*mx - (my*my + mz*mz)); meq[11] = ir * (my*my - mz*mz);
      meq[13] = ir * mx*my; meq[14] = ir * my*mz; meq[15] = ir * mx*mz;
    }
    for (int k = 0; k < Q; k++) { m_out[k * n + s] = m[k]; meq_out[k * n + s] = meq[k]; }
  }
  return 0;
}

/* Velocity-/density-node closures and the regularised boundary step for one orientation
 * (boundary.mako:418-459, 817-835; sym.py:573-682, 750-766, 882-891).
 * mode 0: out_rho = rho/(1 - n.v) for the velocity BC (f already filled)           -> rho_out
 * mode 1: v = -n (rho - par_rho)/par_rho for the density BC                          -> v_out
 * mode 2: non-equilibrium bounce-back of f for (rho, v)                              -> f_out
 * mode 3: regularised populations max(1e-7, feq + w 9/2 Q:Pi_neq) from f, rho, v     -> f_out */
int orc_eval_bc(int lattice, int mode, int orientation, int n, const REAL *f, const REAL *rho, const REAL *v,
                REAL par_rho, REAL *rho_out, REAL *v_out, REAL *f_out) {
  lat_init();
  if (lattice < 0 || lattice > 2) return SLF_EINVAL;
  const Lat *L = &g_lat[lattice]; const int Q = L->Q, D = L->D;
  const int *nn = L->e[L->dir2vec[orientation]];
  for (int s = 0; s < n; s++) {
    REAL vv[3] = {0, 0, 0}, ff[MAXQ], feq[MAXQ];
    for (int d = 0; d < D; d++) vv[d] = v ? v[d * n + s] : 0;
    if (f) for (int i = 0; i < Q; i++) ff[i] = f[i * n + s];
    if (mode == 0) {
      REAL nv = 0; for (int d = 0; d < D; d++) nv += (REAL)nn[d] * vv[d];
      rho_out[s] = rho[s] / ((REAL)1.0 - nv);
    } else if (mode == 1) {
      for (int d = 0; d < D; d++) v_out[d * n + s] = (REAL)(-nn[d]) * (rho[s] - par_rho) / par_rho;
    } else if (mode == 2) {
      feq_all(L, rho[s], vv, feq);
      for (int i = 0; i < Q; i++) {
        int sp = 0; for (int d = 0; d < D; d++) sp += L->e[i][d] * nn[d];
        f_out[i * n + s] = sp > 0 ? ff[L->opp[i]] - feq[L->opp[i]] + feq[i] : ff[i];
      }
    } else {
      feq_all(L, rho[s], vv, feq);
      REAL flux[3][3];
      for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) {
        REAL t = 0; for (int i = 0; i < Q; i++) t += (REAL)(L->e[i][a] * L->e[i][b]) * ff[i];
        flux[a][b] = t - (a == b ? rho[s] * (vv[a]*vv[a] + (REAL)(1.0/3.0)) : rho[s] * vv[a] * vv[b]);
      }
      for (int i = 0; i < Q; i++) {
        REAL t = 0;
        for (int a = 0; a < D; a++) for (int b = 0; b < D; b++)
          t += flux[a][b] * ((REAL)(L->e[i][a] * L->e[i][b]) - (a == b ? (REAL)(1.0/3.0) : (REAL)0));
        REAL r = feq[i] + (REAL)L->w[i] * (REAL)4.5 * t;
        f_out[i * n + s] = r > (REAL)1e-7 ? r : (REAL)1e-7;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Binary Shan-Chen mixture (BASELINE config 5).  Reference: lb_binary.py:375-517 (model class,
 * kernel arguments), templates/models/lb_binary_fluid.mako:88-127 (SetInitialConditions),
 * templates/models/binary_shan_chen.mako:19-141 (ShanChenPrepareMacroFields,
 * ShanChenCollideAndPropagate{0,1}), templates/shan_chen.mako:9-84 (force),
 * relaxation.mako:99-181 with relaxation_common.mako (BGK_relaxate{0,1} + Guo forcing),
 * kernel_utils.mako:410-474 (ApplyMacroPeriodicBoundaryConditions).
 * ------------------------------------------------------------------------------------ */
static inline REAL sc_ppot(int pot, REAL x) {       /* sym.py:896-908 */
  return pot == SLF_SC_CLASSIC ? (REAL)1.0 * ((REAL)1.0 - (REAL)exp(-(double)x / 1.0)) : x;
}

static inline int is_wet(int cls) {
  return cls == SLF_NT_FLUID || cls == SLF_NT_HALFBB_WALL || cls == SLF_NT_EQUILIBRIUM_VELOCITY ||
         cls == SLF_NT_EQUILIBRIUM_DENSITY || cls == SLF_NT_REGULARIZED_VELOCITY ||
         cls == SLF_NT_REGULARIZED_DENSITY || cls == SLF_NT_DO_NOTHING || cls == SLF_NT_ZOUHE_VELOCITY ||
         cls == SLF_NT_ZOUHE_DENSITY || cls == SLF_NT_COPY || cls == SLF_NT_YU_OUTFLOW || cls == SLF_NT_WALL_TMS;
}

int orc_sc_set_initial_conditions(const SlfParams *P, const uint32_t *map, REAL *d1, REAL *d2, const REAL *ivx,
                                  const REAL *ivy, const REAL *ivz, const REAL *irho, const REAL *iphi) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L;
  for (int z = 0; z < C->nz; z++) for (int y = 0; y < C->ny; y++) for (int x = 0; x < C->nx; x++) {
    long long gi = x + C->stride[1] * y + C->stride[2] * z;
    int cls = nclass(C, map[gi] & C->tmask);
    if (!is_wet(cls) && cls != SLF_NT_FULLBB_WALL) {
      for (int i = 0; i < L->Q; i++) { d1[gi + C->DS * i] = (REAL)INFINITY; d2[gi + C->DS * i] = (REAL)INFINITY; }
      continue;
    }
    REAL feq[MAXQ], v[3] = {ivx[gi], ivy[gi], L->D == 3 ? ivz[gi] : (REAL)0};
    feq_all(L, irho[gi], v, feq);
    for (int i = 0; i < L->Q; i++) d1[gi + C->DS * i] = feq[i];
    feq_all(L, iphi[gi], v, feq);
    for (int i = 0; i < L->Q; i++) d2[gi + C->DS * i] = feq[i];
  }
  return 0;
}

int orc_sc_prepare_macro_fields(const SlfParams *P, const uint32_t *map, const REAL *d1, const REAL *d2,
                                REAL *orho, REAL *ophi, REAL *ovx, REAL *ovy, REAL *ovz, int options,
                                uint32_t iteration, int z_begin, int z_end) {
  (void)options;
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L; const int Q = L->Q, D = L->D;
  REAL *ov[3] = {ovx, ovy, ovz};
  const int odd = P->access_pattern == SLF_AA && (iteration & 1);
  const REAL tau0_inv = (REAL)1.0 / C->tau, tau1_inv = (REAL)1.0 / (REAL)P->tau_phi;
  const int outer_n = D == 3 ? C->nz : C->ny;
  if (z_end < 0) z_end = outer_n;
  long long off[MAXQ];
  for (int i = 0; i < Q; i++) off[i] = L->e[i][0] + C->stride[1] * L->e[i][1] + C->stride[2] * L->e[i][2];
  const int rows_per_o = D == 3 ? C->ny : 1;
#pragma omp parallel for schedule(static)
  for (long long row = (long long)z_begin * rows_per_o; row < (long long)z_end * rows_per_o; row++) {
    const int o = (int)(row / rows_per_o), z = D == 3 ? o : 0, y = D == 3 ? (int)(row % rows_per_o) : o;
    for (int x = 0; x < C->nx; x++) {
      const long long gi = x + C->stride[1] * y + C->stride[2] * z;
      const int cls = nclass(C, map[gi] & C->tmask);
      if (!is_wet(cls)) continue;
      if (cls == SLF_NT_EQUILIBRIUM_DENSITY) continue;   /* pressure nodes keep their density */
      REAL v[3] = {0, 0, 0}, rho[2];
      const REAL *dd[2] = {d1, d2}; const REAL ti[2] = {tau0_inv, tau1_inv};
      for (int g = 0; g < 2; g++) {
        REAL f[MAXQ], r = 0;
        for (int i = 0; i < Q; i++) f[i] = odd ? dd[g][gi - off[i] + C->DS * L->opp[i]] : dd[g][gi + C->DS * i];
        for (int i = 0; i < Q; i++) r += f[i];
        rho[g] = r;
        for (int d = 0; d < D; d++) {
          REAL s = 0; for (int i = 0; i < Q; i++) s += (REAL)L->e[i][d] * f[i];
          v[d] = g == 0 ? ti[g] * s : v[d] + ti[g] * s;      /* compute_1st_moment(fi, v, add, factor) */
        }
      }
      orho[gi] = rho[0]; ophi[gi] = rho[1];
      const REAL total = tau0_inv * rho[0] + tau1_inv * rho[1];
      for (int d = 0; d < D; d++) ov[d][gi] = v[d] / total;
    }
  }
  return 0;
}

int orc_sc_collide_and_propagate(const SlfParams *P, int grid_idx, const uint32_t *map, const REAL *din, REAL *dout,
                                 const REAL *rho_f, const REAL *phi_f, const REAL *ovx, const REAL *ovy,
                                 const REAL *ovz, int options, uint32_t iteration, int z_begin, int z_end) {
  (void)options;
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const Lat *L = C->L; const int Q = L->Q, D = L->D;
  const REAL *ov[3] = {ovx, ovy, ovz};
  const int aa = P->access_pattern == SLF_AA, odd = aa && (iteration & 1);
  const int outer_n = D == 3 ? C->nz : C->ny;
  if (z_end < 0) z_end = outer_n;
  long long off[MAXQ];
  for (int i = 0; i < Q; i++) off[i] = L->e[i][0] + C->stride[1] * L->e[i][1] + C->stride[2] * L->e[i][2];
  const REAL *own = grid_idx == 0 ? rho_f : phi_f;
  const REAL *fields[2] = {rho_f, phi_f};
  const REAL G[2] = {(REAL)P->sc_g[2 * grid_idx + 0], (REAL)P->sc_g[2 * grid_idx + 1]};
  const double tau_g = grid_idx == 0 ? (6.0 * P->visc + 1.0) / 2.0 : P->tau_phi;
  const REAL omega = (REAL)(1.0 / tau_g), tauf = (REAL)tau_g;
  const int rows_per_o = D == 3 ? C->ny : 1;
#pragma omp parallel for schedule(static)
  for (long long row = (long long)z_begin * rows_per_o; row < (long long)z_end * rows_per_o; row++) {
    const int o = (int)(row / rows_per_o), z = D == 3 ? o : 0, y = D == 3 ? (int)(row % rows_per_o) : o;
    for (int x = 0; x < C->nx; x++) {
      const long long gi = x + C->stride[1] * y + C->stride[2] * z;
      const int cls = nclass(C, map[gi] & C->tmask);
      if (cls < 0 || cls == SLF_NT_GHOST || cls == SLF_NT_UNUSED || cls == SLF_NT_PROPAGATION_ONLY) continue;
      const int wet = is_wet(cls);
      const REAL gm = own[gi];
      REAL a[3] = {0, 0, 0};
      if (wet) {                                     /* sc_calculate_force, shan_chen.mako:9-25 */
        for (int k = 0; k < 2; k++) {
          if (G[k] == (REAL)0) continue;
          REAL force[3] = {0, 0, 0};
          for (int i = 1; i < Q; i++) {               /* shan_chen_force, :45-84 */
            const REAL psi = sc_ppot(P->sc_potential, fields[k][gi + off[i]]);
            for (int d = 0; d < D; d++) if (L->e[i][d]) force[d] += psi * (REAL)(L->e[i][d] * L->w[i]);
          }
          const REAL psi0 = sc_ppot(P->sc_potential, gm);
          for (int d = 0; d < D; d++) a[d] += force[d] * (-psi0 * G[k]);
        }
        for (int d = 0; d < D; d++) a[d] /= gm;
        /* body_force(), relaxation_common.mako:24-33: the constant acceleration of this component is
         * added to the Shan-Chen one */
        for (int d = 0; d < D; d++)
          a[d] += grid_idx == 0 ? (P->force_enabled ? (REAL)P->accel[d] : (REAL)0) : (REAL)P->accel1[d];
      }
      REAL f[MAXQ], v[3] = {0, 0, 0};
      for (int i = 0; i < Q; i++) f[i] = odd ? din[gi - off[i] + C->DS * L->opp[i]] : din[gi + C->DS * i];
      for (int d = 0; d < D; d++) v[d] = ov[d][gi];
      if (cls == SLF_NT_FULLBB_WALL) bounce_back(L, f);
      if (wet && P->relaxation_enabled) {             /* BGK_relaxate{grid_idx} with Guo forcing */
        REAL v0[3] = {0, 0, 0}, feq[MAXQ];
        for (int d = 0; d < D; d++) v0[d] = v[d] + (REAL)0.5 * a[d];
        feq_all(L, gm, v0, feq);
        for (int i = 0; i < Q; i++) f[i] += omega * (feq[i] - f[i]);
        const REAL pref = gm * (REAL)3.0 * ((REAL)1.0 - (REAL)1.0 / ((REAL)2.0 * tauf));
        for (int i = 0; i < Q; i++) {
          REAL ev = 0, s = 0;
          for (int d = 0; d < D; d++) ev += (REAL)L->e[i][d] * v0[d];
          for (int d = 0; d < D; d++) s += ((REAL)L->e[i][d] - v0[d] + (REAL)3.0 * ev * (REAL)L->e[i][d]) * a[d];
          f[i] += pref * (REAL)L->w[i] * s;
        }
      }
      if (!P->propagation_enabled) {
        for (int i = 0; i < Q; i++) dout[gi + C->DS * i] = f[i];
      } else if (aa && !odd) {
        for (int i = 0; i < Q; i++) dout[gi + C->DS * L->opp[i]] = f[i];
      } else {
        for (int i = 0; i < Q; i++) if (in_bounds_target(C, L->e[i], x, y, z)) dout[gi + off[i] + C->DS * i] = f[i];
      }
    }
  }
  return 0;
}

int orc_apply_macro_periodic_boundary_conditions_linked(const SlfParams *P, REAL *field, int axis, int link_lo,
                                                        int link_hi) {
  Ctx Cs; int rc = ctx_init(&Cs, P); if (rc) return rc;
  const Ctx *C = &Cs; const int D = C->L->D;
  if (axis < 0 || axis >= D) return D == 2 && axis == 2 ? 0 : SLF_EINVAL;
  const int lim[3] = {C->nx, C->ny, C->nz};
  int o1 = axis == 0 ? 1 : 0, o2 = axis == 2 ? 1 : 2;
  if (D == 2) { o1 = 1 - axis; o2 = 2; }
  const int n1 = lim[o1], n2 = D == 3 ? lim[o2] : 1;
  for (int i2 = 0; i2 < n2; i2++) for (int i1 = 0; i1 < n1; i1++) {
    int c[3] = {0, 0, 0};
    c[o1] = i1; if (D == 3) c[o2] = i2;
    /* ghost layers across a split face are delivered by the neighbour (slf_halo_push_field) */
    const int so = D - 1, io = o1 == so ? i1 : ((D == 3 && o2 == so) ? i2 : -1);
    if (axis != so && io >= 0 && ((link_lo && io == 0) || (link_hi && io == lim[so] - 1))) continue;
    long long g0, g1;
    c[axis] = 0; g0 = gidx(C, c); c[axis] = lim[axis] - 2; g1 = gidx(C, c);
    if (isfinite(field[g1])) field[g0] = field[g1];
    c[axis] = 1; g0 = gidx(C, c); c[axis] = lim[axis] - 1; g1 = gidx(C, c);
    if (isfinite(field[g0])) field[g1] = field[g0];
  }
  return 0;
}

int orc_apply_macro_periodic_boundary_conditions(const SlfParams *P, REAL *field, int axis) {
  return orc_apply_macro_periodic_boundary_conditions_linked(P, field, axis, 0, 0);
}
