/* vlct_oracle.c -- TEST INFRASTRUCTURE ONLY, not part of the product.
 *
 * A plain-C (C99) CPU restatement of the reference's VL+CT hydro/MHD block
 * update, EnzoMethodMHDVlct. It is the checker the CUDA path is compared with
 * in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing
 * in the product (enzo-e_b200/) may include, link or call it.
 *
 * PARITY PINNED: this file is validated bit-for-bit against
 * oracle/_ref/libvlct_ref.so (the reference's own sources compiled against
 * oracle/ref_shim) and against the golden L1 norms of the reference's vlct
 * answer tests (input/vlct/run_*_test.py) -- see tests/test_oracle_*.py.
 *
 * The reference performs one full-array pass per sub-step; so does this file
 * (simple to audit). Floating-point expressions keep the reference's operand
 * order and parenthesisation so that, built without FMA contraction, results
 * are bit-identical to the reference's default (value-safe) build.
 *
 * Citations "X.cpp:a-b" are relative to /root/reference/src/Enzo/ unless they
 * start with Cello/.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vlct.h"

/* ------------------------------------------------------------------------ */
/* small helpers                                                            */
/* ------------------------------------------------------------------------ */

/* quantity slots used for primitive, integration, flux and dU arrays */
enum { Q_RHO = 0, Q_VX, Q_VY, Q_VZ, Q_EN /* etot | pressure */,
       Q_BX, Q_BY, Q_BZ, Q_EINT, Q_SC0, Q_MAX = Q_SC0 + VLCT_MAX_PASSIVE };

typedef struct { double *p; int n0, n1, n2; } arr3;  /* shape (z, y, x) */

#define AT(a, k, j, i) \
  ((a).p[((size_t)(k) * (size_t)(a).n1 + (size_t)(j)) * (size_t)(a).n2 + (size_t)(i)])

static arr3 arr3_alloc(int n0, int n1, int n2)
{
  arr3 a;
  a.n0 = n0; a.n1 = n1; a.n2 = n2;
  /* zero-initialised like CelloView allocations (Cello/view_CelloView.hpp:722) */
  a.p = (double *) calloc((size_t) n0 * (size_t) n1 * (size_t) n2, sizeof(double));
  if (a.p == NULL) { fprintf(stderr, "vlct_oracle: out of memory\n"); abort(); }
  return a;
}

static arr3 arr3_wrap(double *p, int n0, int n1, int n2)
{
  arr3 a; a.p = p; a.n0 = n0; a.n1 = n1; a.n2 = n2; return a;
}

/* utils/utils.hpp:71-74 -- parenthesised on purpose */
static inline double sq3(double i, double j, double k)
{ return ((i * i) + ((j * j) + (k * k))); }

/* utils/utils.hpp:82-89 */
static inline double min3(double a, double b, double c)
{
  if (a < b) { return (c < a) ? c : a; }
  else       { return (c < b) ? c : b; }
}

/* std::max(value, floor)  (utils/utils.hpp:105-118) */
static inline double apply_floor(double value, double floor_)
{ return (value < floor_) ? floor_ : value; }

static inline double std_min(double a, double b) { return (b < a) ? b : a; }
static inline double std_max(double a, double b) { return (a < b) ? b : a; }

typedef struct vlct_oracle {
  vlct_config cfg;
  int mhd, de, nsc;
  int gx, gy, gz;
  int mx, my, mz;           /* cell-centred extents incl. ghosts; 0 = not yet */
  /* scratch (EnzoMethodMHDVlct.hpp:244-273, EnzoBfieldMethodCT.cpp:41-76) */
  arr3 temp[Q_MAX];         /* temp_integration_map                          */
  arr3 prim[Q_MAX];         /* primitive_map                                 */
  arr3 wl[Q_MAX], wr[Q_MAX];/* priml_map / primr_map (cell-shaped)           */
  arr3 flux[3][Q_MAX];      /* xflux / yflux / zflux                          */
  arr3 dU[Q_MAX];           /* dUcons_map                                    */
  arr3 vbar;                /* interface_vel_arr                             */
  arr3 tbi[3];              /* temp_bfieldi_l_                               */
  arr3 weight[3];           /* weight_l_                                     */
  arr3 edge[3];             /* edge_efield_l_                                */
  arr3 ecen;                /* center_efield_                                */
} vlct_oracle;

/* which quantity slots are in use */
static int has_q(const vlct_oracle *o, int q, int is_prim, int is_dU)
{
  if (q <= Q_EN) return 1;
  if (q >= Q_BX && q <= Q_BZ) return o->mhd && !is_dU; /* skip_B_update */
  if (q == Q_EINT) return o->de && !is_prim;
  return (q - Q_SC0) < o->nsc;
}

static void alloc_scratch(vlct_oracle *o, int mx, int my, int mz)
{
  o->mx = mx; o->my = my; o->mz = mz;
  for (int q = 0; q < Q_MAX; q++) {
    if (has_q(o, q, 0, 0)) {
      o->temp[q] = arr3_alloc(mz, my, mx);
      o->flux[0][q] = arr3_alloc(mz, my, mx - 1);
      o->flux[1][q] = arr3_alloc(mz, my - 1, mx);
      o->flux[2][q] = arr3_alloc(mz - 1, my, mx);
    }
    if (has_q(o, q, 1, 0)) {
      o->prim[q] = arr3_alloc(mz, my, mx);
      o->wl[q] = arr3_alloc(mz, my, mx);
      o->wr[q] = arr3_alloc(mz, my, mx);
    }
    if (has_q(o, q, 0, 1)) o->dU[q] = arr3_alloc(mz, my, mx);
  }
  if (o->de) o->vbar = arr3_alloc(mz, my, mx);
  if (o->mhd) {
    o->tbi[0] = arr3_alloc(mz, my, mx + 1);
    o->tbi[1] = arr3_alloc(mz, my + 1, mx);
    o->tbi[2] = arr3_alloc(mz + 1, my, mx);
    o->weight[0] = arr3_alloc(mz, my, mx - 1);
    o->weight[1] = arr3_alloc(mz, my - 1, mx);
    o->weight[2] = arr3_alloc(mz - 1, my, mx);
    o->edge[0] = arr3_alloc(mz - 1, my - 1, mx);
    o->edge[1] = arr3_alloc(mz - 1, my, mx - 1);
    o->edge[2] = arr3_alloc(mz, my - 1, mx - 1);
    o->ecen = arr3_alloc(mz, my, mx);
  }
}

static void free_scratch(vlct_oracle *o)
{
  for (int q = 0; q < Q_MAX; q++) {
    free(o->temp[q].p); free(o->prim[q].p); free(o->wl[q].p); free(o->wr[q].p);
    free(o->dU[q].p);
    for (int d = 0; d < 3; d++) free(o->flux[d][q].p);
  }
  free(o->vbar.p); free(o->ecen.p);
  for (int d = 0; d < 3; d++) {
    free(o->tbi[d].p); free(o->weight[d].p); free(o->edge[d].p);
  }
}

/* the integration map of a block (EnzoMethodMHDVlct.cpp:201-215) */
static void wrap_block(const vlct_oracle *o, const vlct_block *b, arr3 *u,
                       arr3 *bi)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy,
            mz = b->nz + 2 * b->gz;
  memset(u, 0, sizeof(arr3) * Q_MAX);
  u[Q_RHO] = arr3_wrap(b->density, mz, my, mx);
  u[Q_VX] = arr3_wrap(b->velocity_x, mz, my, mx);
  u[Q_VY] = arr3_wrap(b->velocity_y, mz, my, mx);
  u[Q_VZ] = arr3_wrap(b->velocity_z, mz, my, mx);
  u[Q_EN] = arr3_wrap(b->total_energy, mz, my, mx);
  if (o->mhd) {
    u[Q_BX] = arr3_wrap(b->bfield_x, mz, my, mx);
    u[Q_BY] = arr3_wrap(b->bfield_y, mz, my, mx);
    u[Q_BZ] = arr3_wrap(b->bfield_z, mz, my, mx);
    if (bi != NULL) {
      bi[0] = arr3_wrap(b->bfieldi_x, mz, my, mx + 1);
      bi[1] = arr3_wrap(b->bfieldi_y, mz, my + 1, mx);
      bi[2] = arr3_wrap(b->bfieldi_z, mz + 1, my, mx);
    }
  }
  if (o->de) u[Q_EINT] = arr3_wrap(b->internal_energy, mz, my, mx);
  for (int s = 0; s < o->nsc; s++)
    u[Q_SC0 + s] = arr3_wrap(b->passive[s], mz, my, mx);
}

/* ------------------------------------------------------------------------ */
/* equation of state (fluid-props/EnzoEOSIdeal.hpp:58-140)                   */
/* ------------------------------------------------------------------------ */

static inline double eos_cs2(double gamma, double rho, double p)
{ return gamma * p / rho; }

static inline double eos_specific_eint(double gamma, double rho, double p)
{ return p / ((gamma - 1.0) * rho); }

/* fast_magnetosonic_speed<-1>: EnzoEOSIdeal.hpp:111-140 (last branch) */
static inline double eos_cfast(double gamma, double rho, double p,
                               double bi, double bj, double bk)
{
  const double B2 = sq3(bi, bj, bk);
  const double cs2 = eos_cs2(gamma, rho, p);
  const double inv_density = 1.0 / rho;
  const double va2 = B2 * inv_density;
  const double va2_cos2 = (bi * bi) * inv_density;
  const double t = cs2 + va2;
  return sqrt(0.5 * (va2 + cs2 + sqrt(t * t - 4. * cs2 * va2_cos2)));
}

/* fast_magnetosonic_speed<0>: EnzoEOSIdeal.hpp:124-126 */
static inline double eos_cfast_max(double gamma, double rho, double p,
                                   double bi, double bj, double bk)
{
  const double B2 = sq3(bi, bj, bk);
  const double cs2 = eos_cs2(gamma, rho, p);
  const double va2 = B2 / rho;
  return sqrt(va2 + cs2);
}

/* riemann/EnzoRiemannUtils.hpp:224-249 */
static inline double passive_flux(double left, double right, double dflux)
{
  double upwind = (dflux > 0) * left + (dflux <= 0) * right;
  return upwind * dflux;
}

static inline double passive_eint_flux(double gamma, double rho_l, double p_l,
                                       double rho_r, double p_r, double dflux)
{
  double eint_l = eos_specific_eint(gamma, rho_l, p_l);
  double eint_r = eos_specific_eint(gamma, rho_r, p_r);
  return passive_flux(eint_l, eint_r, dflux);
}

/* ------------------------------------------------------------------------ */
/* Riemann solvers. States are in the permuted (i,j,k) frame:                */
/*   w = {rho, vi, vj, vk, p, bi, bj, bk}                                    */
/*   F = {rho, mom_i, mom_j, mom_k, etot_dens, (B_i: 0), B_j, B_k}           */
/* ------------------------------------------------------------------------ */
enum { W_RHO = 0, W_VI, W_VJ, W_VK, W_P, W_BI, W_BJ, W_BK, W_N };

/* riemann/EnzoRiemannHLLD.hpp:40-448 */
static void riemann_hlld(double gamma, const double *wli, const double *wri,
                         double *F, double *eint_flux, double *vbar)
{
  const double SMALL_NUMBER = 1.0e-8;
  const double igm1 = 1.0 / (gamma - 1.0);
  double spd[5];
  struct cons1d { double d, mx, my, mz, e, by, bz; };
  struct cons1d ul, ur, ulst, uldst, urdst, urst, fl, fr;

  const double pressure_l = wli[W_P], pressure_r = wri[W_P];
  const double bxi = wli[W_BI];
  double bxsq = bxi * bxi;
  double pbl = 0.5 * (bxsq + (wli[W_BJ] * wli[W_BJ] + wli[W_BK] * wli[W_BK]));
  double pbr = 0.5 * (bxsq + (wri[W_BJ] * wri[W_BJ] + wri[W_BK] * wri[W_BK]));
  double kel = 0.5 * wli[W_RHO] * (wli[W_VI] * wli[W_VI] +
                                   (wli[W_VJ] * wli[W_VJ] + wli[W_VK] * wli[W_VK]));
  double ker = 0.5 * wri[W_RHO] * (wri[W_VI] * wri[W_VI] +
                                   (wri[W_VJ] * wri[W_VJ] + wri[W_VK] * wri[W_VK]));

  ul.d = wli[W_RHO];
  ul.mx = wli[W_VI] * ul.d;
  ul.my = wli[W_VJ] * ul.d;
  ul.mz = wli[W_VK] * ul.d;
  ul.e = pressure_l * igm1 + kel + pbl;
  ul.by = wli[W_BJ];
  ul.bz = wli[W_BK];

  ur.d = wri[W_RHO];
  ur.mx = wri[W_VI] * ur.d;
  ur.my = wri[W_VJ] * ur.d;
  ur.mz = wri[W_VK] * ur.d;
  ur.e = pressure_r * igm1 + ker + pbr;
  ur.by = wri[W_BJ];
  ur.bz = wri[W_BK];

  /* step 2: outer wave speeds (HLLD.hpp:126-131) */
  double cfl = eos_cfast(gamma, wli[W_RHO], pressure_l, wli[W_BI], wli[W_BJ], wli[W_BK]);
  double cfr = eos_cfast(gamma, wri[W_RHO], pressure_r, wri[W_BI], wri[W_BJ], wri[W_BK]);
  spd[0] = std_min(wli[W_VI] - cfl, wri[W_VI] - cfr);
  spd[4] = std_max(wli[W_VI] + cfl, wri[W_VI] + cfr);

  /* step 3: L/R fluxes (HLLD.hpp:145-164) */
  double ptl = pressure_l + pbl;
  double ptr = pressure_r + pbr;

  fl.d = ul.mx;
  fl.mx = ul.mx * wli[W_VI] + ptl - bxsq;
  fl.my = ul.my * wli[W_VI] - bxi * ul.by;
  fl.mz = ul.mz * wli[W_VI] - bxi * ul.bz;
  fl.e = wli[W_VI] * (ul.e + ptl - bxsq) - bxi * (wli[W_VJ] * ul.by + wli[W_VK] * ul.bz);
  fl.by = ul.by * wli[W_VI] - bxi * wli[W_VJ];
  fl.bz = ul.bz * wli[W_VI] - bxi * wli[W_VK];

  fr.d = ur.mx;
  fr.mx = ur.mx * wri[W_VI] + ptr - bxsq;
  fr.my = ur.my * wri[W_VI] - bxi * ur.by;
  fr.mz = ur.mz * wri[W_VI] - bxi * ur.bz;
  fr.e = wri[W_VI] * (ur.e + ptr - bxsq) - bxi * (wri[W_VJ] * ur.by + wri[W_VK] * ur.bz);
  fr.by = ur.by * wri[W_VI] - bxi * wri[W_VJ];
  fr.bz = ur.bz * wri[W_VI] - bxi * wri[W_VK];

  /* step 4: middle and Alfven speeds (HLLD.hpp:168-189) */
  double sdl = spd[0] - wli[W_VI];
  double sdr = spd[4] - wri[W_VI];
  spd[2] = (sdr * ur.mx - sdl * ul.mx + (ptl - ptr)) / (sdr * ur.d - sdl * ul.d);

  double sdml = spd[0] - spd[2];
  double sdmr = spd[4] - spd[2];
  double sdml_inv = 1.0 / sdml;
  double sdmr_inv = 1.0 / sdmr;
  ulst.d = ul.d * sdl * sdml_inv;
  urst.d = ur.d * sdr * sdmr_inv;
  double ulst_d_inv = 1.0 / ulst.d;
  double urst_d_inv = 1.0 / urst.d;
  double sqrtdl = sqrt(ulst.d);
  double sqrtdr = sqrt(urst.d);

  spd[1] = spd[2] - fabs(bxi) / sqrtdl;
  spd[3] = spd[2] + fabs(bxi) / sqrtdr;

  /* step 5: intermediate states (HLLD.hpp:194-306) */
  double ptstl = ptl + ul.d * sdl * (spd[2] - wli[W_VI]);
  double ptstr = ptr + ur.d * sdr * (spd[2] - wri[W_VI]);
  double ptst = 0.5 * (ptstr + ptstl);

  ulst.mx = ulst.d * spd[2];
  if (fabs(ul.d * sdl * sdml - bxsq) < (SMALL_NUMBER) * ptst) {
    ulst.my = ulst.d * wli[W_VJ];
    ulst.mz = ulst.d * wli[W_VK];
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double tmp = bxi * (sdl - sdml) / (ul.d * sdl * sdml - bxsq);
    ulst.my = ulst.d * (wli[W_VJ] - ul.by * tmp);
    ulst.mz = ulst.d * (wli[W_VK] - ul.bz * tmp);
    tmp = (ul.d * (sdl * sdl) - bxsq) / (ul.d * sdl * sdml - bxsq);
    ulst.by = ul.by * tmp;
    ulst.bz = ul.bz * tmp;
  }
  double vbstl = (ulst.mx * bxi + (ulst.my * ulst.by + ulst.mz * ulst.bz)) * ulst_d_inv;
  ulst.e = (sdl * ul.e - ptl * wli[W_VI] + ptst * spd[2] +
            bxi * (wli[W_VI] * bxi + (wli[W_VJ] * ul.by + wli[W_VK] * ul.bz)
                   - vbstl)) * sdml_inv;

  urst.mx = urst.d * spd[2];
  if (fabs(ur.d * sdr * sdmr - bxsq) < (SMALL_NUMBER) * ptst) {
    urst.my = urst.d * wri[W_VJ];
    urst.mz = urst.d * wri[W_VK];
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double tmp = bxi * (sdr - sdmr) / (ur.d * sdr * sdmr - bxsq);
    urst.my = urst.d * (wri[W_VJ] - ur.by * tmp);
    urst.mz = urst.d * (wri[W_VK] - ur.bz * tmp);
    tmp = (ur.d * (sdr * sdr) - bxsq) / (ur.d * sdr * sdmr - bxsq);
    urst.by = ur.by * tmp;
    urst.bz = ur.bz * tmp;
  }
  double vbstr = (urst.mx * bxi + (urst.my * urst.by + urst.mz * urst.bz)) * urst_d_inv;
  urst.e = (sdr * ur.e - ptr * wri[W_VI] + ptst * spd[2] +
            bxi * (wri[W_VI] * bxi + (wri[W_VJ] * ur.by + wri[W_VK] * ur.bz)
                   - vbstr)) * sdmr_inv;

  if (0.5 * bxsq < (SMALL_NUMBER) * ptst) {
    uldst = ulst;
    urdst = urst;
  } else {
    double invsumd = 1.0 / (sqrtdl + sqrtdr);
    double bxsig = (bxi > 0.0 ? 1.0 : -1.0);

    uldst.d = ulst.d;
    urdst.d = urst.d;
    uldst.mx = ulst.mx;
    urdst.mx = urst.mx;

    double tmp = invsumd * (sqrtdl * (ulst.my * ulst_d_inv) +
                            sqrtdr * (urst.my * urst_d_inv) +
                            bxsig * (urst.by - ulst.by));
    uldst.my = uldst.d * tmp;
    urdst.my = urdst.d * tmp;

    tmp = invsumd * (sqrtdl * (ulst.mz * ulst_d_inv) +
                     sqrtdr * (urst.mz * urst_d_inv) +
                     bxsig * (urst.bz - ulst.bz));
    uldst.mz = uldst.d * tmp;
    urdst.mz = urdst.d * tmp;

    tmp = invsumd * (sqrtdl * urst.by + sqrtdr * ulst.by +
                     bxsig * sqrtdl * sqrtdr * ((urst.my * urst_d_inv) -
                                                (ulst.my * ulst_d_inv)));
    uldst.by = urdst.by = tmp;

    tmp = invsumd * (sqrtdl * urst.bz + sqrtdr * ulst.bz +
                     bxsig * sqrtdl * sqrtdr * ((urst.mz * urst_d_inv) -
                                                (ulst.mz * ulst_d_inv)));
    uldst.bz = urdst.bz = tmp;

    tmp = spd[2] * bxi + (uldst.my * uldst.by + uldst.mz * uldst.bz) / uldst.d;
    uldst.e = ulst.e - sqrtdl * bxsig * (vbstl - tmp);
    urdst.e = urst.e + sqrtdr * bxsig * (vbstr - tmp);
  }

  /* step 6: flux (HLLD.hpp:309-395) */
  uldst.d = spd[1] * (uldst.d - ulst.d);
  uldst.mx = spd[1] * (uldst.mx - ulst.mx);
  uldst.my = spd[1] * (uldst.my - ulst.my);
  uldst.mz = spd[1] * (uldst.mz - ulst.mz);
  uldst.e = spd[1] * (uldst.e - ulst.e);
  uldst.by = spd[1] * (uldst.by - ulst.by);
  uldst.bz = spd[1] * (uldst.bz - ulst.bz);

  ulst.d = spd[0] * (ulst.d - ul.d);
  ulst.mx = spd[0] * (ulst.mx - ul.mx);
  ulst.my = spd[0] * (ulst.my - ul.my);
  ulst.mz = spd[0] * (ulst.mz - ul.mz);
  ulst.e = spd[0] * (ulst.e - ul.e);
  ulst.by = spd[0] * (ulst.by - ul.by);
  ulst.bz = spd[0] * (ulst.bz - ul.bz);

  urdst.d = spd[3] * (urdst.d - urst.d);
  urdst.mx = spd[3] * (urdst.mx - urst.mx);
  urdst.my = spd[3] * (urdst.my - urst.my);
  urdst.mz = spd[3] * (urdst.mz - urst.mz);
  urdst.e = spd[3] * (urdst.e - urst.e);
  urdst.by = spd[3] * (urdst.by - urst.by);
  urdst.bz = spd[3] * (urdst.bz - urst.bz);

  urst.d = spd[4] * (urst.d - ur.d);
  urst.mx = spd[4] * (urst.mx - ur.mx);
  urst.my = spd[4] * (urst.my - ur.my);
  urst.mz = spd[4] * (urst.mz - ur.mz);
  urst.e = spd[4] * (urst.e - ur.e);
  urst.by = spd[4] * (urst.by - ur.by);
  urst.bz = spd[4] * (urst.bz - ur.bz);

  struct cons1d f;
  if (spd[0] >= 0.0) {
    f = fl;
  } else if (spd[4] <= 0.0) {
    f = fr;
  } else if (spd[1] >= 0.0) {
    f.d = fl.d + ulst.d;    f.mx = fl.mx + ulst.mx;
    f.my = fl.my + ulst.my; f.mz = fl.mz + ulst.mz;
    f.e = fl.e + ulst.e;    f.by = fl.by + ulst.by;  f.bz = fl.bz + ulst.bz;
  } else if (spd[2] >= 0.0) {
    f.d = fl.d + ulst.d + uldst.d;     f.mx = fl.mx + ulst.mx + uldst.mx;
    f.my = fl.my + ulst.my + uldst.my; f.mz = fl.mz + ulst.mz + uldst.mz;
    f.e = fl.e + ulst.e + uldst.e;     f.by = fl.by + ulst.by + uldst.by;
    f.bz = fl.bz + ulst.bz + uldst.bz;
  } else if (spd[3] > 0.0) {
    f.d = fr.d + urst.d + urdst.d;     f.mx = fr.mx + urst.mx + urdst.mx;
    f.my = fr.my + urst.my + urdst.my; f.mz = fr.mz + urst.mz + urdst.mz;
    f.e = fr.e + urst.e + urdst.e;     f.by = fr.by + urst.by + urdst.by;
    f.bz = fr.bz + urst.bz + urdst.bz;
  } else {
    f.d = fr.d + urst.d;    f.mx = fr.mx + urst.mx;
    f.my = fr.my + urst.my; f.mz = fr.mz + urst.mz;
    f.e = fr.e + urst.e;    f.by = fr.by + urst.by;  f.bz = fr.bz + urst.bz;
  }
  F[W_RHO] = f.d; F[W_VI] = f.mx; F[W_VJ] = f.my; F[W_VK] = f.mz;
  F[W_P] = f.e; F[W_BI] = 0.0; F[W_BJ] = f.by; F[W_BK] = f.bz;

  /* dual-energy extras (HLLD.hpp:409-447) */
  *eint_flux = passive_eint_flux(gamma, wli[W_RHO], pressure_l, wri[W_RHO],
                                 pressure_r, F[W_RHO]);
  const double S_M = spd[2], S_l = spd[0], S_r = spd[4];
  const double l_coef = (S_l - wli[W_VI]) / (S_l - S_M);
  const double r_coef = (S_r - wri[W_VI]) / (S_r - S_M);
  if (S_l > 0)        *vbar = wli[W_VI];
  else if (S_r < 0)   *vbar = wri[W_VI];
  else if (S_M >= 0)  *vbar = S_M * l_coef;
  else                *vbar = S_M * r_coef;
}

/* compute_conserved (riemann/EnzoRiemannUtils.hpp:48-82): total energy density */
static inline double cons_etot(double gamma, const double *w, int mhd)
{
  double internal_edens = w[W_P] / (gamma - 1.0);
  double kinetic_edens = 0.5 * w[W_RHO] * sq3(w[W_VI], w[W_VJ], w[W_VK]);
  double magnetic_edens = mhd ? 0.5 * sq3(w[W_BI], w[W_BJ], w[W_BK])
                              : 0.5 * sq3(0., 0., 0.);
  return internal_edens + kinetic_edens + magnetic_edens;
}

/* EinfeldtWavespeed (riemann/EnzoRiemannHLL.hpp:44-172) */
static void einfeldt_speeds(double gamma, int mhd, const double *wl,
                            const double *wr, double etot_l, double etot_r,
                            double *bp, double *bm)
{
  const double pressure_l = wl[W_P], pressure_r = wr[W_P];
  double c_l, c_r;
  if (mhd) {
    c_l = eos_cfast(gamma, wl[W_RHO], pressure_l, wl[W_BI], wl[W_BJ], wl[W_BK]);
    c_r = eos_cfast(gamma, wr[W_RHO], pressure_r, wr[W_BI], wr[W_BJ], wr[W_BK]);
  } else {
    c_l = sqrt(eos_cs2(gamma, wl[W_RHO], pressure_l));
    c_r = sqrt(eos_cs2(gamma, wr[W_RHO], pressure_r));
  }
  double left_speed = (wl[W_VI] - c_l);
  double right_speed = (wr[W_VI] + c_r);

  double sqrtrho_l = sqrt(wl[W_RHO]);
  double sqrtrho_r = sqrt(wr[W_RHO]);
  double inv_sqrtrho_tot = 1.0 / (sqrtrho_l + sqrtrho_r);

  double vi_roe = (sqrtrho_l * wl[W_VI] + sqrtrho_r * wr[W_VI]) * inv_sqrtrho_tot;
  double vj_roe = (sqrtrho_l * wl[W_VJ] + sqrtrho_r * wr[W_VJ]) * inv_sqrtrho_tot;
  double vk_roe = (sqrtrho_l * wl[W_VK] + sqrtrho_r * wr[W_VK]) * inv_sqrtrho_tot;
  double v_roe2 = vi_roe * vi_roe + vj_roe * vj_roe + vk_roe * vk_roe;

  double ptot_l = pressure_l, ptot_r = pressure_r;
  if (mhd) {
    ptot_l += 0.5 * sq3(wl[W_BI], wl[W_BJ], wl[W_BK]);
    ptot_r += 0.5 * sq3(wr[W_BI], wr[W_BJ], wr[W_BK]);
  }
  double h_l = (etot_l + ptot_l) / wl[W_RHO];
  double h_r = (etot_r + ptot_r) / wr[W_RHO];
  double h_roe = (sqrtrho_l * h_l + sqrtrho_r * h_r) * inv_sqrtrho_tot;

  double c_roe;
  if (mhd) {
    /* roe_cfast_ (HLL.hpp:134-172) */
    double rho_roe = sqrtrho_l * sqrtrho_r;
    double bi_roe = wl[W_BI];
    double bj_roe = (sqrtrho_l * wr[W_BJ] + sqrtrho_r * wl[W_BJ]) * inv_sqrtrho_tot;
    double bk_roe = (sqrtrho_l * wr[W_BK] + sqrtrho_r * wl[W_BK]) * inv_sqrtrho_tot;
    double b_roe2 = bi_roe * bi_roe + bj_roe * bj_roe + bk_roe * bk_roe;
    double gamma_prime = gamma - 1.;
    double dbj = wl[W_BJ] - wr[W_BJ], dbk = wl[W_BK] - wr[W_BK];
    double x_prime = ((dbj * dbj + dbk * dbk) * 0.5 * (gamma_prime - 1) * inv_sqrtrho_tot);
    double y_prime = ((gamma_prime - 1) * (wl[W_RHO] + wr[W_RHO]) * 0.5 / rho_roe);
    double tilde_a2 = (gamma_prime * (h_roe - 0.5 * v_roe2 - b_roe2 / rho_roe) - x_prime);
    double tilde_vai2 = bi_roe * bi_roe / rho_roe;
    double tilde_va2 = (tilde_vai2 + (gamma_prime - y_prime) *
                        (bj_roe * bj_roe + bk_roe * bk_roe) / rho_roe);
    double t = tilde_a2 + tilde_va2;
    c_roe = sqrt(0.5 * (tilde_a2 + tilde_va2 +
                        sqrt(t * t - 4 * tilde_a2 * tilde_vai2)));
  } else {
    /* roe_cs_ (HLL.hpp:123-130) */
    double temp = h_roe - 0.5 * v_roe2;
    c_roe = sqrt((gamma - 1) * std_max(temp, 0.));
  }
  *bp = fmax(vi_roe + c_roe, right_speed);
  *bm = fmin(vi_roe - c_roe, left_speed);
}

/* HLLKernel<EinfeldtWavespeed<MHDLUT>> (riemann/EnzoRiemannHLL.hpp:238-345)
 * with enzo_riemann_utils::active_fluxes (EnzoRiemannUtils.hpp:112-149) */
static void riemann_hlle_mhd(double gamma, const double *wl, const double *wr,
                             double *F, double *eint_flux, double *vbar)
{
  /* LUT order (EnzoRiemannImpl.hpp:40-56): rho, Bi,Bj,Bk, vi,vj,vk, etot */
  double Ul[W_N], Ur[W_N], Fl[W_N], Fr[W_N];
  const double *w[2] = { wl, wr };
  double *U[2] = { Ul, Ur };
  double *Fx[2] = { Fl, Fr };
  for (int s = 0; s < 2; s++) {
    const double *p = w[s];
    U[s][W_RHO] = p[W_RHO];
    U[s][W_BI] = p[W_BI]; U[s][W_BJ] = p[W_BJ]; U[s][W_BK] = p[W_BK];
    U[s][W_VI] = p[W_VI] * p[W_RHO];
    U[s][W_VJ] = p[W_VJ] * p[W_RHO];
    U[s][W_VK] = p[W_VK] * p[W_RHO];
    U[s][W_P] = cons_etot(gamma, p, 1);

    double vi = p[W_VI], vj = p[W_VJ], vk = p[W_VK];
    double Bi = p[W_BI], Bj = p[W_BJ], Bk = p[W_BK];
    double etot = U[s][W_P];
    double ptot = p[W_P] + 0.5 * sq3(Bi, Bj, Bk);
    double mom_i = U[s][W_VI];
    Fx[s][W_RHO] = mom_i;
    Fx[s][W_VI] = mom_i * vi - Bi * Bi + ptot;
    Fx[s][W_VJ] = mom_i * vj - Bj * Bi;
    Fx[s][W_VK] = mom_i * vk - Bk * Bi;
    Fx[s][W_P] = ((etot + ptot) * vi - (Bi * vi + (Bj * vj + Bk * vk)) * Bi);
    Fx[s][W_BI] = 0;
    Fx[s][W_BJ] = Bj * vi - Bi * vj;
    Fx[s][W_BK] = Bk * vi - Bi * vk;
  }
  double bp, bm;
  einfeldt_speeds(gamma, 1, wl, wr, Ul[W_P], Ur[W_P], &bp, &bm);
  bp = fmax(bp, 0.0);
  bm = fmin(bm, 0.0);
  double inv_speed_diff = 1. / (bp - bm);
  for (int f = 0; f < W_N; f++) {
    F[f] = ((bp * Fl[f] - bm * Fr[f] + (Ur[f] - Ul[f]) * bp * bm) * inv_speed_diff);
  }
  *eint_flux = passive_eint_flux(gamma, wl[W_RHO], wl[W_P], wr[W_RHO], wr[W_P],
                                 F[W_RHO]);
  *vbar = (bp * wl[W_VI] - bm * wr[W_VI]) * inv_speed_diff;
}

/* HLLCKernel (riemann/EnzoRiemannHLLC.hpp:34-172), hydro only */
static void riemann_hllc(double gamma, const double *wl, const double *wr,
                         double *F, double *eint_flux, double *vbar)
{
  const double pressure_l = wl[W_P], pressure_r = wr[W_P];
  const double etot_l = cons_etot(gamma, wl, 0);
  const double etot_r = cons_etot(gamma, wr, 0);
  const double momi_l = wl[W_VI] * wl[W_RHO];
  const double momi_r = wr[W_VI] * wr[W_RHO];

  double cs_l, cs_r;
  /* called as (..., &cs_r, &cs_l): bp -> cs_r, bm -> cs_l (HLLC.hpp:70-73) */
  einfeldt_speeds(gamma, 0, wl, wr, etot_l, etot_r, &cs_r, &cs_l);

  double bm = fmin(cs_l, 0.0);
  double bp = fmax(cs_r, 0.0);

  double tl = (pressure_l - (cs_l - wl[W_VI]) * wl[W_RHO] * wl[W_VI]);
  double tr = (pressure_r - (cs_r - wr[W_VI]) * wr[W_RHO] * wr[W_VI]);
  double dl = wl[W_RHO] * (cs_l - wl[W_VI]);
  double dr = -wr[W_RHO] * (cs_r - wr[W_VI]);
  double q1 = 1.0 / (dl + dr);
  double cw = (tr - tl) * q1;
  double cp = (dl * tr + dr * tl) * q1;

  double sl, sr, sm;
  if (cw >= 0.) {
    sl = cw / (cw - bm);
    sr = 0.;
    sm = -bm / (cw - bm);
  } else {
    sl = 0.;
    sr = -cw / (bp - cw);
    sm = bp / (bp - cw);
  }
  cp = std_max(cp, 0.);

  double dfl, dfr, ufl, ufr, vfl, vfr, wfl, wfr, efl, efr;
  dfl = momi_l - bm * wl[W_RHO];
  dfr = momi_r - bp * wr[W_RHO];
  ufl = momi_l * (wl[W_VI] - bm) + pressure_l;
  ufr = momi_r * (wr[W_VI] - bp) + pressure_r;
  vfl = (wl[W_RHO] * wl[W_VJ] * (wl[W_VI] - bm));
  vfr = (wr[W_RHO] * wr[W_VJ] * (wr[W_VI] - bp));
  wfl = (wl[W_RHO] * wl[W_VK] * (wl[W_VI] - bm));
  wfr = (wr[W_RHO] * wr[W_VK] * (wr[W_VI] - bp));
  efl = (etot_l * (wl[W_VI] - bm) + pressure_l * wl[W_VI]);
  efr = (etot_r * (wr[W_VI] - bp) + pressure_r * wr[W_VI]);

  F[W_RHO] = sl * dfl + sr * dfr;
  F[W_VI] = sl * ufl + sr * ufr;
  F[W_VJ] = sl * vfl + sr * vfr;
  F[W_VK] = sl * wfl + sr * wfr;
  F[W_P] = sl * efl + sr * efr;
  F[W_VI] += (sm * cp);
  F[W_P] += (sm * cp * cw);
  F[W_BI] = F[W_BJ] = F[W_BK] = 0.0;

  *eint_flux = passive_eint_flux(gamma, wl[W_RHO], pressure_l, wr[W_RHO],
                                 pressure_r, F[W_RHO]);
  *vbar = (sl * (wl[W_VI] - bm) + sr * (wr[W_VI] - bp));
}

/* ------------------------------------------------------------------------ */
/* slope limiters (toolkit/EnzoReconstructorPLM.hpp:253-346)                 */
/* ------------------------------------------------------------------------ */
static inline double sign_(double val) { return (double) ((0.0 < val) - (val < 0.0)); }

static inline double limiter_enzo(double vm1, double v, double vp1, double theta)
{
  double dv_c = 0.5 * (vp1 - vm1);
  double dv_l = (v - vm1) * theta;
  double dv_r = (vp1 - v) * theta;
  return (0.5 * (sign_(dv_l) + sign_(dv_r))) * min3(fabs(dv_l), fabs(dv_r), fabs(dv_c));
}

static inline double limiter_athena(double vm1, double v, double vp1)
{
  double dv_l = (v - vm1);
  double dv_r = (vp1 - v);
  double temp = dv_l * dv_r;
  if (temp <= 0.) { return 0.; }
  return 2. * temp / (dv_l + dv_r);
}

/* ------------------------------------------------------------------------ */
/* one stage: EnzoMHDIntegratorStageCommands::compute_update_stage           */
/* (hydro-mhd/EnzoMHDIntegratorStageCommands.cpp:102-261)                    */
/* ------------------------------------------------------------------------ */

/* EnzoPhysicsFluidProps::pressure_from_integration
 * (fluid-props/EnzoPhysicsFluidProps.cpp:142-156,
 *  fluid-props/EnzoComputePressure.cpp:82-198, rank 3) */
static void pressure_from_integration(const vlct_oracle *o, const arr3 *u,
                                      arr3 p, int s)
{
  const double gm1 = o->cfg.gamma - 1.0;
  const int mz = p.n0, my = p.n1, mx = p.n2;
  for (int k = s; k < mz - s; k++)
    for (int j = s; j < my - s; j++)
      for (int i = s; i < mx - s; i++) {
        if (o->de) {
          AT(p, k, j, i) = gm1 * AT(u[Q_RHO], k, j, i) * AT(u[Q_EINT], k, j, i);
        } else {
          double vx = AT(u[Q_VX], k, j, i), vy = AT(u[Q_VY], k, j, i),
                 vz = AT(u[Q_VZ], k, j, i);
          double ke = 0.5 * (vx * vx + vy * vy + vz * vz);
          double me_den = 0.;
          if (o->mhd) {
            double bx = AT(u[Q_BX], k, j, i), by = AT(u[Q_BY], k, j, i),
                   bz = AT(u[Q_BZ], k, j, i);
            me_den = 0.5 * (bx * bx + by * by + bz * bz);
          }
          AT(p, k, j, i) = gm1 * (AT(u[Q_RHO], k, j, i) *
                                  (AT(u[Q_EN], k, j, i) - ke) - me_den);
        }
      }
}

/* EnzoPhysicsFluidProps::primitive_from_integration
 * (fluid-props/EnzoPhysicsFluidProps.cpp:64-138) */
static void primitive_from_integration(vlct_oracle *o, const arr3 *cur, int s)
{
  const int mz = o->mz, my = o->my, mx = o->mx;
  for (int q = 0; q < Q_MAX; q++) {
    if (q == Q_EN || q == Q_EINT || !has_q(o, q, 1, 0)) continue;
    const int passive = (q >= Q_SC0);
    for (int k = s; k < mz - s; k++)
      for (int j = s; j < my - s; j++)
        for (int i = s; i < mx - s; i++) {
          AT(o->prim[q], k, j, i) = passive
            ? AT(cur[q], k, j, i) / AT(cur[Q_RHO], k, j, i)
            : AT(cur[q], k, j, i);
        }
  }
  pressure_from_integration(o, cur, o->prim[Q_EN], s);
}

/* reconstruct -> fix B -> Riemann -> passive fluxes, for one dimension.
 * Face f along `dim` sits between cells f and f+1. */
static void compute_flux_dim(vlct_oracle *o, int dim, int recon, int stale,
                             const arr3 *bi_cur)
{
  const int mz = o->mz, my = o->my, mx = o->mx;
  const int di = (dim == 0), dj = (dim == 1), dk = (dim == 2);
  const double theta = o->cfg.theta_limiter;
  const double gamma = o->cfg.gamma;
  const int jd = (dim + 1) % 3, kd = (dim + 2) % 3;
  const int qv[3] = { Q_VX, Q_VY, Q_VZ }, qb[3] = { Q_BX, Q_BY, Q_BZ };
  int cur_stale;

  /* ---- reconstruction --------------------------------------------------- */
  if (recon == VLCT_RECON_NN) {
    /* toolkit/EnzoReconstructorNN.cpp:14-48: staling 0 */
    for (int q = 0; q < Q_MAX; q++) {
      if (!has_q(o, q, 1, 0)) continue;
      for (int k = stale; k < mz - stale - dk; k++)
        for (int j = stale; j < my - stale - dj; j++)
          for (int i = stale; i < mx - stale - di; i++) {
            AT(o->wl[q], k, j, i) = AT(o->prim[q], k, j, i);
            AT(o->wr[q], k, j, i) = AT(o->prim[q], k + dk, j + dj, i + di);
          }
    }
    cur_stale = stale;
  } else {
    /* toolkit/EnzoReconstructorPLM.hpp:166-248: staling 1 */
    for (int q = 0; q < Q_MAX; q++) {
      if (!has_q(o, q, 1, 0)) continue;
      int use_floor = 0; double prim_floor = 0;
      if (q == Q_RHO) { use_floor = 1; prim_floor = o->cfg.density_floor; }
      if (q == Q_EN)  { use_floor = 1; prim_floor = o->cfg.pressure_floor; }
      /* centre cell c runs over [stale+1, ext-stale-1) along dim */
      for (int k = stale + dk; k < mz - stale - dk; k++)
        for (int j = stale + dj; j < my - stale - dj; j++)
          for (int i = stale + di; i < mx - stale - di; i++) {
            double vm1 = AT(o->prim[q], k - dk, j - dj, i - di);
            double val = AT(o->prim[q], k, j, i);
            double vp1 = AT(o->prim[q], k + dk, j + dj, i + di);
            double dv = (recon == VLCT_RECON_PLM_ATHENA)
              ? limiter_athena(vm1, val, vp1)
              : limiter_enzo(vm1, val, vp1, theta);
            double half_dv = dv * 0.5;
            double right_val, left_val;
            if (use_floor) {
              right_val = apply_floor(val - half_dv, prim_floor);
              left_val = apply_floor(val + half_dv, prim_floor);
            } else {
              right_val = val - half_dv;
              left_val = val + half_dv;
            }
            AT(o->wr[q], k - dk, j - dj, i - di) = right_val; /* face c-1 */
            AT(o->wl[q], k, j, i) = left_val;                 /* face c   */
          }
    }
    cur_stale = stale + 1;
  }

  /* face-array extents for this dim and the non-stale face region */
  const int fx = mx - di, fy = my - dj, fz = mz - dk;
  const int s = cur_stale;

  /* ---- longitudinal B from the face-centred field ----------------------- */
  /* toolkit/EnzoBfieldMethodCT.cpp:122-166: face f <-> bfieldi index f+1 */
  if (o->mhd) {
    for (int k = s; k < fz - s; k++)
      for (int j = s; j < fy - s; j++)
        for (int i = s; i < fx - s; i++) {
          double b = AT(bi_cur[dim], k + dk, j + dj, i + di);
          AT(o->wl[qb[dim]], k, j, i) = b;
          AT(o->wr[qb[dim]], k, j, i) = b;
        }
  }

  /* ---- Riemann solve (riemann/EnzoRiemannImpl.hpp:266-338) -------------- */
  arr3 *F = o->flux[dim];
  for (int k = s; k < fz - s; k++)
    for (int j = s; j < fy - s; j++)
      for (int i = s; i < fx - s; i++) {
        double wl[W_N], wr[W_N], f[W_N], ef, vb;
        wl[W_RHO] = AT(o->wl[Q_RHO], k, j, i);   wr[W_RHO] = AT(o->wr[Q_RHO], k, j, i);
        wl[W_VI] = AT(o->wl[qv[dim]], k, j, i);  wr[W_VI] = AT(o->wr[qv[dim]], k, j, i);
        wl[W_VJ] = AT(o->wl[qv[jd]], k, j, i);   wr[W_VJ] = AT(o->wr[qv[jd]], k, j, i);
        wl[W_VK] = AT(o->wl[qv[kd]], k, j, i);   wr[W_VK] = AT(o->wr[qv[kd]], k, j, i);
        wl[W_P] = AT(o->wl[Q_EN], k, j, i);      wr[W_P] = AT(o->wr[Q_EN], k, j, i);
        if (o->mhd) {
          wl[W_BI] = AT(o->wl[qb[dim]], k, j, i); wr[W_BI] = AT(o->wr[qb[dim]], k, j, i);
          wl[W_BJ] = AT(o->wl[qb[jd]], k, j, i);  wr[W_BJ] = AT(o->wr[qb[jd]], k, j, i);
          wl[W_BK] = AT(o->wl[qb[kd]], k, j, i);  wr[W_BK] = AT(o->wr[qb[kd]], k, j, i);
        } else {
          wl[W_BI] = wl[W_BJ] = wl[W_BK] = 0.; wr[W_BI] = wr[W_BJ] = wr[W_BK] = 0.;
        }
        switch (o->cfg.riemann_solver) {
        case VLCT_RIEMANN_HLLD: riemann_hlld(gamma, wl, wr, f, &ef, &vb); break;
        case VLCT_RIEMANN_HLLE: riemann_hlle_mhd(gamma, wl, wr, f, &ef, &vb); break;
        default:                riemann_hllc(gamma, wl, wr, f, &ef, &vb); break;
        }
        AT(F[Q_RHO], k, j, i) = f[W_RHO];
        AT(F[qv[dim]], k, j, i) = f[W_VI];
        AT(F[qv[jd]], k, j, i) = f[W_VJ];
        AT(F[qv[kd]], k, j, i) = f[W_VK];
        AT(F[Q_EN], k, j, i) = f[W_P];
        if (o->mhd) {
          AT(F[qb[dim]], k, j, i) = f[W_BI];
          AT(F[qb[jd]], k, j, i) = f[W_BJ];
          AT(F[qb[kd]], k, j, i) = f[W_BK];
        }
        if (o->de) {
          AT(F[Q_EINT], k, j, i) = ef;
          AT(o->vbar, k, j, i) = vb;
        }
        /* passive scalars (riemann/EnzoRiemannUtils.hpp:267-314) */
        for (int sc = 0; sc < o->nsc; sc++) {
          AT(F[Q_SC0 + sc], k, j, i) =
            passive_flux(AT(o->wl[Q_SC0 + sc], k, j, i),
                         AT(o->wr[Q_SC0 + sc], k, j, i), f[W_RHO]);
        }
      }
}

/* dU -= dt/dx (F_{c+1/2} - F_{c-1/2}) plus the dual-energy source
 * (toolkit/EnzoIntegrationQuanUpdate.cpp:105-147,
 *  toolkit/EnzoSourceInternalEnergy.cpp:16-92) and the upwind weights
 * (toolkit/EnzoBfieldMethodCT.cpp:170-216) */
static void accumulate_dim(vlct_oracle *o, int dim, double dt, double width,
                           int s)
{
  const int mz = o->mz, my = o->my, mx = o->mx;
  const int di = (dim == 0), dj = (dim == 1), dk = (dim == 2);
  const double dtdx_i = dt / width;
  arr3 *F = o->flux[dim];

  for (int q = 0; q < Q_MAX; q++) {
    if (!has_q(o, q, 0, 1)) continue;
    for (int k = s + dk; k < mz - s - dk; k++)
      for (int j = s + dj; j < my - s - dj; j++)
        for (int i = s + di; i < mx - s - di; i++) {
          double fr = AT(F[q], k, j, i);
          double fl = AT(F[q], k - dk, j - dj, i - di);
          AT(o->dU[q], k, j, i) -= dtdx_i * (fr - fl);
        }
  }

  if (o->de) {
    const double p_floor = o->cfg.pressure_floor;
    const double dtdx = dt / width;
    for (int k = s + dk; k < mz - s - dk; k++)
      for (int j = s + dj; j < my - s - dj; j++)
        for (int i = s + di; i < mx - s - di; i++) {
          double p = apply_floor(AT(o->prim[Q_EN], k, j, i), p_floor);
          double vr = AT(o->vbar, k, j, i);
          double vl = AT(o->vbar, k - dk, j - dj, i - di);
          AT(o->dU[Q_EINT], k, j, i) -= dtdx * p * (vr - vl);
        }
  }

  if (o->mhd) {
    const int fx = mx - di, fy = my - dj, fz = mz - dk;
    for (int k = s; k < fz - s; k++)
      for (int j = s; j < fy - s; j++)
        for (int i = s; i < fx - s; i++) {
          double df = AT(F[Q_RHO], k, j, i);
          double w;
          if (df > 0) w = 1.0; else if (df < 0) w = 0.0; else w = 0.5;
          AT(o->weight[dim], k, j, i) = w;
        }
  }
}

/* EnzoSourceGravity::calculate_source (toolkit/EnzoSourceGravity.cpp:17-69) */
static void gravity_source(vlct_oracle *o, const arr3 *u0, const vlct_block *b,
                           double dt, int s)
{
  const int mz = o->mz, my = o->my, mx = o->mx;
  arr3 ax = arr3_wrap(b->acceleration_x, mz, my, mx);
  arr3 ay = arr3_wrap(b->acceleration_y, mz, my, mx);
  arr3 az = arr3_wrap(b->acceleration_z, mz, my, mx);
  for (int k = s; k < mz - s; k++)
    for (int j = s; j < my - s; j++)
      for (int i = s; i < mx - s; i++) {
        double rho = AT(u0[Q_RHO], k, j, i);
        double gx = AT(ax, k, j, i), gy = AT(ay, k, j, i), gz = AT(az, k, j, i);
        AT(o->dU[Q_VX], k, j, i) += dt * rho * gx;
        AT(o->dU[Q_VY], k, j, i) += dt * rho * gy;
        AT(o->dU[Q_VZ], k, j, i) += dt * rho * gz;
        AT(o->dU[Q_EN], k, j, i) += dt * rho * ((AT(u0[Q_VX], k, j, i) * gx) +
                                                (AT(u0[Q_VY], k, j, i) * gy) +
                                                (AT(u0[Q_VZ], k, j, i) * gz));
      }
}

/* constrained transport: EnzoBfieldMethodCT::update_all_bfield_components
 * (toolkit/EnzoBfieldMethodCT.cpp:220-728) */
static void ct_update(vlct_oracle *o, const arr3 *cur, const arr3 *bi0,
                      arr3 *bi_out, arr3 *out, double dt, const double *width,
                      int s)
{
  const int qv[3] = { Q_VX, Q_VY, Q_VZ }, qb[3] = { Q_BX, Q_BY, Q_BZ };
  const int m[3] = { o->mx, o->my, o->mz };

  for (int d = 0; d < 3; d++) {
    const int jd = (d + 1) % 3, kd = (d + 2) % 3;
    /* unit vectors of the j and k axes in (x,y,z) index space */
    const int jx = (jd == 0), jy = (jd == 1), jz = (jd == 2);
    const int kx = (kd == 0), ky = (kd == 1), kz = (kd == 2);

    /* cell-centred E_d = -v_j B_k + v_k B_j  (CT.cpp:267-292) */
    for (int k = s; k < m[2] - s; k++)
      for (int j = s; j < m[1] - s; j++)
        for (int i = s; i < m[0] - s; i++) {
          AT(o->ecen, k, j, i) = (-AT(cur[qv[jd]], k, j, i) * AT(cur[qb[kd]], k, j, i)
                                  + AT(cur[qv[kd]], k, j, i) * AT(cur[qb[jd]], k, j, i));
        }

    /* edge E_d (CT.cpp:384-557); edge (k,j,i) sits at +1/2 along jd and kd.
     * E on j-faces is -F_j(B_k) (negation applied inline: negate_Ej = true),
     * E on k-faces is +F_k(B_j). */
    const arr3 Fj = o->flux[jd][qb[kd]];
    const arr3 Fk = o->flux[kd][qb[jd]];
    const arr3 Wj = o->weight[jd], Wk = o->weight[kd];
    /* trimmed cell extents are m-2s; loop starts at 1 along d, 0 along j,k
     * and stops at (extent-1) on every axis (CT.cpp:548-556) */
    int lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      lo[a] = s + ((a == d) ? 1 : 0);
      hi[a] = m[a] - s - 1;
    }
    for (int k = lo[2]; k < hi[2]; k++)
      for (int j = lo[1]; j < hi[1]; j++)
        for (int i = lo[0]; i < hi[0]; i++) {
          double Ec = AT(o->ecen, k, j, i);
          double Ec_jp1 = AT(o->ecen, k + jz, j + jy, i + jx);
          double Ec_kp1 = AT(o->ecen, k + kz, j + ky, i + kx);
          double Ec_jkp1 = AT(o->ecen, k + jz + kz, j + jy + ky, i + jx + kx);
          double Ej = AT(Fj, k, j, i);
          double Ej_kp1 = AT(Fj, k + kz, j + ky, i + kx);
          double Ek = AT(Fk, k, j, i);
          double Ek_jp1 = AT(Fk, k + jz, j + jy, i + jx);
          double wj = AT(Wj, k, j, i);
          double wj_kp1 = AT(Wj, k + kz, j + ky, i + kx);
          double wk = AT(Wk, k, j, i);
          double wk_jp1 = AT(Wk, k + jz, j + jy, i + jx);

          double dEdj_r = wk_jp1 * (Ec_jp1 + Ej) + (1 - wk_jp1) * (Ec_jkp1 + Ej_kp1);
          double dEdj_l = wk * (-Ej - Ec) + (1 - wk) * (-Ej_kp1 - Ec_kp1);
          double dEdk_r = wj_kp1 * (Ec_kp1 - Ek) + (1 - wj_kp1) * (Ec_jkp1 - Ek_jp1);
          double dEdk_l = wj * (Ek - Ec) + (1 - wj) * (Ek_jp1 - Ec_jp1);

          double Ej_sum = Ej + Ej_kp1;
          Ej_sum *= -1;
          double Ek_sum = Ek + Ek_jp1;
          AT(o->edge[d], k, j, i) = 0.25 * (Ej_sum + Ek_sum + (dEdj_l - dEdj_r) +
                                            (dEdk_l - dEdk_r));
        }
  }

  /* face-B update (CT.cpp:617-687): face index f of bfieldi_d <-> f-1/2 */
  for (int d = 0; d < 3; d++) {
    const int jd = (d + 1) % 3, kd = (d + 2) % 3;
    const int jx = (jd == 0), jy = (jd == 1), jz = (jd == 2);
    const int kx = (kd == 0), ky = (kd == 1), kz = (kd == 2);
    const int dx_ = (d == 0), dy_ = (d == 1), dz_ = (d == 2);
    const double dtdj = dt / width[jd];
    const double dtdk = dt / width[kd];
    const arr3 E_j = o->edge[jd], E_k = o->edge[kd];
    int lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      /* interior faces along d: trimmed face extent m+1-2s, slice (1,-1);
       * inner cells along j,k: trimmed extent m-2s, slice (1,-1) */
      lo[a] = s + 1;
      hi[a] = (a == d) ? (m[a] + 1 - s - 1) : (m[a] - s - 1);
    }
    for (int k = lo[2]; k < hi[2]; k++)
      for (int j = lo[1]; j < hi[1]; j++)
        for (int i = lo[0]; i < hi[0]; i++) {
          /* edge arrays: index e along a face-centred axis <-> e+1/2; the
           * face f along d is edge index f-1 */
          const int ek = k - dz_, ej = j - dy_, ei = i - dx_;
          double ek_Rj = AT(E_k, ek, ej, ei);
          double ek_Lj = AT(E_k, ek - jz, ej - jy, ei - jx);
          double ej_Rk = AT(E_j, ek, ej, ei);
          double ej_Lk = AT(E_j, ek - kz, ej - ky, ei - kx);
          double E_k_term = dtdj * (ek_Rj - ek_Lj);
          double E_j_term = dtdk * (ej_Rk - ej_Lk);
          AT(bi_out[d], k, j, i) = AT(bi0[d], k, j, i) - E_k_term + E_j_term;
        }
  }

  /* cell-centred B = average of the two faces (CT.cpp:702-728) */
  for (int d = 0; d < 3; d++) {
    const int dx_ = (d == 0), dy_ = (d == 1), dz_ = (d == 2);
    for (int k = s; k < m[2] - s; k++)
      for (int j = s; j < m[1] - s; j++)
        for (int i = s; i < m[0] - s; i++) {
          AT(out[qb[d]], k, j, i) = 0.5 * (AT(bi_out[d], k, j, i) +
                                           AT(bi_out[d], k + dz_, j + dy_, i + dx_));
        }
  }
}

/* EnzoPhysicsFluidProps::apply_floor_to_energy_and_sync
 * (fluid-props/EnzoPhysicsFluidProps.cpp:162-290) */
static void floor_energy_and_sync(const vlct_oracle *o, arr3 *u, int s)
{
  const int mz = u[Q_RHO].n0, my = u[Q_RHO].n1, mx = u[Q_RHO].n2;
  const double gamma = o->cfg.gamma;
  const double eta = o->de ? o->cfg.dual_energy_eta : 0.0;
  float ggm1 = gamma * (gamma - 1.);            /* single precision: cpp:234 */
  double pressure_floor = o->cfg.pressure_floor;
  double inv_gm1 = 1. / (gamma - 1.);
  const double half_factor = (eta != 0.) ? 0.5 : 0.;

  for (int k = s; k < mz - s; k++)
    for (int j = s; j < my - s; j++)
      for (int i = s; i < mx - s; i++) {
        double inv_rho = 1. / AT(u[Q_RHO], k, j, i);
        double eint_floor = pressure_floor * inv_gm1 * inv_rho;
        double vx = AT(u[Q_VX], k, j, i), vy = AT(u[Q_VY], k, j, i),
               vz = AT(u[Q_VZ], k, j, i);
        double v2 = (vx * vx + vy * vy + vz * vz);
        double non_thermal_e = 0.5 * v2;
        double b2 = 0;
        if (o->mhd) {
          double bx = AT(u[Q_BX], k, j, i), by = AT(u[Q_BY], k, j, i),
                 bz = AT(u[Q_BZ], k, j, i);
          b2 = (bx * bx + by * by + bz * bz);
          non_thermal_e += (0.5 * b2 * inv_rho);
        }
        if (o->de) {
          double eint_1 = AT(u[Q_EN], k, j, i) - non_thermal_e;
          double cur_eint = AT(u[Q_EINT], k, j, i);
          double cs2_1 = fmax(0., ggm1 * eint_1);
          if ((cs2_1 > fmax(eta * v2, eta * b2 * inv_rho)) &&
              (eint_1 > half_factor * cur_eint)) {
            cur_eint = eint_1;
          }
          cur_eint = apply_floor(cur_eint, eint_floor);
          AT(u[Q_EINT], k, j, i) = cur_eint;
          AT(u[Q_EN], k, j, i) = cur_eint + non_thermal_e;
        } else {
          double etot_floor = eint_floor + non_thermal_e;
          AT(u[Q_EN], k, j, i) = apply_floor(AT(u[Q_EN], k, j, i), etot_floor);
        }
      }
}

/* EnzoIntegrationQuanUpdate::update_quantities
 * (toolkit/EnzoIntegrationQuanUpdate.cpp:183-268) */
static void update_quantities(vlct_oracle *o, const arr3 *u0, arr3 *out, int s)
{
  const int mz = o->mz, my = o->my, mx = o->mx;
  const double density_floor = o->cfg.density_floor;

  for (int sc = 0; sc < o->nsc; sc++) {
    const int q = Q_SC0 + sc;
    for (int k = s + 1; k < mz - s - 1; k++)
      for (int j = s + 1; j < my - s - 1; j++)
        for (int i = s + 1; i < mx - s - 1; i++)
          AT(out[q], k, j, i) = AT(u0[q], k, j, i) + AT(o->dU[q], k, j, i);
  }

  for (int k = s + 1; k < mz - s - 1; k++)
    for (int j = s + 1; j < my - s - 1; j++)
      for (int i = s + 1; i < mx - s - 1; i++) {
        double old_rho = AT(u0[Q_RHO], k, j, i);
        double new_rho = old_rho + AT(o->dU[Q_RHO], k, j, i);
        new_rho = apply_floor(new_rho, density_floor);
        AT(out[Q_RHO], k, j, i) = new_rho;
        double inv_new_rho = 1. / new_rho;
        const int specific[5] = { Q_VX, Q_VY, Q_VZ, Q_EN, Q_EINT };
        for (int n = 0; n < 5; n++) {
          const int q = specific[n];
          if (q == Q_EINT && !o->de) continue;
          AT(out[q], k, j, i) =
            (AT(u0[q], k, j, i) * old_rho + AT(o->dU[q], k, j, i)) * inv_new_rho;
        }
      }

  floor_energy_and_sync(o, out, s + 1);
}

static int total_staling(int recon) { return (recon == VLCT_RECON_NN) ? 1 : 2; }
static int immediate_staling(int recon) { return (recon == VLCT_RECON_NN) ? 0 : 1; }

/* ------------------------------------------------------------------------ */
/* public entry points                                                       */
/* ------------------------------------------------------------------------ */

vlct_oracle *vlct_oracle_create(const vlct_config *cfg, int gx, int gy, int gz)
{
  vlct_oracle *o = (vlct_oracle *) calloc(1, sizeof(vlct_oracle));
  o->cfg = *cfg;
  if (o->cfg.courant < 0)
    o->cfg.courant = (cfg->time_scheme == VLCT_TIME_VL) ? 0.3 : 1.0;
  o->mhd = (cfg->mhd_choice == VLCT_MHD_CONSTRAINED_TRANSPORT);
  o->de = (cfg->dual_energy != VLCT_DE_DISABLED);
  o->nsc = cfg->n_passive;
  o->gx = gx; o->gy = gy; o->gz = gz;
  return o;
}

void vlct_oracle_destroy(vlct_oracle *o)
{
  if (o == NULL) return;
  if (o->mx) free_scratch(o);
  free(o);
}

/* EnzoMethodMHDVlct::compute (hydro-mhd/EnzoMethodMHDVlct.cpp:356-500) */
int vlct_oracle_compute(vlct_oracle *o, const vlct_block *b, double dt)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  if (o->mx == 0) alloc_scratch(o, mx, my, mz);
  if (mx != o->mx || my != o->my || mz != o->mz) return 2;

  arr3 ext[Q_MAX], bi[3];
  wrap_block(o, b, ext, bi);
  const double width[3] = { b->dx, b->dy, b->dz };
  const int nstages = (o->cfg.time_scheme == VLCT_TIME_EULER) ? 1 : 2;
  int stale = 0;

  for (int stage = 0; stage < nstages; stage++) {
    const int final = (stage + 1) == nstages;
    const double cur_dt = (!final) ? dt / 2. : dt;
    const int recon = (nstages == 2 && stage == 0) ? VLCT_RECON_NN
                                                   : o->cfg.reconstruct_method;
    const arr3 *cur = (stage == 0) ? ext : o->temp;
    arr3 *out = final ? ext : o->temp;
    /* CT state machine (toolkit/EnzoBfieldMethodCT.cpp:122-137,220-240) */
    const arr3 *bi_cur = (stage == 0) ? bi : o->tbi;
    arr3 *bi_out = (stage == 1 || nstages == 1) ? bi : o->tbi;

    /* clear dU (toolkit/EnzoIntegrationQuanUpdate.cpp:83-101) */
    for (int q = 0; q < Q_MAX; q++)
      if (has_q(o, q, 0, 1))
        memset(o->dU[q].p, 0, sizeof(double) * (size_t) mx * my * mz);

    primitive_from_integration(o, cur, stale);

    for (int dim = 0; dim < 3; dim++) {
      compute_flux_dim(o, dim, recon, stale, bi_cur);
      accumulate_dim(o, dim, cur_dt, width[dim], stale + immediate_staling(recon));
    }
    int s = stale + immediate_staling(recon);

    if (stage == 1 && o->cfg.has_acceleration && b->acceleration_x != NULL)
      gravity_source(o, ext, b, cur_dt, s);

    if (o->mhd) ct_update(o, cur, bi, bi_out, out, cur_dt, width, s);

    update_quantities(o, ext, out, s);

    stale += total_staling(recon);
  }
  return 0;
}

/* EnzoMethodMHDVlct::save_fluxes_for_corrections_
 * (hydro-mhd/EnzoMethodMHDVlct.cpp:250-330): dt/dx times the final-stage flux
 * through the block's two faces along every dimension, over the active
 * transverse extent, for the "conserved" fields. Uses the flux arrays the last
 * vlct_oracle_compute left behind. out[dim][side][field]: packed 2-D arrays,
 * slower axis first -- (z,y) for dim 0, (z,x) for dim 1, (y,x) for dim 2;
 * field slots as in vlct_face_fluxes (include/vlct.h); NULL = skip. */
int vlct_oracle_face_fluxes(const vlct_oracle *o, const vlct_block *b, double dt,
                            double *const out[3][2][6 + VLCT_MAX_PASSIVE])
{
  if (o->mx == 0) return 2;
  const int slot_q[6] = { Q_RHO, Q_VX, Q_VY, Q_VZ, Q_EN, Q_EINT };
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  const int m[3] = { o->mx, o->my, o->mz };
  const double width[3] = { b->dx, b->dy, b->dz };
  for (int dim = 0; dim < 3; dim++) {
    const double dt_dxi = dt / width[dim];
    const int left = g[dim] - 1, right = m[dim] - g[dim] - 1;
    for (int f = 0; f < 6 + o->nsc; f++) {
      const int q = (f < 6) ? slot_q[f] : Q_SC0 + (f - 6);
      if (q == Q_EINT && !o->de) continue;
      const arr3 F = o->flux[dim][q];
      for (int side = 0; side < 2; side++) {
        double *dst = out[dim][side][f];
        if (dst == NULL) continue;
        const int at = side ? right : left;
        /* transverse axes: (a1 slower, a0 faster) */
        const int a0 = (dim == 0) ? 1 : 0, a1 = (dim == 2) ? 1 : 2;
        for (int i1 = 0; i1 < n[a1]; i1++)
          for (int i0 = 0; i0 < n[a0]; i0++) {
            int idx[3];
            idx[dim] = at; idx[a0] = g[a0] + i0; idx[a1] = g[a1] + i1;
            dst[(size_t) i1 * n[a0] + i0] =
              dt_dxi * F.p[((size_t) idx[2] * F.n1 + idx[1]) * F.n2 + idx[0]];
          }
      }
    }
  }
  return 0;
}

/* EnzoMethodMHDVlct::timestep (hydro-mhd/EnzoMethodMHDVlct.cpp:551-588) and
 * EnzoMHDIntegratorStageCommands::timestep (StageCommands.cpp:299-366) */
int vlct_oracle_timestep(vlct_oracle *o, const vlct_block *b, double *dt_out)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  arr3 u[Q_MAX];
  wrap_block(o, b, u, NULL);
  if (o->de) floor_energy_and_sync(o, u, 0);
  arr3 p = arr3_wrap(b->pressure, mz, my, mx);
  pressure_from_integration(o, u, p, 0);

  const double gamma = o->cfg.gamma;
  const double dx = b->dx, dy = b->dy, dz = b->dz;
  double dtBaryons = 1.7976931348623157e308;
  for (int k = 0; k < mz; k++)
    for (int j = 0; j < my; j++)
      for (int i = 0; i < mx; i++) {
        double c;
        if (o->mhd) {
          c = eos_cfast_max(gamma, AT(u[Q_RHO], k, j, i), AT(p, k, j, i),
                            AT(u[Q_BX], k, j, i), AT(u[Q_BY], k, j, i),
                            AT(u[Q_BZ], k, j, i));
        } else {
          c = sqrt(eos_cs2(gamma, AT(u[Q_RHO], k, j, i), AT(p, k, j, i)));
        }
        double local_dt = min3(dx / (fabs(AT(u[Q_VX], k, j, i)) + c),
                               dy / (fabs(AT(u[Q_VY], k, j, i)) + c),
                               dz / (fabs(AT(u[Q_VZ], k, j, i)) + c));
        dtBaryons = std_min(dtBaryons, local_dt);
      }
  *dt_out = dtBaryons * o->cfg.courant;
  return 0;
}
