/* vlct_oracle_ic.c -- TEST INFRASTRUCTURE ONLY, not part of the product.
 *
 * CPU restatements of the reference's problem initialisers that the vlct
 * answer tests use, so that the golden L1 norms of input/vlct/run_*_test.py
 * can be reproduced on raw arrays -- also on the GPU box, where the reference
 * tree does not exist. Each is pinned bit for bit against the reference's own
 * Initial class compiled into oracle/_ref (tests/test_oracle_golden.py:
 * test_*_restatement_equals_compiled_reference):
 *
 *   vlct_ic_inclined_wave   initial/EnzoInitialInclinedWave.cpp
 *   vlct_ic_shock_tube      initial/EnzoInitialShockTube.cpp
 *   vlct_ic_cloud[_perturbed] initial/EnzoInitialCloud.cpp (cloud in a wind, optional density perturbation)
 *   (face-B from a vector potential, centred B)  initial/EnzoInitialBCenter.cpp
 *
 * plus the periodic ghost-zone refresh of a single block
 * (vlct_oracle_refresh_periodic), which stands in for the Charm++ refresh
 * phase (Cello/control_refresh.cpp, Cello/data_FieldFace.cpp) when
 * Mesh:root_blocks = [1,1,1] and Boundary:type = "periodic".
 *
 * Citations are relative to /root/reference/src/Enzo/ unless noted.
 * sin/cos come from the platform libm, as in the reference.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vlct.h"

static const double cello_pi = 3.14159265358979324; /* Cello/cello.hpp:623 */

/* ---- Rotation (EnzoInitialInclinedWave.cpp:54-117) ---------------------- */
typedef struct { double m[3][3]; } rotation;

static rotation rotation_make(double a, double b)
{
  rotation r;
  r.m[0][0] = cos(a) * cos(b);
  r.m[0][1] = cos(a) * sin(b);
  r.m[0][2] = sin(a);
  r.m[1][0] = -1. * sin(b);
  r.m[1][1] = cos(b);
  r.m[1][2] = 0.;
  r.m[2][0] = -1. * sin(a) * cos(b);
  r.m[2][1] = -1. * sin(a) * sin(b);
  r.m[2][2] = cos(a);
  return r;
}

static void rot_fwd(const rotation *r, double v0, double v1, double v2,
                    double *o0, double *o1, double *o2)
{
  *o0 = r->m[0][0] * v0 + r->m[0][1] * v1 + r->m[0][2] * v2;
  *o1 = r->m[1][0] * v0 + r->m[1][1] * v1 + r->m[1][2] * v2;
  *o2 = r->m[2][0] * v0 + r->m[2][1] * v1 + r->m[2][2] * v2;
}

static void rot_inv(const rotation *r, double r0, double r1, double r2,
                    double *v0, double *v1, double *v2)
{
  *v0 = r->m[0][0] * r0 + r->m[1][0] * r1 + r->m[2][0] * r2;
  *v1 = r->m[0][1] * r0 + r->m[1][1] * r1 + r->m[2][1] * r2;
  *v2 = r->m[0][2] * r0 + r->m[1][2] * r1 + r->m[2][2] * r2;
}

/* ---- the wave description ------------------------------------------------ */
typedef struct {
  rotation rot;
  double amplitude, lambda;
  /* linear (conserved-form) initialisers */
  double rho_back, rho_ev, etot_back, etot_ev;
  double mom_back[3], mom_ev[3];
  int use_cosine;
  /* vector potential */
  int has_a, circ;               /* circ: circularly polarised Alfven wave */
  double b_back[3], b1_ev, b2_ev;
} wave_t;

/* LinearScalarInit (cpp:140-168) */
static double linear_scalar(const wave_t *w, double back, double ev, double x0)
{
  double tmp = x0 * 2. * cello_pi / w->lambda;
  return (back + w->amplitude * ev *
          (w->use_cosine * cos(tmp) + (1 - w->use_cosine) * sin(tmp)));
}

static double rotated_scalar(const wave_t *w, double back, double ev,
                             double x, double y, double z)
{
  double x0, x1, x2;
  rot_fwd(&w->rot, x, y, z, &x0, &x1, &x2);
  return linear_scalar(w, back, ev, x0);
}

/* RotatedVectorInit(LinearVectorInit) (cpp:232-299) */
static void rotated_momentum(const wave_t *w, double x, double y, double z,
                             double *v0, double *v1, double *v2)
{
  double x0, x1, x2, r0, r1, r2;
  rot_fwd(&w->rot, x, y, z, &x0, &x1, &x2);
  double tmp = x0 * 2. * cello_pi / w->lambda;
  double trig_term = (w->use_cosine * cos(tmp) + (1 - w->use_cosine) * sin(tmp));
  r0 = w->mom_back[0] + w->amplitude * w->mom_ev[0] * trig_term;
  r1 = w->mom_back[1] + w->amplitude * w->mom_ev[1] * trig_term;
  r2 = w->mom_back[2] + w->amplitude * w->mom_ev[2] * trig_term;
  rot_inv(&w->rot, r0, r1, r2, v0, v1, v2);
}

/* RotatedVectorInit(LinearVectorPotentialInit) (cpp:303-336) or the
 * circularly polarised variant (cpp:899-908) */
static void rotated_vector_potential(const wave_t *w, double x, double y,
                                     double z, double *a0, double *a1,
                                     double *a2)
{
  double x0, x1, x2, r0, r1, r2;
  rot_fwd(&w->rot, x, y, z, &x0, &x1, &x2);
  if (w->circ) {
    r0 = (x2 * 0.1 * sin(2. * cello_pi * x0 / w->lambda) -
          x1 * 0.1 * cos(2. * cello_pi * x0 / w->lambda));
    r1 = 0.0;
    r2 = x1;
  } else {
    r0 = (x2 * w->amplitude * w->b1_ev * cos(2. * cello_pi * x0 / w->lambda) -
          x1 * w->amplitude * w->b2_ev * cos(2. * cello_pi * x0 / w->lambda));
    r1 = w->b_back[2] * x0;
    r2 = w->b_back[0] * x1 - w->b_back[1] * x0;
  }
  rot_inv(&w->rot, r0, r1, r2, a0, a1, a2);
}

#define C3(p, k, j, i, n1, n2) \
  ((p)[((size_t)(k) * (size_t)(n1) + (size_t)(j)) * (size_t)(n2) + (size_t)(i)])

/* cell-centred B = face average
 * (hydro-mhd/toolkit/EnzoBfieldMethodCT.cpp:702-728 with stale_depth 0) */
static void center_bfield(const vlct_block *b)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  for (int k = 0; k < mz; k++)
    for (int j = 0; j < my; j++)
      for (int i = 0; i < mx; i++) {
        C3(b->bfield_x, k, j, i, my, mx) =
          0.5 * (C3(b->bfieldi_x, k, j, i, my, mx + 1) +
                 C3(b->bfieldi_x, k, j, i + 1, my, mx + 1));
        C3(b->bfield_y, k, j, i, my, mx) =
          0.5 * (C3(b->bfieldi_y, k, j, i, my + 1, mx) +
                 C3(b->bfieldi_y, k, j + 1, i, my + 1, mx));
        C3(b->bfield_z, k, j, i, my, mx) =
          0.5 * (C3(b->bfieldi_z, k, j, i, my, mx) +
                 C3(b->bfieldi_z, k + 1, j, i, my, mx));
      }
}

int vlct_ic_center_bfield(const vlct_block *b) { center_bfield(b); return 0; }

/* setup_bfield (EnzoInitialInclinedWave.cpp:426-490) +
 * EnzoInitialBCenter::initialize_bfield_interface (EnzoInitialBCenter.cpp:43-127) */
static void setup_bfield(const vlct_block *b, const double *lower,
                         const wave_t *w)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const double dx = b->dx, dy = b->dy, dz = b->dz;
  if (!w->has_a) {
    memset(b->bfieldi_x, 0, sizeof(double) * (size_t) mz * my * (mx + 1));
    memset(b->bfieldi_y, 0, sizeof(double) * (size_t) mz * (my + 1) * mx);
    memset(b->bfieldi_z, 0, sizeof(double) * (size_t) (mz + 1) * my * mx);
    center_bfield(b);
    return;
  }
  double *Ax = (double *) calloc((size_t) (mz + 1) * (my + 1) * mx, sizeof(double));
  double *Ay = (double *) calloc((size_t) (mz + 1) * my * (mx + 1), sizeof(double));
  double *Az = (double *) calloc((size_t) mz * (my + 1) * (mx + 1), sizeof(double));

  for (int k = 0; k < mz + 1; k++)
    for (int j = 0; j < my + 1; j++)
      for (int i = 0; i < mx + 1; i++) {
        /* MeshPos (cpp:375-422) */
        double xc = lower[0] + dx * (0.5 + (double) (i - b->gx));
        double yc = lower[1] + dy * (0.5 + (double) (j - b->gy));
        double zc = lower[2] + dz * (0.5 + (double) (k - b->gz));
        double xf = lower[0] + dx * (double) (i - b->gx);
        double yf = lower[1] + dy * (double) (j - b->gy);
        double zf = lower[2] + dz * (double) (k - b->gz);
        double t0, t1, t2;
        if (i != mx) {
          rotated_vector_potential(w, xc, yf, zf, &t0, &t1, &t2);
          C3(Ax, k, j, i, my + 1, mx) = t0;
        }
        if (j != my) {
          rotated_vector_potential(w, xf, yc, zf, &t0, &t1, &t2);
          C3(Ay, k, j, i, my, mx + 1) = t1;
        }
        if (k != mz) {
          rotated_vector_potential(w, xf, yf, zc, &t0, &t1, &t2);
          C3(Az, k, j, i, my + 1, mx + 1) = t2;
        }
      }

  /* B_i = dA_k/dj - dA_j/dk on the face (EnzoInitialBCenter.cpp:43-96) */
  for (int k = 0; k < mz; k++)
    for (int j = 0; j < my; j++)
      for (int i = 0; i < mx + 1; i++)
        C3(b->bfieldi_x, k, j, i, my, mx + 1) =
          ((C3(Az, k, j + 1, i, my + 1, mx + 1) - C3(Az, k, j, i, my + 1, mx + 1)) / dy -
           (C3(Ay, k + 1, j, i, my, mx + 1) - C3(Ay, k, j, i, my, mx + 1)) / dz);
  for (int k = 0; k < mz; k++)
    for (int j = 0; j < my + 1; j++)
      for (int i = 0; i < mx; i++)
        C3(b->bfieldi_y, k, j, i, my + 1, mx) =
          ((C3(Ax, k + 1, j, i, my + 1, mx) - C3(Ax, k, j, i, my + 1, mx)) / dz -
           (C3(Az, k, j, i + 1, my + 1, mx + 1) - C3(Az, k, j, i, my + 1, mx + 1)) / dx);
  for (int k = 0; k < mz + 1; k++)
    for (int j = 0; j < my; j++)
      for (int i = 0; i < mx; i++)
        C3(b->bfieldi_z, k, j, i, my, mx) =
          ((C3(Ay, k, j, i + 1, my, mx + 1) - C3(Ay, k, j, i, my, mx + 1)) / dx -
           (C3(Ax, k, j + 1, i, my + 1, mx) - C3(Ax, k, j, i, my + 1, mx)) / dy);
  free(Ax); free(Ay); free(Az);
  center_bfield(b);
}

/* EnzoInitialInclinedWave::enforce_block and helpers
 * (EnzoInitialInclinedWave.cpp:722-772, 885-1147).
 * parallel_vel: pass DBL_MIN for "not specified" (cpp:655-656). */
int vlct_ic_inclined_wave(const vlct_block *b, const double *lower, double gamma,
                          const char *wave_type, double alpha, double beta,
                          double amplitude, double lambda, int positive_vel,
                          double parallel_vel)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int mhd = (b->bfield_x != NULL);
  wave_t w;
  memset(&w, 0, sizeof(w));
  w.rot = rotation_make(alpha, beta);
  w.amplitude = amplitude;
  w.lambda = lambda;
  w.use_cosine = 1;
  const double wsign = positive_vel ? 1. : -1.;
  const int is_hd = (!strcmp(wave_type, "sound") || !strcmp(wave_type, "hd_entropy") ||
                     !strcmp(wave_type, "hd_transv_entropy_v1") ||
                     !strcmp(wave_type, "hd_transv_entropy_v2"));
  int primitive_form = 0;

  if (is_hd) {
    /* prepare_HD_initializers_ (cpp:1022-1122) */
    double v0_back = 0, v1_back = 0, v2_back = 0;
    if (parallel_vel != DBL_MIN) v0_back = parallel_vel;
    else if (strcmp(wave_type, "sound") != 0) v0_back = wsign;
    double squared_v_back = v0_back * v0_back + v1_back * v1_back + v2_back * v2_back;
    w.rho_back = 1;
    w.mom_back[0] = v0_back; w.mom_back[1] = v1_back; w.mom_back[2] = v2_back;
    w.etot_back = ((1. / gamma) / (gamma - 1.) + 0.5 * squared_v_back);
    if (!strcmp(wave_type, "sound")) {
      double h_back = 1 / (gamma - 1.) + 0.5 * squared_v_back;
      double signed_cs = wsign * 1;
      w.rho_ev = 1;
      w.mom_ev[0] = v0_back + signed_cs;
      w.mom_ev[1] = v1_back;
      w.mom_ev[2] = v2_back;
      w.etot_ev = h_back + v0_back * signed_cs;
    } else if (!strcmp(wave_type, "hd_entropy")) {
      w.rho_ev = 1;
      w.mom_ev[0] = v0_back; w.mom_ev[1] = v1_back; w.mom_ev[2] = v2_back;
      w.etot_ev = 0.5 * squared_v_back;
    } else if (!strcmp(wave_type, "hd_transv_entropy_v1")) {
      w.rho_ev = 0; w.mom_ev[0] = 0; w.mom_ev[1] = 1; w.mom_ev[2] = 0;
      w.etot_ev = v1_back;
    } else {
      w.rho_ev = 0; w.mom_ev[0] = 0; w.mom_ev[1] = 0; w.mom_ev[2] = 1;
      w.etot_ev = v2_back;
    }
    w.has_a = 0;
  } else if (!strcmp(wave_type, "circ_alfven")) {
    if (!mhd) return 1;
    w.has_a = 1; w.circ = 1;
    primitive_form = 1;
  } else {
    /* prepare_MHD_initializers_ (cpp:941-1017) */
    if (!mhd) return 1;
    w.rho_back = 1;
    w.mom_back[0] = 0; w.mom_back[1] = 0; w.mom_back[2] = 0;
    w.b_back[0] = 1.; w.b_back[1] = 1.5; w.b_back[2] = 0.0;
    w.etot_back = (1. / gamma) / (gamma - 1.) + 1.625;
    if (!strcmp(wave_type, "mhd_entropy")) {
      w.mom_back[0] = wsign;
      w.etot_back += 0.5;
    }
    if (!strcmp(wave_type, "fast")) {
      double coef = 0.5 / sqrt(5.);
      w.rho_ev = 2. * coef;
      w.mom_ev[0] = wsign * 4. * coef;
      w.mom_ev[1] = -1. * wsign * 2. * coef;
      w.mom_ev[2] = 0;
      w.etot_ev = 9. * coef;
      w.b1_ev = 4. * coef;
      w.b2_ev = 0;
    } else if (!strcmp(wave_type, "alfven")) {
      w.rho_ev = 0;
      w.mom_ev[0] = 0; w.mom_ev[1] = 0; w.mom_ev[2] = -1. * wsign * 1;
      w.etot_ev = 0;
      w.b1_ev = 0; w.b2_ev = 1.;
    } else if (!strcmp(wave_type, "slow")) {
      double coef = 0.5 / sqrt(5.);
      w.rho_ev = 4. * coef;
      w.mom_ev[0] = wsign * 2. * coef;
      w.mom_ev[1] = wsign * 4. * coef;
      w.mom_ev[2] = 0;
      w.etot_ev = 3. * coef;
      w.b1_ev = -2. * coef;
      w.b2_ev = 0;
    } else if (!strcmp(wave_type, "mhd_entropy")) {
      double coef = 0.5;
      w.rho_ev = 2. * coef;
      w.mom_ev[0] = 2. * coef * wsign;
      w.mom_ev[1] = 0; w.mom_ev[2] = 0;
      w.etot_ev = 1. * coef;
      w.b1_ev = 0; w.b2_ev = 0;
    } else {
      return 2; /* unknown wave type */
    }
    w.has_a = 1;
  }

  if (mhd) setup_bfield(b, lower, &w);

  /* setup_fluid_ (cpp:543-643) */
  for (int k = 0; k < mz; k++)
    for (int j = 0; j < my; j++)
      for (int i = 0; i < mx; i++) {
        double x = lower[0] + b->dx * (0.5 + (double) (i - b->gx));
        double y = lower[1] + b->dy * (0.5 + (double) (j - b->gy));
        double z = lower[2] + b->dz * (0.5 + (double) (k - b->gz));
        if (!primitive_form) {
          double rho = rotated_scalar(&w, w.rho_back, w.rho_ev, x, y, z);
          double px, py, pz;
          C3(b->density, k, j, i, my, mx) = rho;
          rotated_momentum(&w, x, y, z, &px, &py, &pz);
          C3(b->velocity_x, k, j, i, my, mx) = (px / rho);
          C3(b->velocity_y, k, j, i, my, mx) = (py / rho);
          C3(b->velocity_z, k, j, i, my, mx) = (pz / rho);
          double etot_dens = rotated_scalar(&w, w.etot_back, w.etot_ev, x, y, z);
          C3(b->total_energy, k, j, i, my, mx) = (etot_dens / rho);
        } else {
          /* circularly polarised Alfven wave (cpp:924-940, 582-642) */
          double rho = 1.0, pressure = 0.1;
          double x0, x1, x2, vx, vy, vz;
          rot_fwd(&w.rot, x, y, z, &x0, &x1, &x2);
          double r0 = 0.0;
          double r1 = 0.1 * sin(2. * cello_pi * x0 / lambda);
          double r2 = 0.1 * cos(2. * cello_pi * x0 / lambda);
          rot_inv(&w.rot, r0, r1, r2, &vx, &vy, &vz);
          C3(b->density, k, j, i, my, mx) = rho;
          C3(b->velocity_x, k, j, i, my, mx) = vx;
          C3(b->velocity_y, k, j, i, my, mx) = vy;
          C3(b->velocity_z, k, j, i, my, mx) = vz;
          const double inv_gm1 = 1.0 / (gamma - 1.0);
          double inv_rho = 1.0 / rho;
          double eint = inv_gm1 * pressure * inv_rho;
          double non_thermal_e = 0.5 * ((vx * vx) + (vy * vy) + (vz * vz));
          double bx = C3(b->bfield_x, k, j, i, my, mx);
          double by = C3(b->bfield_y, k, j, i, my, mx);
          double bz = C3(b->bfield_z, k, j, i, my, mx);
          double b2 = ((bx * bx) + (by * by) + (bz * bz));
          non_thermal_e += 0.5 * b2 * inv_rho;
          C3(b->total_energy, k, j, i, my, mx) = (eint + non_thermal_e);
          if (b->internal_energy) C3(b->internal_energy, k, j, i, my, mx) = eint;
        }
      }

  /* setup_eint_ (cpp:495-539) */
  if (!primitive_form && b->internal_energy != NULL) {
    for (int k = 0; k < mz; k++)
      for (int j = 0; j < my; j++)
        for (int i = 0; i < mx; i++) {
          double vx = C3(b->velocity_x, k, j, i, my, mx);
          double vy = C3(b->velocity_y, k, j, i, my, mx);
          double vz = C3(b->velocity_z, k, j, i, my, mx);
          double kinetic = 0.5 * (vx * vx + vy * vy + vz * vz);
          double magnetic = 0.;
          if (mhd) {
            double bx = C3(b->bfield_x, k, j, i, my, mx);
            double by = C3(b->bfield_y, k, j, i, my, mx);
            double bz = C3(b->bfield_z, k, j, i, my, mx);
            magnetic = 0.5 * (bx * bx + by * by + bz * bz) / C3(b->density, k, j, i, my, mx);
          }
          C3(b->internal_energy, k, j, i, my, mx) =
            C3(b->total_energy, k, j, i, my, mx) - kinetic - magnetic;
        }
  }
  return 0;
}

/* EnzoInitialShockTube::enforce_block (initial/EnzoInitialShockTube.cpp:37-330)
 * setup: "rj2a" or "sod"; aligned_ax 0/1/2 */
int vlct_ic_shock_tube(const vlct_block *b, const double *lower, double gamma,
                       const char *setup, int aligned_ax, double axis_velocity_,
                       double trans_velocity_, int flipped)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int mhd = (b->bfieldi_x != NULL);
  /* {density, pressure, v0, v1, v2, b1, b2} */
  double L[7], R[7], b0;
  if (!strcmp(setup, "rj2a")) {
    const double l[7] = { 1.08, 0.95, 1.2, 0.01, 0.5, 1.0155412503859613, 0.5641895835477563 };
    const double r[7] = { 1., 1.0, 0.0, 0.0, 0.0, 1.1283791670955126, 0.5641895835477563 };
    memcpy(L, l, sizeof(l)); memcpy(R, r, sizeof(r));
    b0 = 0.5641895835477563;
  } else if (!strcmp(setup, "sod")) {
    const double l[7] = { 1.0, 1.0, 0., 0., 0., 0., 0. };
    const double r[7] = { 0.125, 0.1, 0., 0., 0., 0., 0. };
    memcpy(L, l, sizeof(l)); memcpy(R, r, sizeof(r));
    b0 = 0.0;
  } else {
    return 1;
  }
  if (flipped) {
    double t[7];
    memcpy(t, L, sizeof(t)); memcpy(L, R, sizeof(t)); memcpy(R, t, sizeof(t));
    for (int n = 2; n < 7; n++) { L[n] = -1. * L[n]; R[n] = -1. * R[n]; }
  }
  const double flip = flipped ? -1. : 1.;
  const double aligned_bfield_val = flip * b0;
  const double axis_velocity = flip * axis_velocity_;
  const double trans_velocity = flip * trans_velocity_;

  const int m[3] = { mx, my, mz }, g[3] = { b->gx, b->gy, b->gz };
  const double h[3] = { b->dx, b->dy, b->dz };
  const int mi = m[aligned_ax], gi = g[aligned_ax];
  int shock_ind = (int) ceil((0.5 - lower[aligned_ax]) / h[aligned_ax] - 0.5 + (double) gi);
  if (shock_ind < 0) shock_ind = 0;
  if (shock_ind > mi) shock_ind = mi;

  double *vel[3] = { b->velocity_x, b->velocity_y, b->velocity_z };
  double *bfi[3] = { b->bfieldi_x, b->bfieldi_y, b->bfieldi_z };
  const int ia = aligned_ax, ja = (aligned_ax + 1) % 3, ka = (aligned_ax + 2) % 3;

  for (int side = 0; side < 2; side++) {
    const double *v = side ? R : L;
    const int lo = side ? shock_ind : 0, hi = side ? mi : shock_ind;
    if (lo >= hi) continue;
    double velocity_0 = v[2] + axis_velocity;
    double velocity_1 = v[3] + trans_velocity;
    double velocity_2 = v[4];
    double eint = (v[1] / ((gamma - 1.) * v[0]));
    double v2 = (velocity_0 * velocity_0 + velocity_1 * velocity_1 + velocity_2 * velocity_2);
    double b2 = (aligned_bfield_val * aligned_bfield_val + v[5] * v[5] + v[6] * v[6]);
    double etot = (eint + 0.5 * (v2 + b2 / v[0]));
    for (int k = 0; k < mz; k++)
      for (int j = 0; j < my; j++)
        for (int i = 0; i < mx; i++) {
          const int idx[3] = { i, j, k };
          if (idx[ia] < lo || idx[ia] >= hi) continue;
          C3(b->density, k, j, i, my, mx) = v[0];
          C3(vel[ia], k, j, i, my, mx) = velocity_0;
          C3(vel[ja], k, j, i, my, mx) = velocity_1;
          C3(vel[ka], k, j, i, my, mx) = velocity_2;
          if (b->internal_energy) C3(b->internal_energy, k, j, i, my, mx) = eint;
          C3(b->total_energy, k, j, i, my, mx) = etot;
        }
    if (mhd) {
      /* transverse face fields: the slice [lo,hi) along the aligned axis is
       * applied to the face-centred arrays as they are (cpp:246-250,381-390) */
      for (int t = 0; t < 2; t++) {
        const int ax = t ? ka : ja;
        const double val = t ? v[6] : v[5];
        const int n2 = mx + (ax == 0), n1 = my + (ax == 1), n0 = mz + (ax == 2);
        for (int k = 0; k < n0; k++)
          for (int j = 0; j < n1; j++)
            for (int i = 0; i < n2; i++) {
              const int idx[3] = { i, j, k };
              if (idx[ia] < lo || idx[ia] >= hi) continue;
              C3(bfi[ax], k, j, i, n1, n2) = val;
            }
      }
    }
  }
  if (mhd) {
    const int n2 = mx + (ia == 0), n1 = my + (ia == 1), n0 = mz + (ia == 2);
    for (size_t n = 0; n < (size_t) n0 * n1 * n2; n++) bfi[ia][n] = aligned_bfield_val;
    center_bfield(b);
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* periodic refresh of a single block                                        */
/* ------------------------------------------------------------------------ */

/* Fill the ghost layers along one axis from the block's own active zone.
 * cen = 1 if the field is face-centred along `axis`. A face-centred field has
 * n+1 active faces; the first and last coincide under periodicity and both
 * are owned (Cello/data_FieldFace.cpp: faces shared by two blocks are sent by
 * both and overwritten with the neighbour's identical value). */
static void wrap_axis(double *p, int n0, int n1, int n2, int axis, int n, int g,
                      int cen)
{
  const int ext[3] = { n2, n1, n0 };   /* extent along x,y,z */
  const int m = ext[axis];             /* = n + 2g + cen */
  (void) m;
  for (int k = 0; k < n0; k++)
    for (int j = 0; j < n1; j++)
      for (int i = 0; i < n2; i++) {
        int idx[3] = { i, j, k };
        int a = idx[axis];
        int src;
        if (a < g) src = a + n;                    /* lower ghosts */
        else if (a >= g + n + cen) src = a - n;    /* upper ghosts */
        else continue;
        idx[axis] = src;
        p[((size_t) k * n1 + j) * n2 + i] =
          p[((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0]];
      }
}

int vlct_oracle_refresh_periodic(const vlct_block *b, int n_passive, int axes)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  double *cell[32];
  int nc = 0;
  double *cands[] = { b->density, b->velocity_x, b->velocity_y, b->velocity_z,
                      b->total_energy, b->internal_energy, b->bfield_x,
                      b->bfield_y, b->bfield_z, b->pressure, b->acceleration_x,
                      b->acceleration_y, b->acceleration_z };
  for (size_t c = 0; c < sizeof(cands) / sizeof(cands[0]); c++)
    if (cands[c]) cell[nc++] = cands[c];
  for (int s = 0; s < n_passive; s++) cell[nc++] = b->passive[s];
  double *face[3] = { b->bfieldi_x, b->bfieldi_y, b->bfieldi_z };

  for (int axis = 0; axis < 3; axis++) {
    if (!(axes & (1 << axis))) continue;
    for (int c = 0; c < nc; c++)
      wrap_axis(cell[c], mz, my, mx, axis, n[axis], g[axis], 0);
    for (int f = 0; f < 3; f++) {
      if (!face[f]) continue;
      wrap_axis(face[f], mz + (f == 2), my + (f == 1), mx + (f == 0), axis,
                n[axis], g[axis], f == axis);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* non-periodic domain boundaries of a single block                          */
/* ------------------------------------------------------------------------ */

/* EnzoBoundary::enforce_outflow_precision_ / enforce_reflecting_precision_
 * (src/Enzo/enzo-core/EnzoBoundary.cpp:164-283, 352-466) for one field.
 * n0,n1,n2: array extents (z,y,x) incl. ghosts and centering; n, g: active
 * cells and ghost depth along `axis`; cen = 1 if the field is face-centred
 * along `axis`; side 0 = lower, 1 = upper; type 0 = outflow, 1 = reflecting;
 * sign: -1 for the vector component along `axis` (reflecting only). */
static void boundary_axis(double *p, int n0, int n1, int n2, int axis, int n,
                          int g, int cen, int side, int type, double sign)
{
  const int ext[3] = { n2, n1, n0 };
  for (int ig = 0; ig < g; ig++) {
    int src, dst;
    if (type == 0) {                 /* outflow: copy the outermost active value */
      if (side == 0) { src = g;               dst = g - ig - 1; }
      else           { src = n + g - 1 + cen; dst = src + ig + 1; }
    } else {                         /* reflecting: mirror image, signed */
      if (side == 0) { src = g + cen + ig;    dst = g - ig - 1; }
      else           { src = n + g - 1 - ig;  dst = n + g + ig + cen; }
    }
    int lim[3] = { ext[0], ext[1], ext[2] };
    lim[axis] = 1;
    for (int k = 0; k < lim[2]; k++)
      for (int j = 0; j < lim[1]; j++)
        for (int i = 0; i < lim[0]; i++) {
          int is[3] = { i, j, k }, id[3] = { i, j, k };
          is[axis] = src; id[axis] = dst;
          const double v = p[((size_t) is[2] * n1 + is[1]) * n2 + is[0]];
          p[((size_t) id[2] * n1 + id[1]) * n2 + id[0]] =
            (type == 0) ? v : sign * v;
        }
  }
}

/* Block::update_boundary_ (src/Cello/mesh_Block.cpp:1057-1077) for one face of
 * the domain: every permanent field of the block. type 0 = "outflow",
 * 1 = "reflecting". */
int vlct_oracle_boundary(const vlct_block *b, int n_passive, int axis, int side,
                         int type)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  if (axis < 0 || axis > 2 || (side != 0 && side != 1) || (type != 0 && type != 1))
    return 1;
  /* vector components: has_vector_name_ (EnzoBoundary.cpp:80-86) */
  struct { double *p; int comp; } cell[] = {
    { b->density, -1 }, { b->velocity_x, 0 }, { b->velocity_y, 1 },
    { b->velocity_z, 2 }, { b->total_energy, -1 }, { b->internal_energy, -1 },
    { b->bfield_x, 0 }, { b->bfield_y, 1 }, { b->bfield_z, 2 },
    { b->pressure, -1 }, { b->acceleration_x, -1 }, { b->acceleration_y, -1 },
    { b->acceleration_z, -1 } };
  for (size_t c = 0; c < sizeof(cell) / sizeof(cell[0]); c++)
    if (cell[c].p)
      boundary_axis(cell[c].p, mz, my, mx, axis, n[axis], g[axis], 0, side, type,
                    cell[c].comp == axis ? -1.0 : 1.0);
  for (int s = 0; s < n_passive; s++)
    boundary_axis(b->passive[s], mz, my, mx, axis, n[axis], g[axis], 0, side,
                  type, 1.0);
  double *face[3] = { b->bfieldi_x, b->bfieldi_y, b->bfieldi_z };
  for (int f = 0; f < 3; f++)
    if (face[f])
      boundary_axis(face[f], mz + (f == 2), my + (f == 1), mx + (f == 0), axis,
                    n[axis], g[axis], f == axis, side, type,
                    f == axis ? -1.0 : 1.0);
  return 0;
}

/* BoundaryValue::enforce (Cello/problem_BoundaryValue.cpp:131-273) for constant
 * value-expressions: the g outermost layers of one side take the value; the
 * other axes run over their full extent. next = extent along `axis`. */
static void inflow_axis(double *p, int n0, int n1, int n2, int axis, int g,
                        int side, double value)
{
  const int ext[3] = { n2, n1, n0 };
  int lo[3] = { 0, 0, 0 }, hi[3] = { n2, n1, n0 };
  if (side == 0) hi[axis] = g; else lo[axis] = ext[axis] - g;
  for (int k = lo[2]; k < hi[2]; k++)
    for (int j = lo[1]; j < hi[1]; j++)
      for (int i = lo[0]; i < hi[0]; i++)
        p[((size_t) k * n1 + j) * n2 + i] = value;
}

/* One "inflow" Boundary object applied to one face of the domain: the fields
 * with a non-NULL pointer in b and a non-NaN entry in v form its field list. */
int vlct_oracle_boundary_inflow(const vlct_block *b, int n_passive, int axis,
                                int side, const vlct_inflow_values *v)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int g[3] = { b->gx, b->gy, b->gz };
  if (axis < 0 || axis > 2 || (side != 0 && side != 1) || v == NULL) return 1;
  struct { double *p; double value; int face; } item[] = {
    { b->density, v->density, -1 },
    { b->velocity_x, v->velocity_x, -1 }, { b->velocity_y, v->velocity_y, -1 },
    { b->velocity_z, v->velocity_z, -1 },
    { b->total_energy, v->total_energy, -1 },
    { b->internal_energy, v->internal_energy, -1 },
    { b->bfield_x, v->bfield_x, -1 }, { b->bfield_y, v->bfield_y, -1 },
    { b->bfield_z, v->bfield_z, -1 },
    { b->bfieldi_x, v->bfieldi_x, 0 }, { b->bfieldi_y, v->bfieldi_y, 1 },
    { b->bfieldi_z, v->bfieldi_z, 2 },
    { b->pressure, v->pressure, -1 } };
  for (size_t c = 0; c < sizeof(item) / sizeof(item[0]); c++) {
    if (item[c].p == NULL || isnan(item[c].value)) continue;
    const int f = item[c].face;
    inflow_axis(item[c].p, mz + (f == 2), my + (f == 1), mx + (f == 0), axis,
                g[axis], side, item[c].value);
  }
  for (int s = 0; s < n_passive; s++)
    if (b->passive[s] && !isnan(v->passive[s]))
      inflow_axis(b->passive[s], mz, my, mx, axis, g[axis], side, v->passive[s]);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* spherical cloud in a wind                                                 */
/* ------------------------------------------------------------------------ */

/* SphereRegion::check_point (initial/EnzoInitialCloud.cpp:194-200) */
static int cloud_check_point(const double *center, double sqr_radius, double x,
                             double y, double z)
{
  double dx = x - center[0];
  double dy = y - center[1];
  double dz = z - center[2];
  return (dx * dx + dy * dy + dz * dz) <= sqr_radius;
}

/* ---- the optional density perturbation (cpp:16-163) ------------------------- */
/* std::minstd_rand: x <- 48271 x mod (2^31 - 1), a zero seed becomes 1 */
typedef struct { unsigned long long x; } minstd;
#define MINSTD_MAX 2147483646.0
static void minstd_seed(minstd *g, unsigned int seed)
{ g->x = seed % 2147483647ULL; if (g->x == 0) g->x = 1; }
static double minstd_next(minstd *g)
{ g->x = (g->x * 48271ULL) % 2147483647ULL; return (double) g->x; }

/* uniform_dist_transform_ (cpp:16-44) */
static double cloud_uniform(minstd *g, int include_zero, int include_one)
{
  double raw = minstd_next(g);
  double range;
  if (include_zero && include_one) { range = MINSTD_MAX - 1.; raw--; }
  else if (include_zero)           { range = MINSTD_MAX;      raw--; }
  else if (include_one)            { range = MINSTD_MAX; }
  else                             { range = MINSTD_MAX + 1.; }
  return raw / range;
}

/* normal_dist_transform_ (Box-Muller, cpp:51-59) */
static void cloud_normal_pair(minstd *g, double *a, double *b)
{
  double x1 = cloud_uniform(g, 0, 0);
  double x2 = cloud_uniform(g, 0, 0);
  double coef = sqrt(-2. * log(x1));
  *a = coef * cos(2. * cello_pi * x2);
  *b = coef * sin(2. * cello_pi * x2);
}

#define CLOUD_MAX_WAVES 1024
typedef struct {
  int nwaves;
  double amplitude;
  double kx[CLOUD_MAX_WAVES], ky[CLOUD_MAX_WAVES], kz[CLOUD_MAX_WAVES],
         phi[CLOUD_MAX_WAVES];
} cloud_waves;

/* WavePerturbation::WavePerturbation (cpp:91-119) */
static void cloud_waves_init(cloud_waves *w, int nwaves, unsigned int seed,
                             double amplitude, double min_lambda, double max_lambda)
{
  w->nwaves = nwaves;
  w->amplitude = amplitude;
  if (!(nwaves > 0 && amplitude > 0.)) { w->nwaves = (nwaves > 0) ? nwaves : 0; }
  if (nwaves > 0 && amplitude > 0.) {
    minstd g;
    minstd_seed(&g, seed);
    for (int i = 0; i < nwaves; i++) {
      double lambda = min_lambda + (max_lambda - min_lambda) * cloud_uniform(&g, 1, 1);
      /* sample_sphere_points_ (cpp:61-84) */
      double x, y, z, unused;
      for (;;) {
        cloud_normal_pair(&g, &x, &y);
        cloud_normal_pair(&g, &z, &unused);
        if ((x != 0.) || (y != 0.) || (z != 0.)) break;
      }
      double magnitude = sqrt(x * x + y * y + z * z);
      w->kx[i] = (x / magnitude) * 2 * cello_pi / lambda;
      w->ky[i] = (y / magnitude) * 2 * cello_pi / lambda;
      w->kz[i] = (z / magnitude) * 2 * cello_pi / lambda;
      w->phi[i] = cello_pi * cloud_uniform(&g, 1, 0);
    }
  } else {
    w->nwaves = 0;     /* operator() then sums nothing: exactly 0 */
  }
}

/* WavePerturbation::operator() (cpp:135-160): volume average over a cell */
static double cloud_waves_eval(const cloud_waves *w, double xc, double yc, double zc,
                               double hx, double hy, double hz)
{
  double total = 0.;
  double alpha = 8. * w->amplitude / (hx * hy * hz);
  for (int i = 0; i < w->nwaves; i++) {
    double ci = (sin(w->kx[i] * hx * 0.5) * sin(w->ky[i] * hy * 0.5) *
                 sin(w->kz[i] * hz * 0.5)) / (w->kx[i] * w->ky[i] * w->kz[i]);
    total += ci * cos(w->kx[i] * xc + w->ky[i] * yc + w->kz[i] * zc + w->phi[i]);
  }
  return alpha * total;
}

/* CloudInitHelper::query_cell (cpp:327-390): fraction of the cell [left, right]
 * enclosed by the sphere and the average density perturbation of that part */
static double cloud_frac_enclosed(const double *center, double sqr_radius,
                                  const double *left, const double *right,
                                  int nsub, double num_subsampled_cells,
                                  const double *off[3], const cloud_waves *w,
                                  const double *h, double *perturbation)
{
  /* SphereRegion::check_intersect (cpp:204-237) */
  double nearest[3], furthest[3];
  for (int i = 0; i < 3; i++) {
    if (center[i] <= left[i]) {
      nearest[i] = left[i]; furthest[i] = right[i];
    } else if (center[i] >= right[i]) {
      nearest[i] = right[i]; furthest[i] = left[i];
    } else {
      nearest[i] = center[i];
      if ((center[i] - left[i]) > (right[i] - center[i])) furthest[i] = left[i];
      else furthest[i] = right[i];
    }
  }
  if (cloud_check_point(center, sqr_radius, furthest[0], furthest[1], furthest[2])) {
    *perturbation = cloud_waves_eval(w, 0.5 * (left[0] + right[0]),
                                     0.5 * (left[1] + right[1]),
                                     0.5 * (left[2] + right[2]), h[0], h[1], h[2]);
    return 1.0;                                   /* enclosed_cell */
  }
  *perturbation = 0.;
  if (!cloud_check_point(center, sqr_radius, nearest[0], nearest[1], nearest[2]))
    return 0.0;                                   /* no_overlap */
  int n_enclosed = 0;                             /* partial_overlap */
  double perturb_sum = 0.;
  const double n_axis = (double) nsub;
  const double sub_h[3] = { h[0] / n_axis, h[1] / n_axis, h[2] / n_axis };   /* cpp:295-296 */
  for (int sz = 0; sz < nsub; sz++) {
    double sub_zc = left[2] + off[2][sz];
    for (int sy = 0; sy < nsub; sy++) {
      double sub_yc = left[1] + off[1][sy];
      for (int sx = 0; sx < nsub; sx++) {
        double sub_xc = left[0] + off[0][sx];
        if (cloud_check_point(center, sqr_radius, sub_xc, sub_yc, sub_zc)) {
          n_enclosed++;
          perturb_sum += cloud_waves_eval(w, sub_xc, sub_yc, sub_zc, sub_h[0],
                                          sub_h[1], sub_h[2]);
        }
      }
    }
  }
  if (n_enclosed > 0) *perturbation = perturb_sum / (double) n_enclosed;
  return (double) n_enclosed / num_subsampled_cells;   /* cpp:377 */
}

/* EnzoInitialCloud::enforce_block (initial/EnzoInitialCloud.cpp:606-748), no
 * "color" fields (the input/vlct/dual_energy_cloud files define no Group:color, so
 * cloud_dye / metal_density are not touched), magnetic fields pre-initialised
 * by the caller and uniform (MHDHandler, cpp:452-523).
 * p[] = { cloud_radius, center_x, center_y, center_z, cloud_density,
 *         wind_density, wind_velocity, wind_total_energy, wind_internal_energy }
 * lower: domain coordinate of the block's first active cell. */
int vlct_ic_cloud_perturbed(const vlct_block *b, const double *lower, int subsample_n,
                            const double *p, int nwaves, unsigned int seed,
                            double amplitude, double min_lambda, double max_lambda);

/* perturb_Nwaves = 0 (the default; the perturbation is then exactly 0) */
int vlct_ic_cloud(const vlct_block *b, const double *lower, int subsample_n,
                  const double *p)
{ return vlct_ic_cloud_perturbed(b, lower, subsample_n, p, 0, 0u, 0., 0., 0.); }

/* ... with the density perturbation of Initial:cloud:perturb_* (a sum of
 * `nwaves` inclined plane waves with wavelengths in [min_lambda, max_lambda],
 * drawn from std::minstd_rand(seed); cpp:86-163, hpp:39-58) */
int vlct_ic_cloud_perturbed(const vlct_block *b, const double *lower, int subsample_n,
                            const double *p, int nwaves, unsigned int seed,
                            double amplitude, double min_lambda, double max_lambda)
{
  const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
  const int m[3] = { mx, my, mz }, g[3] = { b->gx, b->gy, b->gz };
  const double h[3] = { b->dx, b->dy, b->dz };
  const double center[3] = { p[1], p[2], p[3] };
  const double sqr_radius = p[0] * p[0];
  const double density_cloud = p[4], density_wind = p[5], velocity_wind = p[6],
               etot_wind = p[7], eint_wind = p[8];
  if (subsample_n < 0 || p[0] <= 0.) return 1;
  if (nwaves > CLOUD_MAX_WAVES) return 3;
  if (nwaves > 0 && !(amplitude > 0. && min_lambda > 0. && max_lambda >= min_lambda))
    return 4;                                        /* hpp:95-106 */
  cloud_waves *waves = (cloud_waves *) malloc(sizeof(cloud_waves));
  cloud_waves_init(waves, nwaves, seed, amplitude, min_lambda, max_lambda);

  /* Data::field_cell_faces with cx = cy = cz = 1 (Cello/data_Data.cpp:91-121) */
  double *xf[3];
  for (int a = 0; a < 3; a++) {
    xf[a] = (double *) malloc(sizeof(double) * (size_t) (m[a] + 1));
    for (int i = -g[a]; i < m[a] - g[a] + 1; i++)
      xf[a][i + g[a]] = lower[a] + (i + 0.) * h[a];
  }
  /* prep_subcell_offsets_ (cpp:246-256) */
  const int nsub = (int) pow(2, subsample_n);
  double *off[3];
  for (int a = 0; a < 3; a++) {
    off[a] = (double *) malloc(sizeof(double) * (size_t) nsub);
    double cur_frac = 1. / pow(2, subsample_n + 1);
    off[a][0] = cur_frac * h[a];
    for (int i = 1; i < nsub; i++) {
      cur_frac += 1. / pow(2, subsample_n);
      off[a][i] = cur_frac * h[a];
    }
  }
  const double *coff[3] = { off[0], off[1], off[2] };
  const double num_subsampled_cells = pow(pow(2, subsample_n), 3);   /* cpp:269 */

  /* MHDHandler: B is assumed uniform over the active zone (cpp:498-511) */
  const int mhd = (b->bfield_x != NULL);
  double magnetic_edens_wind = 0.;
  if (mhd) {
    const size_t c = ((size_t) b->gz * my + b->gy) * mx + b->gx;
    magnetic_edens_wind = 0.5 * (b->bfield_x[c] * b->bfield_x[c] +
                                 b->bfield_y[c] * b->bfield_y[c] +
                                 b->bfield_z[c] * b->bfield_z[c]);
  }
  const int dual_energy = (b->internal_energy != NULL);
  double eint_density;
  if (dual_energy) {
    eint_density = eint_wind * density_wind;
  } else {
    eint_density = ((etot_wind - 0.5 * velocity_wind * velocity_wind)
                    * density_wind - magnetic_edens_wind);
  }
  if (!(eint_density > 0)) return 2;

  for (int iz = 0; iz < mz; iz++)
    for (int iy = 0; iy < my; iy++)
      for (int ix = 0; ix < mx; ix++) {
        const size_t c = ((size_t) iz * my + iy) * mx + ix;
        b->velocity_y[c] = 0.;
        b->velocity_z[c] = 0.;
        const double left[3] = { xf[0][ix], xf[1][iy], xf[2][iz] };
        const double right[3] = { xf[0][ix + 1], xf[1][iy + 1], xf[2][iz + 1] };
        double perturbation = 0.;
        double frac_enclosed = cloud_frac_enclosed(center, sqr_radius, left, right,
                                                   nsub, num_subsampled_cells, coff,
                                                   waves, h, &perturbation);
        perturbation += 1.;
        double avg_density = (frac_enclosed * density_cloud * perturbation +
                              (1. - frac_enclosed) * density_wind);
        b->density[c] = avg_density;
        double wind_to_average_ratio = density_wind / avg_density;
        double wind_mass_weight = (1. - frac_enclosed) * wind_to_average_ratio;
        b->velocity_x[c] = (wind_mass_weight * velocity_wind);
        if (dual_energy) b->internal_energy[c] = eint_wind * wind_to_average_ratio;
        if (frac_enclosed == 0) {
          b->total_energy[c] = etot_wind;
        } else {
          double magnetic_edens = 0.;
          if (mhd)
            magnetic_edens = 0.5 * (b->bfield_x[c] * b->bfield_x[c] +
                                    b->bfield_y[c] * b->bfield_y[c] +
                                    b->bfield_z[c] * b->bfield_z[c]);
          b->total_energy[c] = ((eint_density + magnetic_edens) / avg_density +
                                0.5 * b->velocity_x[c] * b->velocity_x[c]);
        }
      }
  for (int a = 0; a < 3; a++) { free(xf[a]); free(off[a]); }
  free(waves);
  return 0;
}
