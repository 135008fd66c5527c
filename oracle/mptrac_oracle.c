/*
 * mptrac_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mptrac_oracle.h).
 *
 * CPU restatement of the time-step path of slcs-jsc/mptrac.  Every function names the reference
 * lines (src/mptrac.c unless noted) whose behaviour it restates.  It is organised like the
 * reference -- one pass over all parcels per module, a materialised random-number array, a dt
 * array -- which is deliberately NOT how the CUDA engine is organised (one fused kernel, in-register
 * counters), so the two are independent statements of the same arithmetic.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared  (no FMA contraction: the reference's
 * x86-64 build has none either).
 */
#include "mptrac_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* constants, src/mptrac.h:265-340 */
#define C_G0 9.80665
#define C_H0 7.0
#define C_KB 1.3806504e-23
#define C_MA 28.9644
#define C_P0 1013.25
#define C_RI 8.3144598
#define C_RA (1e3 * C_RI / C_MA)
#define C_RE 6367.421
#define C_MAIR 4.8096e-26

/* ---------------------------------------------------------------------------------------------
 * scalar helpers
 * ------------------------------------------------------------------------------------------- */

/* FMOD macro, src/mptrac.h:1121: quotient truncated through an int cast */
static double tmod(double x, double y) { return x - (int)(x / y) * y; }

/* LIN macro, src/mptrac.h:1351 */
static double linear(double x0, double y0, double x1, double y1, double x) {
  return y0 + (y1 - y0) / (x1 - x0) * (x - x0);
}

/* DZ2DP, src/mptrac.h:941 */
static double km2hpa(double dz, double p) { return -dz * p / C_H0; }

/* DX2COORD / DX2DEG, src/mptrac.h:904-906, 966 */
static double east_m_to_coord(int coord_type, double dx, double lat) {
  if (coord_type != 0) return dx;
  const double km = dx / 1000.0;
  if (lat < -89.999 || lat > 89.999) return 0;
  return km * 180. / (M_PI * C_RE * cos(lat * (M_PI / 180.0)));
}

/* DY2COORD / DY2DEG, src/mptrac.h:922-923, 989 */
static double north_m_to_coord(int coord_type, double dy) {
  if (coord_type != 0) return dy;
  return (dy / 1000.0) * 180. / (M_PI * C_RE);
}

/* locate_irr, 3495-3521 */
static int bisect(const double *xx, int n, double x) {
  int lo = 0, hi = n - 1;
  const int m = (hi + lo) >> 1;
  if (xx[m] < xx[m + 1]) {
    while (hi > lo + 1) {
      const int i = (hi + lo) >> 1;
      if (xx[i] > x) hi = i; else lo = i;
    }
  } else {
    while (hi > lo + 1) {
      const int i = (hi + lo) >> 1;
      if (xx[i] <= x) hi = i; else lo = i;
    }
  }
  return lo;
}

/* locate_reg, 3559-3574 */
static int regular_index(const double *xx, int n, double x) {
  const int i = (int)((x - xx[0]) / (xx[1] - xx[0]));
  if (i < 0) return 0;
  if (i > n - 2) return n - 2;
  return i;
}

static double clampd(double x, double a, double b) {  /* MIN(MAX(x, lo), hi) with lo = min(a, b) */
  const double lo = a < b ? a : b, hi = a < b ? b : a;
  double r = x > lo ? x : lo;
  return r < hi ? r : hi;
}

/* intpol_check_lon_lat / intpol_check_cartesian, 2755-2803 */
static void horizontal_check(const orc_met_t *m, double lon, double lat, double *lon2, double *lat2) {
  if (m->coord_type == 0) {
    double l = tmod(lon, 360.);
    if (l < m->lon[0]) l += 360;
    else if (l > m->lon[m->nx - 1]) l -= 360;
    *lon2 = l;
  } else {
    *lon2 = clampd(lon, m->lon[0], m->lon[m->nx - 1]);
  }
  *lat2 = clampd(lat, m->lat[0], m->lat[m->ny - 1]);
}

/* ---------------------------------------------------------------------------------------------
 * interpolation; the cache (ci, cw) of the reference becomes an explicit struct
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int ip, ix, iy;        /* ci[0], ci[1], ci[2] */
  double wp, wx, wy;     /* cw[0], cw[1], cw[2]: weight of the lower-index node */
} cell_t;

static const cell_t CELL_ZERO = {0, 0, 0, 0.0, 0.0, 0.0};  /* INTPOL_INIT, src/mptrac.h:1174 */

static size_t at3(const orc_met_t *m, int ix, int iy, int ip) {
  return ((size_t)ix * (size_t)m->ny + (size_t)iy) * (size_t)m->np + (size_t)ip;
}
static size_t at2(const orc_met_t *m, int ix, int iy) { return (size_t)ix * (size_t)m->ny + (size_t)iy; }

/* the init part of intpol_met_space_3d / _2d, 2997-3021 and 3061-3078 */
static void locate_cell(const orc_met_t *m, int with_p, double p, double lon, double lat, cell_t *c) {
  double lon2, lat2;
  horizontal_check(m, lon, lat, &lon2, &lat2);
  if (with_p) {
    c->ip = bisect(m->p, m->np, p);
    c->wp = (m->p[c->ip + 1] - p) / (m->p[c->ip + 1] - m->p[c->ip]);
  }
  c->ix = regular_index(m->lon, m->nx, lon2);
  c->iy = bisect(m->lat, m->ny, lat2);
  c->wx = (m->lon[c->ix + 1] - lon2) / (m->lon[c->ix + 1] - m->lon[c->ix]);
  c->wy = (m->lat[c->iy + 1] - lat2) / (m->lat[c->iy + 1] - m->lat[c->iy]);
}

/* vertical leg: the difference is taken in float before promotion, 3023-3038 */
static double column_value(const orc_met_t *m, const float *f, const cell_t *c, int dx, int dy) {
  const float lo = f[at3(m, c->ix + dx, c->iy + dy, c->ip)];
  const float hi = f[at3(m, c->ix + dx, c->iy + dy, c->ip + 1)];
  return c->wp * (lo - hi) + hi;
}

/* intpol_met_space_3d body, 3023-3043 */
static double space3(const orc_met_t *m, const float *f, const cell_t *c) {
  const double v00 = column_value(m, f, c, 0, 0), v01 = column_value(m, f, c, 0, 1);
  const double v10 = column_value(m, f, c, 1, 0), v11 = column_value(m, f, c, 1, 1);
  const double a0 = c->wy * (v00 - v01) + v01;
  const double a1 = c->wy * (v10 - v11) + v11;
  return c->wx * (a0 - a1) + a1;
}

/* intpol_met_space_2d body, 3080-3107 */
static double space2(const orc_met_t *m, const float *f, const cell_t *c) {
  const double v00 = f[at2(m, c->ix, c->iy)], v01 = f[at2(m, c->ix, c->iy + 1)];
  const double v10 = f[at2(m, c->ix + 1, c->iy)], v11 = f[at2(m, c->ix + 1, c->iy + 1)];
  if (isfinite(v00) && isfinite(v01) && isfinite(v10) && isfinite(v11)) {
    const double a0 = c->wy * (v00 - v01) + v01;
    const double a1 = c->wy * (v10 - v11) + v11;
    return c->wx * (a0 - a1) + a1;
  }
  if (c->wy < 0.5) return c->wx < 0.5 ? v11 : v01;
  return c->wx < 0.5 ? v10 : v00;
}

/* intpol_met_time_3d, 3112-3137 (init: locate on met0, reuse for met1) */
static double time3(const orc_met_t *m0, const float *f0, const orc_met_t *m1, const float *f1, double ts,
                    double p, double lon, double lat, cell_t *c, int init) {
  if (init) locate_cell(m0, 1, p, lon, lat, c);
  const double a = space3(m0, f0, c), b = space3(m1, f1, c);
  const double wt = (m1->time - ts) / (m1->time - m0->time);
  return wt * (a - b) + b;
}

/* intpol_met_time_2d, 3141-3170 */
static double time2(const orc_met_t *m0, const float *f0, const orc_met_t *m1, const float *f1, double ts,
                    double lon, double lat, cell_t *c, int init) {
  if (init) locate_cell(m0, 0, 0.0, lon, lat, c);
  const double a = space2(m0, f0, c), b = space2(m1, f1, c);
  const double wt = (m1->time - ts) / (m1->time - m0->time);
  if (isfinite(a) && isfinite(b)) return wt * (a - b) + b;
  return wt < 0.5 ? b : a;
}

void orc_intpol_met_time_3d(const orc_met_t *met0, const orc_met_t *met1, const float *f0, const float *f1,
                            double ts, double p, double lon, double lat, double *var) {
  cell_t c = CELL_ZERO;
  *var = time3(met0, f0, met1, f1, ts, p, lon, lat, &c, 1);
}

/* ---------------------------------------------------------------------------------------------
 * module_timesteps, 5999-6042
 * ------------------------------------------------------------------------------------------- */
void orc_module_timesteps(const orc_ctl_t *ctl, const orc_met_t *met0, orc_atm_t *atm, double t) {
  double latmin = met0->lat[0], latmax = met0->lat[0];
  for (int i = 1; i < met0->ny; i++) {
    if (met0->lat[i] < latmin) latmin = met0->lat[i];
    if (met0->lat[i] > latmax) latmax = met0->lat[i];
  }
  const int local = fabs(met0->lon[met0->nx - 1] - met0->lon[0] - 360.0) >= 0.01;
  const double lon_w = met0->lon[0], lon_e = met0->lon[met0->nx - 1];
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double tp = atm->time[ip];
    const int active = ctl->direction * (tp - ctl->t_start) >= 0 && ctl->direction * (tp - ctl->t_stop) <= 0 &&
                       ctl->direction * (tp - t) < 0;
    double dt = active ? t - tp : 0.0;
    if (local && (atm->lon[ip] <= lon_w || atm->lon[ip] >= lon_e || atm->lat[ip] <= latmin || atm->lat[ip] >= latmax))
      dt = 0.0;
    atm->dt[ip] = dt;
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_position, 5435-5489.  The surface-pressure lookup at 5483 runs with init = 0 on a freshly
 * zeroed cache, i.e. on cell (0,0) with zero weights: it yields node [1][1].  Restated as written.
 * ------------------------------------------------------------------------------------------- */
void orc_module_position(const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  const double ptop = met0->p[met0->np - 1];
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (atm->dt[ip] == 0) continue;
    double lon = atm->lon[ip], lat = atm->lat[ip], p = atm->p[ip];
    if (met0->coord_type == 0) {
      lon = tmod(lon, 360.);
      lat = tmod(lat, 360.);
      while (lat < -90 || lat > 90) {
        if (lat > 90) { lat = 180 - lat; lon += 180; }
        if (lat < -90) { lat = -180 - lat; lon += 180; }
      }
      while (lon < -180) lon += 360;
      while (lon >= 180) lon -= 360;
    } else {
      const double x = clampd(lon, met0->lon[0], met0->lon[met0->nx - 1]);
      const double y = clampd(lat, met0->lat[0], met0->lat[met0->ny - 1]);
      lon = x; lat = y;
    }
    if (p < ptop) {
      p = ptop * ptop / p;
    } else if (p > 300.) {
      cell_t c = CELL_ZERO;
      const double ps = time2(met0, met0->ps, met1, met1->ps, atm->time[ip], lon, lat, &c, 0);
      if (p > ps) p = ps * ps / p;
    }
    atm->lon[ip] = lon; atm->lat[ip] = lat; atm->p[ip] = p;
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_advect, pressure-level branch, 3612-3677
 * ------------------------------------------------------------------------------------------- */
/* ---------------------------------------------------------------------------------------------
 * interpolation on model levels: intpol_met_4d_zeta 2808-2981, locate_vert 3578-3594,
 * locate_irr_float 3525-3555.  NB the weight convention differs from the pressure-level routines
 * (weight of the UPPER-index node), and time is interpolated FIRST.
 * ------------------------------------------------------------------------------------------- */
static size_t atl(const orc_met_t *m, int ix, int iy, int k) {
  return ((size_t)ix * (size_t)m->ny + (size_t)iy) * (size_t)m->npl + (size_t)k;
}

static int bisect_float(const float *xx, int n, double x, int ig) {
  int lo = 0, hi = n - 1, i = (hi + lo) >> 1;
  if ((xx[ig] <= x && x < xx[ig + 1]) || (xx[ig] >= x && x > xx[ig + 1])) return ig;
  if (xx[i] < xx[i + 1]) {
    while (hi > lo + 1) { i = (hi + lo) >> 1; if (xx[i] > x) hi = i; else lo = i; }
  } else {
    while (hi > lo + 1) { i = (hi + lo) >> 1; if (xx[i] <= x) hi = i; else lo = i; }
  }
  return lo;
}

typedef struct {
  int ix, iy, iz;            /* ci[0], ci[1], ci[2] */
  double wx, wy, wz, wt;     /* cw[0], cw[1], cw[2], cw[3] */
} zcell_t;

/* time-then-bilinear value of a model-level field at level k of the column quartet (the "height_bot/top" blocks) */
static double level_value(const orc_met_t *m, const float *f0, const float *f1, const zcell_t *c, int k) {
  const double v00 = c->wt * (f1[atl(m, c->ix, c->iy, k)] - f0[atl(m, c->ix, c->iy, k)]) + f0[atl(m, c->ix, c->iy, k)];
  const double v01 = c->wt * (f1[atl(m, c->ix, c->iy + 1, k)] - f0[atl(m, c->ix, c->iy + 1, k)]) + f0[atl(m, c->ix, c->iy + 1, k)];
  const double v10 = c->wt * (f1[atl(m, c->ix + 1, c->iy, k)] - f0[atl(m, c->ix + 1, c->iy, k)]) + f0[atl(m, c->ix + 1, c->iy, k)];
  const double v11 = c->wt * (f1[atl(m, c->ix + 1, c->iy + 1, k)] - f0[atl(m, c->ix + 1, c->iy + 1, k)]) + f0[atl(m, c->ix + 1, c->iy + 1, k)];
  const double a0 = c->wy * (v01 - v00) + v00;
  const double a1 = c->wy * (v11 - v10) + v10;
  return c->wx * (a1 - a0) + a0;
}

/* the init part: 2825-2938 */
static void zeta_locate(const orc_met_t *m0, const float *h0, const orc_met_t *m1, const float *h1, double ts, double height,
                        double lon, double lat, zcell_t *c) {
  double lon2, lat2;
  horizontal_check(m0, lon, lat, &lon2, &lat2);
  c->ix = regular_index(m0->lon, m0->nx, lon2);
  c->iy = bisect(m0->lat, m0->ny, lat2);
  int ind[2][4];
  const float *hh[2] = {h0, h1};
  const int npl[2] = {m0->npl, m1->npl};
  for (int t = 0; t < 2; t++) {   /* locate_vert: each column starts from the previous column's answer */
    ind[t][0] = bisect_float(hh[t] + atl(m0, c->ix, c->iy, 0), npl[t], height, 0);
    ind[t][1] = bisect_float(hh[t] + atl(m0, c->ix + 1, c->iy, 0), npl[t], height, ind[t][0]);
    ind[t][2] = bisect_float(hh[t] + atl(m0, c->ix, c->iy + 1, 0), npl[t], height, ind[t][1]);
    ind[t][3] = bisect_float(hh[t] + atl(m0, c->ix + 1, c->iy + 1, 0), npl[t], height, ind[t][2]);
  }
  c->iz = ind[0][0];
  int k_max = ind[0][0];
  for (int t = 0; t < 2; t++)
    for (int j = 0; j < 4; j++) {
      if (c->iz > ind[t][j]) c->iz = ind[t][j];
      if (k_max < ind[t][j]) k_max = ind[t][j];
    }
  c->wt = (ts - m0->time) / (m1->time - m0->time);
  c->wx = (lon2 - m0->lon[c->ix]) / (m0->lon[c->ix + 1] - m0->lon[c->ix]);
  c->wy = (lat2 - m0->lat[c->iy]) / (m0->lat[c->iy + 1] - m0->lat[c->iy]);
  double bot = level_value(m0, h0, h1, c, c->iz), top = level_value(m0, h0, h1, c, c->iz + 1);
  const int descending = h0[0] > h0[1], ascending = h0[0] < h0[1];   /* heights0[0][0][0] vs [0][0][1], 2914-2921 */
  while ((descending && ((bot <= height) || (top > height)) && (bot >= height) && (c->iz < k_max)) ||
         (ascending && ((bot >= height) || (top < height)) && (bot <= height) && (c->iz < k_max))) {
    c->iz++;
    bot = top;
    top = level_value(m0, h0, h1, c, c->iz + 1);
  }
  c->wz = (height - bot) / (top - bot);
}

/* the value part: 2941-2980 (time first, then lon, lat, level) */
static double zeta_value(const orc_met_t *m, const float *a0, const float *a1, const zcell_t *c) {
#define TV(dx, dy, dk) (c->wt * (a1[atl(m, c->ix + dx, c->iy + dy, c->iz + dk)] - a0[atl(m, c->ix + dx, c->iy + dy, c->iz + dk)]) \
                        + a0[atl(m, c->ix + dx, c->iy + dy, c->iz + dk)])
  const double v000 = TV(0, 0, 0), v100 = TV(1, 0, 0), v010 = TV(0, 1, 0), v110 = TV(1, 1, 0);
  const double v001 = TV(0, 0, 1), v101 = TV(1, 0, 1), v011 = TV(0, 1, 1), v111 = TV(1, 1, 1);
#undef TV
  const double b00 = c->wx * (v100 - v000) + v000, b10 = c->wx * (v110 - v010) + v010;
  const double b01 = c->wx * (v101 - v001) + v001, b11 = c->wx * (v111 - v011) + v011;
  const double e0 = c->wy * (b10 - b00) + b00, e1 = c->wy * (b11 - b01) + b01;
  return c->wz * (e1 - e0) + e0;
}

void orc_intpol_met_4d_zeta(const orc_met_t *met0, const float *h0, const float *a0, const orc_met_t *met1, const float *h1,
                            const float *a1, double ts, double height, double lon, double lat, double *var) {
  zcell_t c;
  zeta_locate(met0, h0, met1, h1, ts, height, lon, lat, &c);
  *var = zeta_value(met0, a0, a1, &c);
}

/* module_advect_init, 3762-3785: pressure consistent with zeta (ADVECT_VERT_COORD 1 only, every parcel) */
void orc_module_advect_init(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  if (ctl->advect_vert_coord != 1) return;
  double *zeta = atm->q + (size_t)ctl->qnt_zeta * (size_t)atm->q_stride;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++)
    orc_intpol_met_4d_zeta(met0, met0->zetal, met0->pl, met1, met1->zetal, met1->pl, atm->time[ip], zeta[ip], atm->lon[ip],
                           atm->lat[ip], &atm->p[ip]);
}

/* module_advect with the vertical velocity of a model-level coordinate: branch A with ADVECT_VERT_COORD 2 (3646-3657: omega on
 * model levels, heights = pl) and branch B (3680-3757: zeta / eta, the parcel's coordinate lives in a quantity) */
static void advect_model_levels(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  const int n = ctl->advect, ct = met0->coord_type, vc = ctl->advect_vert_coord;
  const float *h0 = vc == 2 ? met0->pl : met0->zetal, *h1 = vc == 2 ? met1->pl : met1->zetal;
  const float *w0 = vc == 2 ? met0->wl : met0->zeta_dotl, *w1 = vc == 2 ? met1->wl : met1->zeta_dotl;
  double *zq = vc == 2 ? NULL : atm->q + (size_t)(vc == 1 ? ctl->qnt_zeta : ctl->qnt_eta) * (size_t)atm->q_stride;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double dt = atm->dt[ip];
    if (dt == 0) continue;
    const double lon0 = atm->lon[ip], lat0 = atm->lat[ip], t0 = atm->time[ip];
    if (zq)   /* pressure -> zeta / eta, 3690-3695 */
      orc_intpol_met_4d_zeta(met0, met0->pl, met0->zetal, met1, met1->pl, met1->zetal, t0, atm->p[ip], lon0, lat0, &zq[ip]);
    const double z0 = zq ? zq[ip] : atm->p[ip];
    double u[4], v[4], w[4], um = 0, vm = 0, wm = 0, pos[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) {
      double dts = 0.0;
      if (i == 0) {
        pos[0] = lon0; pos[1] = lat0; pos[2] = z0;
      } else {
        dts = (i == 3 ? 1.0 : 0.5) * dt;
        pos[0] = lon0 + east_m_to_coord(ct, dts * u[i - 1], lat0);
        pos[1] = lat0 + north_m_to_coord(ct, dts * v[i - 1]);
        pos[2] = z0 + dts * w[i - 1];
      }
      zcell_t c;
      zeta_locate(met0, h0, met1, h1, t0 + dts, pos[2], pos[0], pos[1], &c);
      u[i] = zeta_value(met0, met0->ul, met1->ul, &c);
      v[i] = zeta_value(met0, met0->vl, met1->vl, &c);
      w[i] = zeta_value(met0, w0, w1, &c);
      double k = 1.0;
      if (n == 2) k = (i == 0 ? 0.0 : 1.0);
      else if (n == 4) k = (i == 0 || i == 3 ? 1.0 / 6.0 : 2.0 / 6.0);
      um += k * u[i]; vm += k * v[i]; wm += k * w[i];
    }
    atm->time[ip] = t0 + dt;
    atm->lon[ip] = lon0 + east_m_to_coord(ct, dt * um, n == 2 ? pos[1] : lat0);
    atm->lat[ip] = lat0 + north_m_to_coord(ct, dt * vm);
    if (zq) {   /* 3746-3755 */
      zq[ip] = z0 + dt * wm;
      orc_intpol_met_4d_zeta(met0, met0->zetal, met0->pl, met1, met1->zetal, met1->pl, atm->time[ip], zq[ip], atm->lon[ip],
                             atm->lat[ip], &atm->p[ip]);
    } else {
      atm->p[ip] = z0 + dt * wm;
    }
  }
}

void orc_module_advect(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  if (ctl->advect_vert_coord != 0) { advect_model_levels(ctl, met0, met1, atm); return; }
  const int n = ctl->advect, ct = met0->coord_type;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double dt = atm->dt[ip];
    if (dt == 0) continue;
    const double lon0 = atm->lon[ip], lat0 = atm->lat[ip], p0 = atm->p[ip], t0 = atm->time[ip];
    double u[4], v[4], w[4], um = 0, vm = 0, wm = 0, pos[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) {
      double dts = 0.0;
      if (i == 0) {
        pos[0] = lon0; pos[1] = lat0; pos[2] = p0;
      } else {
        dts = (i == 3 ? 1.0 : 0.5) * dt;
        pos[0] = lon0 + east_m_to_coord(ct, dts * u[i - 1], lat0);
        pos[1] = lat0 + north_m_to_coord(ct, dts * v[i - 1]);
        pos[2] = p0 + dts * w[i - 1];
      }
      const double tm = t0 + dts;
      cell_t c = CELL_ZERO;
      u[i] = time3(met0, met0->u, met1, met1->u, tm, pos[2], pos[0], pos[1], &c, 1);
      v[i] = time3(met0, met0->v, met1, met1->v, tm, pos[2], pos[0], pos[1], &c, 0);
      w[i] = time3(met0, met0->w, met1, met1->w, tm, pos[2], pos[0], pos[1], &c, 0);
      double k = 1.0;
      if (n == 2) k = (i == 0 ? 0.0 : 1.0);
      else if (n == 4) k = (i == 0 || i == 3 ? 1.0 / 6.0 : 2.0 / 6.0);
      um += k * u[i]; vm += k * v[i]; wm += k * w[i];
    }
    atm->time[ip] = t0 + dt;
    atm->lon[ip] = lon0 + east_m_to_coord(ct, dt * um, n == 2 ? pos[1] : lat0);
    atm->lat[ip] = lat0 + north_m_to_coord(ct, dt * vm);
    atm->p[ip] = p0 + dt * wm;
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_rng, Squares branch, 5784-5828 (Widynski's squares64 with the reference's fixed key)
 * ------------------------------------------------------------------------------------------- */
static uint64_t rot32(uint64_t x) { return (x >> 32) | (x << 32); }

static double squares_u01(uint64_t counter) {
  const uint64_t key = 0xc8e4fd154ce32f6dULL;
  uint64_t y = counter * key, z = y + key, x = y, t;
  x = rot32(x * x + y);
  x = rot32(x * x + z);
  x = rot32(x * x + y);
  t = x = x * x + z;
  x = rot32(x);
  return (double)(t ^ ((x * x + y) >> 32)) / (double)UINT64_MAX;
}

void orc_module_rng(double *rs, int64_t n, int method, uint64_t *ctr) {
#pragma omp parallel for
  for (int64_t i = 0; i < n + 1; i++) rs[i] = squares_u01(*ctr + (uint64_t)i);
  *ctr += (uint64_t)n + 1;
  if (method == 1) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i += 2) {
      const double r = sqrt(-2.0 * log(rs[i]));
      const double phi = 2.0 * M_PI * rs[i + 1];
      rs[i] = r * cosf((float)phi);
      rs[i + 1] = r * sinf((float)phi);
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * clim_tropo 213-237, pbl_weight 8358-8376, tropo_weight 12748-12770
 * ------------------------------------------------------------------------------------------- */
double orc_clim_tropo(const orc_clim_t *cl, double t, double lat) {
  const double year = 365.25 * 86400.;
  double sec = tmod(t, year);
  while (sec < 0) sec += year;
  const int it = bisect(cl->time, cl->ntime, sec);
  const int il = regular_index(cl->lat, cl->nlat, lat);
  const double *r0 = cl->tropo + (size_t)it * cl->nlat, *r1 = r0 + cl->nlat;
  const double a = linear(cl->lat[il], r0[il], cl->lat[il + 1], r0[il + 1], lat);
  const double b = linear(cl->lat[il], r1[il], cl->lat[il + 1], r1[il + 1], lat);
  return linear(cl->time[it], a, cl->time[it + 1], b, sec);
}

static double ramp(double p_one, double p_zero, double p) {
  if (p > p_one) return 1;
  if (p < p_zero) return 0;
  return linear(p_one, 1.0, p_zero, 0.0, p);
}

static double w_pbl(const orc_ctl_t *ctl, double p, double pbl, double ps) {
  return ramp(pbl, pbl - ctl->turb_pbl_trans * (ps - pbl), p);
}

static double w_tropo(const orc_ctl_t *ctl, const orc_clim_t *cl, double time, double lat, double p) {
  const double pt = orc_clim_tropo(cl, time, ctl->met_coord_type == 0 ? lat : ctl->met_utm_ref_lat);
  return ramp(pt / 0.866877899, pt * 0.866877899, p);
}

/* ---------------------------------------------------------------------------------------------
 * module_diff_turb, 4588-4734
 * ------------------------------------------------------------------------------------------- */
#define REF_MAX(a, b) (((a) > (b)) ? (a) : (b))
#define REF_MIN(a, b) (((a) < (b)) ? (a) : (b))
static double kz_at(const orc_ctl_t *ctl, const orc_clim_t *cl, double time, double lat, double p, double pbl,
                    double ps, double *kx) {
  const double a = w_pbl(ctl, p, pbl, ps);
  const double b = w_tropo(ctl, cl, time, lat, p) * (1.0 - a);
  const double c = 1.0 - a - b;
  if (kx) *kx = a * ctl->turb_dx_pbl + b * ctl->turb_dx_trop + c * ctl->turb_dx_strat;
  return a * ctl->turb_dz_pbl + b * ctl->turb_dz_trop + c * ctl->turb_dz_strat;
}

void orc_module_diff_turb(const orc_ctl_t *ctl, const orc_clim_t *clim, const orc_met_t *met0,
                          const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr) {
  orc_module_rng(atm->rs, 3 * atm->np, 1, ctr);
  const double ptop = met0->p[met0->np - 1];
  const int ct = met0->coord_type;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double dt = atm->dt[ip];
    if (dt == 0) continue;
    cell_t c = CELL_ZERO;
    const double pbl = time2(met0, met0->pbl, met1, met1->pbl, atm->time[ip], atm->lon[ip], atm->lat[ip], &c, 1);
    if (ctl->turb_pbl_scheme > 0 && atm->p[ip] >= pbl) continue;
    const double ps = time2(met0, met0->ps, met1, met1->ps, atm->time[ip], atm->lon[ip], atm->lat[ip], &c, 0);

    double Kx;
    const double Kz = kz_at(ctl, clim, atm->time[ip], atm->lat[ip], atm->p[ip], pbl, ps, &Kx);
    const double dta = fabs(dt);

    if (Kx > 0) {
      const double sh = sqrt(2.0 * Kx * dta);
      atm->lon[ip] += east_m_to_coord(ct, atm->rs[3 * ip] * sh, atm->lat[ip]);
      atm->lat[ip] += north_m_to_coord(ct, atm->rs[3 * ip + 1] * sh);
    }
    if (Kz > 0) {
      const double sz = sqrt(2.0 * Kz * dta) * 1e-3;
      const double p = atm->p[ip], eps = 0.01;
      /* MAX(ptop, MIN(ps, .)) as the reference's ternary macros (src/mptrac.h:1378, 1479): their treatment of a non-finite
         ps (a gap in the surface data) is part of the contract */
      const double pu = REF_MAX(ptop, REF_MIN(ps, p + km2hpa(eps, p)));
      const double pd = REF_MAX(ptop, REF_MIN(ps, p + km2hpa(-eps, p)));
      const double Ku = kz_at(ctl, clim, atm->time[ip], atm->lat[ip], pu, pbl, ps, NULL);
      const double Kd = kz_at(ctl, clim, atm->time[ip], atm->lat[ip], pd, pbl, ps, NULL);
      const double dKdz = (Ku - Kd) / (2.0 * eps * 1e3);
      const double drift = dKdz + Kz * (-1.0 / (1e3 * C_H0));
      const double dz = atm->rs[3 * ip + 2] * sz + drift * dta * 1e-3;
      double pt = p + km2hpa(dz, p);
      for (int it = 0; it < 10; it++) {
        if (pt > ps) pt = ps * ps / pt;
        else if (pt < ptop) pt = ptop * ptop / pt;
        else break;
      }
      atm->p[ip] = REF_MAX(ptop, REF_MIN(ps, pt));
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_diff_meso, 4266-4339: fp32 running sums over the 16 surrounding nodes in the order
 * x-offset, y-offset, z-offset, (met0, met1)
 * ------------------------------------------------------------------------------------------- */
static float sd16(float sum, float sumsq) {
  const float var = sumsq / 16.f - (sum / 16.f) * (sum / 16.f);
  return var > 0 ? sqrtf(var) : 0;
}

void orc_module_diff_meso(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1,
                          orc_atm_t *atm, uint64_t *ctr) {
  orc_module_rng(atm->rs, 3 * atm->np, 1, ctr);
  const int ct = met0->coord_type;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double dt = atm->dt[ip];
    if (dt == 0) continue;
    const int ix = regular_index(met0->lon, met0->nx, atm->lon[ip]);
    const int iy = bisect(met0->lat, met0->ny, atm->lat[ip]);
    const int iz = bisect(met0->p, met0->np, atm->p[ip]);
    float s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
    const float *f0[3] = {met0->u, met0->v, met0->w}, *f1[3] = {met1->u, met1->v, met1->w};
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++)
        for (int k = 0; k < 2; k++) {
          const size_t o = at3(met0, ix + i, iy + j, iz + k);
          for (int f = 0; f < 3; f++) { const float a = f0[f][o]; s[f] += a; s2[f] += a * a; }
          for (int f = 0; f < 3; f++) { const float a = f1[f][o]; s[f] += a; s2[f] += a * a; }
        }
    const float sig[3] = {sd16(s[0], s2[0]), sd16(s[1], s2[1]), sd16(s[2], s2[2])};
    const double r = 1 - 2 * fabs(dt) / ctl->dt_met;
    const double r2 = sqrt(1 - r * r);
    float *uvw = atm->uvwp + 3 * ip;
    if (ctl->turb_mesox > 0) {
      uvw[0] = (float)(r * uvw[0] + r2 * atm->rs[3 * ip] * ctl->turb_mesox * sig[0]);
      atm->lon[ip] += east_m_to_coord(ct, uvw[0] * dt, atm->lat[ip]);
      uvw[1] = (float)(r * uvw[1] + r2 * atm->rs[3 * ip + 1] * ctl->turb_mesox * sig[1]);
      atm->lat[ip] += north_m_to_coord(ct, uvw[1] * dt);
    }
    if (ctl->turb_mesoz > 0) {
      uvw[2] = (float)(r * uvw[2] + r2 * atm->rs[3 * ip + 2] * ctl->turb_mesoz * sig[2]);
      atm->p[ip] += uvw[2] * dt;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * sedi 12506-12535, module_sedi 5859-5883
 * ------------------------------------------------------------------------------------------- */
double orc_sedi(double p, double T, double rp, double rhop) {
  const double r = rp * 1e-6;
  const double rho = 100. * p / (C_RA * T);
  const double eta = 1.8325e-5 * (416.16 / (T + 120.)) * pow(T / 296.16, 1.5);
  const double vth = sqrt(8. * C_KB * T / (M_PI * C_MAIR));
  const double mfp = 2. * eta / (rho * vth);
  const double Kn = mfp / r;
  const double slip = 1. + Kn * (1.249 + 0.42 * exp(-0.87 / Kn));
  return 2. * (r * r) * (rhop - rho) * C_G0 / (9. * eta) * slip;
}

void orc_module_sedi(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  const double *rp = atm->q + (size_t)ctl->qnt_rp * atm->q_stride;
  const double *rhop = atm->q + (size_t)ctl->qnt_rhop * atm->q_stride;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double dt = atm->dt[ip];
    if (dt == 0) continue;
    cell_t c = CELL_ZERO;
    const double T = time3(met0, met0->t, met1, met1->t, atm->time[ip], atm->p[ip], atm->lon[ip], atm->lat[ip], &c, 1);
    const double vs = orc_sedi(atm->p[ip], T, rp[ip], rhop[ip]);
    atm->p[ip] += km2hpa(vs * dt / 1000., atm->p[ip]);
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_sort 5887-5958.  The reference orders by gsl_sort_index (heapsort, NOT stable) or Thrust
 * (stable); the order inside one cell is therefore unspecified.  This restatement uses a stable
 * order (ties by original index), which is what Thrust's radix sort gives.
 * ------------------------------------------------------------------------------------------- */
void orc_sort_keys(const orc_met_t *met0, const orc_atm_t *atm, int32_t *keys) {
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++)
    keys[ip] = (regular_index(met0->lon, met0->nx, atm->lon[ip]) * met0->ny + bisect(met0->lat, met0->ny, atm->lat[ip])) *
                   met0->np + bisect(met0->p, met0->np, atm->p[ip]);
}

typedef struct { int32_t key, idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

void orc_module_sort(const orc_ctl_t *ctl, const orc_met_t *met0, orc_atm_t *atm) {
  const int64_t n = atm->np;
  int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n ? n : 1));
  double *tmp = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1));
  orc_sort_keys(met0, atm, keys);
  for (int64_t i = 0; i < n; i++) { kv[i].key = keys[i]; kv[i].idx = (int32_t)i; }
  qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
  double *arrays[4 + 64];
  int na = 0;
  arrays[na++] = atm->time; arrays[na++] = atm->p; arrays[na++] = atm->lon; arrays[na++] = atm->lat;
  for (int iq = 0; iq < ctl->nq && iq < 64; iq++) arrays[na++] = atm->q + (size_t)iq * atm->q_stride;
  for (int a = 0; a < na; a++) {
    for (int64_t i = 0; i < n; i++) tmp[i] = arrays[a][kv[i].idx];
    memcpy(arrays[a], tmp, sizeof(double) * (size_t)n);
  }
  free(keys); free(kv); free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * module_mixing 5169-5347 (sums run serially in parcel order, like the CPU reference)
 * ------------------------------------------------------------------------------------------- */
static int box_of(double time, double lon, double lat, double p, double t0, double t1, double lon0, double lon1,
                  double lat0, double lat1, double z0, double z1, int nx, int ny, int nz) {
  const double z = C_H0 * log(C_P0 / p);
  if (time < t0 || time > t1 || lon < lon0 || lon >= lon1 || lat < lat0 || lat >= lat1 || z < z0 || z >= z1) return -1;
  const double dlon = (lon1 - lon0) / nx, dlat = (lat1 - lat0) / ny, dz = (z1 - z0) / nz;
  const int ix = (int)((lon - lon0) / dlon), iy = (int)((lat - lat0) / dlat), iz = (int)((z - z0) / dz);
  if (ix >= nx || iy >= ny || iz >= nz) return -1;
  return (ix * ny + iy) * nz + iz;
}

void orc_module_mixing(const orc_ctl_t *ctl, const orc_clim_t *clim, orc_atm_t *atm, double t) {
  const int64_t n = atm->np;
  const int ngrid = ctl->mixing_nx * ctl->mixing_ny * ctl->mixing_nz;
  const int nens = ctl->nens > 0 ? ctl->nens : 1;
  const size_t total = (size_t)ngrid * (size_t)nens;
  int32_t *box = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
  double *mean = (double *)malloc(sizeof(double) * total);
  int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * total);
  const double t0 = t - 0.5 * ctl->dt_mod, t1 = t + 0.5 * ctl->dt_mod;
  const double *ens = (ctl->nens > 0) ? atm->q + (size_t)ctl->qnt_ens * atm->q_stride : NULL;
#pragma omp parallel for
  for (int64_t ip = 0; ip < n; ip++)
    box[ip] = box_of(atm->time[ip], atm->lon[ip], atm->lat[ip], atm->p[ip], t0, t1, ctl->mixing_lon0, ctl->mixing_lon1,
                     ctl->mixing_lat0, ctl->mixing_lat1, ctl->mixing_z0, ctl->mixing_z1, ctl->mixing_nx,
                     ctl->mixing_ny, ctl->mixing_nz);
  for (int k = 0; k < ctl->n_mix_qnt; k++) {
    if (ctl->mix_qnt[k] < 0) continue;
    double *q = atm->q + (size_t)ctl->mix_qnt[k] * atm->q_stride;
    memset(mean, 0, sizeof(double) * total);
    memset(cnt, 0, sizeof(int32_t) * total);
    for (int64_t ip = 0; ip < n; ip++)
      if (box[ip] >= 0) {
        const size_t idx = (size_t)(ens ? (int)ens[ip] : 0) * (size_t)ngrid + (size_t)box[ip];
        mean[idx] += q[ip];
        cnt[idx]++;
      }
    for (size_t i = 0; i < total; i++)
      if (cnt[i] > 0) mean[i] /= cnt[i];
#pragma omp parallel for
    for (int64_t ip = 0; ip < n; ip++)
      if (box[ip] >= 0) {
        double mix = 1.0;
        if (ctl->mixing_trop < 1 || ctl->mixing_strat < 1) {
          const double w = w_tropo(ctl, clim, atm->time[ip], atm->lat[ip], atm->p[ip]);
          mix = w * ctl->mixing_trop + (1.0 - w) * ctl->mixing_strat;
        }
        const size_t idx = (size_t)(ens ? (int)ens[ip] : 0) * (size_t)ngrid + (size_t)box[ip];
        q[ip] += (mean[idx] - q[ip]) * mix;
      }
  }
  free(box); free(mean); free(cnt);
}

/* write_grid binning, 13840-13872, kernel weight 1 */
void orc_grid_bin(const orc_atm_t *atm, int nq, int nx, int ny, int nz, double lon0, double lon1,
                  double lat0, double lat1, double z0, double z1, double t0, double t1,
                  int32_t *count, double *sum, double *sumsq) {
  const size_t nb = (size_t)nx * ny * nz;
  memset(count, 0, sizeof(int32_t) * nb);
  memset(sum, 0, sizeof(double) * nb * (size_t)nq);
  memset(sumsq, 0, sizeof(double) * nb * (size_t)nq);
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const int b = box_of(atm->time[ip], atm->lon[ip], atm->lat[ip], atm->p[ip], t0, t1, lon0, lon1, lat0, lat1, z0, z1, nx, ny, nz);
    if (b < 0) continue;
    count[b]++;
    for (int iq = 0; iq < nq; iq++) {
      const double x = atm->q[(size_t)iq * atm->q_stride + ip];
      sum[(size_t)iq * nb + b] += x;
      sumsq[(size_t)iq * nb + b] += x * x;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * mptrac_run_timestep restricted to the path, 7877-7945
 * ------------------------------------------------------------------------------------------- */
/* ---------------------------------------------------------------------------------------------
 * module_meteo, 5062-5165, restricted to the quantities that derive from T, u, v, w, ps, pbl.
 * INTPOL_TIME_ALL (src/mptrac.h:1278-1316): the first 3-D lookup initialises indices and weights, every
 * other field -- 3-D and 2-D -- reuses them; all parcels are visited (check_dt = 0).
 * ------------------------------------------------------------------------------------------- */
#define C_T0 273.15
#define C_MH2O 18.01528              /* mptrac.h:295 */
#define C_EPS (C_MH2O / C_MA)        /* mptrac.h:260 */
#define C_LV 2501000.                /* mptrac.h:275 */
#define C_CPD 1003.5                 /* mptrac.h:255 */
#define C_KAPPA 0.286
static double theta_of(double p, double t) { return t * pow(1000. / p, C_KAPPA); }   /* THETA, mptrac.h:2124 */

void orc_module_meteo(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  const int32_t *qi = ctl->qnt_meteo;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double tm = atm->time[ip], p = atm->p[ip], lon = atm->lon[ip], lat = atm->lat[ip];
    cell_t c = CELL_ZERO;
    const double t = time3(met0, met0->t, met1, met1->t, tm, p, lon, lat, &c, 1);
    const double u = time3(met0, met0->u, met1, met1->u, tm, p, lon, lat, &c, 0);
    const double v = time3(met0, met0->v, met1, met1->v, tm, p, lon, lat, &c, 0);
    const double w = time3(met0, met0->w, met1, met1->w, tm, p, lon, lat, &c, 0);
    const double ps = time2(met0, met0->ps, met1, met1->ps, tm, lon, lat, &c, 0);
    const double pbl = time2(met0, met0->pbl, met1, met1->pbl, tm, lon, lat, &c, 0);
#define SETQ(slot, val) do { if (qi[slot] >= 0) atm->q[(size_t)qi[slot] * (size_t)atm->q_stride + (size_t)ip] = (val); } while (0)
    SETQ(0, ps);
    SETQ(1, pbl);
    SETQ(2, p);
    SETQ(3, t);
    SETQ(4, 100. * p / (C_RA * t));                       /* RHO, mptrac.h:1961 */
    SETQ(5, u);
    SETQ(6, v);
    SETQ(7, w);
    SETQ(8, sqrt(u * u + v * v));
    SETQ(9, -1e3 * C_H0 / p * w);
    SETQ(10, theta_of(p, t));
    SETQ(11, 6.112 * exp(17.62 * (t - C_T0) / (243.12 + t - C_T0)));   /* PSAT, mptrac.h:1808 */
    SETQ(12, 6.112 * exp(22.46 * (t - C_T0) / (272.62 + t - C_T0)));   /* PSICE, mptrac.h:1832 */
    SETQ(13, (p / ps <= 0.3 ? 1. : sin(M_PI / 2. * (1. - p / ps) / (1. - 0.3))) * theta_of(p, t));   /* ZETA, mptrac.h:2293 */
    /* the other fields of INTPOL_TIME_ALL, same stencil */
    for (int f = 0; f < ORC_NX2; f++)
      if (qi[14 + f] >= 0 && met0->x2[f] && met1->x2[f])
        SETQ(14 + f, time2(met0, met0->x2[f], met1, met1->x2[f], tm, lon, lat, &c, 0));
    for (int f = 0; f < ORC_NX3; f++)
      if (qi[36 + f] >= 0 && met0->x3[f] && met1->x3[f])
        SETQ(36 + f, time3(met0, met0->x3[f], met1, met1->x3[f], tm, p, lon, lat, &c, 0));
    int moist = 0;
    for (int k = 45; k <= 52; k++) moist |= qi[k] >= 0;
    if (moist && met0->x3[2] && met1->x3[2]) {
      const double h2o = time3(met0, met0->x3[2], met1, met1->x3[2], tm, p, lon, lat, &c, 0);
      const double hh = h2o > 0.1e-6 ? h2o : 0.1e-6;                       /* MAX((h2o), 0.1e-6) */
      const double pw = p * hh / (1. + (1. - C_EPS) * hh);                 /* PW, mptrac.h:1859 */
      const double sh = C_EPS * hh;                                        /* SH, mptrac.h:2024 */
      SETQ(45, pw);
      SETQ(46, sh);
      SETQ(47, pw / (6.112 * exp(17.62 * (t - C_T0) / (243.12 + t - C_T0))) * 100.);   /* RH, mptrac.h:1906 */
      SETQ(48, pw / (6.112 * exp(22.46 * (t - C_T0) / (272.62 + t - C_T0))) * 100.);   /* RHICE, mptrac.h:1936 */
      SETQ(49, t * (1. + (1. - C_EPS) * hh));                              /* TVIRT, mptrac.h:2199 */
      {                                                                    /* lapse_rate, src/mptrac.c:3324-3338 */
        const double a = C_RA * (t * t), r = sh / (1. - sh);
        SETQ(50, 1e3 * C_G0 * (a + C_LV * r * t) / (C_CPD * a + (C_LV * C_LV) * r * C_EPS));
      }
      SETQ(51, C_T0 + 243.12 * log(pw / 6.112) / (17.62 - log(pw / 6.112)));   /* TDEW, mptrac.h:2075 */
      SETQ(52, C_T0 + 272.62 * log(pw / 6.112) / (22.46 - log(pw / 6.112)));   /* TICE, mptrac.h:2100 */
    }
#undef SETQ
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_diff_pbl, 4343-4584 (TURB_PBL_SCHEME 1): Hanna / FLEXPART turbulence closure inside the boundary layer -- neutral,
 * unstable and stable regimes from the Monin-Obukhov length --, Langevin update of the three velocity perturbations,
 * horizontal displacement, vertical displacement in geometric height with reflection at the ground and the PBL top
 * ------------------------------------------------------------------------------------------- */
#define O_MAX(a, b) (((a) > (b)) ? (a) : (b))
#define O_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define O_CLAMP(v, lo, hi) (((v) < (lo)) ? (lo) : (((v) > (hi)) ? (hi) : (v)))
#define O_SQR(x) ((x) * (x))
#define O_Z(p) (C_H0 * log(C_P0 / (p)))
void orc_module_diff_pbl(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr) {
  (void)ctl;
  if (!met0->x2[4] || !met0->x2[5] || !met0->x2[6] || !met0->x3[2] || !met1->x2[4] || !met1->x2[5] || !met1->x2[6] || !met1->x3[2]) {
    fprintf(stderr, "orc_module_diff_pbl: the met levels lack ess, nss, shf or h2o\n");
    abort();
  }
  orc_module_rng(atm->rs, 3 * atm->np, 1, ctr);
  const int ct = met0->coord_type;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (atm->dt[ip] == 0) continue;
    const double tm = atm->time[ip], lon = atm->lon[ip], lat = atm->lat[ip];
    double dsigw_dz = 0.0, sig_u = 0.0, sig_v = 0.0, sig_w = 0.0, tau_u = 0.0, tau_v = 0.0, tau_w = 0.0;
    cell_t c = CELL_ZERO;
    const double pbl = time2(met0, met0->pbl, met1, met1->pbl, tm, lon, lat, &c, 1);
    if (atm->p[ip] < pbl) continue;
    const double ps = time2(met0, met0->ps, met1, met1->ps, tm, lon, lat, &c, 0);
    if (!(ps > 0.0 && pbl > 0.0 && ps > pbl)) continue;
    const double p = O_MIN(atm->p[ip], ps);
    const double zs = O_Z(ps);
    const double z_raw = 1e3 * (O_Z(p) - zs);
    const double zi = 1e3 * (O_Z(pbl) - zs);
    if (!(zi > 1.0)) continue;
    const double z = O_CLAMP(z_raw, 0.0, zi);
    const double zeta = O_CLAMP(z / zi, 1e-6, 1.0 - 1e-6);
    const double z_m = O_MAX(z, 1.0);
    const double ess = time2(met0, met0->x2[4], met1, met1->x2[4], tm, lon, lat, &c, 0);
    const double nss = time2(met0, met0->x2[5], met1, met1->x2[5], tm, lon, lat, &c, 0);
    const double t = time3(met0, met0->t, met1, met1->t, tm, p, lon, lat, &c, 1);
    const double h2o = time3(met0, met0->x3[2], met1, met1->x3[2], tm, p, lon, lat, &c, 0);
    const double hh = O_MAX(h2o, 0.1e-6);
    const double tv = t * (1. + (1. - C_EPS) * hh);
    const double thetav = theta_of(p, t) * (1. + (1. - C_EPS) * O_MAX(hh, 0.1e-6));
    const double rho = 100. * p / (C_RA * tv);
    const double tau = sqrt(O_SQR(ess) + O_SQR(nss));
    if (!(rho > 0.0)) continue;
    const double ustar = sqrt(O_MAX(tau / rho, 0.0));
    const double ust = O_MAX(1e-4, ustar);
    const double shf = time2(met0, met0->x2[6], met1, met1->x2[6], tm, lon, lat, &c, 1);
    double ol = 1e12;
    if (fabs(shf) > 1e-6) ol = thetav * rho * C_CPD * O_SQR(ust) * ust / (0.40 * C_G0 * shf);
    if (zi / fabs(ol) < 1.0) {
      const double corr = z_m / ust;
      const double sigw0 = 1.3 * ust * exp(-2e-4 * corr);
      sig_u = O_MAX(2.0 * ust * exp(-3e-4 * corr), 1e-5);
      sig_v = O_MAX(sigw0, 1e-5);
      sig_w = O_MAX(sigw0, 1e-5);
      dsigw_dz = -2e-4 * sigw0 / ust;
      tau_u = 0.5 * z_m / sig_w / (1.0 + 1.5e-3 * corr);
      tau_v = tau_u;
      tau_w = tau_u;
    } else if (ol < 0.0) {
      const double wstar_arg = -C_G0 / thetav * shf / (rho * C_CPD) * zi;
      const double wstar = pow(O_MAX(wstar_arg, 0.0), 1.0 / 3.0);
      double dsigw2_dz = 0.0;
      sig_u = O_MAX(ust * pow(O_MAX(12.0 - 0.5 * zi / ol, 0.0), 1.0 / 3.0), 1e-6);
      sig_v = sig_u;
      if (zeta < 0.03) {
        const double arg = O_MAX(3.0 * zeta - ol / zi, 1e-12);
        sig_w = 0.96 * wstar * pow(arg, 1.0 / 3.0);
        dsigw2_dz = 1.8432 * O_SQR(wstar) / zi * pow(arg, -1.0 / 3.0);
      } else if (zeta < 0.4) {
        const double arg = O_MAX(3.0 * zeta - ol / zi, 1e-12);
        const double s1 = 0.96 * pow(arg, 1.0 / 3.0);
        const double s2 = 0.763 * pow(zeta, 0.175);
        if (s1 < s2) {
          sig_w = wstar * s1;
          dsigw2_dz = 1.8432 * O_SQR(wstar) / zi * pow(arg, -1.0 / 3.0);
        } else {
          sig_w = wstar * s2;
          dsigw2_dz = 0.203759 * O_SQR(wstar) / zi * pow(zeta, -0.65);
        }
      } else if (zeta < 0.96) {
        sig_w = 0.722 * wstar * pow(1.0 - zeta, 0.207);
        dsigw2_dz = -0.215812 * O_SQR(wstar) / zi * pow(1.0 - zeta, -0.586);
      } else {
        sig_w = 0.37 * wstar;
        dsigw2_dz = 0.0;
      }
      sig_w = O_MAX(sig_w, 1e-6);
      dsigw_dz = sig_w > 1e-12 ? 0.5 * dsigw2_dz / sig_w : 0.0;
      tau_u = 0.15 * zi / O_MAX(sig_u, 1e-12);
      tau_v = tau_u;
      if (z_m < fabs(ol)) {
        const double denom = 0.55 - 0.38 * fabs(z_m / ol);
        tau_w = 0.1 * z_m / (sig_w * O_MAX(denom, 0.05));
      } else if (zeta < 0.1)
        tau_w = 0.59 * z_m / sig_w;
      else
        tau_w = 0.15 * zi / sig_w * (1.0 - exp(-5.0 * zeta));
    } else {
      sig_u = O_MAX(2.0 * ust * (1.0 - zeta), 1e-6);
      sig_v = O_MAX(1.3 * ust * (1.0 - zeta), 1e-6);
      sig_w = O_MAX(1.3 * ust * (1.0 - zeta), 1e-6);
      dsigw_dz = -1.3 * ust / zi;
      tau_u = 0.15 * zi / sig_u * sqrt(zeta);
      tau_v = 0.467 * tau_u;
      tau_w = 0.1 * zi / sig_w * pow(zeta, 0.8);
    }
    tau_u = O_MAX(tau_u, 10.0);
    tau_v = O_MAX(tau_v, 10.0);
    tau_w = O_MAX(tau_w, 30.0);
    if (!(sig_u > 0.0 && sig_v > 0.0 && sig_w > 0.0 && tau_u > 0.0 && tau_v > 0.0 && tau_w > 0.0)) continue;
    const double dt = atm->dt[ip], dt_abs = fabs(dt);
    const double ru = exp(-dt_abs / tau_u), ru2 = sqrt(O_MAX(0.0, 1.0 - O_SQR(ru)));
    const double rv = exp(-dt_abs / tau_v), rv2 = sqrt(O_MAX(0.0, 1.0 - O_SQR(rv)));
    float *up = atm->uvwp + 3 * ip;
    up[0] = (float)(up[0] * ru + sig_u * ru2 * atm->rs[3 * ip]);
    up[1] = (float)(up[1] * rv + sig_v * rv2 * atm->rs[3 * ip + 1]);
    const double rw = exp(-dt_abs / tau_w), rw2 = sqrt(O_MAX(0.0, 1.0 - O_SQR(rw)));
    const double rhoaux = -1.0 / (1e3 * C_H0);
    up[2] = (float)(up[2] * rw + sig_w * rw2 * atm->rs[3 * ip + 2]
                    + tau_w * (1.0 - rw) * (2.0 * sig_w * dsigw_dz + rhoaux * O_SQR(sig_w)));
    atm->lon[ip] += east_m_to_coord(ct, up[0] * dt, atm->lat[ip]);
    atm->lat[ip] += north_m_to_coord(ct, up[1] * dt);
    double znew = z + up[2] * dt;
    while (znew < 0.0 || znew > zi) {
      if (znew < 0.0) { znew = -znew; up[2] = -up[2]; }
      if (znew > zi) { znew = 2.0 * zi - znew; up[2] = -up[2]; }
    }
    atm->p[ip] = C_P0 * exp(-(zs + znew / 1000.0) / C_H0);
    atm->p[ip] = O_CLAMP(atm->p[ip], pbl, ps);
  }
}

/* ---------------------------------------------------------------------------------------------
 * module_convection, 4102-4171: one uniform random number per parcel (module_rng method 0), the mixing range from the
 * surface to the PBL top and / or the equilibrium level where CAPE (and CIN) pass their thresholds, the new pressure
 * uniformly distributed in density between the two
 * ------------------------------------------------------------------------------------------- */
void orc_module_convection(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr) {
  if (ctl->conv_cape >= 0 && (!met0->x2[18] || !met0->x2[19] || !met0->x2[20] || !met1->x2[18] || !met1->x2[19] || !met1->x2[20])) {
    fprintf(stderr, "orc_module_convection: the met levels lack cape, cin or pel\n");
    abort();
  }
  orc_module_rng(atm->rs, atm->np, 0, ctr);
  const float *cape0 = met0->x2[19], *cape1 = met1->x2[19], *cin0 = met0->x2[20], *cin1 = met1->x2[20];
  const float *pel0 = met0->x2[18], *pel1 = met1->x2[18];
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (atm->dt[ip] == 0) continue;
    const double tm = atm->time[ip], lon = atm->lon[ip], lat = atm->lat[ip];
    cell_t c = CELL_ZERO;
    const double ps = time2(met0, met0->ps, met1, met1->ps, tm, lon, lat, &c, 1);
    const double pbot = ps;
    double ptop = ps;
    if (ctl->conv_mix_pbl) {
      const double pbl = time2(met0, met0->pbl, met1, met1->pbl, tm, lon, lat, &c, 0);
      ptop = pbl - ctl->conv_pbl_trans * (ps - pbl);
    }
    if (ctl->conv_cape >= 0) {
      const double cape = time2(met0, cape0, met1, cape1, tm, lon, lat, &c, 0);
      const double cin = time2(met0, cin0, met1, cin1, tm, lon, lat, &c, 0);
      const double pel = time2(met0, pel0, met1, pel1, tm, lon, lat, &c, 0);
      if (isfinite(cape) && cape >= ctl->conv_cape && (ctl->conv_cin <= 0 || (isfinite(cin) && cin >= ctl->conv_cin)))
        ptop = ptop < pel ? ptop : pel;   /* GSL_MIN */
    }
    if (ptop != pbot && atm->p[ip] >= ptop) {
      const double tbot = time3(met0, met0->t, met1, met1->t, tm, pbot, lon, lat, &c, 1);
      const double ttop = time3(met0, met0->t, met1, met1->t, tm, ptop, lon, lat, &c, 1);
      const double rhobot = pbot / tbot, rhotop = ptop / ttop;
      const double rho = rhobot + (rhotop - rhobot) * atm->rs[ip];
      atm->p[ip] = linear(rhobot, pbot, rhotop, ptop, rho);
    }
  }
}

/* module_isosurf_init 4886-4952 (modes 1-3; mode 4 reads the balloon file into cache->iso_ts / iso_ps, which the caller
 * provides here) and module_isosurf 4956-5004: every parcel, dt or not */
void orc_module_isosurf_init(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  if (ctl->isosurf < 1 || ctl->isosurf > 3) return;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (ctl->isosurf == 1) { atm->iso_var[ip] = atm->p[ip]; continue; }
    cell_t c = CELL_ZERO;
    const double t = time3(met0, met0->t, met1, met1->t, atm->time[ip], atm->p[ip], atm->lon[ip], atm->lat[ip], &c, 1);
    atm->iso_var[ip] = ctl->isosurf == 2 ? atm->p[ip] / t : theta_of(atm->p[ip], t);
  }
}

void orc_module_isosurf(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (ctl->isosurf == 1) {
      atm->p[ip] = atm->iso_var[ip];
    } else if (ctl->isosurf == 2 || ctl->isosurf == 3) {
      cell_t c = CELL_ZERO;
      const double t = time3(met0, met0->t, met1, met1->t, atm->time[ip], atm->p[ip], atm->lon[ip], atm->lat[ip], &c, 1);
      atm->p[ip] = ctl->isosurf == 2 ? atm->iso_var[ip] * t : 1000. * pow(atm->iso_var[ip] / t, -1. / C_KAPPA);
    } else if (ctl->isosurf == 4) {
      const double tm = atm->time[ip];
      if (tm <= atm->iso_ts[0]) atm->p[ip] = atm->iso_ps[0];
      else if (tm >= atm->iso_ts[atm->iso_n - 1]) atm->p[ip] = atm->iso_ps[atm->iso_n - 1];
      else {
        const int i = bisect(atm->iso_ts, atm->iso_n, tm);
        atm->p[ip] = linear(atm->iso_ts[i], atm->iso_ps[i], atm->iso_ts[i + 1], atm->iso_ps[i + 1], tm);
      }
    }
  }
}

/* module_bound_cond, 3789-3881 (clim_ts 396-410): parcels with dt != 0 inside the latitude / pressure window and, where asked,
 * inside the surface layer get their mass, volume mixing ratio, trace-gas and age-of-air quantities reset */
static const orc_cts_t *g_cts;
void orc_set_cts(const orc_cts_t *cts) { g_cts = cts; }

static double series_at(const orc_cts_t *cts, int k, double t) {
  const double *tm = cts->time[k], *v = cts->vmr[k];
  const int n = cts->n[k];
  if (t <= tm[0]) return v[0];
  if (t >= tm[n - 1]) return v[n - 1];
  const int i = bisect(tm, n, t);
  return linear(tm[i], v[i], tm[i + 1], v[i + 1], t);
}

void orc_module_bound_cond(const orc_ctl_t *ctl, const orc_cts_t *cts, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm) {
  /* (the reference tests qnt_Cccl4 for truth, not for >= 0: src/mptrac.c:3802) */
  if (ctl->qnt_m < 0 && ctl->qnt_vmr < 0 && ctl->qnt_cts[0] && ctl->qnt_cts[1] < 0 && ctl->qnt_cts[2] < 0 && ctl->qnt_cts[3] < 0 &&
      ctl->qnt_cts[4] < 0 && ctl->qnt_aoa < 0)
    return;
  const size_t st = (size_t)atm->q_stride;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (atm->dt[ip] == 0) continue;
    const double tm = atm->time[ip], lon = atm->lon[ip], lat = atm->lat[ip], p = atm->p[ip];
    if (lat < ctl->bound_lat0 || lat > ctl->bound_lat1 || p > ctl->bound_p0 || p < ctl->bound_p1) continue;
    if (ctl->bound_dps > 0 || ctl->bound_dzs > 0 || ctl->bound_zetas > 0 || ctl->bound_pbl) {
      cell_t c = CELL_ZERO;
      const double ps = time2(met0, met0->ps, met1, met1->ps, tm, lon, lat, &c, 1);
      if (ctl->bound_dps > 0 && p < ps - ctl->bound_dps) continue;
      if (ctl->bound_dzs > 0 && C_H0 * log(C_P0 / p) > C_H0 * log(C_P0 / ps) + ctl->bound_dzs) continue;
      if (ctl->bound_zetas > 0) {
        const double t = time3(met0, met0->t, met1, met1->t, tm, p, lon, lat, &c, 1);
        if ((p / ps <= 0.3 ? 1. : sin(M_PI / 2. * (1. - p / ps) / (1. - 0.3))) * theta_of(p, t) > ctl->bound_zetas) continue;
      }
      if (ctl->bound_pbl) {
        const double pbl = time2(met0, met0->pbl, met1, met1->pbl, tm, lon, lat, &c, 0);
        if (p < pbl) continue;
      }
    }
    if (ctl->qnt_m >= 0 && ctl->bound_mass >= 0) atm->q[(size_t)ctl->qnt_m * st + ip] = ctl->bound_mass + ctl->bound_mass_trend * tm;
    if (ctl->qnt_vmr >= 0 && ctl->bound_vmr >= 0) atm->q[(size_t)ctl->qnt_vmr * st + ip] = ctl->bound_vmr + ctl->bound_vmr_trend * tm;
    for (int k = 0; k < ORC_NCTS; k++)
      if (ctl->qnt_cts[k] >= 0 && (ctl->cts_on >> k & 1)) atm->q[(size_t)ctl->qnt_cts[k] * st + ip] = series_at(cts, k, tm);
    if (ctl->qnt_aoa >= 0) atm->q[(size_t)ctl->qnt_aoa * st + ip] = tm;
  }
}

/* module_chem_grid, 3885-4054: mass per box of a regular lon / lat / log-pressure-height grid (summed in parcel order, like the
 * reference's serial loop) -> volume mixing ratio of the box at the temperature of its centre -> quantity Cx of its parcels */
void orc_module_chem_grid(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, double tt) {
  if (ctl->qnt_m < 0 || ctl->qnt_Cx < 0) return;
  const int nx = ctl->chemgrid_nx, ny = ctl->chemgrid_ny, nz = ctl->chemgrid_nz, ngrid = nx * ny * nz;
  const int nens = ctl->nens > 0 ? ctl->nens : 1;
  const size_t st = (size_t)atm->q_stride;
  const double dz = (ctl->chemgrid_z1 - ctl->chemgrid_z0) / nz, dlon = (ctl->chemgrid_lon1 - ctl->chemgrid_lon0) / nx;
  const double dlat = (ctl->chemgrid_lat1 - ctl->chemgrid_lat0) / ny;
  const double t0 = tt - 0.5 * ctl->dt_mod, t1 = tt + 0.5 * ctl->dt_mod;
  double *mass = calloc((size_t)ngrid * (size_t)nens, sizeof(double));
  int *idx = malloc(sizeof(int) * (size_t)(atm->np > 0 ? atm->np : 1));
  for (int64_t ip = 0; ip < atm->np; ip++) {
    const double zpart = C_H0 * log(C_P0 / atm->p[ip]);
    idx[ip] = -1;
    if (atm->time[ip] < t0 || atm->time[ip] > t1 || atm->lon[ip] < ctl->chemgrid_lon0 || atm->lon[ip] >= ctl->chemgrid_lon1 ||
        atm->lat[ip] < ctl->chemgrid_lat0 || atm->lat[ip] >= ctl->chemgrid_lat1 || zpart < ctl->chemgrid_z0 || zpart >= ctl->chemgrid_z1)
      continue;
    const int ix = (int)((atm->lon[ip] - ctl->chemgrid_lon0) / dlon), iy = (int)((atm->lat[ip] - ctl->chemgrid_lat0) / dlat);
    const int iz = (int)((zpart - ctl->chemgrid_z0) / dz);
    if (ix >= nx || iy >= ny || iz >= nz) continue;
    idx[ip] = (ix * ny + iy) * nz + iz;
    int mi = idx[ip];
    if (ctl->nens > 0) mi += (int)atm->q[(size_t)ctl->qnt_ens * st + ip] * ngrid;
    mass[mi] += atm->q[(size_t)ctl->qnt_m * st + ip];
  }
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (idx[ip] < 0) continue;
    const int iz = idx[ip] % nz, iy = (idx[ip] / nz) % ny, ix = idx[ip] / (nz * ny);
    const double z = ctl->chemgrid_z0 + dz * (iz + 0.5), press = C_P0 * exp(-z / C_H0);
    const double lon = ctl->chemgrid_lon0 + dlon * (ix + 0.5), lat = ctl->chemgrid_lat0 + dlat * (iy + 0.5);
    const double area = dlat * dlon * ((C_RE * M_PI / 180.) * (C_RE * M_PI / 180.)) * cos(lat * (M_PI / 180.0));
    cell_t c = CELL_ZERO;
    const double temp = time3(met0, met0->t, met1, met1->t, tt, press, lon, lat, &c, 1);
    int mi = idx[ip];
    if (ctl->nens > 0) mi += (int)atm->q[(size_t)ctl->qnt_ens * st + ip] * ngrid;
    atm->q[(size_t)ctl->qnt_Cx * st + ip] = C_MA / ctl->molmass * mass[mi] / ((100. * press / (C_RA * temp)) * area * dz * 1e9);
  }
  free(mass);
  free(idx);
}

/* module_decay, 4227-4263 (and the reset of the total loss rate that precedes it, 7931-7936) */
void orc_module_decay(const orc_ctl_t *ctl, const orc_clim_t *clim, orc_atm_t *atm) {
  const size_t st = (size_t)atm->q_stride;
#pragma omp parallel for
  for (int64_t ip = 0; ip < atm->np; ip++) {
    if (atm->dt[ip] == 0) continue;
    const double w = w_tropo(ctl, clim, atm->time[ip], atm->lat[ip], atm->p[ip]);
    const double tdec = w * ctl->tdec_trop + (1 - w) * ctl->tdec_strat;
    const double aux = exp(-atm->dt[ip] / tdec);
    if (ctl->qnt_m >= 0) {
      double *m = atm->q + (size_t)ctl->qnt_m * st + ip;
      if (ctl->qnt_mloss_decay >= 0) atm->q[(size_t)ctl->qnt_mloss_decay * st + ip] += *m * (1 - aux);
      *m *= aux;
      if (ctl->qnt_loss_rate >= 0) atm->q[(size_t)ctl->qnt_loss_rate * st + ip] += 1. / tdec;
    }
    if (ctl->qnt_vmr >= 0) atm->q[(size_t)ctl->qnt_vmr * st + ip] *= aux;
  }
}

void orc_run_timestep(const orc_ctl_t *ctl, const orc_clim_t *clim, const orc_met_t *met0,
                      const orc_met_t *met1, orc_atm_t *atm, double t, uint64_t *ctr) {
  if (t == ctl->t_start) {   /* 7863-7873 */
    if (ctl->isosurf >= 1 && ctl->isosurf <= 4) orc_module_isosurf_init(ctl, met0, met1, atm);
    orc_module_advect_init(ctl, met0, met1, atm);
  }
  orc_module_timesteps(ctl, met0, atm, t);
  if (ctl->sort_dt > 0 && fmod(t, ctl->sort_dt) == 0) orc_module_sort(ctl, met0, atm);
  orc_module_position(met0, met1, atm);
  if (ctl->advect > 0) orc_module_advect(ctl, met0, met1, atm);
  if (ctl->diffusion && (ctl->turb_dx_pbl > 0 || ctl->turb_dz_pbl > 0 || ctl->turb_dx_trop > 0 ||
                         ctl->turb_dz_trop > 0 || ctl->turb_dx_strat > 0 || ctl->turb_dz_strat > 0))
    orc_module_diff_turb(ctl, clim, met0, met1, atm, ctr);
  if (ctl->diffusion && ctl->turb_pbl_scheme == 1) orc_module_diff_pbl(ctl, met0, met1, atm, ctr);   /* 7897-7899 */
  if (ctl->diffusion && (ctl->turb_mesox > 0 || ctl->turb_mesoz > 0)) orc_module_diff_meso(ctl, met0, met1, atm, ctr);
  if ((ctl->conv_mix_pbl || ctl->conv_cape >= 0) && (ctl->conv_dt <= 0 || fmod(t, ctl->conv_dt) == 0))   /* 7905-7908 */
    orc_module_convection(ctl, met0, met1, atm, ctr);
  if (ctl->qnt_rp >= 0 && ctl->qnt_rhop >= 0) orc_module_sedi(ctl, met0, met1, atm);
  if (ctl->isosurf >= 1 && ctl->isosurf <= 4) orc_module_isosurf(ctl, met0, met1, atm);   /* 7914-7916 */
  orc_module_position(met0, met1, atm);
  if (ctl->met_dt_out > 0 && (ctl->met_dt_out < ctl->dt_mod || fmod(t, ctl->met_dt_out) == 0)) {   /* 7927-7929 */
    int any = 0;
    for (int i = 0; i < ORC_METEO_SLOTS; i++) any |= ctl->qnt_meteo[i] >= 0;
    if (any) orc_module_meteo(ctl, met0, met1, atm);
  }
  const int bound = ctl->bound_lat0 < ctl->bound_lat1 && ctl->bound_p0 > ctl->bound_p1;
  if (bound) orc_module_bound_cond(ctl, g_cts, met0, met1, atm);   /* 7926-7929 */
  if (ctl->qnt_loss_rate >= 0)   /* 7931-7936 */
    for (int64_t ip = 0; ip < atm->np; ip++)
      if (atm->dt[ip] != 0) atm->q[(size_t)ctl->qnt_loss_rate * (size_t)atm->q_stride + ip] = 0;
  if (ctl->tdec_trop > 0 && ctl->tdec_strat > 0) orc_module_decay(ctl, clim, atm);   /* 7938-7940 */
  if (ctl->mixing_trop >= 0 && ctl->mixing_strat >= 0 && (ctl->mixing_dt <= 0 || fmod(t, ctl->mixing_dt) == 0))
    orc_module_mixing(ctl, clim, atm, t);
  if (ctl->chemgrid) orc_module_chem_grid(ctl, met0, met1, atm, t);   /* 7947-7950 */
  if (bound) orc_module_bound_cond(ctl, g_cts, met0, met1, atm);    /* 7997-8000 */
}
