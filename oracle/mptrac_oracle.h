/*
 * mptrac_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU restatement of the algorithm of MPTRAC's per-particle time-step path
 * (slcs-jsc/mptrac, src/mptrac.c; line numbers in mptrac_oracle.c).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA engine on a
 * machine that has neither the reference tree nor its dependencies.  Nothing under mptrac_b200/
 * includes, links or loads it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this file against the unmodified
 * reference built into oracle/_ref (module by module and through mptrac_run_timestep, on synthetic
 * and ERA-Interim met) and against the reference's own goldens of tests/dt_test, tests/coord_test
 * and tests/tools_test (sedi.tab); tests/golden/ holds vectors generated from the reference for
 * machines without oracle/_ref.
 */
#ifndef MPTRAC_ORACLE_H
#define MPTRAC_ORACLE_H

#include <stdint.h>

#define ORC_MIX_MAXQ 23
#define ORC_METEO_SLOTS 64
#define ORC_NX2 22
#define ORC_NX3 9
/* slots of orc_ctl_t::qnt_meteo:
 *    0-13  ps, pbl, p, t, rho, u, v, w, vh, vz, theta, psat, psice, zeta_d
 *   14-35  the 2-D fields of orc_met_t::x2 in their order: ts, zs, us, vs, ess, nss, shf, lsm, sst, pt, tt, zt, h2ot, pct,
 *          pcb, cl, plcl, plfc, pel, cape, cin, o3c
 *   36-44  the 3-D fields of orc_met_t::x3 in their order: zg (met_t::z), pv, h2o, o3, lwc, rwc, iwc, swc, cc
 *   45-52  pw, sh, rh, rhice, tvirt, lapse, tdew, tice */

/* same field order as mpb_ctl_t so one ctypes structure serves both (the oracle defines its own type
 * on purpose: it must not depend on product headers) */
typedef struct {
  int32_t direction, met_coord_type, advect, advect_vert_coord, rng_type, diffusion, turb_pbl_scheme;
  int32_t nq, qnt_rp, qnt_rhop, qnt_ens, nens;
  int32_t mixing_nx, mixing_ny, mixing_nz;
  int32_t n_mix_qnt;
  int32_t mix_qnt[ORC_MIX_MAXQ];
  int32_t _pad;
  double t_start, t_stop, dt_mod, dt_met, met_utm_ref_lat, sort_dt;
  double turb_dx_pbl, turb_dx_trop, turb_dx_strat, turb_dz_pbl, turb_dz_trop, turb_dz_strat;
  double turb_mesox, turb_mesoz, turb_pbl_trans;
  double mixing_dt, mixing_trop, mixing_strat;
  double mixing_lon0, mixing_lon1, mixing_lat0, mixing_lat1, mixing_z0, mixing_z1;
  double met_dt_out;
  int32_t qnt_meteo[ORC_METEO_SLOTS];   /* quantity index or -1 */
  int32_t qnt_zeta, qnt_eta;            /* vertical coordinate quantity of ADVECT_VERT_COORD 1 / 3, or -1 */
  /* module_convection (src/mptrac.c:4102-4171) and module_decay (4227-4263) */
  double conv_cape, conv_cin, conv_pbl_trans, conv_dt;
  double tdec_trop, tdec_strat;
  int32_t conv_mix_pbl;
  int32_t qnt_m, qnt_vmr, qnt_mloss_decay, qnt_loss_rate;
  int32_t isosurf;                      /* ctl->isosurf: 0 off, 1 pressure, 2 density, 3 potential temperature, 4 balloon */
  /* module_bound_cond (src/mptrac.c:3789-3881) */
  double bound_mass, bound_mass_trend, bound_vmr, bound_vmr_trend;
  double bound_lat0, bound_lat1, bound_p0, bound_p1, bound_dps, bound_dzs, bound_zetas;
  int32_t bound_pbl, qnt_aoa;
  int32_t qnt_cts[5];                   /* qnt_Cccl4, qnt_Cccl3f, qnt_Cccl2f2, qnt_Cn2o, qnt_Csf6 */
  int32_t cts_on;                       /* bit i: the control file names a time series for species i (not "-") */
  /* module_chem_grid (src/mptrac.c:3885-4054) */
  double chemgrid_lon0, chemgrid_lon1, chemgrid_lat0, chemgrid_lat1, chemgrid_z0, chemgrid_z1, molmass;
  int32_t chemgrid_nx, chemgrid_ny, chemgrid_nz, qnt_Cx;
  int32_t chemgrid;                     /* the dispatcher's condition (oh / h2o2 / kpp chemistry on), 7947-7950 */
  int32_t _pad3;
} orc_ctl_t;

#define ORC_NCTS 5
typedef struct {                        /* clim_ts_t of the five species above (src/mptrac.h:3733-3744) */
  int32_t n[ORC_NCTS], _pad;
  const double *time[ORC_NCTS], *vmr[ORC_NCTS];
} orc_cts_t;

/* one met time level, dense: 3-D [nx][ny][np] (z fastest), 2-D [nx][ny] */
typedef struct {
  double time;
  int32_t coord_type, nx, ny, np;
  const double *lon, *lat, *p;
  const float *u, *v, *w, *t, *ps, *pbl;
  /* model-level fields (ADVECT_VERT_COORD 1, 2, 3), dense [nx][ny][npl]; NULL when absent */
  int32_t npl;
  const float *pl, *ul, *vl, *wl, *zetal, *zeta_dotl;
  /* further fields module_meteo interpolates (INTPOL_TIME_ALL, src/mptrac.h:1278-1316); NULL when absent */
  const float *x2[ORC_NX2];   /* [nx][ny] */
  const float *x3[ORC_NX3];   /* [nx][ny][np] */
} orc_met_t;

typedef struct {
  int32_t ntime, nlat;
  const double *time, *lat, *tropo; /* [ntime][nlat] */
} orc_clim_t;

/* parcels + per-parcel cache (atm_t / cache_t) */
typedef struct {
  int64_t np;
  double *time, *p, *lon, *lat;
  double *q;        /* [nq][q_stride] */
  int64_t q_stride;
  double *dt;       /* [np]   cache->dt   */
  float *uvwp;      /* [np][3] cache->uvwp */
  double *rs;       /* [3 np + 1] cache->rs */
  double *iso_var;  /* [np] cache->iso_var (module_isosurf), attached to the array slot like uvwp; may be NULL */
  int32_t iso_n, _pad;              /* balloon time series of ISOSURF 4: cache->iso_n, iso_ts, iso_ps */
  const double *iso_ts, *iso_ps;
} orc_atm_t;

void orc_module_timesteps(const orc_ctl_t *ctl, const orc_met_t *met0, orc_atm_t *atm, double t);
void orc_module_position(const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_advect(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_advect_init(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_intpol_met_4d_zeta(const orc_met_t *met0, const float *h0, const float *a0, const orc_met_t *met1, const float *h1,
                            const float *a1, double ts, double height, double lon, double lat, double *var);
void orc_module_rng(double *rs, int64_t n, int method, uint64_t *ctr);
void orc_module_diff_turb(const orc_ctl_t *ctl, const orc_clim_t *clim, const orc_met_t *met0,
                          const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr);
void orc_module_diff_meso(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1,
                          orc_atm_t *atm, uint64_t *ctr);
void orc_module_sedi(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_sort(const orc_ctl_t *ctl, const orc_met_t *met0, orc_atm_t *atm);
void orc_module_meteo(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_diff_pbl(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr);
void orc_module_convection(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, uint64_t *ctr);
void orc_module_decay(const orc_ctl_t *ctl, const orc_clim_t *clim, orc_atm_t *atm);
void orc_module_bound_cond(const orc_ctl_t *ctl, const orc_cts_t *cts, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_chem_grid(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm, double t);
void orc_set_cts(const orc_cts_t *cts);   /* the series orc_run_timestep's boundary conditions use (NULL = none) */
void orc_module_isosurf_init(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_isosurf(const orc_ctl_t *ctl, const orc_met_t *met0, const orc_met_t *met1, orc_atm_t *atm);
void orc_module_mixing(const orc_ctl_t *ctl, const orc_clim_t *clim, orc_atm_t *atm, double t);
void orc_run_timestep(const orc_ctl_t *ctl, const orc_clim_t *clim, const orc_met_t *met0,
                      const orc_met_t *met1, orc_atm_t *atm, double t, uint64_t *ctr);

double orc_sedi(double p, double T, double rp, double rhop);
double orc_clim_tropo(const orc_clim_t *clim, double t, double lat);
void orc_intpol_met_time_3d(const orc_met_t *met0, const orc_met_t *met1, const float *f0, const float *f1,
                            double ts, double p, double lon, double lat, double *var);
void orc_sort_keys(const orc_met_t *met0, const orc_atm_t *atm, int32_t *keys);
void orc_grid_bin(const orc_atm_t *atm, int nq, int nx, int ny, int nz, double lon0, double lon1,
                  double lat0, double lat1, double z0, double z1, double t0, double t1,
                  int32_t *count, double *sum, double *sumsq);

#endif
