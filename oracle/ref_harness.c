/*
 * ref_harness.c -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference (slcs-jsc/mptrac) on in-memory
 * inputs so the restated oracle and the CUDA engine can be compared with the real thing.
 *
 * It is compiled by oracle/build_ref.sh into oracle/_ref/lib/libref_harness.so and pulls the
 * reference in with `#include "mptrac.c"` straight from /root/reference/src (nothing is copied into
 * this repository); including the translation unit, rather than linking it, is what lets the harness
 * set the file-static Squares counter `rng_ctr` (src/mptrac.c:35).
 *
 * All reference calls go through the reference's own entry points: mptrac_alloc, mptrac_read_ctl
 * (with the "-" pseudo file so that every default is the reference's), clim_tropo_init,
 * mptrac_run_timestep and the module_* functions.
 */
#include "mptrac.c" /* the reference translation unit, found through -I<reference>/src */

#include "mptrac_oracle.h"

static ctl_t *h_ctl;
static cache_t *h_cache;
static clim_t *h_clim;
static met_t *h_met0, *h_met1;
static atm_t *h_atm;
static depo_t *h_depo;
static dd_t *h_dd;

int ref_dims(int *ex, int *ey, int *ep, int *np, int *nq) {
  *ex = EX; *ey = EY; *ep = EP; *np = NP; *nq = NQ;
  return 0;
}

static void ensure_alloc(void) {
  if (!h_ctl) mptrac_alloc(&h_ctl, &h_cache, &h_clim, &h_met0, &h_met1, &h_atm, &h_depo, &h_dd);
}

/* Reference defaults + quantity names (comma separated) + "KEY VALUE" overrides (space separated).
 * Returns the quantity index the reference assigned to `rp`, `rhop`, `m`, `vmr`, `ens` and the 14 module_meteo
 * quantities (slot order of orc_ctl_t::qnt_meteo, out[5..68]) and `zeta`, `eta`, `mloss_decay`, `loss_rate`, `aoa`, `Cccl4`, `Cccl3f`, `Cccl2f2`, `Cn2o`, `Csf6`, `Cx` (out[70..80]); out holds 81 ints. */
int ref_read_ctl(const char *qnt_names, const char *overrides, int *out) {
  ensure_alloc();
  static char buf[8192];
  char *argv[512];
  int argc = 0;
  argv[argc++] = (char *)"ref_harness";
  argv[argc++] = (char *)"-";
  char *w = buf;
  int nq = 0;
  static char names[64][64];
  if (qnt_names && qnt_names[0]) {
    const char *s = qnt_names;
    while (*s && nq < 64) {
      int k = 0;
      while (*s && *s != ',' && k < 63) names[nq][k++] = *s++;
      names[nq][k] = 0;
      if (*s == ',') s++;
      nq++;
    }
  }
  {
    argc = 2; w = buf;
    char *p0 = w; w += sprintf(w, "NQ") + 1; argv[argc++] = p0;
    p0 = w; w += sprintf(w, "%d", nq) + 1; argv[argc++] = p0;
    for (int i = 0; i < nq; i++) {
      p0 = w; w += sprintf(w, "QNT_NAME[%d]", i) + 1; argv[argc++] = p0;
      argv[argc++] = names[i];
      /* a unit is mandatory for names the reference does not know (src/mptrac.c:6970) */
      p0 = w; w += sprintf(w, "QNT_UNIT[%d]", i) + 1; argv[argc++] = p0;
      p0 = w; w += sprintf(w, "-") + 1; argv[argc++] = p0;
    }
    static char ov[4096];
    strncpy(ov, overrides ? overrides : "", sizeof(ov) - 1);
    for (char *tok = strtok(ov, " "); tok && argc < 510; tok = strtok(NULL, " ")) argv[argc++] = tok;
    argv[argc++] = (char *)"";  /* scan_ctl looks at argv[i + 1] up to argc - 2 */
  }
  mptrac_read_ctl("-", argc, argv, h_ctl);
  if (out) {
    out[0] = h_ctl->qnt_rp; out[1] = h_ctl->qnt_rhop; out[2] = h_ctl->qnt_m; out[3] = h_ctl->qnt_vmr;
    out[4] = h_ctl->qnt_ens;
    /* module_meteo quantities in the slot order of orc_ctl_t::qnt_meteo */
    const int mq[53] = {h_ctl->qnt_ps, h_ctl->qnt_pbl, h_ctl->qnt_p, h_ctl->qnt_t, h_ctl->qnt_rho, h_ctl->qnt_u, h_ctl->qnt_v,
                        h_ctl->qnt_w, h_ctl->qnt_vh, h_ctl->qnt_vz, h_ctl->qnt_theta, h_ctl->qnt_psat, h_ctl->qnt_psice,
                        h_ctl->qnt_zeta_d,
                        h_ctl->qnt_ts, h_ctl->qnt_zs, h_ctl->qnt_us, h_ctl->qnt_vs, h_ctl->qnt_ess, h_ctl->qnt_nss, h_ctl->qnt_shf,
                        h_ctl->qnt_lsm, h_ctl->qnt_sst, h_ctl->qnt_pt, h_ctl->qnt_tt, h_ctl->qnt_zt, h_ctl->qnt_h2ot, h_ctl->qnt_pct,
                        h_ctl->qnt_pcb, h_ctl->qnt_cl, h_ctl->qnt_plcl, h_ctl->qnt_plfc, h_ctl->qnt_pel, h_ctl->qnt_cape,
                        h_ctl->qnt_cin, h_ctl->qnt_o3c,
                        h_ctl->qnt_zg, h_ctl->qnt_pv, h_ctl->qnt_h2o, h_ctl->qnt_o3, h_ctl->qnt_lwc, h_ctl->qnt_rwc, h_ctl->qnt_iwc,
                        h_ctl->qnt_swc, h_ctl->qnt_cc,
                        h_ctl->qnt_pw, h_ctl->qnt_sh, h_ctl->qnt_rh, h_ctl->qnt_rhice, h_ctl->qnt_tvirt, h_ctl->qnt_lapse,
                        h_ctl->qnt_tdew, h_ctl->qnt_tice};
    for (int i = 0; i < 64; i++) out[5 + i] = i < 53 ? mq[i] : -1;
    out[70] = h_ctl->qnt_zeta; out[71] = h_ctl->qnt_eta;
    out[72] = h_ctl->qnt_mloss_decay; out[73] = h_ctl->qnt_loss_rate;
    out[74] = h_ctl->qnt_aoa; out[75] = h_ctl->qnt_Cccl4; out[76] = h_ctl->qnt_Cccl3f; out[77] = h_ctl->qnt_Cccl2f2;
    out[78] = h_ctl->qnt_Cn2o; out[79] = h_ctl->qnt_Csf6; out[80] = h_ctl->qnt_Cx;
  }
  return h_ctl->nq;
}

/* the reference's built-in tropopause climatology (src/mptrac.c:241-392) */
int ref_clim_tropo(int *ntime, int *nlat, double *time, double *lat, double *tropo) {
  ensure_alloc();
  clim_tropo_init(h_clim);
  *ntime = h_clim->tropo_ntime; *nlat = h_clim->tropo_nlat;
  for (int i = 0; i < h_clim->tropo_ntime; i++) time[i] = h_clim->tropo_time[i];
  for (int i = 0; i < h_clim->tropo_nlat; i++) lat[i] = h_clim->tropo_lat[i];
  for (int i = 0; i < h_clim->tropo_ntime; i++)
    for (int j = 0; j < h_clim->tropo_nlat; j++) tropo[i * h_clim->tropo_nlat + j] = h_clim->tropo[i][j];
  return 0;
}

static void fill_met(met_t *dst, const orc_met_t *src) {
  if (src->nx > EX || src->ny > EY || src->np > EP) ERRMSG("met grid exceeds the reference build's EX/EY/EP");
  dst->time = src->time; dst->coord_type = src->coord_type;
  dst->nx = src->nx; dst->ny = src->ny; dst->np = src->np; dst->npl = src->np;
  memcpy(dst->lon, src->lon, sizeof(double) * (size_t)src->nx);
  memcpy(dst->lat, src->lat, sizeof(double) * (size_t)src->ny);
  memcpy(dst->p, src->p, sizeof(double) * (size_t)src->np);
#pragma omp parallel for collapse(2)
  for (int ix = 0; ix < src->nx; ix++)
    for (int iy = 0; iy < src->ny; iy++) {
      const size_t o = ((size_t)ix * src->ny + iy) * src->np;
      memcpy(dst->u[ix][iy], src->u + o, sizeof(float) * (size_t)src->np);
      memcpy(dst->v[ix][iy], src->v + o, sizeof(float) * (size_t)src->np);
      memcpy(dst->w[ix][iy], src->w + o, sizeof(float) * (size_t)src->np);
      if (src->t) memcpy(dst->t[ix][iy], src->t + o, sizeof(float) * (size_t)src->np);
      dst->ps[ix][iy] = src->ps ? src->ps[(size_t)ix * src->ny + iy] : 0.f;
      dst->pbl[ix][iy] = src->pbl ? src->pbl[(size_t)ix * src->ny + iy] : 0.f;
    }
  {   /* the further fields module_meteo interpolates (zero where the caller has none, like a file that lacks them) */
    float (*o2[ORC_NX2])[EY] = {dst->ts, dst->zs, dst->us, dst->vs, dst->ess, dst->nss, dst->shf, dst->lsm, dst->sst, dst->pt, dst->tt,
                                dst->zt, dst->h2ot, dst->pct, dst->pcb, dst->cl, dst->plcl, dst->plfc, dst->pel, dst->cape, dst->cin,
                                dst->o3c};
    float (*o3[ORC_NX3])[EY][EP] = {dst->z, dst->pv, dst->h2o, dst->o3, dst->lwc, dst->rwc, dst->iwc, dst->swc, dst->cc};
#pragma omp parallel for collapse(2)
    for (int ix = 0; ix < src->nx; ix++)
      for (int iy = 0; iy < src->ny; iy++) {
        const size_t o = ((size_t)ix * src->ny + iy) * src->np;
        for (int f = 0; f < ORC_NX2; f++) o2[f][ix][iy] = src->x2[f] ? src->x2[f][(size_t)ix * src->ny + iy] : 0.f;
        for (int f = 0; f < ORC_NX3; f++) {
          if (src->x3[f]) memcpy(o3[f][ix][iy], src->x3[f] + o, sizeof(float) * (size_t)src->np);
          else memset(o3[f][ix][iy], 0, sizeof(float) * (size_t)src->np);
        }
      }
  }
  if (src->pl) {   /* model-level fields (ADVECT_VERT_COORD 1, 2, 3) */
    if (src->npl > EP) ERRMSG("model levels exceed the reference build's EP");
    dst->npl = src->npl;
    const float *in[6] = {src->pl, src->ul, src->vl, src->wl, src->zetal, src->zeta_dotl};
    float (*out[6])[EY][EP] = {dst->pl, dst->ul, dst->vl, dst->wl, dst->zetal, dst->zeta_dotl};
#pragma omp parallel for collapse(2)
    for (int ix = 0; ix < src->nx; ix++)
      for (int iy = 0; iy < src->ny; iy++) {
        const size_t o = ((size_t)ix * src->ny + iy) * src->npl;
        for (int f = 0; f < 6; f++)
          if (in[f]) memcpy(out[f][ix][iy], in[f] + o, sizeof(float) * (size_t)src->npl);
      }
  }
}

int ref_set_met(const orc_met_t *m0, const orc_met_t *m1) {
  ensure_alloc();
  fill_met(h_met0, m0);
  fill_met(h_met1, m1);
  return 0;
}

/* overwrite the numeric ctl fields of the path with the caller's values (quantities come from ref_read_ctl) */
static void apply_ctl(const orc_ctl_t *c) {
  h_ctl->direction = c->direction; h_ctl->met_coord_type = c->met_coord_type; h_ctl->advect = c->advect;
  h_ctl->advect_vert_coord = c->advect_vert_coord; h_ctl->rng_type = c->rng_type; h_ctl->diffusion = c->diffusion;
  h_ctl->turb_pbl_scheme = c->turb_pbl_scheme; h_ctl->nens = c->nens;
  h_ctl->mixing_nx = c->mixing_nx; h_ctl->mixing_ny = c->mixing_ny; h_ctl->mixing_nz = c->mixing_nz;
  h_ctl->t_start = c->t_start; h_ctl->t_stop = c->t_stop; h_ctl->dt_mod = c->dt_mod; h_ctl->dt_met = c->dt_met;
  h_ctl->met_utm_ref_lat = c->met_utm_ref_lat; h_ctl->sort_dt = c->sort_dt;
  h_ctl->turb_dx_pbl = c->turb_dx_pbl; h_ctl->turb_dx_trop = c->turb_dx_trop; h_ctl->turb_dx_strat = c->turb_dx_strat;
  h_ctl->turb_dz_pbl = c->turb_dz_pbl; h_ctl->turb_dz_trop = c->turb_dz_trop; h_ctl->turb_dz_strat = c->turb_dz_strat;
  h_ctl->turb_mesox = c->turb_mesox; h_ctl->turb_mesoz = c->turb_mesoz; h_ctl->turb_pbl_trans = c->turb_pbl_trans;
  h_ctl->mixing_dt = c->mixing_dt; h_ctl->mixing_trop = c->mixing_trop; h_ctl->mixing_strat = c->mixing_strat;
  h_ctl->mixing_lon0 = c->mixing_lon0; h_ctl->mixing_lon1 = c->mixing_lon1; h_ctl->mixing_lat0 = c->mixing_lat0;
  h_ctl->mixing_lat1 = c->mixing_lat1; h_ctl->mixing_z0 = c->mixing_z0; h_ctl->mixing_z1 = c->mixing_z1;
  h_ctl->met_dt_out = c->met_dt_out;
  h_ctl->conv_cape = c->conv_cape; h_ctl->conv_cin = c->conv_cin; h_ctl->conv_pbl_trans = c->conv_pbl_trans;
  h_ctl->conv_dt = c->conv_dt; h_ctl->conv_mix_pbl = c->conv_mix_pbl;
  h_ctl->tdec_trop = c->tdec_trop; h_ctl->tdec_strat = c->tdec_strat;
  h_ctl->isosurf = c->isosurf;
  h_ctl->bound_mass = c->bound_mass; h_ctl->bound_mass_trend = c->bound_mass_trend; h_ctl->bound_vmr = c->bound_vmr;
  h_ctl->bound_vmr_trend = c->bound_vmr_trend; h_ctl->bound_lat0 = c->bound_lat0; h_ctl->bound_lat1 = c->bound_lat1;
  h_ctl->bound_p0 = c->bound_p0; h_ctl->bound_p1 = c->bound_p1; h_ctl->bound_dps = c->bound_dps; h_ctl->bound_dzs = c->bound_dzs;
  h_ctl->bound_zetas = c->bound_zetas; h_ctl->bound_pbl = c->bound_pbl;
  h_ctl->chemgrid_lon0 = c->chemgrid_lon0; h_ctl->chemgrid_lon1 = c->chemgrid_lon1; h_ctl->chemgrid_lat0 = c->chemgrid_lat0;
  h_ctl->chemgrid_lat1 = c->chemgrid_lat1; h_ctl->chemgrid_z0 = c->chemgrid_z0; h_ctl->chemgrid_z1 = c->chemgrid_z1;
  h_ctl->chemgrid_nx = c->chemgrid_nx; h_ctl->chemgrid_ny = c->chemgrid_ny; h_ctl->chemgrid_nz = c->chemgrid_nz;
  h_ctl->molmass = c->molmass;
  char *names[5] = {h_ctl->clim_ccl4_timeseries, h_ctl->clim_ccl3f_timeseries, h_ctl->clim_ccl2f2_timeseries,
                    h_ctl->clim_n2o_timeseries, h_ctl->clim_sf6_timeseries};
  for (int k = 0; k < 5; k++) strcpy(names[k], (c->cts_on >> k & 1) ? "given" : "-");
}

/* the five trace-gas time series of clim_t (what mptrac_read_clim reads from the files the control file names) */
int ref_set_cts(const orc_cts_t *cts) {
  ensure_alloc();
  clim_ts_t *dst[5] = {&h_clim->ccl4, &h_clim->ccl3f, &h_clim->ccl2f2, &h_clim->n2o, &h_clim->sf6};
  for (int k = 0; k < 5; k++) {
    if (cts->n[k] > CTS) ERRMSG("time series too long for the reference build's CTS");
    dst[k]->ntime = cts->n[k];
    for (int i = 0; i < cts->n[k]; i++) { dst[k]->time[i] = cts->time[k][i]; dst[k]->vmr[i] = cts->vmr[k][i]; }
  }
  return 0;
}

static void put_atm(const orc_atm_t *a) {
  if (a->np > NP) ERRMSG("too many parcels for the reference build's NP");
  const size_t n = (size_t)a->np;
  h_atm->np = (int)a->np;
  memcpy(h_atm->time, a->time, 8 * n); memcpy(h_atm->p, a->p, 8 * n);
  memcpy(h_atm->lon, a->lon, 8 * n); memcpy(h_atm->lat, a->lat, 8 * n);
  for (int iq = 0; iq < h_ctl->nq; iq++) memcpy(h_atm->q[iq], a->q + (size_t)iq * a->q_stride, 8 * n);
  if (a->dt) memcpy(h_cache->dt, a->dt, 8 * n);
  if (a->uvwp) memcpy(h_cache->uvwp, a->uvwp, 12 * n);
  if (a->iso_var) memcpy(h_cache->iso_var, a->iso_var, 8 * n);
}

static void get_atm(orc_atm_t *a) {
  const size_t n = (size_t)a->np;
  memcpy(a->time, h_atm->time, 8 * n); memcpy(a->p, h_atm->p, 8 * n);
  memcpy(a->lon, h_atm->lon, 8 * n); memcpy(a->lat, h_atm->lat, 8 * n);
  for (int iq = 0; iq < h_ctl->nq; iq++) memcpy(a->q + (size_t)iq * a->q_stride, h_atm->q[iq], 8 * n);
  if (a->dt) memcpy(a->dt, h_cache->dt, 8 * n);
  if (a->uvwp) memcpy(a->uvwp, h_cache->uvwp, 12 * n);
  if (a->rs) memcpy(a->rs, h_cache->rs, 8 * (3 * n + 1));
  if (a->iso_var) memcpy(a->iso_var, h_cache->iso_var, 8 * n);
}

/* what: 0 mptrac_run_timestep, 1 timesteps, 2 position, 3 advect, 4 diff_turb, 5 diff_meso, 6 sedi,
 *       7 sort, 8 mixing, 9 meteo, 10 advect_init.  Met must have been set with ref_set_met, ctl with ref_read_ctl. */
int ref_run(const orc_ctl_t *ctl, orc_atm_t *atm, double t, int what, int nsteps, uint64_t *ctr) {
  ensure_alloc();
  apply_ctl(ctl);
  if (h_clim->tropo_ntime == 0) clim_tropo_init(h_clim);
  put_atm(atm);
  /* ISOSURF 4: module_isosurf_init appends the balloon file (ctl->balloon, set through ref_read_ctl's overrides) to the
     cache's time series without resetting its length */
  if (what == 0 || what == 13) h_cache->iso_n = 0;
  rng_ctr = *ctr;
  switch (what) {
    case 0:
      for (int s = 0; s < nsteps; s++)
        mptrac_run_timestep(h_ctl, h_cache, h_clim, &h_met0, &h_met1, h_atm, h_depo, t + s * ctl->direction * ctl->dt_mod, h_dd);
      break;
    case 1: module_timesteps(h_ctl, h_cache, h_met0, h_atm, t); break;
    case 2: module_position(h_cache, h_met0, h_met1, h_atm); break;
    case 3: module_advect(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 4: module_diff_turb(h_ctl, h_cache, h_clim, h_met0, h_met1, h_atm); break;
    case 5: module_diff_meso(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 6: module_sedi(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 7: module_sort(h_ctl, h_met0, h_atm); break;
    case 8: module_mixing(h_ctl, h_clim, h_atm, t); break;
    case 9: module_meteo(h_ctl, h_cache, h_clim, h_met0, h_met1, h_atm); break;
    case 10: module_advect_init(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 11: module_convection(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 12: module_decay(h_ctl, h_cache, h_clim, h_atm); break;
    case 17: module_chem_grid(h_ctl, h_met0, h_met1, h_atm, t); break;
    case 16: module_bound_cond(h_ctl, h_cache, h_clim, h_met0, h_met1, h_atm); break;
    case 15: module_diff_pbl(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 13: module_isosurf_init(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    case 14: module_isosurf(h_ctl, h_cache, h_met0, h_met1, h_atm); break;
    default: return 1;
  }
  *ctr = rng_ctr;
  get_atm(atm);
  return 0;
}

double ref_sedi(double p, double T, double rp, double rhop) { return sedi(p, T, rp, rhop); }

int ref_module_rng(double *rs, int64_t n, int method, uint64_t *ctr) {
  ensure_alloc();
  h_ctl->rng_type = 1;
  rng_ctr = *ctr;
  module_rng(h_ctl, rs, (size_t)n, method);
  *ctr = rng_ctr;
  return 0;
}

/* physics-group seconds spent so far (the reference's own timers are file-static; we time outside) */
double ref_wtime(void) { return omp_get_wtime(); }

/* ---- file readers, used only by tests/golden/make_golden.py to turn the reference's own test data into
 *      compact fixtures: the reference reads and pre-processes the files, we just copy fields out ---- */
int ref_read_met(const char *filename, int slot) {
  ensure_alloc();
  if (h_clim->tropo_ntime == 0) clim_tropo_init(h_clim);
  return mptrac_read_met(filename, h_ctl, h_clim, slot ? h_met1 : h_met0, h_dd);
}

int ref_met_dims(int slot, int *nx, int *ny, int *np, int *coord_type, double *time) {
  const met_t *m = slot ? h_met1 : h_met0;
  *nx = m->nx; *ny = m->ny; *np = m->np; *coord_type = m->coord_type; *time = m->time;
  return 0;
}

int ref_get_met(int slot, double *lon, double *lat, double *p, float *u, float *v, float *w, float *t, float *ps, float *pbl) {
  met_t *m = slot ? h_met1 : h_met0;
  memcpy(lon, m->lon, 8 * (size_t)m->nx); memcpy(lat, m->lat, 8 * (size_t)m->ny); memcpy(p, m->p, 8 * (size_t)m->np);
  for (int ix = 0; ix < m->nx; ix++)
    for (int iy = 0; iy < m->ny; iy++) {
      const size_t o = ((size_t)ix * m->ny + iy) * m->np;
      memcpy(u + o, m->u[ix][iy], 4 * (size_t)m->np); memcpy(v + o, m->v[ix][iy], 4 * (size_t)m->np);
      memcpy(w + o, m->w[ix][iy], 4 * (size_t)m->np); memcpy(t + o, m->t[ix][iy], 4 * (size_t)m->np);
      ps[(size_t)ix * m->ny + iy] = m->ps[ix][iy]; pbl[(size_t)ix * m->ny + iy] = m->pbl[ix][iy];
    }
  return 0;
}

int ref_read_atm(const char *filename) {
  ensure_alloc();
  if (!mptrac_read_atm(filename, h_ctl, h_atm)) return -1;
  return h_atm->np;
}

int ref_get_atm(double *time, double *p, double *lon, double *lat) {
  const size_t n = (size_t)h_atm->np;
  memcpy(time, h_atm->time, 8 * n); memcpy(p, h_atm->p, 8 * n); memcpy(lon, h_atm->lon, 8 * n); memcpy(lat, h_atm->lat, 8 * n);
  return 0;
}
