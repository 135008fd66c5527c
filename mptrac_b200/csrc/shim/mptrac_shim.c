/*
 * mptrac_shim.c -- the reference-facing side of the drop-in boundary.
 *
 * Built against the reference's OWN header (mptrac.h is found through -I<reference>/src at build time; it
 * is never copied into this repository) and placed in front of the reference's libmptrac.so (link order or
 * LD_PRELOAD).  It defines the symbols the unmodified `trac` driver and the library itself reach through the
 * PLT -- mptrac_run_timestep, mptrac_update_device, mptrac_update_host, mptrac_free -- and forwards them to
 * the C ABI of libmptrac_b200.so.  Host language = C, like the reference.
 *
 *   mptrac_update_device (src/mptrac.c:8005)  -> mpb_set_ctl / set_clim_tropo / set_met / set_atm / set_uvwp
 *   mptrac_update_host   (src/mptrac.c:8061)  -> mpb_get_atm / get_uvwp / get_dt
 *   mptrac_run_timestep  (src/mptrac.c:7851)  -> mpb_run_timestep (module_meteo included when all its quantities derive
 *                                               from T, u, v, w, ps, pbl), or -- when the control file enables modules that
 *                                               are not on the device path -- mpb_run_modules segments with the
 *                                               reference's own CPU modules called in between (hybrid mode)
 *   mptrac_free          (src/mptrac.c:6377)  -> mpb_destroy, then the reference's own mptrac_free
 *
 * Errors follow the reference's convention: message + exit(EXIT_FAILURE) (ERRMSG, src/mptrac.h:2406-2410).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <time.h>

#include "mptrac.h"
#include "mptrac_b200.h"

static mpb_team *g_ctx;      /* the device side: one context per GPU behind one handle (a team of one by default) */
static const met_t *g_slot[2];    /* host met_t mirrored by device slot 0 / 1 */
static int g_np = -1;             /* parcels on the device */
static int g_dev_newer;           /* device parcels are newer than the host copy */
static int g_verbose = -1;

#define MPB(call)                                                        \
  do {                                                                   \
    if ((call) != 0) ERRMSG("mptrac_b200: %s", mpb_last_error());         \
  } while (0)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* The device list of this run: MPTRAC_B200_DEVICES="0,1,2,3" or "0-7" (one simulation over several GPUs), else
 * MPTRAC_B200_DEVICE=n, else GPU 0 (the reference selects ONE device per MPI task, src/trac.c:75-80). */
static int device_list(int *devs) {
  int ndev = 0;
  const char *list = getenv("MPTRAC_B200_DEVICES");
  if (list && *list) {
    const char *q = list;
    while (*q && ndev < MPB_MAX_RANKS) {
      char *end;
      long a = strtol(q, &end, 10), b = a;
      if (end == q) return -1;
      if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); if (end == q || b < a) return -1; }
      for (long d = a; d <= b && ndev < MPB_MAX_RANKS; d++) devs[ndev++] = (int) d;
      if (*end && *end != ',') return -1;
      q = (*end == ',') ? end + 1 : end;
    }
  } else {
    const char *e = getenv("MPTRAC_B200_DEVICE");
    devs[ndev++] = e ? atoi(e) : 0;
  }
  return ndev;
}

/* Creating a CUDA context takes a few hundred milliseconds on a cold process -- as long as `trac` needs to read its control
 * file, parcels and first met files.  A helper thread does it while the driver reads (started when the shim is loaded). */
static void *warmup_thread(void *arg) {
  (void) arg;
  int devs[MPB_MAX_RANKS];
  const int n = device_list(devs);
  for (int i = 0; i < n; i++) mpb_warmup(devs[i]);     /* errors surface later, where they can be reported properly */
  return NULL;
}
__attribute__((constructor)) static void shim_loaded(void) {
  if (getenv("MPTRAC_B200_NO_WARMUP")) return;
  pthread_t th;
  if (pthread_create(&th, NULL, warmup_thread, NULL) == 0) pthread_detach(th);
}

static int verbose(void) {
  if (g_verbose < 0) g_verbose = getenv("MPTRAC_B200_VERBOSE") ? atoi(getenv("MPTRAC_B200_VERBOSE")) : 0;
  return g_verbose;
}

/* The reference draws every random number of a process from ONE file-static counter (rng_ctr, src/mptrac.c:35) that
 * keeps counting across the directories of a `trac` dirlist (mptrac_free / mptrac_alloc per directory, src/trac.c:98-185).
 * The device counter lives in the context, so it is carried over here whenever a context ends. */
static unsigned long long g_rng_ctr;
static const clim_t *g_last_clim;      /* last climatology / cache handed to mptrac_update_device: a re-created context */
static const cache_t *g_last_cache;    /* (more parcels than before) gets them again */

/* The shim reads and writes the driver's structs in place, so it must have been compiled with the dimensions of the
 * library behind it.  A libmptrac.so that exports `const long mptrac_ref_layout[8]` = {NP, NQ, EX, EY, EP, sizeof(atm_t),
 * sizeof(met_t), sizeof(cache_t)} (one extra translation unit) is verified here; one without the symbol
 * cannot be checked: build both with the same -DNP/-DNQ/-DEX/-DEY/-DEP. */
static void check_layout(void) {
  static int done;
  if (done) return;
  done = 1;
  const long *ref = (const long *) dlsym(RTLD_DEFAULT, "mptrac_ref_layout");
  if (!ref) {
    if (verbose()) printf("mptrac_b200: the library behind the shim does not export mptrac_ref_layout: struct layout unchecked\n");
    return;
  }
  const long mine[8] = {(long) NP, (long) NQ, (long) EX, (long) EY, (long) EP, (long) sizeof(atm_t), (long) sizeof(met_t), (long) sizeof(cache_t)};
  const char *what[8] = {"NP", "NQ", "EX", "EY", "EP", "sizeof(atm_t)", "sizeof(met_t)", "sizeof(cache_t)"};
  for (int i = 0; i < 8; i++)
    if (ref[i] != mine[i])
      ERRMSG("mptrac_b200: shim compiled with %s = %ld, libmptrac with %ld -- rebuild both with the same MPTRAC_DEFINES", what[i], mine[i], ref[i]);
}

/* returns 1 when a context was (re-)created */
static int ensure_ctx(const ctl_t *ctl, int np) {
  if (g_ctx && np <= g_np) return 0;
  check_layout();
  if (g_ctx) { g_rng_ctr = mpb_team_get_rng_ctr(g_ctx); MPB(mpb_team_destroy(g_ctx)); }
  /* MPTRAC_B200_DEVICES="0,1,2,3" or "0-7": the parcels are cut into contiguous index ranges over these GPUs, the met data
     is packed once and copied device to device, module_mixing exchanges its box records through peer memory -- all behind
     this one host thread (the reference selects ONE device per MPI task, src/trac.c:75-80).  MPTRAC_B200_DEVICE=n: one GPU. */
  int devs[MPB_MAX_RANKS];
  const int ndev = device_list(devs);
  if (ndev < 1) ERRMSG("mptrac_b200: cannot parse MPTRAC_B200_DEVICES=%s", getenv("MPTRAC_B200_DEVICES"));
  const double t0 = now_s();
  MPB(mpb_team_create(&g_ctx, ndev, devs, np, ctl->nq));
  if (verbose()) printf("mptrac_b200: context(s) created in %.3f s\n", now_s() - t0);
  MPB(mpb_team_set_rng_ctr(g_ctx, g_rng_ctr));
  g_slot[0] = g_slot[1] = NULL;
  g_np = np;
  if (verbose()) {
    printf("mptrac_b200: device context for %d parcels, %d quantities on GPU", np, ctl->nq);
    for (int i = 0; i < ndev; i++) printf("%s%d", i ? "," : " ", devs[i]);
    printf("\n");
  }
  return 1;
}

/* module_meteo quantities that read further met fields (zg, pv, h2o, o3, cloud and surface fields, rh ...) are computed on
 * the device too; the fields they need are uploaded with each met level.  MPTRAC_B200_DEVICE_METEO_FIELDS=0 keeps them on
 * the reference's CPU code (hybrid mode: a parcel round trip per step). */
static int device_meteo_fields(void) {
  static int on = -1;
  if (on < 0) on = getenv("MPTRAC_B200_DEVICE_METEO_FIELDS") ? atoi(getenv("MPTRAC_B200_DEVICE_METEO_FIELDS")) : 1;
  return on;
}
static int g_fields, g_need2[MPB_NX2], g_need3[MPB_NX3];
static int meteo_needs_host(const ctl_t *c);

/* module_diff_pbl, module_convection, module_isosurf and -- when nothing else keeps the tail of the step on the host --
 * module_bound_cond and module_decay run on the device at their places in the dispatcher.  MPTRAC_B200_DEVICE_MODULES=0
 * delegates them to the reference's CPU code instead (hybrid mode). */
static int device_modules(void) {
  static int on = -1;
  if (on < 0) on = getenv("MPTRAC_B200_DEVICE_MODULES") ? atoi(getenv("MPTRAC_B200_DEVICE_MODULES")) : 1;
  return on;
}

/* modules after the final position check that only exist on the host (chemistry, deposition): with any of them on,
 * everything from module_meteo's successor to the end of the step runs through the reference's code, in its order */
static int tail_needs_host(const ctl_t *c) {
  const int wet = (c->wet_depo_ic_a > 0 || c->wet_depo_ic_h[0] > 0) && (c->wet_depo_bc_a > 0 || c->wet_depo_bc_h[0] > 0);
  return c->oh_chem_reaction != 0 || c->h2o2_chem_reaction != 0 || c->kpp_chem || c->tracer_chem || c->radio_decay || c->radio_depo ||
    wet || c->dry_depo_vdep > 0;
}

static int module_timers(void) {
  static int on = -1;
  if (on < 0) on = getenv("MPTRAC_B200_MODULE_TIMERS") ? atoi(getenv("MPTRAC_B200_MODULE_TIMERS")) : 0;
  return on;
}

static int g_levels;   /* the control file advects on model levels: the met uploads carry pl, ul, vl, wl, zetal, zeta_dotl */

static void put_ctl(const ctl_t *c) {
  mpb_ctl_t k;
  memset(&k, 0, sizeof(k));
  k.direction = c->direction; k.met_coord_type = c->met_coord_type; k.advect = c->advect;
  k.advect_vert_coord = c->advect_vert_coord; k.rng_type = c->rng_type; k.diffusion = c->diffusion;
  k.turb_pbl_scheme = c->turb_pbl_scheme; k.nq = c->nq; k.qnt_rp = c->qnt_rp; k.qnt_rhop = c->qnt_rhop;
  k.qnt_ens = c->qnt_ens; k.nens = c->nens;
  k.mixing_nx = c->mixing_nx; k.mixing_ny = c->mixing_ny; k.mixing_nz = c->mixing_nz;
  const int mq[MPB_MIX_MAXQ] = {                      /* order of src/mptrac.c:5222-5230 */
    c->qnt_m, c->qnt_vmr, c->qnt_Ch2o, c->qnt_Co3, c->qnt_Cco, c->qnt_Coh, c->qnt_Ch, c->qnt_Cho2,
    c->qnt_Ch2o2, c->qnt_Co1d, c->qnt_Co3p, c->qnt_Cccl4, c->qnt_Cccl3f, c->qnt_Cccl2f2, c->qnt_Cn2o,
    c->qnt_Csf6, c->qnt_aoa, c->qnt_Arn222, c->qnt_Apb210, c->qnt_Abe7, c->qnt_Acs137, c->qnt_Ai131, c->qnt_Axe133};
  k.n_mix_qnt = MPB_MIX_MAXQ;
  for (int i = 0; i < MPB_MIX_MAXQ; i++) k.mix_qnt[i] = mq[i];
  k.t_start = c->t_start; k.t_stop = c->t_stop; k.dt_mod = c->dt_mod; k.dt_met = c->dt_met;
  k.met_utm_ref_lat = c->met_utm_ref_lat; k.sort_dt = c->sort_dt;
  k.turb_dx_pbl = c->turb_dx_pbl; k.turb_dx_trop = c->turb_dx_trop; k.turb_dx_strat = c->turb_dx_strat;
  k.turb_dz_pbl = c->turb_dz_pbl; k.turb_dz_trop = c->turb_dz_trop; k.turb_dz_strat = c->turb_dz_strat;
  k.turb_mesox = c->turb_mesox; k.turb_mesoz = c->turb_mesoz; k.turb_pbl_trans = c->turb_pbl_trans;
  k.mixing_dt = c->mixing_dt; k.mixing_trop = c->mixing_trop; k.mixing_strat = c->mixing_strat;
  k.mixing_lon0 = c->mixing_lon0; k.mixing_lon1 = c->mixing_lon1; k.mixing_lat0 = c->mixing_lat0;
  k.mixing_lat1 = c->mixing_lat1; k.mixing_z0 = c->mixing_z0; k.mixing_z1 = c->mixing_z1;
  k.met_dt_out = c->met_dt_out;
  for (int i = 0; i < MPB_METEO_SLOTS; i++) k.qnt_meteo[i] = -1;
  k.qnt_meteo[MPB_Q_PS] = c->qnt_ps; k.qnt_meteo[MPB_Q_PBL] = c->qnt_pbl; k.qnt_meteo[MPB_Q_P] = c->qnt_p;
  k.qnt_meteo[MPB_Q_T] = c->qnt_t; k.qnt_meteo[MPB_Q_RHO] = c->qnt_rho; k.qnt_meteo[MPB_Q_U] = c->qnt_u;
  k.qnt_meteo[MPB_Q_V] = c->qnt_v; k.qnt_meteo[MPB_Q_W] = c->qnt_w; k.qnt_meteo[MPB_Q_VH] = c->qnt_vh;
  k.qnt_meteo[MPB_Q_VZ] = c->qnt_vz; k.qnt_meteo[MPB_Q_THETA] = c->qnt_theta; k.qnt_meteo[MPB_Q_PSAT] = c->qnt_psat;
  k.qnt_meteo[MPB_Q_PSICE] = c->qnt_psice; k.qnt_meteo[MPB_Q_ZETA_D] = c->qnt_zeta_d;
  k.qnt_zeta = c->qnt_zeta; k.qnt_eta = c->qnt_eta;
  /* the rank-4 modules are switched off on the device side here and switched on below when they are routed to it */
  k.conv_cape = k.conv_cin = k.conv_dt = -999; k.conv_pbl_trans = 0; k.conv_mix_pbl = 0;
  k.tdec_trop = k.tdec_strat = 0;
  k.qnt_m = k.qnt_vmr = k.qnt_mloss_decay = k.qnt_loss_rate = -1;
  k.isosurf = 0;    /* module_isosurf likewise (iso_cpu below) */
  k.bound_lat0 = k.bound_lat1 = k.bound_p0 = k.bound_p1 = -999;   /* module_bound_cond likewise (bound_cpu below) */
  k.bound_mass = k.bound_vmr = k.bound_dps = k.bound_dzs = k.bound_zetas = -999;
  k.qnt_aoa = -1; k.cts_on = 0;
  for (int i = 0; i < 5; i++) k.qnt_cts[i] = -1;
  k.chemgrid = 0; k.qnt_Cx = -1;   /* module_chem_grid feeds the chemistry modules, which are host-only: stays on the host */
  g_levels = c->advect_vert_coord != 0;
  g_fields = device_meteo_fields() || device_modules();
  memset(g_need2, 0, sizeof(g_need2)); memset(g_need3, 0, sizeof(g_need3));
  if (device_modules()) {
    k.conv_cape = c->conv_cape; k.conv_cin = c->conv_cin; k.conv_dt = c->conv_dt; k.conv_pbl_trans = c->conv_pbl_trans;
    k.conv_mix_pbl = c->conv_mix_pbl;
    k.isosurf = c->isosurf;
    if (c->conv_cape >= 0) g_need2[MPB_F2_CAPE] = g_need2[MPB_F2_CIN] = g_need2[MPB_F2_PEL] = 1;
    if (c->diffusion && c->turb_pbl_scheme == 1) g_need2[MPB_F2_ESS] = g_need2[MPB_F2_NSS] = g_need2[MPB_F2_SHF] = g_need3[MPB_F3_H2O] = 1;
    if (!tail_needs_host(c) && !meteo_needs_host(c)) {
      k.tdec_trop = c->tdec_trop; k.tdec_strat = c->tdec_strat;
      k.qnt_m = c->qnt_m; k.qnt_vmr = c->qnt_vmr; k.qnt_mloss_decay = c->qnt_mloss_decay; k.qnt_loss_rate = c->qnt_loss_rate;
      k.bound_mass = c->bound_mass; k.bound_mass_trend = c->bound_mass_trend; k.bound_vmr = c->bound_vmr;
      k.bound_vmr_trend = c->bound_vmr_trend; k.bound_lat0 = c->bound_lat0; k.bound_lat1 = c->bound_lat1; k.bound_p0 = c->bound_p0;
      k.bound_p1 = c->bound_p1; k.bound_dps = c->bound_dps; k.bound_dzs = c->bound_dzs; k.bound_zetas = c->bound_zetas;
      k.bound_pbl = c->bound_pbl; k.qnt_aoa = c->qnt_aoa;
      const int qc[5] = {c->qnt_Cccl4, c->qnt_Cccl3f, c->qnt_Cccl2f2, c->qnt_Cn2o, c->qnt_Csf6};
      const char *nm[5] = {c->clim_ccl4_timeseries, c->clim_ccl3f_timeseries, c->clim_ccl2f2_timeseries, c->clim_n2o_timeseries,
                           c->clim_sf6_timeseries};
      for (int i = 0; i < 5; i++) { k.qnt_cts[i] = qc[i]; if (nm[i][0] != '-') k.cts_on |= 1 << i; }
    }
  }
  if (device_meteo_fields()) {
    const int q2[MPB_NX2] = {c->qnt_ts, c->qnt_zs, c->qnt_us, c->qnt_vs, c->qnt_ess, c->qnt_nss, c->qnt_shf, c->qnt_lsm, c->qnt_sst,
                             c->qnt_pt, c->qnt_tt, c->qnt_zt, c->qnt_h2ot, c->qnt_pct, c->qnt_pcb, c->qnt_cl, c->qnt_plcl, c->qnt_plfc,
                             c->qnt_pel, c->qnt_cape, c->qnt_cin, c->qnt_o3c};
    const int q3[MPB_NX3] = {c->qnt_zg, c->qnt_pv, c->qnt_h2o, c->qnt_o3, c->qnt_lwc, c->qnt_rwc, c->qnt_iwc, c->qnt_swc, c->qnt_cc};
    const int qm[8] = {c->qnt_pw, c->qnt_sh, c->qnt_rh, c->qnt_rhice, c->qnt_tvirt, c->qnt_lapse, c->qnt_tdew, c->qnt_tice};
    for (int f = 0; f < MPB_NX2; f++) { k.qnt_meteo[MPB_Q_TS + f] = q2[f]; g_need2[f] |= q2[f] >= 0; }
    for (int f = 0; f < MPB_NX3; f++) { k.qnt_meteo[MPB_Q_ZG + f] = q3[f]; g_need3[f] |= q3[f] >= 0; }
    for (int i = 0; i < 8; i++) { k.qnt_meteo[MPB_Q_PW + i] = qm[i]; if (qm[i] >= 0) g_need3[MPB_F3_H2O] = 1; }
  }
  MPB(mpb_team_set_ctl(g_ctx, &k));
}

/* Does module_meteo (src/mptrac.c:5062-5165) set any quantity of this control file that the device cannot compute
 * (anything that needs met fields beyond T, u, v, w, ps, pbl, or a climatology)?  Then it runs on the host. */
static int meteo_needs_host(const ctl_t *c) {
  static int force = -1;   /* MPTRAC_B200_HOST_METEO=1 keeps module_meteo on the reference's CPU code (hybrid mode) */
  if (force < 0) force = getenv("MPTRAC_B200_HOST_METEO") ? atoi(getenv("MPTRAC_B200_HOST_METEO")) : 0;
  if (force) return 1;
  if (device_meteo_fields()) {
    const int clim_q[] = {c->qnt_hno3, c->qnt_oh, c->qnt_h2o2, c->qnt_ho2, c->qnt_o1d, c->qnt_tnat, c->qnt_tsts};
    for (size_t i = 0; i < sizeof(clim_q) / sizeof(clim_q[0]); i++) if (clim_q[i] >= 0) return 1;
    return 0;
  }
  const int other[] = {
    c->qnt_ts, c->qnt_zs, c->qnt_us, c->qnt_vs, c->qnt_ess, c->qnt_nss, c->qnt_shf, c->qnt_lsm, c->qnt_sst, c->qnt_pt,
    c->qnt_tt, c->qnt_zt, c->qnt_h2ot, c->qnt_zg, c->qnt_h2o, c->qnt_o3, c->qnt_lwc, c->qnt_rwc, c->qnt_iwc, c->qnt_swc,
    c->qnt_cc, c->qnt_pct, c->qnt_pcb, c->qnt_cl, c->qnt_plcl, c->qnt_plfc, c->qnt_pel, c->qnt_cape, c->qnt_cin, c->qnt_o3c,
    c->qnt_hno3, c->qnt_oh, c->qnt_h2o2, c->qnt_ho2, c->qnt_o1d, c->qnt_pw, c->qnt_sh, c->qnt_rh, c->qnt_rhice,
    c->qnt_tvirt, c->qnt_lapse, c->qnt_pv, c->qnt_tdew, c->qnt_tice, c->qnt_tnat, c->qnt_tsts};
  for (size_t i = 0; i < sizeof(other) / sizeof(other[0]); i++) if (other[i] >= 0) return 1;
  return 0;
}

static void put_met(met_t *m) {
  int s = -1;
  for (int i = 0; i < 2; i++) if (g_slot[i] == m) s = i;
  if (s < 0) s = (g_slot[0] == NULL) ? 0 : (g_slot[1] == NULL ? 1 : 0);
  mpb_met_view_t v;
  memset(&v, 0, sizeof(v));
  v.time = m->time; v.coord_type = m->coord_type; v.nx = m->nx; v.ny = m->ny; v.np = m->np;
  v.lon = m->lon; v.lat = m->lat; v.p = m->p;
  v.u = &m->u[0][0][0]; v.v = &m->v[0][0][0]; v.w = &m->w[0][0][0]; v.t = &m->t[0][0][0];
  v.ps = &m->ps[0][0]; v.pbl = &m->pbl[0][0];
  v.sx = (int64_t) EY * EP; v.sy = EP; v.sx2 = EY;
  v.npl = 0; v._pad = 0; v.pl = v.ul = v.vl = v.wl = v.zetal = v.zeta_dotl = NULL;
  v.sxl = (int64_t) EY * EP; v.syl = EP;
  if (g_levels) {
    v.npl = m->npl;
    v.pl = &m->pl[0][0][0]; v.ul = &m->ul[0][0][0]; v.vl = &m->vl[0][0][0]; v.wl = &m->wl[0][0][0];
    v.zetal = &m->zetal[0][0][0]; v.zeta_dotl = &m->zeta_dotl[0][0][0];
  }
  if (g_fields) {   /* the further fields the control file's module_meteo quantities read (device_meteo_fields()) */
    const float *f2[MPB_NX2] = {&m->ts[0][0], &m->zs[0][0], &m->us[0][0], &m->vs[0][0], &m->ess[0][0], &m->nss[0][0], &m->shf[0][0],
                                &m->lsm[0][0], &m->sst[0][0], &m->pt[0][0], &m->tt[0][0], &m->zt[0][0], &m->h2ot[0][0], &m->pct[0][0],
                                &m->pcb[0][0], &m->cl[0][0], &m->plcl[0][0], &m->plfc[0][0], &m->pel[0][0], &m->cape[0][0],
                                &m->cin[0][0], &m->o3c[0][0]};
    const float *f3[MPB_NX3] = {&m->z[0][0][0], &m->pv[0][0][0], &m->h2o[0][0][0], &m->o3[0][0][0], &m->lwc[0][0][0],
                                &m->rwc[0][0][0], &m->iwc[0][0][0], &m->swc[0][0][0], &m->cc[0][0][0]};
    for (int f = 0; f < MPB_NX2; f++) if (g_need2[f]) v.x2[f] = f2[f];
    for (int f = 0; f < MPB_NX3; f++) if (g_need3[f]) v.x3[f] = f3[f];
  }
  const double t0 = now_s();
  MPB(mpb_team_set_met(g_ctx, s, &v));
  g_slot[s] = m;
  if (verbose()) printf("mptrac_b200: met level t=%.0f (%d x %d x %d) -> device slot %d in %.3f s\n", m->time, m->nx, m->ny, m->np, s, now_s() - t0);
}

static void put_atm(const atm_t *atm) {
  const double t0 = now_s();
  MPB(mpb_team_set_atm(g_ctx, atm->np, atm->time, atm->p, atm->lon, atm->lat, &atm->q[0][0], NP));
  g_dev_newer = 0;
  if (verbose() > 1) { MPB(mpb_team_sync(g_ctx)); printf("mptrac_b200: %d parcels -> device in %.3f s\n", atm->np, now_s() - t0); }
}

static void get_atm(atm_t *atm) {
  if (!g_ctx || !g_dev_newer) return;
  MPB(mpb_team_get_atm(g_ctx, atm->time, atm->p, atm->lon, atm->lat, &atm->q[0][0], NP));
  g_dev_newer = 0;
}

/* ------------------------------------------------------------------------------------------------ */

void mptrac_update_device(const ctl_t *ctl, const cache_t *cache, const clim_t *clim, met_t **met0, met_t **met1,
                          const atm_t *atm) {
  SELECT_TIMER("UPDATE_DEVICE", "MEMORY");
  static const ctl_t *last_ctl;
  if (ctl) last_ctl = ctl;
  if (clim) g_last_clim = clim;
  if (cache) g_last_cache = cache;
  if (atm) {
    if (!last_ctl) ERRMSG("mptrac_b200: mptrac_update_device(atm) before the control parameters are known");
    if (ensure_ctx(last_ctl, atm->np)) {   /* a fresh context knows nothing yet */
      if (!clim) clim = g_last_clim;
      if (!cache) cache = g_last_cache;
    }
  }
  if (!g_ctx) return;              /* nothing to mirror yet (e.g. tools that never step) */
  if (ctl) put_ctl(ctl);
  else if (atm && last_ctl) put_ctl(last_ctl);
  if (clim && clim->tropo_ntime > 0)
    MPB(mpb_team_set_clim_tropo(g_ctx, clim->tropo_ntime, clim->tropo_nlat, clim->tropo_time, clim->tropo_lat, &clim->tropo[0][0]));
  if (clim && device_modules()) {
    const clim_ts_t *ts[5] = {&clim->ccl4, &clim->ccl3f, &clim->ccl2f2, &clim->n2o, &clim->sf6};
    for (int i = 0; i < 5; i++)
      if (ts[i]->ntime > 0) MPB(mpb_team_set_clim_ts(g_ctx, i, ts[i]->ntime, ts[i]->time, ts[i]->vmr));
  }
  if (met0 && *met0) put_met(*met0);
  if (met1 && *met1) put_met(*met1);
  if (atm) put_atm(atm);
  if (cache && g_np >= 0 && mpb_team_get_np(g_ctx) > 0) MPB(mpb_team_set_uvwp(g_ctx, &cache->uvwp[0][0]));
}

void mptrac_update_host(const ctl_t *ctl, const cache_t *cache, const clim_t *clim, met_t **met0, met_t **met1,
                        const atm_t *atm) {
  (void) ctl; (void) clim; (void) met0; (void) met1;   /* never modified on the device */
  SELECT_TIMER("UPDATE_HOST", "MEMORY");
  if (!g_ctx) return;
  if (atm) get_atm((atm_t *) atm);
  if (cache && mpb_team_get_np(g_ctx) > 0) {
    MPB(mpb_team_get_uvwp(g_ctx, (float *) &cache->uvwp[0][0]));
    MPB(mpb_team_get_dt(g_ctx, (double *) cache->dt));
  }
}

void mptrac_free(ctl_t *ctl, cache_t *cache, clim_t *clim, met_t *met0, met_t *met1, atm_t *atm, depo_t *depo, dd_t *dd) {
  if (g_ctx) {
    if (verbose()) printf("mptrac_b200: %lld kernel launches\n", (long long) mpb_team_launch_count(g_ctx));
    g_rng_ctr = mpb_team_get_rng_ctr(g_ctx);     /* the next directory of the dirlist continues the stream */
    MPB(mpb_team_destroy(g_ctx));
    g_ctx = NULL; g_np = -1;
  }
  if (clim == g_last_clim) g_last_clim = NULL;
  if (cache == g_last_cache) g_last_cache = NULL;
  void (*real)(ctl_t *, cache_t *, clim_t *, met_t *, met_t *, atm_t *, depo_t *, dd_t *) = dlsym(RTLD_NEXT, "mptrac_free");
  if (real) real(ctl, cache, clim, met0, met1, atm, depo, dd);
}

/* bring device slots in line with the (possibly swapped) host pointers, src/mptrac.c:6489-6491 */
static void align_met(met_t *m0, met_t *m1) {
  if (g_slot[0] == m1 && g_slot[1] == m0) {
    MPB(mpb_team_swap_met(g_ctx));
    const met_t *t = g_slot[0]; g_slot[0] = g_slot[1]; g_slot[1] = t;
  }
  if (g_slot[0] != m0) put_met(m0), align_met(m0, m1);
  if (g_slot[1] != m1) put_met(m1);
}

/* The reference draws all random numbers from ONE counter (file-static rng_ctr, src/mptrac.c:35) in module order.  In
 * hybrid mode some draws happen on the device (diff_turb, diff_meso) and some inside reference modules on the host
 * (convection, diff_pbl), whose counter the shim cannot set -- but it can advance it: before a host module that draws,
 * the host counter is brought up to the device's by drawing (and discarding) the difference; afterwards the device
 * counter is set past the host module's draws.  The stream every module sees is then the reference's. */
static unsigned long long g_host_rng;    /* what the reference's static counter is right now */

static void host_rng_catch_up(const ctl_t *ctl, cache_t *cache) {
  const unsigned long long want = mpb_team_get_rng_ctr(g_ctx);
  while (g_host_rng < want) {
    unsigned long long n = want - g_host_rng;              /* module_rng(n - 1) consumes n counters */
    if (n > 3ull * NP + 1ull) n = 3ull * NP + 1ull;        /* capacity of cache->rs */
    module_rng(ctl, cache->rs, (size_t) (n - 1), 0);
    g_host_rng += n;
  }
}

static void host_rng_drew(unsigned long long n) {
  g_host_rng += n;
  MPB(mpb_team_set_rng_ctr(g_ctx, g_host_rng));
}

/* run a reference CPU module with the host copy current; the device copy is refreshed afterwards */
#define ON_HOST(stmt)                                                         \
  do {                                                                        \
    get_atm(atm);                                                             \
    MPB(mpb_team_get_uvwp(g_ctx, &cache->uvwp[0][0]));                             \
    MPB(mpb_team_get_dt(g_ctx, cache->dt));                                        \
    stmt;                                                                     \
    host_dirty = 1;                                                           \
  } while (0)

#define FLUSH_HOST()                                                          \
  do {                                                                        \
    if (host_dirty) { put_atm(atm); MPB(mpb_team_set_uvwp(g_ctx, &cache->uvwp[0][0])); host_dirty = 0; } \
  } while (0)

void mptrac_run_timestep(ctl_t *ctl, cache_t *cache, clim_t *clim, met_t **met0, met_t **met1, atm_t *atm,
                         depo_t *depo, double t, dd_t *dd) {
  (void) dd;
  if (!g_ctx) ERRMSG("mptrac_b200: mptrac_run_timestep before mptrac_init / mptrac_update_device");
  if (ctl->rng_type != 1)
    ERRMSG("mptrac_b200: only RNG_TYPE 1 runs on the device");
  SELECT_TIMER("MODULE_B200_STEP", "PHYSICS");
  align_met(*met0, *met1);
  int host_dirty = 0;

  /* which reference modules does this control file enable that are NOT on the device path?
     (conditions copied in meaning from the dispatcher, src/mptrac.c:7863-8000) */
  const int init_cpu = (t == ctl->t_start) && ((ctl->isosurf >= 1 && ctl->isosurf <= 4) || 1 /* chem_init is unconditional */);
  const int dm = device_modules();                                          /* route the rank-4 modules to the device */
  const int dm_tail = dm && !tail_needs_host(ctl) && !meteo_needs_host(ctl);   /* ... bound_cond and decay among them (put_ctl) */
  const int pbl_cpu = !dm && ctl->diffusion && ctl->turb_pbl_scheme == 1;
  const int conv_cpu = !dm && (ctl->conv_mix_pbl || ctl->conv_cape >= 0) && (ctl->conv_dt <= 0 || fmod(t, ctl->conv_dt) == 0);
  const int iso_cpu = !dm && ctl->isosurf >= 1 && ctl->isosurf <= 4;
  const int meteo_cpu = ctl->met_dt_out > 0 && (ctl->met_dt_out < ctl->dt_mod || fmod(t, ctl->met_dt_out) == 0) && meteo_needs_host(ctl);
  const int bound_cpu = !dm_tail && (ctl->bound_lat0 < ctl->bound_lat1) && (ctl->bound_p0 > ctl->bound_p1);
  const int decay_cpu = !dm_tail && ctl->tdec_trop > 0 && ctl->tdec_strat > 0;
  const int loss_cpu = !dm_tail && ctl->qnt_loss_rate >= 0;
  const int chemgrid_cpu = ctl->oh_chem_reaction != 0 || ctl->h2o2_chem_reaction != 0 || (ctl->kpp_chem && fmod(t, ctl->dt_kpp) == 0);
  const int wet_cpu = (ctl->wet_depo_ic_a > 0 || ctl->wet_depo_ic_h[0] > 0) && (ctl->wet_depo_bc_a > 0 || ctl->wet_depo_bc_h[0] > 0);
  const int tail_cpu = meteo_cpu || bound_cpu || loss_cpu || decay_cpu || chemgrid_cpu || ctl->oh_chem_reaction != 0 ||
    ctl->h2o2_chem_reaction != 0 || ctl->tracer_chem || ctl->radio_decay || ctl->radio_depo || ctl->kpp_chem || wet_cpu || ctl->dry_depo_vdep > 0;

  if (init_cpu) {
    ON_HOST({
      /* (routed: the device computes iso_var itself; ISOSURF 4 still reads the balloon file here) */
      if (ctl->isosurf >= 1 && ctl->isosurf <= 4 && (!dm || ctl->isosurf == 4)) module_isosurf_init(ctl, cache, *met0, *met1, atm);
      module_chem_init(ctl, cache, clim, *met0, *met1, atm);
    });
    FLUSH_HOST();
    if (dm && ctl->isosurf == 4) MPB(mpb_team_set_balloon(g_ctx, cache->iso_n, cache->iso_ts, cache->iso_ps));
  }

  if (!pbl_cpu && !conv_cpu && !iso_cpu && !tail_cpu && module_timers()) {
    /* MPTRAC_B200_MODULE_TIMERS=1: the step in the reference's timer groups (src/mptrac.c:5893, 5071, 5176, ...), each group
       waited for, so that the reference's own timing report lines (TIMER_MODULE_SORT / _METEO / _MIXING ...) stay comparable;
       the per-parcel modules between the two position checks remain ONE launch and report as MODULE_B200_STEP */
    const unsigned parcel = MPB_MOD_POSITION0 | MPB_MOD_ADVECT | MPB_MOD_DIFF_TURB | MPB_MOD_DIFF_PBL | MPB_MOD_DIFF_MESO |
      MPB_MOD_CONVECTION | MPB_MOD_SEDI | MPB_MOD_ISOSURF | MPB_MOD_POSITION1;
    const struct { const char *name; unsigned mask; } part[] = {
      {"MODULE_TIMESTEPS", MPB_MOD_TIMESTEPS}, {"MODULE_SORT", MPB_MOD_SORT}, {"MODULE_B200_STEP", parcel},
      {"MODULE_METEO", MPB_MOD_METEO}, {"MODULE_BOUND_COND", MPB_MOD_BOUND0}, {"MODULE_DECAY", MPB_MOD_DECAY},
      {"MODULE_MIXING", MPB_MOD_MIXING}, {"MODULE_CHEM_GRID", MPB_MOD_CHEMGRID}, {"MODULE_BOUND_COND", MPB_MOD_BOUND1}};
    for (size_t i = 0; i < sizeof(part) / sizeof(part[0]); i++) {
      SELECT_TIMER(part[i].name, "PHYSICS");
      MPB(mpb_team_run_modules(g_ctx, t, part[i].mask));
      MPB(mpb_team_sync(g_ctx));
    }
  } else if (!pbl_cpu && !conv_cpu && !iso_cpu && !tail_cpu) {
    MPB(mpb_team_run_timestep(g_ctx, t));            /* the whole step is one fused launch (+ sort / mixing kernels) */
  } else {
    unsigned seg = MPB_MOD_TIMESTEPS | MPB_MOD_SORT | MPB_MOD_POSITION0 | MPB_MOD_ADVECT | MPB_MOD_DIFF_TURB;
    if (dm) seg |= MPB_MOD_DIFF_PBL;
    if (pbl_cpu) {
      MPB(mpb_team_run_modules(g_ctx, t, seg)); seg = 0; g_dev_newer = 1;
      ON_HOST({
        host_rng_catch_up(ctl, cache);
        module_diff_pbl(ctl, cache, *met0, *met1, atm);
        host_rng_drew(3ull * (unsigned long long) atm->np + 1ull);
      });
      FLUSH_HOST();
    }
    seg |= MPB_MOD_DIFF_MESO;
    if (dm) seg |= MPB_MOD_CONVECTION;
    if (conv_cpu) {
      MPB(mpb_team_run_modules(g_ctx, t, seg)); seg = 0; g_dev_newer = 1;
      ON_HOST({
        host_rng_catch_up(ctl, cache);
        module_convection(ctl, cache, *met0, *met1, atm);
        host_rng_drew((unsigned long long) atm->np + 1ull);
      });
      FLUSH_HOST();
    }
    seg |= MPB_MOD_SEDI;
    if (dm) seg |= MPB_MOD_ISOSURF;
    if (iso_cpu) {
      MPB(mpb_team_run_modules(g_ctx, t, seg)); seg = 0; g_dev_newer = 1;
      ON_HOST(module_isosurf(ctl, cache, *met0, *met1, atm));
      FLUSH_HOST();
    }
    seg |= MPB_MOD_POSITION1;
    if (!meteo_cpu) seg |= MPB_MOD_METEO;       /* every quantity module_meteo sets here is available on the device */
    if (dm_tail) seg |= MPB_MOD_BOUND0 | MPB_MOD_DECAY;
    MPB(mpb_team_run_modules(g_ctx, t, seg));
    g_dev_newer = 1;
    if (tail_cpu) {
      /* everything after the final position check runs through the reference's own CPU code, in its order */
      const int mix = ctl->mixing_trop >= 0 && ctl->mixing_strat >= 0 && (ctl->mixing_dt <= 0 || fmod(t, ctl->mixing_dt) == 0);
      ON_HOST({
        if (meteo_cpu) module_meteo(ctl, cache, clim, *met0, *met1, atm);
        if (bound_cpu) module_bound_cond(ctl, cache, clim, *met0, *met1, atm);
        if (ctl->qnt_loss_rate >= 0)
          for (int ip = 0; ip < atm->np; ip++) if (cache->dt[ip] != 0) atm->q[ctl->qnt_loss_rate][ip] = 0;
        if (decay_cpu) module_decay(ctl, cache, clim, atm);
        if (mix) module_mixing(ctl, clim, atm, t);
        if (chemgrid_cpu) module_chem_grid(ctl, *met0, *met1, atm, t);
        if (ctl->oh_chem_reaction != 0) module_oh_chem(ctl, cache, clim, *met0, *met1, atm);
        if (ctl->h2o2_chem_reaction != 0) module_h2o2_chem(ctl, cache, clim, *met0, *met1, atm);
        if (ctl->tracer_chem) module_tracer_chem(ctl, cache, clim, *met0, *met1, atm);
        if (ctl->radio_decay) module_radio_decay(ctl, cache, atm);
        if (ctl->radio_depo) module_radio_depo(ctl, cache, *met0, *met1, atm, depo);
        if (ctl->kpp_chem && fmod(t, ctl->dt_kpp) == 0) ERRMSG("mptrac_b200: KPP chemistry is not available");
        if (wet_cpu) module_wet_depo(ctl, cache, *met0, *met1, atm);
        if (ctl->dry_depo_vdep > 0) module_dry_depo(ctl, cache, *met0, *met1, atm);
        if (bound_cpu) module_bound_cond(ctl, cache, clim, *met0, *met1, atm);
      });
      FLUSH_HOST();
      g_dev_newer = 0;     /* host and device hold the same parcels now */
    } else {
      MPB(mpb_team_run_modules(g_ctx, t, MPB_MOD_MIXING | (dm_tail ? MPB_MOD_BOUND1 : 0u)));
    }
  }
  if (!tail_cpu) g_dev_newer = 1;
  if (!getenv("MPTRAC_B200_ASYNC")) MPB(mpb_team_sync(g_ctx));   /* keeps the reference's wall-clock timers honest */
}
