/*
 * mptrac_b200.h -- C ABI of the B200-native particle time-step engine.
 *
 * This is the drop-in boundary for MPTRAC's per-particle time-step path.  Every entry
 * point takes plain pointers / sizes (no torch, no CUDA types) and is what a binding of
 * the reference (its C driver `trac`, the Fortran `bind(c)` wrapper, or ctypes) links
 * against.  The reference interface each function stands in for is cited as
 * file:line relative to the slcs-jsc/mptrac tree.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mpb_last_error()
 *     then holds a one-line message (the reference itself aborts through ERRMSG,
 *     src/mptrac.h:2406-2410; the shim in mptrac_b200/csrc/shim turns a non-zero
 *     status into exactly that behaviour).
 *   - all work is enqueued on the context's CUDA stream and is asynchronous unless
 *     stated otherwise; mpb_sync() waits for it.
 *   - there is no CPU fallback: without a CUDA device mpb_create() fails.
 */
#ifndef MPTRAC_B200_H
#define MPTRAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI history (mpb_abi_version()): 2 module_meteo quantities of the resident fields; 3 model-level fields and the zeta / eta
 * quantities (ADVECT_VERT_COORD 1, 2, 3); 4 further met fields (x2 / x3), 64 meteo slots, module_convection, module_decay,
 * module_isosurf, module_diff_pbl, module_bound_cond, module_chem_grid (their control fields at the end of mpb_ctl_t, MPB_MOD_* bits);
 * 5 module_mixing accumulates all quantities at once (box records), ranks exchange through peer memory (mpb_peer_*), several
 * devices behind one host thread (mpb_team_*) */
#define MPB_ABI_VERSION 5
#define MPB_MAX_RANKS 16         /* ranks / team members that can exchange box records */
#define MPB_IPC_HANDLE_BYTES 64  /* size of the handle mpb_peer_init exports (a CUDA IPC memory handle) */
#define MPB_MIX_MAXQ 23  /* number of mixable quantities, src/mptrac.c:5222-5230 */

/* Quantities module_meteo (src/mptrac.c:5062-5165) can set on the device, slot order of mpb_ctl_t::qnt_meteo:
 *   PS .. ZETA_D  from the met fields the path keeps resident (T, u, v, w on pressure levels; ps, pbl);
 *   TS .. O3C     the further 2-D fields of INTPOL_TIME_ALL (src/mptrac.h:1278-1316), slot = MPB_Q_TS + MPB_F2_*;
 *   ZG .. CC      its further 3-D fields, slot = MPB_Q_ZG + MPB_F3_*;
 *   PW .. TICE    derived from T and H2O (PW, SH, RH, RHICE, TVIRT, lapse_rate, TDEW, TICE).
 * A further-field quantity needs that field in both met levels (mpb_met_view_t::x2 / x3).  Quantities that read a
 * climatology (hno3, oh, h2o2, ho2, o1d, tnat, tsts) are not on the device. */
enum {
  MPB_Q_PS, MPB_Q_PBL, MPB_Q_P, MPB_Q_T, MPB_Q_RHO, MPB_Q_U, MPB_Q_V, MPB_Q_W, MPB_Q_VH, MPB_Q_VZ, MPB_Q_THETA,
  MPB_Q_PSAT, MPB_Q_PSICE, MPB_Q_ZETA_D,
  MPB_Q_TS, MPB_Q_ZS, MPB_Q_US, MPB_Q_VS, MPB_Q_ESS, MPB_Q_NSS, MPB_Q_SHF, MPB_Q_LSM, MPB_Q_SST, MPB_Q_PT, MPB_Q_TT, MPB_Q_ZT,
  MPB_Q_H2OT, MPB_Q_PCT, MPB_Q_PCB, MPB_Q_CL, MPB_Q_PLCL, MPB_Q_PLFC, MPB_Q_PEL, MPB_Q_CAPE, MPB_Q_CIN, MPB_Q_O3C,
  MPB_Q_ZG, MPB_Q_PV, MPB_Q_H2O, MPB_Q_O3, MPB_Q_LWC, MPB_Q_RWC, MPB_Q_IWC, MPB_Q_SWC, MPB_Q_CC,
  MPB_Q_PW, MPB_Q_SH, MPB_Q_RH, MPB_Q_RHICE, MPB_Q_TVIRT, MPB_Q_LAPSE, MPB_Q_TDEW, MPB_Q_TICE, MPB_NMETEO
};
#define MPB_METEO_SLOTS 64
/* further met fields: met_t::ts ... o3c ([EX][EY]) and met_t::z ... cc ([EX][EY][EP]), src/mptrac.h:3886-3995 */
enum {
  MPB_F2_TS, MPB_F2_ZS, MPB_F2_US, MPB_F2_VS, MPB_F2_ESS, MPB_F2_NSS, MPB_F2_SHF, MPB_F2_LSM, MPB_F2_SST, MPB_F2_PT, MPB_F2_TT,
  MPB_F2_ZT, MPB_F2_H2OT, MPB_F2_PCT, MPB_F2_PCB, MPB_F2_CL, MPB_F2_PLCL, MPB_F2_PLFC, MPB_F2_PEL, MPB_F2_CAPE, MPB_F2_CIN,
  MPB_F2_O3C, MPB_NX2
};
enum { MPB_F3_Z, MPB_F3_PV, MPB_F3_H2O, MPB_F3_O3, MPB_F3_LWC, MPB_F3_RWC, MPB_F3_IWC, MPB_F3_SWC, MPB_F3_CC, MPB_NX3 };

typedef struct mpb_ctx mpb_ctx;

/* Scalar control parameters read on the path.  Field names and meaning are those of
 * ctl_t (src/mptrac.h:2494 ff.); defaults are set by mptrac_read_ctl (src/mptrac.c:6723 ff.). */
typedef struct mpb_ctl {
  int32_t direction;          /* +1 forward, -1 backward                    ctl->direction       */
  int32_t met_coord_type;     /* 0 lon/lat degrees, 1 Cartesian metres       ctl->met_coord_type  */
  int32_t advect;             /* 0 off, 1 Euler, 2 midpoint, 4 RK4           ctl->advect          */
  int32_t advect_vert_coord;  /* only 0 (pressure levels, omega) is on device ctl->advect_vert_coord */
  int32_t rng_type;           /* only 1 (Squares counter RNG) is on device    ctl->rng_type        */
  int32_t diffusion;          /* master switch of both diffusion modules      ctl->diffusion       */
  int32_t turb_pbl_scheme;    /* >0: diff_turb skips parcels inside the PBL   ctl->turb_pbl_scheme */
  int32_t nq;                 /* number of quantities                         ctl->nq              */
  int32_t qnt_rp;             /* quantity index of particle radius or -1      ctl->qnt_rp          */
  int32_t qnt_rhop;           /* quantity index of particle density or -1     ctl->qnt_rhop        */
  int32_t qnt_ens;            /* quantity index of ensemble id or -1          ctl->qnt_ens         */
  int32_t nens;               /* number of ensembles (0 = off)                ctl->nens            */
  int32_t mixing_nx, mixing_ny, mixing_nz;
  int32_t n_mix_qnt;                  /* how many entries of mix_qnt are valid */
  int32_t mix_qnt[MPB_MIX_MAXQ];      /* quantity indices relaxed by module_mixing, in reference order */
  int32_t _pad;
  double t_start, t_stop;     /* ctl->t_start (after module_timesteps_init), ctl->t_stop */
  double dt_mod, dt_met;
  double met_utm_ref_lat;
  double sort_dt;
  double turb_dx_pbl, turb_dx_trop, turb_dx_strat;
  double turb_dz_pbl, turb_dz_trop, turb_dz_strat;
  double turb_mesox, turb_mesoz;
  double turb_pbl_trans;
  double mixing_dt, mixing_trop, mixing_strat;
  double mixing_lon0, mixing_lon1, mixing_lat0, mixing_lat1, mixing_z0, mixing_z1;
  double met_dt_out;                        /* module_meteo every met_dt_out seconds (<= 0: never)  ctl->met_dt_out */
  int32_t qnt_meteo[MPB_METEO_SLOTS];       /* quantity index per MPB_Q_* slot or -1               ctl->qnt_ps ... */
  int32_t qnt_zeta, qnt_eta;                /* the parcel's model-level coordinate (ADVECT_VERT_COORD 1 / 3) or -1:
                                               ctl->qnt_zeta, ctl->qnt_eta (src/mptrac.c:3683-3687) */
  /* module_convection (src/mptrac.c:4102-4171; needs the met fields MPB_F2_CAPE, _CIN, _PEL when conv_cape >= 0) and
   * module_decay (4227-4263) with the reset of the total loss rate before it (7931-7936) */
  double conv_cape, conv_cin, conv_pbl_trans, conv_dt;   /* ctl->conv_*: off with conv_cape < 0 and conv_mix_pbl 0 */
  double tdec_trop, tdec_strat;                          /* ctl->tdec_*: decay runs when both are > 0 */
  int32_t conv_mix_pbl;
  int32_t qnt_m, qnt_vmr, qnt_mloss_decay, qnt_loss_rate;   /* quantity indices or -1 */
  int32_t isosurf;            /* ctl->isosurf: module_isosurf (src/mptrac.c:4956-5004), 0 = off, 1 pressure, 2 density,
                                 3 potential temperature, 4 balloon time series (mpb_set_balloon) */
  /* module_bound_cond (src/mptrac.c:3789-3881): on when bound_lat0 < bound_lat1 and bound_p0 > bound_p1 (7926, 7997) */
  double bound_mass, bound_mass_trend, bound_vmr, bound_vmr_trend;   /* ctl->bound_* */
  double bound_lat0, bound_lat1, bound_p0, bound_p1, bound_dps, bound_dzs, bound_zetas;
  int32_t bound_pbl, qnt_aoa;
  int32_t qnt_cts[5];         /* ctl->qnt_Cccl4, qnt_Cccl3f, qnt_Cccl2f2, qnt_Cn2o, qnt_Csf6 */
  int32_t cts_on;             /* bit i: ctl->clim_*_timeseries of species i is not "-" (series through mpb_set_clim_ts) */
  /* module_chem_grid (src/mptrac.c:3885-4054): volume mixing ratio of the parcel's chemistry-grid box into quantity Cx */
  double chemgrid_lon0, chemgrid_lon1, chemgrid_lat0, chemgrid_lat1, chemgrid_z0, chemgrid_z1, molmass;   /* ctl->chemgrid_*, molmass */
  int32_t chemgrid_nx, chemgrid_ny, chemgrid_nz, qnt_Cx;
  int32_t chemgrid;           /* run it (the reference does for its OH / H2O2 / KPP chemistry, src/mptrac.c:7947-7950) */
  int32_t _pad3;
} mpb_ctl_t;

/* Host view of one met_t time level (src/mptrac.h:3844-4014).  3-D element (ix,iy,iz) lives at
 * base[ix*sx + iy*sy + iz]; 2-D element (ix,iy) at base[ix*sx2 + iy].  For the reference structs
 * sx = EY*EP, sy = EP, sx2 = EY.  t and pbl may be NULL (treated as 0). */
typedef struct mpb_met_view {
  double time;
  int32_t coord_type, nx, ny, np;
  const double *lon, *lat, *p;
  const float *u, *v, *w, *t;
  const float *ps, *pbl;
  int64_t sx, sy, sx2;
  /* fields on model levels (met_t::pl, ul, vl, wl, zetal, zeta_dotl, src/mptrac.h:3540-3556), element (ix,iy,k) at
   * base[ix*sxl + iy*syl + k] with k < npl (the reference structs: sxl = EY*EP, syl = EP); npl = 0 and null pointers
   * when the run does not advect on model levels */
  int32_t npl, _pad;
  const float *pl, *ul, *vl, *wl, *zetal, *zeta_dotl;
  int64_t sxl, syl;
  /* further fields for module_meteo, MPB_F2_* (strides like ps) and MPB_F3_* (strides like u); NULL = not given */
  const float *x2[MPB_NX2];
  const float *x3[MPB_NX3];
} mpb_met_view_t;

/* Parameters of the gridded-output binning (write_grid, src/mptrac.c:13752 ff.). */
typedef struct mpb_grid {
  int32_t nx, ny, nz, _pad;
  double lon0, lon1, lat0, lat1, z0, z1;
  double t0, t1;              /* time window t -/+ 0.5 dt_mod, src/mptrac.c:13824-13825 */
} mpb_grid_t;

const char *mpb_last_error(void);
int mpb_abi_version(void);
int mpb_device_count(void);
int mpb_warmup(int device);   /* create the device's CUDA context now (thread-safe): lets a driver hide it behind its file input */

/* --- lifetime: stands in for mptrac_alloc / mptrac_free (src/mptrac.c:6294, :6377) --- */
int mpb_create(mpb_ctx **ctx, int device, int64_t np_max, int nq);
int mpb_destroy(mpb_ctx *ctx);
int mpb_set_stream(mpb_ctx *ctx, void *cuda_stream);   /* run on a caller-owned stream (0 = own stream) */
int mpb_sync(mpb_ctx *ctx);

/* --- host -> device: stands in for mptrac_update_device (src/mptrac.c:8005) --- */
int mpb_set_ctl(mpb_ctx *ctx, const mpb_ctl_t *ctl);
int mpb_set_clim_tropo(mpb_ctx *ctx, int ntime, int nlat, const double *time,
                       const double *lat, const double *tropo /* [ntime][nlat] */);
int mpb_set_met(mpb_ctx *ctx, int slot /* 0 = met0, 1 = met1 */, const mpb_met_view_t *met);
/* A met level straight from one of the reference's uncompressed binary files (MET_TYPE 1: write_met_bin src/mptrac.c:14204,
 * read_met_bin :8887-9181) into the device layout, without a met_t in between; 3-D fields are clamped like read_met_bin_3d
 * does; all_fields != 0 also uploads the further fields (mpb_met_view_t::x2 / x3).  Needs mpb_set_ctl first (MET_COORD_TYPE). */
int mpb_set_met_bin(mpb_ctx *ctx, int slot, const char *path, int all_fields);
int mpb_swap_met(mpb_ctx *ctx);   /* the pointer swap of mptrac_get_met, src/mptrac.c:6489-6491 */
int mpb_set_atm(mpb_ctx *ctx, int64_t np, const double *time, const double *p,
                const double *lon, const double *lat, const double *q, int64_t q_stride);
int mpb_set_uvwp(mpb_ctx *ctx, const float *uvwp /* [np][3], cache_t::uvwp */);

/* --- device -> host: stands in for mptrac_update_host (src/mptrac.c:8061) --- */
int mpb_get_atm(mpb_ctx *ctx, double *time, double *p, double *lon, double *lat,
                double *q, int64_t q_stride);
int mpb_get_uvwp(mpb_ctx *ctx, float *uvwp);
/* module_isosurf (src/mptrac.c:4886-5004): cache_t::iso_var [np] stays attached to the array slot (module_sort does not move
 * it, src/mptrac.c:5944-5949); the balloon series is what module_isosurf_init reads from ctl->balloon for ISOSURF 4 */
int mpb_set_iso_var(mpb_ctx *ctx, const double *iso_var);
int mpb_get_iso_var(mpb_ctx *ctx, double *iso_var);
int mpb_set_balloon(mpb_ctx *ctx, int n, const double *ts, const double *ps);
/* trace-gas time series of module_bound_cond, clim_t::ccl4, ccl3f, ccl2f2, n2o, sf6 (species 0 .. 4; src/mptrac.h:3821-3833) */
int mpb_set_clim_ts(mpb_ctx *ctx, int species, int n, const double *time, const double *vmr);
int mpb_get_dt(mpb_ctx *ctx, double *dt /* cache_t::dt */);
int64_t mpb_get_np(mpb_ctx *ctx);

/* --- multi-GPU: this context owns parcels [offset, offset+np) of a run with global_np parcels.
 *     Random numbers are addressed by GLOBAL parcel index, so a sharded run reproduces the
 *     single-device stream (src/mptrac.c:5797-5826). --- */
int mpb_set_shard(mpb_ctx *ctx, int64_t global_offset, int64_t global_np);
int mpb_set_rng_ctr(mpb_ctx *ctx, uint64_t ctr);   /* static rng_ctr, src/mptrac.c:35 */
uint64_t mpb_get_rng_ctr(mpb_ctx *ctx);

/* --- the step: stands in for mptrac_run_timestep (src/mptrac.c:7851-8001) restricted to the
 *     modules on the path; all enabled modules run fused in ONE kernel per step. --- */
int mpb_run_timestep(mpb_ctx *ctx, double t);

/* The same step for parcels that live in HOST memory (the driver's atm_t): stands in for the sequence
 * mptrac_update_device(atm) -> mptrac_run_timestep -> mptrac_update_host(atm) (src/mptrac.c:8005, :7851, :8061) that a
 * caller needs when it wants the parcels back after every step; the result is identical to the three calls.
 * Pinned (page-locked) host arrays are read and written by the step kernel itself through their device mapping, so
 * both PCIe directions run for the whole launch with no copy engine in between; pageable arrays are cut into chunks
 * that are uploaded, stepped and downloaded in a software pipeline over several streams.  Steps with a global phase
 * (cell sort, mixing) take the plain three-call sequence.  Synchronous: the arrays are valid on return.
 * Only the quantities the path reads (rp, rhop) are transferred; none is written back (the path modifies none). */
int mpb_run_timestep_host(mpb_ctx *ctx, double t, int64_t np, double *time, double *p, double *lon, double *lat,
                          double *q, int64_t q_stride);
/* bytes the last mpb_run_timestep_host call moved across the host link in each direction (when all parcels carry the same
 * time -- the usual case -- time[] does not travel: every parcel takes the same time step, the host array is filled with the
 * common result by the host itself; 24 instead of 32 bytes per parcel and direction) */
int mpb_host_step_bytes(mpb_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* A sub-sequence of the step, for callers that interleave modules of their own (the shim does this for
 * reference modules that are not on the device path).  `mask` selects modules, which still run in the
 * reference's order and only if the control parameters enable them; adjacent selected modules are fused into one
 * kernel.  Without MPB_MOD_TIMESTEPS the per-parcel dt is read from device memory (cache_t::dt). */
#define MPB_MOD_TIMESTEPS 0x001
#define MPB_MOD_SORT      0x002
#define MPB_MOD_POSITION0 0x004
#define MPB_MOD_ADVECT    0x008
#define MPB_MOD_DIFF_TURB 0x010
#define MPB_MOD_DIFF_MESO 0x020
#define MPB_MOD_SEDI      0x040
#define MPB_MOD_POSITION1 0x080
#define MPB_MOD_MIXING    0x100
#define MPB_MOD_METEO     0x200   /* between POSITION1 and MIXING, like the reference (src/mptrac.c:7927-7945) */
#define MPB_MOD_CONVECTION 0x400  /* between DIFF_MESO and SEDI (src/mptrac.c:7905-7908) */
#define MPB_MOD_DECAY     0x800   /* reset of the total loss rate + module_decay, between METEO and MIXING (7931-7940) */
#define MPB_MOD_ISOSURF   0x1000  /* between SEDI and POSITION1 (src/mptrac.c:7914-7916); its init runs at t_start */
#define MPB_MOD_DIFF_PBL  0x2000  /* TURB_PBL_SCHEME 1, between DIFF_TURB and DIFF_MESO (src/mptrac.c:7897-7899) */
#define MPB_MOD_BOUND0    0x4000  /* module_bound_cond after METEO (src/mptrac.c:7926-7929) ... */
#define MPB_MOD_BOUND1    0x8000  /* ... and again at the end of the step (7997-8000) */
#define MPB_MOD_CHEMGRID  0x10000 /* module_chem_grid, after MIXING (src/mptrac.c:7947-7950) */
#define MPB_MOD_ALL       0x1ffff
int mpb_run_modules(mpb_ctx *ctx, double t, unsigned mask);
/* The launches mpb_run_modules(ctx, t, mask) would make for this control structure, as text ("step(advect=4,phys=0x0,mod=0x43)
 * meteo ..."): pure host logic, needs neither a context nor a device.  mod bits: 0x01 timesteps, 0x02 position (initial),
 * 0x40 position (final), 0x80 dt written to cache_t::dt; phys bits: 1 diff_turb, 2 diff_meso, 4 sedi. */
int mpb_plan_modules(const mpb_ctl_t *ctl, double t, unsigned mask, char *buf, int len);

/* --- single modules (same symbols the reference exports, src/mptrac.h:6140-7132); each is the
 *     same fused kernel restricted to one module and reads cache->dt from device memory. --- */
int mpb_module_timesteps(mpb_ctx *ctx, double t);   /* src/mptrac.c:5999 */
int mpb_module_position(mpb_ctx *ctx);              /* src/mptrac.c:5435 */
int mpb_module_advect(mpb_ctx *ctx);                /* src/mptrac.c:3598 */
int mpb_module_diff_turb(mpb_ctx *ctx);             /* src/mptrac.c:4588 */
int mpb_module_diff_meso(mpb_ctx *ctx);             /* src/mptrac.c:4266 */
int mpb_module_sedi(mpb_ctx *ctx);                  /* src/mptrac.c:5859 */
int mpb_module_sort(mpb_ctx *ctx);                  /* src/mptrac.c:5887 */
int mpb_module_meteo(mpb_ctx *ctx);                 /* src/mptrac.c:5062, the MPB_Q_* quantities; every parcel (check_dt = 0) */
int mpb_module_mixing(mpb_ctx *ctx, double t);      /* src/mptrac.c:5169 (single device) */
int mpb_module_rng(mpb_ctx *ctx, double *rs_host, int64_t n, int method); /* src/mptrac.c:5753; fills host array, advances counter */

/* --- module_mixing (src/mptrac.c:5169-5347) and the binning of write_grid (13840-13872) over several ranks ---
 * The box records of module_mixing -- per box {count, sum of every mixed quantity}, doubles -- are accumulated for ALL mixed
 * quantities in one pass.  Three ways to run the exchange step:
 *  (1) ranks attached through peer memory (mpb_peer_init + mpb_peer_attach, one process per GPU): mpb_run_timestep /
 *      mpb_module_mixing do everything -- the box space is cut into one slice per rank; every parcel's contribution is
 *      routed into the inbox of its box's OWNER (coalesced stores over NVLink), the owner folds its inboxes into its
 *      records (local atomics) and routes the finished record back to each contribution's sender; two barriers in stream
 *      order; 8 x (quantities + 1) bytes per parcel and direction cross the link;
 *  (2) a team (mpb_team_*, several devices behind one host thread): the same through mpb_team_run_timestep;
 *  (3) any other transport: mpb_mixing_accumulate_all, sum mpb_device_ptr("mix_rec") [mpb_mixing_rec_len() doubles] over the
 *      ranks (one all-reduce), mpb_mixing_apply_all.
 * Gridded output: mpb_grid_accumulate leaves partial arrays on every rank; mpb_grid_reduce adds them up on rank 0 (attached
 * ranks; otherwise reduce mpb_device_ptr("grid_cnt" / "grid_sum" / "grid_sq") yourself); mpb_grid_fetch copies them out. */
int mpb_mixing_accumulate_all(mpb_ctx *ctx, double t);             /* box index per parcel, records zeroed, local contributions */
int mpb_mixing_apply_all(mpb_ctx *ctx);                            /* box means + relaxation of every mixed quantity */
int64_t mpb_mixing_nbox(mpb_ctx *ctx);
int64_t mpb_mixing_rec_len(mpb_ctx *ctx);                          /* boxes x (mixed quantities + 1, rounded up to even) doubles */
int mpb_grid_accumulate(mpb_ctx *ctx, const mpb_grid_t *grid);      /* -> count[nbox], sum[nq][nbox], sumsq[nq][nbox] on device */
int mpb_grid_reduce(mpb_ctx *ctx);                                  /* attached ranks: sum over ranks onto rank 0 */
int mpb_grid_fetch(mpb_ctx *ctx, int *count, double *sum, double *sumsq);   /* copy them to the host (any may be NULL) */

/* Ranks that exchange through peer memory.  mpb_peer_init allocates this rank's exchange area -- barrier flags, the inboxes and
 * outboxes of the mixing exchange (mix_bytes >= 256 + 16 x nranks x P x (mixed quantities + 1), P = the largest number of
 * parcels any rank holds; the SAME value on every rank), its partial output arrays (grid_bytes >= boxes x (16 x nq + 4)) --
 * and exports a handle (MPB_IPC_HANDLE_BYTES) other processes open with mpb_peer_attach (handles
 * of all ranks, rank order; the caller gathers them, e.g. with its MPI / torch.distributed, and synchronises the ranks once
 * after attaching).  Contexts of one process attach each other's mpb_peer_area directly.  All ranks must then issue the same
 * sequence of exchange steps.  A barrier that waits ~5 s for a missing rank gives up; mpb_sync reports it. */
int mpb_peer_init(mpb_ctx *ctx, int rank, int nranks, int64_t mix_bytes, int64_t grid_bytes, void *ipc_handle_out /* or NULL */);
int mpb_peer_attach(mpb_ctx *ctx, const void *ipc_handles /* [nranks][MPB_IPC_HANDLE_BYTES] */);
int mpb_peer_attach_local(mpb_ctx *ctx, void *const *areas /* [nranks] */, const int *devices /* [nranks] */);
void *mpb_peer_area(mpb_ctx *ctx);
int mpb_peer_barrier(mpb_ctx *ctx);

/* --- a team: several devices behind ONE host thread, for the reference's single-process driver (device selection:
 *     src/trac.c:75-80; met broadcast: src/mptrac.c:45-69).  Parcels are cut into contiguous index ranges, one per member;
 *     the met data is uploaded and packed once and copied device to device; each call below is the per-context call of
 *     the same name applied to the whole parcel set. --- */
typedef struct mpb_team mpb_team;
int mpb_team_create(mpb_team **team, int ndev, const int *devices, int64_t np_max, int nq);
int mpb_team_destroy(mpb_team *team);
int mpb_team_size(mpb_team *team);
mpb_ctx *mpb_team_member(mpb_team *team, int i);
int mpb_team_set_ctl(mpb_team *team, const mpb_ctl_t *ctl);
int mpb_team_set_clim_tropo(mpb_team *team, int ntime, int nlat, const double *time, const double *lat, const double *tropo);
int mpb_team_set_clim_ts(mpb_team *team, int species, int n, const double *time, const double *vmr);
int mpb_team_set_balloon(mpb_team *team, int n, const double *ts, const double *ps);
int mpb_team_set_met(mpb_team *team, int slot, const mpb_met_view_t *met);
int mpb_team_swap_met(mpb_team *team);
int mpb_team_set_atm(mpb_team *team, int64_t np, const double *time, const double *p, const double *lon, const double *lat,
                     const double *q, int64_t q_stride);
int mpb_team_get_atm(mpb_team *team, double *time, double *p, double *lon, double *lat, double *q, int64_t q_stride);
int mpb_team_set_uvwp(mpb_team *team, const float *uvwp);
int mpb_team_get_uvwp(mpb_team *team, float *uvwp);
int mpb_team_get_dt(mpb_team *team, double *dt);
int mpb_team_set_iso_var(mpb_team *team, const double *iso_var);
int mpb_team_get_iso_var(mpb_team *team, double *iso_var);
int64_t mpb_team_get_np(mpb_team *team);
int mpb_team_set_rng_ctr(mpb_team *team, uint64_t ctr);
uint64_t mpb_team_get_rng_ctr(mpb_team *team);
int mpb_team_run_timestep(mpb_team *team, double t);
int mpb_team_run_modules(mpb_team *team, double t, unsigned mask);
int mpb_team_sync(mpb_team *team);
int64_t mpb_team_launch_count(mpb_team *team);
int mpb_team_grid_accumulate(mpb_team *team, const mpb_grid_t *grid);   /* partial arrays everywhere, summed on member 0 */
int mpb_team_grid_fetch(mpb_team *team, int *count, double *sum, double *sumsq);

/* --- introspection --- */
void *mpb_device_ptr(mpb_ctx *ctx, const char *name);   /* "time","p","lon","lat","q","dt","uvwp","mix_rec","grid_cnt","grid_sum","grid_sq" */
int64_t mpb_launch_count(mpb_ctx *ctx);                  /* kernels launched by this context so far */
int mpb_met_bytes(mpb_ctx *ctx, int64_t *bytes);         /* packed device bytes of both met levels */

#ifdef __cplusplus
}
#endif
#endif /* MPTRAC_B200_H */
