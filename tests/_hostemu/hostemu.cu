// TEST-ONLY host emulation of the device physics.  The development container has no GPU, so this file
// runs the *same* __host__ __device__ source the kernels are built from (mptrac_b200/csrc/physics.cuh)
// in a plain CPU loop, to debug the parcel arithmetic against the oracle before GPU time is spent.
// It is never part of libmptrac_b200.so, never imported by the package, and proves nothing about the
// GPU build: the parity tests proper are the `-m gpu` tests that call through the C ABI.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mptrac_b200/csrc/met_tables.hpp"
#include "../../mptrac_b200/csrc/physics.cuh"

using namespace mpb;

struct EmuMet {
  double time;
  int coord_type, nx, ny, np;
  const double *lon, *lat, *p;
  const float *u, *v, *w, *t, *ps, *pbl;
  int npl;
  const float *pl, *ul, *vl, *wl, *zetal, *zeta_dotl;
  const float *x2[22], *x3[9];
};

struct EmuCtl {
  double t, t_start, t_stop, dt_met, utm_ref_lat;
  double dx_pbl, dx_trop, dx_strat, dz_pbl, dz_trop, dz_strat, mesox, mesoz, pbl_trans;
  uint64_t ctr_turb, ctr_meso;
  int direction, pbl_scheme, advect, phys;
  unsigned modules;
};

static std::vector<Node> pack_nodes(const EmuMet &m0, const EmuMet &m1) {
  const size_t n = (size_t)m0.nx * m0.ny * m0.np;
  std::vector<Node> o(n);
  for (size_t i = 0; i < n; i++) {
    o[i].u0 = m0.u[i]; o[i].v0 = m0.v[i]; o[i].w0 = m0.w[i]; o[i].t0 = m0.t ? m0.t[i] : 0.f;
    o[i].u1 = m1.u[i]; o[i].v1 = m1.v[i]; o[i].w1 = m1.w[i]; o[i].t1 = m1.t ? m1.t[i] : 0.f;
  }
  return o;
}
static std::vector<float4> pack_surf(const EmuMet &m0, const EmuMet &m1) {
  const size_t n = (size_t)m0.nx * m0.ny;
  std::vector<float4> o(n);
  for (size_t i = 0; i < n; i++)
    o[i] = make_float4(m0.ps ? m0.ps[i] : 0.f, m0.pbl ? m0.pbl[i] : 0.f, m1.ps ? m1.ps[i] : 0.f, m1.pbl ? m1.pbl[i] : 0.f);
  return o;
}

struct HostMet {
  std::vector<Node> f;
  std::vector<float4> s;
  AxisTables t;
  MetView g;
};

static void make_view(HostMet &h, const EmuMet *m0, const EmuMet *m1, bool with_fields) {
  if (with_fields) { h.f = pack_nodes(*m0, *m1); h.s = pack_surf(*m0, *m1); }
  h.t = build_axis_tables(m0->lon, m0->nx, m0->lat, m0->ny, m0->p, m0->np);
  MetView &g = h.g;
  std::memset(&g, 0, sizeof(g));
  g.f = h.f.data(); g.s = h.s.data();
  g.lon = m0->lon; g.lat = m0->lat; g.p = m0->p;
  g.lonc = h.t.lonc.data(); g.latc = h.t.latc.data(); g.pc = h.t.pc.data(); g.p_lut = h.t.p_lut.data();
  fill_axis_scalars(g, m0->lon, m0->nx, m0->lat, m0->ny, m0->p, m0->np, m0->coord_type, m0->time, m1 ? m1->time : m0->time + 1,
                    h.t);
}

template <int ADVECT, bool DIFF>
static void run(const MetView &g, const ClimView &cl, const CtlView &c, const EmuCtl &e, long long np, long long ig0,
                double *time, double *lon, double *lat, double *p, double *dtarr, float *uvwp, const double *rp,
                const double *rhop) {
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    double dt;
    if (e.modules & MOD_TIMESTEPS) {
      dt = parcel_dt(g, c, a);
      if (e.modules & MOD_STORE_DT) dtarr[ip] = dt;
    } else dt = dtarr[ip];
    if (dt == 0) continue;
    const uint64_t ig = (uint64_t)(ig0 + ip);
    CubeT<DIFF> cube;     // like the kernels: difference cube unless the mesoscale module needs the raw corners
    cube_reset(cube);
    if (e.modules & MOD_POS_PRE) fix_position(g, a);
#if MPB_CUBE_F64
    if (ADVECT > 0) { WindCube wc; cube_reset(wc); advect<(ADVECT > 0 ? ADVECT : 1)>(g, dt, a, wc); }
#else
    if (ADVECT > 0) advect<(ADVECT > 0 ? ADVECT : 1)>(g, dt, a, cube);
#endif
    if (e.phys & 1) diffuse_turbulent(g, cl, c, dt, ig, a);
    if constexpr (!DIFF) {
      if (e.phys & 2) diffuse_mesoscale(g, c, dt, ig, a, uvwp[3 * ip], uvwp[3 * ip + 1], uvwp[3 * ip + 2], cube);
    }
    if (e.phys & 4) sediment(g, dt, rp[ip], rhop[ip], a, cube);
    if (e.modules & MOD_POS_POST) fix_position(g, a);
    time[ip] = a.time; lon[ip] = a.lon; lat[ip] = a.lat; p[ip] = a.p;
  }
}

template <int ADVECT>
static void run(const MetView &g, const ClimView &cl, const CtlView &c, const EmuCtl &e, long long np, long long ig0,
                double *time, double *lon, double *lat, double *p, double *dtarr, float *uvwp, const double *rp,
                const double *rhop) {
  if (e.phys & 2) run<ADVECT, false>(g, cl, c, e, np, ig0, time, lon, lat, p, dtarr, uvwp, rp, rhop);
  else run<ADVECT, true>(g, cl, c, e, np, ig0, time, lon, lat, p, dtarr, uvwp, rp, rhop);
}

extern "C" int emu_step(const EmuMet *m0, const EmuMet *m1, const EmuCtl *e, int ntime, int nlat, const double *cl_time,
                        const double *cl_lat, const double *cl_tropo, long long np, long long ig0, double *time,
                        double *lon, double *lat, double *p, double *dt, float *uvwp, const double *rp,
                        const double *rhop) {
  HostMet h;
  make_view(h, m0, m1, true);
  const MetView &g = h.g;
  ClimView cl = {cl_time, cl_lat, cl_tropo, ntime, nlat};
  CtlView c;
  c.t = e->t; c.t_start = e->t_start; c.t_stop = e->t_stop; c.dt_met = e->dt_met; c.utm_ref_lat = e->utm_ref_lat;
  c.dx_pbl = e->dx_pbl; c.dx_trop = e->dx_trop; c.dx_strat = e->dx_strat;
  c.dz_pbl = e->dz_pbl; c.dz_trop = e->dz_trop; c.dz_strat = e->dz_strat;
  c.mesox = e->mesox; c.mesoz = e->mesoz; c.pbl_trans = e->pbl_trans;
  c.ctr_turb = e->ctr_turb; c.ctr_meso = e->ctr_meso; c.direction = e->direction; c.pbl_scheme = e->pbl_scheme;
  switch (e->advect) {
    case 0: run<0>(g, cl, c, *e, np, ig0, time, lon, lat, p, dt, uvwp, rp, rhop); break;
    case 1: run<1>(g, cl, c, *e, np, ig0, time, lon, lat, p, dt, uvwp, rp, rhop); break;
    case 2: run<2>(g, cl, c, *e, np, ig0, time, lon, lat, p, dt, uvwp, rp, rhop); break;
    case 4: run<4>(g, cl, c, *e, np, ig0, time, lon, lat, p, dt, uvwp, rp, rhop); break;
    default: return 1;
  }
  return 0;
}

extern "C" void emu_sort_keys(const EmuMet *m0, long long np, const double *lon, const double *lat, const double *p, int *keys) {
  HostMet h;
  make_view(h, m0, nullptr, false);
  for (long long i = 0; i < np; i++) keys[i] = cell_key(h.g, lon[i], lat[i], p[i]);
}

// model levels: the packing of mpb_set_met restated for the host
struct HostLevels {
  std::vector<LevelNode> lp, lz;
  std::vector<float4> pz;
};
static void pack_levels(HostLevels &L, MetView &g, const EmuMet &m0, const EmuMet &m1) {
  const size_t n = (size_t)m0.nx * m0.ny * m0.npl;
  L.lp.resize(n); L.lz.resize(n); L.pz.resize(n);
  auto at = [](const float *f, size_t i) { return f ? f[i] : 0.f; };
  for (size_t i = 0; i < n; i++) {
    L.lp[i] = {m0.pl[i], m0.ul[i], m0.vl[i], at(m0.wl, i), m1.pl[i], m1.ul[i], m1.vl[i], at(m1.wl, i)};
    L.lz[i] = {at(m0.zetal, i), m0.ul[i], m0.vl[i], at(m0.zeta_dotl, i), at(m1.zetal, i), m1.ul[i], m1.vl[i], at(m1.zeta_dotl, i)};
    L.pz[i] = make_float4(m0.pl[i], at(m0.zetal, i), m1.pl[i], at(m1.zetal, i));
  }
  g.lp = L.lp.data(); g.lz = L.lz.data(); g.pz = L.pz.data(); g.npl = m0.npl;
}

extern "C" int emu_advect_levels(const EmuMet *m0, const EmuMet *m1, int vert_coord, int order, long long np, double *time,
                                 double *lon, double *lat, double *p, const double *dt, double *zq) {
  HostMet h;
  HostLevels L;
  make_view(h, m0, m1, false);
  pack_levels(L, h.g, *m0, *m1);
  const MetView &g = h.g;
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    if (dt[ip] == 0) continue;
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    double z = 0;
    double *zp = zq ? &z : nullptr;
    if (order == 1) advect_on_levels<1>(g, vert_coord, dt[ip], a, zp);
    else if (order == 2) advect_on_levels<2>(g, vert_coord, dt[ip], a, zp);
    else advect_on_levels<4>(g, vert_coord, dt[ip], a, zp);
    time[ip] = a.time; lon[ip] = a.lon; lat[ip] = a.lat; p[ip] = a.p;
    if (zq) zq[ip] = z;
  }
  return 0;
}

extern "C" int emu_advect_init(const EmuMet *m0, const EmuMet *m1, long long np, const double *time, const double *lon,
                               const double *lat, double *p, const double *zq) {
  HostMet h;
  HostLevels L;
  make_view(h, m0, m1, false);
  pack_levels(L, h.g, *m0, *m1);
  for (long long ip = 0; ip < np; ip++) p[ip] = pressure_of_zeta(h.g, time[ip], zq[ip], lon[ip], lat[ip]);
  return 0;
}

// module_meteo: the 14 quantities of the resident fields (meteo_kernel's arithmetic) and the further fields / moist
// quantities (meteo_fields_kernel's), slot order of mpb_ctl_t::qnt_meteo
extern "C" int emu_meteo(const EmuMet *m0, const EmuMet *m1, long long np, const double *time, const double *lon,
                         const double *lat, const double *p, double *q, long long q_stride, const int *qnt) {
  HostMet h;
  make_view(h, m0, m1, true);
  const MetView &g = h.g;
  const size_t nnode = (size_t)m0->nx * m0->ny * m0->np, ncol = (size_t)m0->nx * m0->ny;
  std::vector<float2> x2[22], x3[9];
  for (int f = 0; f < 22; f++)
    if (m0->x2[f] && m1->x2[f]) { x2[f].resize(ncol); for (size_t i = 0; i < ncol; i++) x2[f][i] = make_float2(m0->x2[f][i], m1->x2[f][i]); }
  for (int f = 0; f < 9; f++)
    if (m0->x3[f] && m1->x3[f]) { x3[f].resize(nnode); for (size_t i = 0; i < nnode; i++) x3[f][i] = make_float2(m0->x3[f][i], m1->x3[f][i]); }
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    auto set = [&](int slot, double v) { if (qnt[slot] >= 0) q[(long long)qnt[slot] * q_stride + ip] = v; };
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    CubeT<true> cube;
    cube_reset(cube);
    MeteoValues m;
    meteo_at(g, a, cube, m);
    set(0, m.ps); set(1, m.pbl); set(2, a.p); set(3, m.t); set(4, 100. * a.p / (kRA * m.t)); set(5, m.u); set(6, m.v); set(7, m.w);
    set(8, sqrt(m.u * m.u + m.v * m.v)); set(9, -1e3 * kH0 / a.p * m.w); set(10, potential_temperature(a.p, m.t));
    set(11, saturation_pressure(m.t)); set(12, saturation_pressure_ice(m.t)); set(13, zeta_diagnosed(m.ps, a.p, m.t));
    Stencil s;
    locate(g, a.lon, a.lat, a.p, cube, s);
    const double wt = time_weight(g, a.time);
    for (int f = 0; f < 22; f++) if (qnt[14 + f] >= 0 && !x2[f].empty()) set(14 + f, field2_at(g, x2[f].data(), s, wt));
    for (int f = 0; f < 9; f++) if (qnt[36 + f] >= 0 && !x3[f].empty()) set(36 + f, field3_at(g, x3[f].data(), s, wt));
    bool moist = false;
    for (int k = 45; k <= 52; k++) moist |= qnt[k] >= 0;
    if (moist && !x3[2].empty()) {
      MoistValues w;
      moist_at(a.p, m.t, field3_at(g, x3[2].data(), s, wt), w);
      set(45, w.pw); set(46, w.sh); set(47, w.rh); set(48, w.rhice); set(49, w.tvirt); set(50, w.lapse); set(51, w.tdew); set(52, w.tice);
    }
  }
  return 0;
}

// module_convection: r[ip] are the uniform random numbers of the module's module_rng call
extern "C" int emu_convection(const EmuMet *m0, const EmuMet *m1, double conv_cape, double conv_cin, double conv_pbl_trans,
                              int conv_mix_pbl, long long np, const double *time, const double *lon, const double *lat,
                              double *p, const double *dt, const double *r) {
  HostMet h;
  make_view(h, m0, m1, true);
  const size_t ncol = (size_t)m0->nx * m0->ny;
  std::vector<float2> f[3];
  const int idx[3] = {19, 20, 18};   // cape, cin, pel
  for (int k = 0; k < 3; k++)
    if (m0->x2[idx[k]] && m1->x2[idx[k]]) {
      f[k].resize(ncol);
      for (size_t i = 0; i < ncol; i++) f[k][i] = make_float2(m0->x2[idx[k]][i], m1->x2[idx[k]][i]);
    }
  ConvView k = {conv_cape, conv_cin, conv_pbl_trans, conv_mix_pbl, f[0].data(), f[1].data(), f[2].data()};
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    if (dt[ip] == 0) continue;
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    convect(h.g, k, r[ip], a);
    p[ip] = a.p;
  }
  return 0;
}

extern "C" int emu_decay(int coord_type, double utm_ref_lat, double tdec_trop, double tdec_strat, int ntime, int nlat,
                         const double *cl_time, const double *cl_lat, const double *cl_tropo, long long np, const double *time,
                         const double *lat, const double *p, const double *dt, double *aux, double *tdec) {
  ClimView cl = {cl_time, cl_lat, cl_tropo, ntime, nlat};
  for (long long ip = 0; ip < np; ip++) {
    Parcel a = {time[ip], 0.0, lat[ip], p[ip]};
    aux[ip] = decay_factor(cl, coord_type, utm_ref_lat, tdec_trop, tdec_strat, a, dt[ip], tdec[ip]);
  }
  return 0;
}

extern "C" int emu_isosurf(const EmuMet *m0, const EmuMet *m1, int mode, int init, long long np, const double *time, const double *lon,
                           const double *lat, double *p, double *iso_var, const double *ts, const double *ps, int n) {
  HostMet h;
  make_view(h, m0, m1, true);
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    if (init) iso_var[ip] = isosurf_variable(h.g, mode, a);
    else p[ip] = isosurf_pressure(h.g, mode, iso_var[ip], a, ts, ps, n);
  }
  return 0;
}

extern "C" int emu_diff_pbl(const EmuMet *m0, const EmuMet *m1, unsigned long long ctr, long long np, double *time, double *lon,
                            double *lat, double *p, const double *dt, float *uvwp) {
  HostMet h;
  make_view(h, m0, m1, true);
  const size_t nnode = (size_t)m0->nx * m0->ny * m0->np, ncol = (size_t)m0->nx * m0->ny;
  std::vector<float2> e(ncol), n(ncol), sh(ncol), w(nnode);
  for (size_t i = 0; i < ncol; i++) {
    e[i] = make_float2(m0->x2[4][i], m1->x2[4][i]); n[i] = make_float2(m0->x2[5][i], m1->x2[5][i]);
    sh[i] = make_float2(m0->x2[6][i], m1->x2[6][i]);
  }
  for (size_t i = 0; i < nnode; i++) w[i] = make_float2(m0->x3[2][i], m1->x3[2][i]);
  const PblFields f = {e.data(), n.data(), sh.data(), w.data()};
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++) {
    if (dt[ip] == 0) continue;
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    diffuse_pbl(h.g, f, ctr, dt[ip], (uint64_t)ip, a, uvwp[3 * ip], uvwp[3 * ip + 1], uvwp[3 * ip + 2]);
    lon[ip] = a.lon; lat[ip] = a.lat; p[ip] = a.p;
  }
  return 0;
}

extern "C" int emu_bound_applies(const EmuMet *m0, const EmuMet *m1, const double *k7, int pbl, long long np, const double *time,
                                 const double *lon, const double *lat, const double *p, int *hit) {
  HostMet h;
  make_view(h, m0, m1, true);
  const BoundView k = {k7[0], k7[1], k7[2], k7[3], k7[4], k7[5], k7[6], pbl};
  for (long long ip = 0; ip < np; ip++) {
    Parcel a = {time[ip], lon[ip], lat[ip], p[ip]};
    hit[ip] = bound_applies(h.g, k, a) ? 1 : 0;
  }
  return 0;
}
extern "C" double emu_series_at(const double *tm, const double *v, int n, double t) { return series_at(tm, v, n, t); }

// module_chem_grid with the box masses summed in parcel order (the device sums them with atomics)
extern "C" int emu_chem_grid(const EmuMet *m0, const EmuMet *m1, const double *k13, int nx, int ny, int nz, int nens, long long np,
                             const double *time, const double *lon, const double *lat, const double *p, const double *m,
                             const double *ens, double *cx) {
  HostMet h;
  make_view(h, m0, m1, true);
  ChemGrid k = {k13[0], k13[1], k13[2], k13[3], k13[4], k13[5], k13[6], k13[7], k13[8], k13[9], k13[10], k13[11], k13[12], nx, ny, nz};
  const int ngrid = nx * ny * nz;
  std::vector<double> mass((size_t)ngrid * (nens > 0 ? nens : 1), 0.0);
  std::vector<int> box((size_t)np);
  for (long long ip = 0; ip < np; ip++) {
    box[ip] = chem_box(k, time[ip], lon[ip], lat[ip], p[ip]);
    if (box[ip] >= 0) mass[box[ip] + (nens > 0 ? (int)ens[ip] * ngrid : 0)] += m[ip];
  }
#pragma omp parallel for
  for (long long ip = 0; ip < np; ip++)
    if (box[ip] >= 0) cx[ip] = chem_vmr(h.g, k, box[ip], mass[box[ip] + (nens > 0 ? (int)ens[ip] * ngrid : 0)]);
  return 0;
}
