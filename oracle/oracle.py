"""TEST INFRASTRUCTURE -- ctypes bindings of the CPU oracle (oracle/_build/libmptrac_oracle.so) and, where it
was built (oracle/_ref), of the harness around the unmodified reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing in mptrac_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "libmptrac_oracle.so"
HARNESS_SO = HERE / "_ref" / "lib" / "libref_harness.so"
REF_BIN = HERE / "_ref" / "bin"
MIX_MAXQ = 23

_INT_FIELDS = ("direction", "met_coord_type", "advect", "advect_vert_coord", "rng_type", "diffusion",
               "turb_pbl_scheme", "nq", "qnt_rp", "qnt_rhop", "qnt_ens", "nens",
               "mixing_nx", "mixing_ny", "mixing_nz", "n_mix_qnt")
_DBL_FIELDS = ("t_start", "t_stop", "dt_mod", "dt_met", "met_utm_ref_lat", "sort_dt",
               "turb_dx_pbl", "turb_dx_trop", "turb_dx_strat", "turb_dz_pbl", "turb_dz_trop", "turb_dz_strat",
               "turb_mesox", "turb_mesoz", "turb_pbl_trans", "mixing_dt", "mixing_trop", "mixing_strat",
               "mixing_lon0", "mixing_lon1", "mixing_lat0", "mixing_lat1", "mixing_z0", "mixing_z1", "met_dt_out")
# module_convection / module_decay block at the end of orc_ctl_t, with the reference's defaults
_TAIL_DBL = ("conv_cape", "conv_cin", "conv_pbl_trans", "conv_dt", "tdec_trop", "tdec_strat")
_TAIL_INT = ("conv_mix_pbl", "qnt_m", "qnt_vmr", "qnt_mloss_decay", "qnt_loss_rate", "isosurf")
_TAIL_DEFAULT = dict(conv_cape=-999.0, conv_cin=-999.0, conv_pbl_trans=0.0, conv_dt=-999.0, tdec_trop=0.0, tdec_strat=0.0,
                     conv_mix_pbl=0, qnt_m=-1, qnt_vmr=-1, qnt_mloss_decay=-1, qnt_loss_rate=-1, isosurf=0)
# module_bound_cond block
_BOUND_DBL = ("bound_mass", "bound_mass_trend", "bound_vmr", "bound_vmr_trend", "bound_lat0", "bound_lat1", "bound_p0", "bound_p1",
              "bound_dps", "bound_dzs", "bound_zetas")
_BOUND_DEFAULT = dict(bound_mass=-999.0, bound_mass_trend=0.0, bound_vmr=-999.0, bound_vmr_trend=0.0, bound_lat0=-999.0, bound_lat1=-999.0,
                      bound_p0=-999.0, bound_p1=-999.0, bound_dps=-999.0, bound_dzs=-999.0, bound_zetas=-999.0)
_CHEM_DBL = ("chemgrid_lon0", "chemgrid_lon1", "chemgrid_lat0", "chemgrid_lat1", "chemgrid_z0", "chemgrid_z1", "molmass")
_CHEM_INT = ("chemgrid_nx", "chemgrid_ny", "chemgrid_nz", "qnt_Cx", "chemgrid")
_CHEM_DEFAULT = dict(chemgrid_lon0=-180.0, chemgrid_lon1=180.0, chemgrid_lat0=-90.0, chemgrid_lat1=90.0, chemgrid_z0=-5.0, chemgrid_z1=85.0,
                     molmass=-999.0, chemgrid_nx=360, chemgrid_ny=180, chemgrid_nz=90, qnt_Cx=-1, chemgrid=0)
CTS_SPECIES = ("Cccl4", "Cccl3f", "Cccl2f2", "Cn2o", "Csf6")


class OrcCts(C.Structure):
    _fields_ = [("n", C.c_int32 * 5), ("_pad", C.c_int32), ("time", C.c_void_p * 5), ("vmr", C.c_void_p * 5)]


def cts_struct(series):
    """series: dict species -> (time, vmr); returns (OrcCts, keep-alive list)"""
    s, keep = OrcCts(), []
    for k, name in enumerate(CTS_SPECIES):
        if series and name in series:
            t, v = (np.ascontiguousarray(x, np.float64) for x in series[name])
            keep += [t, v]
            s.n[k], s.time[k], s.vmr[k] = t.size, t.ctypes.data, v.ctypes.data
    return s, keep


# slots of orc_ctl_t::qnt_meteo (mptrac_oracle.h): 14 from the path's own fields, the 22 2-D and 9 3-D further fields of
# INTPOL_TIME_ALL, 8 derived from t and h2o
METEO_QNT = ("ps", "pbl", "p", "t", "rho", "u", "v", "w", "vh", "vz", "theta", "psat", "psice", "zeta_d",
             "ts", "zs", "us", "vs", "ess", "nss", "shf", "lsm", "sst", "pt", "tt", "zt", "h2ot", "pct", "pcb", "cl", "plcl", "plfc",
             "pel", "cape", "cin", "o3c",
             "zg", "pv", "h2o", "o3", "lwc", "rwc", "iwc", "swc", "cc",
             "pw", "sh", "rh", "rhice", "tvirt", "lapse", "tdew", "tice")
METEO_SLOTS = 64
MET_X2 = ("ts", "zs", "us", "vs", "ess", "nss", "shf", "lsm", "sst", "pt", "tt", "zt", "h2ot", "pct", "pcb", "cl", "plcl", "plfc", "pel", "cape", "cin", "o3c")   # orc_met_t::x2, [nx][ny]
MET_X3 = ("z", "pv", "h2o", "o3", "lwc", "rwc", "iwc", "swc", "cc")   # orc_met_t::x3, [nx][ny][np]


class OrcCtl(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in _INT_FIELDS] + [("mix_qnt", C.c_int32 * MIX_MAXQ), ("_pad", C.c_int32)]
                + [(n, C.c_double) for n in _DBL_FIELDS] + [("qnt_meteo", C.c_int32 * METEO_SLOTS)]
                + [("qnt_zeta", C.c_int32), ("qnt_eta", C.c_int32)]
                + [(n, C.c_double) for n in _TAIL_DBL] + [(n, C.c_int32) for n in _TAIL_INT]
                + [(n, C.c_double) for n in _BOUND_DBL] + [("bound_pbl", C.c_int32), ("qnt_aoa", C.c_int32),
                                                          ("qnt_cts", C.c_int32 * 5), ("cts_on", C.c_int32)]
                + [(n, C.c_double) for n in _CHEM_DBL] + [(n, C.c_int32) for n in _CHEM_INT] + [("_pad3", C.c_int32)])


_LEVEL_FIELDS = ("pl", "ul", "vl", "wl", "zetal", "zeta_dotl")   # model-level fields, [nx][ny][npl]


class OrcMet(C.Structure):
    _fields_ = [("time", C.c_double), ("coord_type", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("np", C.c_int32)] + [
        (n, C.c_void_p) for n in ("lon", "lat", "p", "u", "v", "w", "t", "ps", "pbl")] + [("npl", C.c_int32)] + [
        (n, C.c_void_p) for n in _LEVEL_FIELDS] + [("x2", C.c_void_p * len(MET_X2)), ("x3", C.c_void_p * len(MET_X3))]


class OrcClim(C.Structure):
    _fields_ = [("ntime", C.c_int32), ("nlat", C.c_int32), ("time", C.c_void_p), ("lat", C.c_void_p), ("tropo", C.c_void_p)]


class OrcAtm(C.Structure):
    _fields_ = [("np", C.c_int64), ("time", C.c_void_p), ("p", C.c_void_p), ("lon", C.c_void_p), ("lat", C.c_void_p),
                ("q", C.c_void_p), ("q_stride", C.c_int64), ("dt", C.c_void_p), ("uvwp", C.c_void_p), ("rs", C.c_void_p),
                ("iso_var", C.c_void_p), ("iso_n", C.c_int32), ("_pad", C.c_int32), ("iso_ts", C.c_void_p), ("iso_ps", C.c_void_p)]


def ctl_struct(ctl) -> OrcCtl:
    """Build the C struct from any object with the ctl field names (e.g. mptrac_b200.Ctl) or a dict."""
    get = (lambda k: ctl[k]) if isinstance(ctl, dict) else (lambda k: getattr(ctl, k))
    s = OrcCtl()
    for n in _INT_FIELDS:
        if n != "n_mix_qnt":
            setattr(s, n, int(get(n)))
    for n in _DBL_FIELDS:
        setattr(s, n, float(get(n)))
    mq = list(get("mix_qnt"))
    s.n_mix_qnt = len(mq)
    for i, v in enumerate(mq):
        s.mix_qnt[i] = int(v)
    try:
        qm = dict(get("qnt_meteo"))
    except (KeyError, AttributeError):
        qm = {}
    for i in range(METEO_SLOTS):
        s.qnt_meteo[i] = int(qm.get(METEO_QNT[i], -1)) if i < len(METEO_QNT) else -1
    for n in ("qnt_zeta", "qnt_eta"):
        try:
            setattr(s, n, int(get(n)))
        except (KeyError, AttributeError):
            setattr(s, n, -1)
    for n in _TAIL_DBL + _TAIL_INT:
        try:
            v = get(n)
        except (KeyError, AttributeError):
            v = _TAIL_DEFAULT[n]
        setattr(s, n, float(v) if n in _TAIL_DBL else int(v))
    for n in _BOUND_DBL:
        try:
            setattr(s, n, float(get(n)))
        except (KeyError, AttributeError):
            setattr(s, n, _BOUND_DEFAULT[n])
    for n, d in (("bound_pbl", 0), ("qnt_aoa", -1), ("cts_on", 0)):
        try:
            setattr(s, n, int(get(n)))
        except (KeyError, AttributeError):
            setattr(s, n, d)
    try:
        qc = list(get("qnt_cts"))
    except (KeyError, AttributeError):
        qc = [-1] * 5
    for k in range(5):
        s.qnt_cts[k] = int(qc[k])
    for n in _CHEM_DBL + _CHEM_INT:
        try:
            v = get(n)
        except (KeyError, AttributeError):
            v = _CHEM_DEFAULT[n]
        setattr(s, n, float(v) if n in _CHEM_DBL else int(v))
    return s


def met_struct(met):
    """met: object with time, lon, lat, p, u, v, w, t, ps, pbl, coord_type (dense numpy arrays)."""
    s = OrcMet()
    s.time, s.coord_type = float(met.time), int(met.coord_type)
    s.nx, s.ny, s.np = met.lon.size, met.lat.size, met.p.size
    for n in ("lon", "lat", "p", "u", "v", "w", "t", "ps", "pbl"):
        a = getattr(met, n)
        setattr(s, n, a.ctypes.data if a is not None else None)
    s.npl = 0
    for n in _LEVEL_FIELDS:
        a = getattr(met, n, None)
        setattr(s, n, a.ctypes.data if a is not None else None)
        if a is not None:
            s.npl = a.shape[2]
    extra = getattr(met, "extra", None) or {}
    for i, n in enumerate(MET_X2):
        s.x2[i] = extra[n].ctypes.data if n in extra else None
    for i, n in enumerate(MET_X3):
        s.x3[i] = extra[n].ctypes.data if n in extra else None
    return s


class Parcels:
    """Host copy of atm_t + cache_t for np parcels."""

    def __init__(self, time, p, lon, lat, q: Optional[np.ndarray] = None, uvwp=None, dt=None):
        self.time = np.array(time, np.float64)
        self.p = np.array(p, np.float64)
        self.lon = np.array(lon, np.float64)
        self.lat = np.array(lat, np.float64)
        n = self.time.size
        self.q = np.zeros((0, n)) if q is None else np.array(q, np.float64).reshape(-1, n)
        self.uvwp = np.zeros((n, 3), np.float32) if uvwp is None else np.array(uvwp, np.float32)
        self.dt = np.zeros(n) if dt is None else np.array(dt, np.float64)
        self.rs = np.zeros(3 * n + 1)
        self.iso_var = np.zeros(n)      # cache->iso_var (module_isosurf)
        self.balloon = None             # (ts, ps) of ISOSURF 4

    @property
    def np(self):
        return self.time.size

    def copy(self) -> "Parcels":
        c = Parcels(self.time, self.p, self.lon, self.lat, self.q, self.uvwp, self.dt)
        c.iso_var[:] = self.iso_var
        c.balloon = self.balloon
        return c

    def struct(self) -> OrcAtm:
        s = OrcAtm()
        s.np = self.time.size
        for n in ("time", "p", "lon", "lat", "dt", "uvwp", "rs", "iso_var"):
            setattr(s, n, getattr(self, n).ctypes.data)
        if self.balloon is not None:
            self._bal = [np.ascontiguousarray(x, np.float64) for x in self.balloon]
            s.iso_n, s.iso_ts, s.iso_ps = self._bal[0].size, self._bal[0].ctypes.data, self._bal[1].ctypes.data
        s.q = self.q.ctypes.data if self.q.size else None
        s.q_stride = self.q.shape[1] if self.q.size else 0
        return s


def clim_struct(clim):
    """clim: (time[ntime], lat[nlat], tropo[ntime][nlat]) float64 arrays."""
    t, la, tr = clim
    s = OrcClim()
    s.ntime, s.nlat = t.size, la.size
    s.time, s.lat, s.tropo = t.ctypes.data, la.ctypes.data, tr.ctypes.data
    return s


def build_oracle():
    subprocess.run(["make", "-C", str(HERE), "-s"], check=True)


class Oracle:
    """The CPU restatement."""

    def __init__(self):
        if not ORACLE_SO.exists():
            build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        P = C.POINTER
        L.orc_sedi.restype = C.c_double
        L.orc_sedi.argtypes = [C.c_double] * 4
        L.orc_clim_tropo.restype = C.c_double
        L.orc_clim_tropo.argtypes = [P(OrcClim), C.c_double, C.c_double]
        L.orc_module_rng.argtypes = [C.c_void_p, C.c_int64, C.c_int, P(C.c_uint64)]
        self._cts = cts_struct(None)
        L.orc_set_cts(C.byref(self._cts[0]))
        L.orc_run_timestep.argtypes = [P(OrcCtl), P(OrcClim), P(OrcMet), P(OrcMet), P(OrcAtm), C.c_double, P(C.c_uint64)]
        L.orc_module_timesteps.argtypes = [P(OrcCtl), P(OrcMet), P(OrcAtm), C.c_double]
        L.orc_module_position.argtypes = [P(OrcMet), P(OrcMet), P(OrcAtm)]
        L.orc_module_advect.argtypes = [P(OrcCtl), P(OrcMet), P(OrcMet), P(OrcAtm)]
        L.orc_module_diff_turb.argtypes = [P(OrcCtl), P(OrcClim), P(OrcMet), P(OrcMet), P(OrcAtm), P(C.c_uint64)]
        L.orc_module_diff_meso.argtypes = [P(OrcCtl), P(OrcMet), P(OrcMet), P(OrcAtm), P(C.c_uint64)]
        L.orc_module_sedi.argtypes = [P(OrcCtl), P(OrcMet), P(OrcMet), P(OrcAtm)]
        L.orc_module_sort.argtypes = [P(OrcCtl), P(OrcMet), P(OrcAtm)]
        L.orc_module_meteo.argtypes = [P(OrcCtl), P(OrcMet), P(OrcMet), P(OrcAtm)]
        L.orc_module_advect_init.argtypes = [P(OrcCtl), P(OrcMet), P(OrcMet), P(OrcAtm)]
        L.orc_module_advect_init.restype = None
        L.orc_module_mixing.argtypes = [P(OrcCtl), P(OrcClim), P(OrcAtm), C.c_double]
        L.orc_sort_keys.argtypes = [P(OrcMet), P(OrcAtm), C.c_void_p]
        L.orc_intpol_met_time_3d.argtypes = [P(OrcMet), P(OrcMet), C.c_void_p, C.c_void_p] + [C.c_double] * 4 + [P(C.c_double)]
        L.orc_grid_bin.argtypes = [P(OrcAtm), C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_double] * 8 + [C.c_void_p] * 3
        for f in ("orc_module_rng", "orc_run_timestep", "orc_module_timesteps", "orc_module_position", "orc_module_advect",
                  "orc_module_diff_turb", "orc_module_diff_meso", "orc_module_sedi", "orc_module_sort", "orc_module_meteo",
                  "orc_module_mixing", "orc_sort_keys", "orc_intpol_met_time_3d", "orc_grid_bin"):
            getattr(L, f).restype = None
        self.L = L
        self.ctr = 0

    def sedi(self, p, T, rp, rhop):
        return self.L.orc_sedi(p, T, rp, rhop)

    def clim_tropo(self, clim, t, lat):
        s = clim_struct(clim)
        return self.L.orc_clim_tropo(C.byref(s), t, lat)

    def module_rng(self, n, method):
        rs = np.zeros(n + 1)
        c = C.c_uint64(self.ctr)
        self.L.orc_module_rng(rs.ctypes.data, n, method, C.byref(c))
        self.ctr = c.value
        return rs

    def intpol_met_time_3d(self, met0, met1, field, ts, p, lon, lat):
        m0, m1 = met_struct(met0), met_struct(met1)
        out = C.c_double()
        self.L.orc_intpol_met_time_3d(C.byref(m0), C.byref(m1), getattr(met0, field).ctypes.data,
                                      getattr(met1, field).ctypes.data, ts, p, lon, lat, C.byref(out))
        return out.value

    def set_cts(self, series):
        """trace-gas time series of module_bound_cond: dict species (CTS_SPECIES) -> (time, vmr)"""
        self._cts = cts_struct(series)
        self.L.orc_set_cts(C.byref(self._cts[0]))

    def run(self, what, ctl, clim, met0, met1, atm: Parcels, t=0.0, nsteps=1):
        """what: 'timestep' | 'timesteps' | 'position' | 'advect' | 'diff_turb' | 'diff_meso' | 'sedi' | 'sort' | 'mixing'."""
        c = ctl_struct(ctl)
        m0, m1 = met_struct(met0), met_struct(met1)
        cl = clim_struct(clim) if clim is not None else OrcClim()
        a = atm.struct()
        ctr = C.c_uint64(self.ctr)
        L = self.L
        if what == "timestep":
            for s in range(nsteps):
                L.orc_run_timestep(C.byref(c), C.byref(cl), C.byref(m0), C.byref(m1), C.byref(a),
                                   t + s * ctl.direction * ctl.dt_mod, C.byref(ctr))
        elif what == "timesteps":
            L.orc_module_timesteps(C.byref(c), C.byref(m0), C.byref(a), t)
        elif what == "position":
            L.orc_module_position(C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "advect":
            L.orc_module_advect(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "diff_turb":
            L.orc_module_diff_turb(C.byref(c), C.byref(cl), C.byref(m0), C.byref(m1), C.byref(a), C.byref(ctr))
        elif what == "diff_meso":
            L.orc_module_diff_meso(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a), C.byref(ctr))
        elif what == "sedi":
            L.orc_module_sedi(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "sort":
            L.orc_module_sort(C.byref(c), C.byref(m0), C.byref(a))
        elif what == "mixing":
            L.orc_module_mixing(C.byref(c), C.byref(cl), C.byref(a), t)
        elif what == "meteo":
            L.orc_module_meteo(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "advect_init":
            L.orc_module_advect_init(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "convection":
            L.orc_module_convection(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a), C.byref(ctr))
        elif what == "decay":
            L.orc_module_decay(C.byref(c), C.byref(cl), C.byref(a))
        elif what == "chem_grid":
            L.orc_module_chem_grid(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a), C.c_double(t))
        elif what == "bound_cond":
            L.orc_module_bound_cond(C.byref(c), C.byref(self._cts[0]), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "diff_pbl":
            L.orc_module_diff_pbl(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a), C.byref(ctr))
        elif what == "isosurf_init":
            L.orc_module_isosurf_init(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        elif what == "isosurf":
            L.orc_module_isosurf(C.byref(c), C.byref(m0), C.byref(m1), C.byref(a))
        else:
            raise ValueError(what)
        self.ctr = ctr.value
        return atm

    def sort_keys(self, met0, atm: Parcels):
        keys = np.zeros(atm.np, np.int32)
        m0 = met_struct(met0)
        a = atm.struct()
        self.L.orc_sort_keys(C.byref(m0), C.byref(a), keys.ctypes.data)
        return keys

    def grid_bin(self, atm: Parcels, nx, ny, nz, lon0, lon1, lat0, lat1, z0, z1, t0, t1):
        nq = atm.q.shape[0]
        nb = nx * ny * nz
        cnt = np.zeros(nb, np.int32)
        s = np.zeros((nq, nb))
        sq = np.zeros((nq, nb))
        a = atm.struct()
        self.L.orc_grid_bin(C.byref(a), nq, nx, ny, nz, lon0, lon1, lat0, lat1, z0, z1, t0, t1,
                            cnt.ctypes.data, s.ctypes.data, sq.ctypes.data)
        return cnt, s, sq


_WHAT = {"timestep": 0, "timesteps": 1, "position": 2, "advect": 3, "diff_turb": 4, "diff_meso": 5, "sedi": 6,
         "sort": 7, "mixing": 8, "meteo": 9, "advect_init": 10, "convection": 11, "decay": 12, "isosurf_init": 13, "isosurf": 14, "diff_pbl": 15, "bound_cond": 16, "chem_grid": 17}


def reference_available() -> bool:
    return HARNESS_SO.exists()


class Reference:
    """The unmodified reference, driven in memory through oracle/_ref/lib/libref_harness.so."""

    def __init__(self):
        if not reference_available():
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)")
        L = C.CDLL(str(HARNESS_SO))
        P = C.POINTER
        L.ref_read_ctl.argtypes = [C.c_char_p, C.c_char_p, P(C.c_int)]
        L.ref_clim_tropo.argtypes = [P(C.c_int), P(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_set_met.argtypes = [P(OrcMet), P(OrcMet)]
        L.ref_run.argtypes = [P(OrcCtl), P(OrcAtm), C.c_double, C.c_int, C.c_int, P(C.c_uint64)]
        L.ref_sedi.restype = C.c_double
        L.ref_sedi.argtypes = [C.c_double] * 4
        L.ref_module_rng.argtypes = [C.c_void_p, C.c_int64, C.c_int, P(C.c_uint64)]
        L.ref_dims.argtypes = [P(C.c_int)] * 5
        self.L = L
        self.ctr = 0
        self.qnt = {}

    def dims(self):
        v = [C.c_int() for _ in range(5)]
        self.L.ref_dims(*[C.byref(x) for x in v])
        return dict(zip(("EX", "EY", "EP", "NP", "NQ"), (x.value for x in v)))

    def read_ctl(self, qnt_names=(), overrides=""):
        out = (C.c_int * 81)()
        nq = self.L.ref_read_ctl(",".join(qnt_names).encode(), overrides.encode(), out)
        self.qnt = dict(zip(("rp", "rhop", "m", "vmr", "ens"), list(out)[:5]))
        self.qnt["zeta"], self.qnt["eta"], self.qnt["mloss_decay"], self.qnt["loss_rate"] = out[70], out[71], out[72], out[73]
        self.qnt["aoa"], self.qnt["Cx"] = out[74], out[80]
        self.qnt["cts"] = [out[75 + k] for k in range(5)]
        self.qnt_meteo = {n: i for n, i in zip(METEO_QNT, list(out)[5:5 + len(METEO_QNT)]) if i >= 0}   # name -> index the reference assigned
        return nq

    def set_cts(self, series):
        s, keep = cts_struct(series)
        self.L.ref_set_cts(C.byref(s))

    def clim_tropo(self):
        nt, nl = C.c_int(), C.c_int()
        t, la, tr = np.zeros(12), np.zeros(73), np.zeros(12 * 73)
        self.L.ref_clim_tropo(C.byref(nt), C.byref(nl), t.ctypes.data, la.ctypes.data, tr.ctypes.data)
        return t[:nt.value].copy(), la[:nl.value].copy(), tr[:nt.value * nl.value].reshape(nt.value, nl.value).copy()

    def set_met(self, met0, met1):
        m0, m1 = met_struct(met0), met_struct(met1)
        self.L.ref_set_met(C.byref(m0), C.byref(m1))

    def run(self, what, ctl, atm: Parcels, t=0.0, nsteps=1):
        c = ctl_struct(ctl)
        a = atm.struct()
        ctr = C.c_uint64(self.ctr)
        rc = self.L.ref_run(C.byref(c), C.byref(a), t, _WHAT[what], nsteps, C.byref(ctr))
        if rc:
            raise RuntimeError("ref_run failed")
        self.ctr = ctr.value
        return atm

    def sedi(self, p, T, rp, rhop):
        return self.L.ref_sedi(p, T, rp, rhop)

    def read_met(self, path, slot, met_cls):
        """Let the reference read + pre-process a met file (mptrac_read_met) and copy the fields of the path out as a
        dense ``met_cls`` (mptrac_b200.Met)."""
        L = self.L
        L.ref_read_met.argtypes = [C.c_char_p, C.c_int]
        if L.ref_read_met(str(path).encode(), slot) != 1:
            raise RuntimeError(f"reference could not read {path}")
        nx, ny, nz, ct = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        tm = C.c_double()
        L.ref_met_dims(slot, C.byref(nx), C.byref(ny), C.byref(nz), C.byref(ct), C.byref(tm))
        nx, ny, nz = nx.value, ny.value, nz.value
        lon, lat, p = np.zeros(nx), np.zeros(ny), np.zeros(nz)
        f3 = [np.zeros((nx, ny, nz), np.float32) for _ in range(4)]
        f2 = [np.zeros((nx, ny), np.float32) for _ in range(2)]
        L.ref_get_met.argtypes = [C.c_int] + [C.c_void_p] * 9
        L.ref_get_met(slot, *[a.ctypes.data for a in (lon, lat, p, *f3, *f2)])
        return met_cls(time=tm.value, lon=lon, lat=lat, p=p, u=f3[0], v=f3[1], w=f3[2], t=f3[3], ps=f2[0], pbl=f2[1],
                       coord_type=ct.value)

    def read_atm(self, path):
        """mptrac_read_atm -> (time, p, lon, lat)"""
        L = self.L
        L.ref_read_atm.argtypes = [C.c_char_p]
        n = L.ref_read_atm(str(path).encode())
        if n < 0:
            raise RuntimeError(f"reference could not read {path}")
        arrs = [np.zeros(n) for _ in range(4)]
        L.ref_get_atm.argtypes = [C.c_void_p] * 4
        L.ref_get_atm(*[a.ctypes.data for a in arrs])
        return arrs

    def module_rng(self, n, method):
        rs = np.zeros(n + 1)
        c = C.c_uint64(self.ctr)
        self.L.ref_module_rng(rs.ctypes.data, n, method, C.byref(c))
        self.ctr = c.value
        return rs
