"""Host-side mirror of the reference interface for the time-step path (ctypes over the C ABI).

Names follow the reference: :class:`Ctl` carries the ``ctl_t`` fields the path reads (defaults are the
ones ``mptrac_read_ctl`` sets, src/mptrac.c:6723 ff.), :class:`Met` is one ``met_t`` time level,
:class:`Engine` owns the device mirror of ``atm_t``/``cache_t``/``met_t`` and exposes
``run_timestep`` (= ``mptrac_run_timestep``, src/mptrac.c:7851) and the single ``module_*`` calls.

Nothing here computes: every method forwards to ``libmptrac_b200.so``.  If that library or a CUDA
device is missing the constructor raises :class:`MpbError` -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

MIX_MAXQ = 23
# quantities module_meteo can set on the device, in the slot order of mpb_ctl_t::qnt_meteo (MPB_Q_*)
METEO_QNT = ("ps", "pbl", "p", "t", "rho", "u", "v", "w", "vh", "vz", "theta", "psat", "psice", "zeta_d",
             "ts", "zs", "us", "vs", "ess", "nss", "shf", "lsm", "sst", "pt", "tt", "zt", "h2ot", "pct", "pcb", "cl", "plcl", "plfc",
             "pel", "cape", "cin", "o3c",
             "zg", "pv", "h2o", "o3", "lwc", "rwc", "iwc", "swc", "cc",
             "pw", "sh", "rh", "rhice", "tvirt", "lapse", "tdew", "tice")
METEO_SLOTS = 64
# further met fields module_meteo interpolates (mpb_met_view_t::x2 / x3, MPB_F2_* / MPB_F3_*): Met.extra[name]
MET_X2 = ("ts", "zs", "us", "vs", "ess", "nss", "shf", "lsm", "sst", "pt", "tt", "zt", "h2ot", "pct", "pcb", "cl", "plcl", "plfc", "pel", "cape", "cin", "o3c")   # [nx][ny]
MET_X3 = ("z", "pv", "h2o", "o3", "lwc", "rwc", "iwc", "swc", "cc")   # [nx][ny][np]; "z" is the geopotential height that quantity zg reports
# module bits of mpb_run_modules (include/mptrac_b200.h MPB_MOD_*)
(MOD_TIMESTEPS, MOD_SORT, MOD_POSITION0, MOD_ADVECT, MOD_DIFF_TURB, MOD_DIFF_MESO, MOD_SEDI, MOD_POSITION1, MOD_MIXING,
 MOD_METEO, MOD_CONVECTION, MOD_DECAY, MOD_ISOSURF, MOD_DIFF_PBL, MOD_BOUND0, MOD_BOUND1, MOD_CHEMGRID) = (1 << i for i in range(17))
MOD_ALL = 0x1ffff
_LIBDIR = Path(__file__).resolve().parent / "_lib"


class MpbError(RuntimeError):
    pass


class _CtlStruct(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in (
            "direction", "met_coord_type", "advect", "advect_vert_coord", "rng_type", "diffusion",
            "turb_pbl_scheme", "nq", "qnt_rp", "qnt_rhop", "qnt_ens", "nens",
            "mixing_nx", "mixing_ny", "mixing_nz", "n_mix_qnt")]
        + [("mix_qnt", C.c_int32 * MIX_MAXQ), ("_pad", C.c_int32)]
        + [(n, C.c_double) for n in (
            "t_start", "t_stop", "dt_mod", "dt_met", "met_utm_ref_lat", "sort_dt",
            "turb_dx_pbl", "turb_dx_trop", "turb_dx_strat", "turb_dz_pbl", "turb_dz_trop", "turb_dz_strat",
            "turb_mesox", "turb_mesoz", "turb_pbl_trans",
            "mixing_dt", "mixing_trop", "mixing_strat",
            "mixing_lon0", "mixing_lon1", "mixing_lat0", "mixing_lat1", "mixing_z0", "mixing_z1", "met_dt_out")]
        + [("qnt_meteo", C.c_int32 * METEO_SLOTS), ("qnt_zeta", C.c_int32), ("qnt_eta", C.c_int32)]
        + [(n, C.c_double) for n in ("conv_cape", "conv_cin", "conv_pbl_trans", "conv_dt", "tdec_trop", "tdec_strat")]
        + [(n, C.c_int32) for n in ("conv_mix_pbl", "qnt_m", "qnt_vmr", "qnt_mloss_decay", "qnt_loss_rate", "isosurf")]
        + [(n, C.c_double) for n in ("bound_mass", "bound_mass_trend", "bound_vmr", "bound_vmr_trend", "bound_lat0", "bound_lat1",
                                     "bound_p0", "bound_p1", "bound_dps", "bound_dzs", "bound_zetas")]
        + [("bound_pbl", C.c_int32), ("qnt_aoa", C.c_int32), ("qnt_cts", C.c_int32 * 5), ("cts_on", C.c_int32)]
        + [(n, C.c_double) for n in ("chemgrid_lon0", "chemgrid_lon1", "chemgrid_lat0", "chemgrid_lat1", "chemgrid_z0", "chemgrid_z1",
                                     "molmass")]
        + [(n, C.c_int32) for n in ("chemgrid_nx", "chemgrid_ny", "chemgrid_nz", "qnt_Cx", "chemgrid", "_pad3")]
    )


class _MetViewStruct(C.Structure):
    _fields_ = [
        ("time", C.c_double),
        ("coord_type", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("np", C.c_int32),
        ("lon", C.c_void_p), ("lat", C.c_void_p), ("p", C.c_void_p),
        ("u", C.c_void_p), ("v", C.c_void_p), ("w", C.c_void_p), ("t", C.c_void_p),
        ("ps", C.c_void_p), ("pbl", C.c_void_p),
        ("sx", C.c_int64), ("sy", C.c_int64), ("sx2", C.c_int64),
        ("npl", C.c_int32), ("_pad", C.c_int32),
        ("pl", C.c_void_p), ("ul", C.c_void_p), ("vl", C.c_void_p), ("wl", C.c_void_p),
        ("zetal", C.c_void_p), ("zeta_dotl", C.c_void_p),
        ("sxl", C.c_int64), ("syl", C.c_int64),
        ("x2", C.c_void_p * len(MET_X2)), ("x3", C.c_void_p * len(MET_X3)),
    ]


class _GridStruct(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("_pad", C.c_int32)] + [
        (n, C.c_double) for n in ("lon0", "lon1", "lat0", "lat1", "z0", "z1", "t0", "t1")]


@dataclass
class Ctl:
    """The ``ctl_t`` fields read on the path, with the reference defaults."""
    direction: int = 1
    met_coord_type: int = 0
    advect: int = 2
    advect_vert_coord: int = 0
    rng_type: int = 1
    diffusion: int = 0
    turb_pbl_scheme: int = 0
    nq: int = 0
    qnt_rp: int = -1
    qnt_rhop: int = -1
    qnt_ens: int = -1
    nens: int = 0
    mixing_nx: int = 360
    mixing_ny: int = 180
    mixing_nz: int = 90
    mix_qnt: Sequence[int] = field(default_factory=list)
    t_start: float = 0.0
    t_stop: float = 1e100
    dt_mod: float = 180.0
    dt_met: float = 3600.0
    met_utm_ref_lat: float = 0.0
    sort_dt: float = -999.0
    turb_dx_pbl: float = 50.0
    turb_dx_trop: float = 50.0
    turb_dx_strat: float = 0.0
    turb_dz_pbl: float = 0.0
    turb_dz_trop: float = 0.0
    turb_dz_strat: float = 0.1
    turb_mesox: float = 0.16
    turb_mesoz: float = 0.16
    turb_pbl_trans: float = 0.0
    mixing_dt: float = 3600.0
    mixing_trop: float = -999.0
    mixing_strat: float = -999.0
    mixing_lon0: float = -180.0
    mixing_lon1: float = 180.0
    mixing_lat0: float = -90.0
    mixing_lat1: float = 90.0
    mixing_z0: float = -5.0
    mixing_z1: float = 85.0
    met_dt_out: float = 0.0                                   # module_meteo off (NB the reference's default is 0.1 = every step)
    qnt_zeta: int = -1                                        # quantity holding zeta (ADVECT_VERT_COORD 1)
    qnt_eta: int = -1                                         # quantity holding eta (ADVECT_VERT_COORD 3)
    conv_cape: float = -999.0          # module_convection: CAPE threshold [J/kg], < 0 = off
    conv_cin: float = -999.0
    conv_pbl_trans: float = 0.0
    conv_dt: float = -999.0
    conv_mix_pbl: int = 0
    tdec_trop: float = 0.0             # module_decay: e-folding times [s], both > 0 = on
    tdec_strat: float = 0.0
    qnt_m: int = -1
    qnt_vmr: int = -1
    qnt_mloss_decay: int = -1
    qnt_loss_rate: int = -1
    isosurf: int = 0                   # module_isosurf: 1 pressure, 2 density, 3 potential temperature, 4 balloon series
    bound_mass: float = -999.0         # module_bound_cond: on when bound_lat0 < bound_lat1 and bound_p0 > bound_p1
    bound_mass_trend: float = 0.0
    bound_vmr: float = -999.0
    bound_vmr_trend: float = 0.0
    bound_lat0: float = -999.0
    bound_lat1: float = -999.0
    bound_p0: float = -999.0
    bound_p1: float = -999.0
    bound_dps: float = -999.0
    bound_dzs: float = -999.0
    bound_zetas: float = -999.0
    bound_pbl: int = 0
    qnt_aoa: int = -1
    qnt_cts: Sequence[int] = (-1, -1, -1, -1, -1)   # Cccl4, Cccl3f, Cccl2f2, Cn2o, Csf6
    cts_on: int = 0                    # bit i: species i has a time series (Engine.set_clim_ts)
    chemgrid_lon0: float = -180.0      # module_chem_grid: on with chemgrid = 1 (the reference runs it for its chemistry modules)
    chemgrid_lon1: float = 180.0
    chemgrid_lat0: float = -90.0
    chemgrid_lat1: float = 90.0
    chemgrid_z0: float = -5.0
    chemgrid_z1: float = 85.0
    molmass: float = -999.0
    chemgrid_nx: int = 360
    chemgrid_ny: int = 180
    chemgrid_nz: int = 90
    qnt_Cx: int = -1
    chemgrid: int = 0
    qnt_meteo: Dict[str, int] = field(default_factory=dict)   # quantity name (METEO_QNT) -> index, e.g. {"t": 0, "u": 1}

    def to_struct(self) -> _CtlStruct:
        s = _CtlStruct()
        for name, _ in _CtlStruct._fields_:
            if name in ("mix_qnt", "_pad", "_pad3", "n_mix_qnt", "qnt_meteo", "qnt_cts"):
                continue
            setattr(s, name, getattr(self, name))
        unknown = set(self.qnt_meteo) - set(METEO_QNT)
        if unknown:
            raise ValueError(f"module_meteo quantities not available on the device: {sorted(unknown)}")
        for i in range(METEO_SLOTS):
            s.qnt_meteo[i] = int(self.qnt_meteo.get(METEO_QNT[i], -1)) if i < len(METEO_QNT) else -1
        for i in range(5):
            s.qnt_cts[i] = int(self.qnt_cts[i])
        mq = list(self.mix_qnt)
        if len(mq) > MIX_MAXQ:
            raise ValueError("too many mixing quantities")
        s.n_mix_qnt = len(mq)
        for i, v in enumerate(mq):
            s.mix_qnt[i] = int(v)
        return s


@dataclass
class Met:
    """One met_t time level as dense host arrays: 3-D fields [nx][ny][np] float32, 2-D [nx][ny]."""
    time: float
    lon: np.ndarray
    lat: np.ndarray
    p: np.ndarray
    u: np.ndarray
    v: np.ndarray
    w: np.ndarray
    t: Optional[np.ndarray] = None
    ps: Optional[np.ndarray] = None
    pbl: Optional[np.ndarray] = None
    coord_type: int = 0
    # model-level fields [nx][ny][npl] (ADVECT_VERT_COORD 1, 2, 3): pressure, winds, omega, zeta (or eta) and its tendency
    pl: Optional[np.ndarray] = None
    ul: Optional[np.ndarray] = None
    vl: Optional[np.ndarray] = None
    wl: Optional[np.ndarray] = None
    zetal: Optional[np.ndarray] = None
    zeta_dotl: Optional[np.ndarray] = None
    # further fields for module_meteo, by name (MET_X2: [nx][ny], MET_X3: [nx][ny][np])
    extra: Dict[str, np.ndarray] = field(default_factory=dict)

    def __post_init__(self):
        self.lon = np.ascontiguousarray(self.lon, dtype=np.float64)
        self.lat = np.ascontiguousarray(self.lat, dtype=np.float64)
        self.p = np.ascontiguousarray(self.p, dtype=np.float64)
        shp3 = (self.lon.size, self.lat.size, self.p.size)
        for n in ("u", "v", "w", "t"):
            a = getattr(self, n)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                if a.shape != shp3:
                    raise ValueError(f"met field {n} has shape {a.shape}, expected {shp3}")
                setattr(self, n, a)
        for n in ("ps", "pbl"):
            a = getattr(self, n)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                if a.shape != shp3[:2]:
                    raise ValueError(f"met field {n} has shape {a.shape}, expected {shp3[:2]}")
                setattr(self, n, a)
        npl = None
        for n in ("pl", "ul", "vl", "wl", "zetal", "zeta_dotl"):
            a = getattr(self, n)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                npl = a.shape[2] if npl is None else npl
                if a.ndim != 3 or a.shape != (shp3[0], shp3[1], npl):
                    raise ValueError(f"model-level field {n} has shape {a.shape}, expected {(shp3[0], shp3[1], npl)}")
                setattr(self, n, a)
        for n in list(self.extra):
            if n not in MET_X2 and n not in MET_X3:
                raise ValueError(f"unknown met field {n!r} (known: {MET_X2 + MET_X3})")
            want = shp3 if n in MET_X3 else shp3[:2]
            a = np.ascontiguousarray(self.extra[n], dtype=np.float32)
            if a.shape != want:
                raise ValueError(f"met field {n} has shape {a.shape}, expected {want}")
            self.extra[n] = a

    def view(self) -> _MetViewStruct:
        nx, ny, nz = self.lon.size, self.lat.size, self.p.size
        s = _MetViewStruct()
        s.time = float(self.time)
        s.coord_type, s.nx, s.ny, s.np = int(self.coord_type), nx, ny, nz
        for n in ("lon", "lat", "p", "u", "v", "w", "t", "ps", "pbl"):
            a = getattr(self, n)
            setattr(s, n, a.ctypes.data if a is not None else None)
        s.sx, s.sy, s.sx2 = ny * nz, nz, ny
        npl = 0
        for n in ("pl", "ul", "vl", "wl", "zetal", "zeta_dotl"):
            a = getattr(self, n)
            setattr(s, n, a.ctypes.data if a is not None else None)
            npl = a.shape[2] if a is not None else npl
        s.npl, s.sxl, s.syl = npl, ny * npl, npl
        for i, n in enumerate(MET_X2):
            s.x2[i] = self.extra[n].ctypes.data if n in self.extra else None
        for i, n in enumerate(MET_X3):
            s.x3[i] = self.extra[n].ctypes.data if n in self.extra else None
        return s


def plan_modules(ctl: "Ctl", t: float, mask: int = MOD_ALL) -> str:
    """The launches ``Engine.run_modules(t, mask)`` would make for ``ctl`` (``mpb_plan_modules``: no device needed)."""
    lib = load_library()
    buf = C.create_string_buffer(4096)
    s = ctl.to_struct()
    if lib.mpb_plan_modules(C.byref(s), float(t), int(mask), buf, len(buf)) != 0:
        raise MpbError(lib.mpb_last_error().decode())
    return buf.value.decode()


_lib_cache = {}
IPC_HANDLE_BYTES = 64      # MPB_IPC_HANDLE_BYTES
MAX_RANKS = 16             # MPB_MAX_RANKS


def exchange_area_bytes(ctl: "Ctl", nranks: int, nq: int, np_max: int, grid_boxes: int = 0):
    """(mix_bytes, grid_bytes) of mpb_peer_init for a control structure and ``np_max`` parcels per rank (the same value on every
    rank): inboxes + outboxes of the routed mixing exchange -- one entry of (mixed quantities + 1) doubles per parcel and
    peer --; count + sum + sum of squares per box of the gridded output"""
    nmix = sum(1 for i in ctl.mix_qnt if i >= 0)
    mixing = ctl.mixing_trop >= 0 and ctl.mixing_strat >= 0 and nmix > 0
    mix = 256 + 2 * 8 * nranks * int(np_max) * (nmix + 1) if mixing else 0
    return mix, grid_boxes * (16 * max(nq, 1) + 4)


def load_library(strict: bool = False) -> C.CDLL:
    """dlopen the in-tree C-ABI library (``strict`` = the -fmad=false build used for tight parity)."""
    name = "libmptrac_b200_strict.so" if strict else "libmptrac_b200.so"
    if name in _lib_cache:
        return _lib_cache[name]
    path = Path(os.environ.get("MPTRAC_B200_LIBDIR", _LIBDIR)) / name
    if not path.exists():
        raise MpbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(str(path))
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    sig = {
        "mpb_last_error": (C.c_char_p, []),
        "mpb_abi_version": (i32, []),
        "mpb_device_count": (i32, []),
        "mpb_warmup": (i32, [i32]),
        "mpb_create": (i32, [P(vp), i32, i64, i32]),
        "mpb_destroy": (i32, [vp]),
        "mpb_set_stream": (i32, [vp, vp]),
        "mpb_sync": (i32, [vp]),
        "mpb_set_ctl": (i32, [vp, P(_CtlStruct)]),
        "mpb_set_clim_tropo": (i32, [vp, i32, i32, vp, vp, vp]),
        "mpb_set_met": (i32, [vp, i32, P(_MetViewStruct)]),
        "mpb_set_met_bin": (i32, [vp, i32, C.c_char_p, i32]),
        "mpb_swap_met": (i32, [vp]),
        "mpb_set_atm": (i32, [vp, i64, vp, vp, vp, vp, vp, i64]),
        "mpb_set_uvwp": (i32, [vp, vp]),
        "mpb_get_atm": (i32, [vp, vp, vp, vp, vp, vp, i64]),
        "mpb_get_uvwp": (i32, [vp, vp]),
        "mpb_set_iso_var": (i32, [vp, vp]),
        "mpb_get_iso_var": (i32, [vp, vp]),
        "mpb_set_balloon": (i32, [vp, i32, vp, vp]),
        "mpb_set_clim_ts": (i32, [vp, i32, i32, vp, vp]),
        "mpb_plan_modules": (i32, [vp, dbl, C.c_uint, C.c_char_p, i32]),
        "mpb_get_dt": (i32, [vp, vp]),
        "mpb_get_np": (i64, [vp]),
        "mpb_set_shard": (i32, [vp, i64, i64]),
        "mpb_set_rng_ctr": (i32, [vp, u64]),
        "mpb_get_rng_ctr": (u64, [vp]),
        "mpb_run_timestep": (i32, [vp, dbl]),
        "mpb_run_timestep_host": (i32, [vp, dbl, i64, vp, vp, vp, vp, vp, i64]),
        "mpb_host_step_bytes": (i32, [vp, P(i64), P(i64)]),
        "mpb_run_modules": (i32, [vp, dbl, C.c_uint]),
        "mpb_module_timesteps": (i32, [vp, dbl]),
        "mpb_module_position": (i32, [vp]),
        "mpb_module_advect": (i32, [vp]),
        "mpb_module_diff_turb": (i32, [vp]),
        "mpb_module_diff_meso": (i32, [vp]),
        "mpb_module_sedi": (i32, [vp]),
        "mpb_module_sort": (i32, [vp]),
        "mpb_module_meteo": (i32, [vp]),
        "mpb_module_mixing": (i32, [vp, dbl]),
        "mpb_module_rng": (i32, [vp, vp, i64, i32]),
        "mpb_mixing_accumulate_all": (i32, [vp, dbl]),
        "mpb_mixing_apply_all": (i32, [vp]),
        "mpb_mixing_nbox": (i64, [vp]),
        "mpb_mixing_rec_len": (i64, [vp]),
        "mpb_grid_accumulate": (i32, [vp, P(_GridStruct)]),
        "mpb_grid_reduce": (i32, [vp]),
        "mpb_grid_fetch": (i32, [vp, vp, vp, vp]),
        "mpb_peer_init": (i32, [vp, i32, i32, i64, i64, vp]),
        "mpb_peer_attach": (i32, [vp, vp]),
        "mpb_peer_attach_local": (i32, [vp, P(vp), P(i32)]),
        "mpb_peer_area": (vp, [vp]),
        "mpb_peer_barrier": (i32, [vp]),
        "mpb_team_create": (i32, [P(vp), i32, P(i32), i64, i32]),
        "mpb_team_destroy": (i32, [vp]),
        "mpb_team_size": (i32, [vp]),
        "mpb_team_member": (vp, [vp, i32]),
        "mpb_team_set_ctl": (i32, [vp, P(_CtlStruct)]),
        "mpb_team_set_clim_tropo": (i32, [vp, i32, i32, vp, vp, vp]),
        "mpb_team_set_clim_ts": (i32, [vp, i32, i32, vp, vp]),
        "mpb_team_set_balloon": (i32, [vp, i32, vp, vp]),
        "mpb_team_set_met": (i32, [vp, i32, P(_MetViewStruct)]),
        "mpb_team_swap_met": (i32, [vp]),
        "mpb_team_set_atm": (i32, [vp, i64, vp, vp, vp, vp, vp, i64]),
        "mpb_team_get_atm": (i32, [vp, vp, vp, vp, vp, vp, i64]),
        "mpb_team_set_uvwp": (i32, [vp, vp]),
        "mpb_team_get_uvwp": (i32, [vp, vp]),
        "mpb_team_get_dt": (i32, [vp, vp]),
        "mpb_team_set_iso_var": (i32, [vp, vp]),
        "mpb_team_get_iso_var": (i32, [vp, vp]),
        "mpb_team_get_np": (i64, [vp]),
        "mpb_team_set_rng_ctr": (i32, [vp, u64]),
        "mpb_team_get_rng_ctr": (u64, [vp]),
        "mpb_team_run_timestep": (i32, [vp, dbl]),
        "mpb_team_run_modules": (i32, [vp, dbl, C.c_uint]),
        "mpb_team_sync": (i32, [vp]),
        "mpb_team_launch_count": (i64, [vp]),
        "mpb_team_grid_accumulate": (i32, [vp, P(_GridStruct)]),
        "mpb_team_grid_fetch": (i32, [vp, vp, vp, vp]),
        "mpb_device_ptr": (vp, [vp, C.c_char_p]),
        "mpb_launch_count": (i64, [vp]),
        "mpb_met_bytes": (i32, [vp, P(i64)]),
    }
    for fn, (res, args) in sig.items():
        f = getattr(lib, fn)   # AttributeError here = the library does not export what the header declares
        f.restype, f.argtypes = res, args
    lib._mpb_symbols = tuple(sig)
    _lib_cache[name] = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Engine:
    """Device mirror of one simulation's ``atm_t`` / ``cache_t`` / ``met_t`` pair (one per GPU)."""

    def __init__(self, np_max: int, nq: int = 0, device: int = 0, strict: bool = False):
        self._lib = load_library(strict)
        self._h = C.c_void_p()
        self.nq = int(nq)
        self.np_max = int(np_max)
        self.device = int(device)
        rc = self._lib.mpb_create(C.byref(self._h), int(device), int(np_max), int(nq))
        if rc:
            self._h = C.c_void_p()
            raise MpbError(self._lib.mpb_last_error().decode())

    # -- plumbing -------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc:
            raise MpbError(self._lib.mpb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.mpb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.mpb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self._lib.mpb_sync(self._h))

    # -- update_device --------------------------------------------------------------------------
    def set_ctl(self, ctl: Ctl):
        if ctl.nq != self.nq:
            raise ValueError("ctl.nq must equal the engine's nq")
        s = ctl.to_struct()
        self._ck(self._lib.mpb_set_ctl(self._h, C.byref(s)))
        self.ctl = ctl

    def set_clim_tropo(self, time: np.ndarray, lat: np.ndarray, tropo: np.ndarray):
        time = np.ascontiguousarray(time, np.float64)
        lat = np.ascontiguousarray(lat, np.float64)
        tropo = np.ascontiguousarray(tropo, np.float64)
        if tropo.shape != (time.size, lat.size):
            raise ValueError("tropo must be [ntime][nlat]")
        self._ck(self._lib.mpb_set_clim_tropo(self._h, time.size, lat.size, _ptr(time), _ptr(lat), _ptr(tropo)))

    def set_met(self, slot: int, met: Met):
        v = met.view()
        self._ck(self._lib.mpb_set_met(self._h, int(slot), C.byref(v)))

    def set_met_bin(self, slot: int, path, all_fields: bool = False):
        """a met level straight from one of the reference's uncompressed binary files (MET_TYPE 1)"""
        self._ck(self._lib.mpb_set_met_bin(self._h, int(slot), str(path).encode(), int(bool(all_fields))))

    def swap_met(self):
        self._ck(self._lib.mpb_swap_met(self._h))

    def set_atm(self, time, p, lon, lat, q: Optional[np.ndarray] = None):
        """Upload parcels.  Arrays may be pinned torch-backed numpy views; q is [nq][np]."""
        arrs = [np.ascontiguousarray(a, np.float64) for a in (time, p, lon, lat)]
        n = arrs[0].size
        if any(a.size != n for a in arrs):
            raise ValueError("ragged parcel arrays")
        stride = 0
        if self.nq:
            if q is None:
                raise ValueError("q is required when nq > 0")
            q = np.asarray(q, np.float64)
            if q.ndim != 2 or q.shape[0] != self.nq or q.shape[1] < n or q.strides[1] != 8:
                raise ValueError("q must be [nq][>=np] float64 with contiguous rows")
            stride = q.strides[0] // 8
        self._keep = (arrs, q)
        self._ck(self._lib.mpb_set_atm(self._h, n, *[_ptr(a) for a in arrs], _ptr(q) if self.nq else None, stride))

    def set_uvwp(self, uvwp: np.ndarray):
        uvwp = np.ascontiguousarray(uvwp, np.float32)
        if uvwp.shape != (self.np, 3):
            raise ValueError("uvwp must be [np][3]")
        self._ck(self._lib.mpb_set_uvwp(self._h, _ptr(uvwp)))
        self.sync()

    def set_shard(self, global_offset: int, global_np: int):
        self._ck(self._lib.mpb_set_shard(self._h, int(global_offset), int(global_np)))

    # -- update_host ----------------------------------------------------------------------------
    @property
    def np(self) -> int:
        return int(self._lib.mpb_get_np(self._h))

    def get_atm(self, out=None):
        """Download parcels -> dict(time, p, lon, lat, q).  ``out`` may hold preallocated (pinned) arrays."""
        n = self.np
        if out is None:
            out = {k: np.empty(n, np.float64) for k in ("time", "p", "lon", "lat")}
            out["q"] = np.empty((self.nq, n), np.float64)
        q = out.get("q")
        stride = q.strides[0] // 8 if (q is not None and self.nq) else 0
        self._ck(self._lib.mpb_get_atm(self._h, _ptr(out["time"]), _ptr(out["p"]), _ptr(out["lon"]), _ptr(out["lat"]),
                                       _ptr(q) if self.nq else None, stride))
        return out

    def get_uvwp(self) -> np.ndarray:
        a = np.empty((self.np, 3), np.float32)
        self._ck(self._lib.mpb_get_uvwp(self._h, _ptr(a)))
        return a

    def set_iso_var(self, iso_var):
        a = np.ascontiguousarray(iso_var, np.float64)
        self._ck(self._lib.mpb_set_iso_var(self._h, _ptr(a)))

    def get_iso_var(self) -> np.ndarray:
        a = np.empty(self.np, np.float64)
        self._ck(self._lib.mpb_get_iso_var(self._h, _ptr(a)))
        return a

    def set_balloon(self, ts, ps):
        ts, ps = np.ascontiguousarray(ts, np.float64), np.ascontiguousarray(ps, np.float64)
        if ts.size != ps.size or ts.size < 1:
            raise ValueError("balloon series: ts and ps must have the same length >= 1")
        self._ck(self._lib.mpb_set_balloon(self._h, int(ts.size), _ptr(ts), _ptr(ps)))

    def set_clim_ts(self, species: int, time, vmr):
        """trace-gas time series of module_bound_cond; species 0 .. 4 = Cccl4, Cccl3f, Cccl2f2, Cn2o, Csf6"""
        time, vmr = np.ascontiguousarray(time, np.float64), np.ascontiguousarray(vmr, np.float64)
        if time.size != vmr.size or time.size < 1:
            raise ValueError("time series: time and vmr must have the same length >= 1")
        self._ck(self._lib.mpb_set_clim_ts(self._h, int(species), int(time.size), _ptr(time), _ptr(vmr)))

    def get_dt(self) -> np.ndarray:
        a = np.empty(self.np, np.float64)
        self._ck(self._lib.mpb_get_dt(self._h, _ptr(a)))
        return a

    # -- rng ------------------------------------------------------------------------------------
    @property
    def rng_ctr(self) -> int:
        return int(self._lib.mpb_get_rng_ctr(self._h))

    @rng_ctr.setter
    def rng_ctr(self, v: int):
        self._ck(self._lib.mpb_set_rng_ctr(self._h, int(v)))

    def module_rng(self, n: int, method: int) -> np.ndarray:
        rs = np.empty(n + 1, np.float64)
        self._ck(self._lib.mpb_module_rng(self._h, _ptr(rs), int(n), int(method)))
        return rs

    # -- the step -------------------------------------------------------------------------------
    def run_timestep(self, t: float):
        self._ck(self._lib.mpb_run_timestep(self._h, float(t)))

    def run_timestep_host(self, t: float, time, p, lon, lat, q: Optional[np.ndarray] = None):
        """One step for parcels in HOST arrays (float64, contiguous, ideally pinned), updated in place: upload, step and
        download are pipelined chunk by chunk (``mpb_run_timestep_host``)."""
        arrs = (time, p, lon, lat)
        n = arrs[0].size
        for a in arrs:
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.size != n:
                raise ValueError("parcel arrays must be contiguous float64 of one length")
        stride = 0
        if self.nq:
            if q is None or q.dtype != np.float64 or q.ndim != 2 or q.shape[0] != self.nq or q.strides[1] != 8:
                raise ValueError("q must be [nq][>=np] float64 with contiguous rows")
            stride = q.strides[0] // 8
        self._ck(self._lib.mpb_run_timestep_host(self._h, float(t), n, *[_ptr(a) for a in arrs],
                                                 _ptr(q) if self.nq else None, stride))

    @property
    def host_step_bytes(self):
        """(host -> device, device -> host) bytes of the last run_timestep_host call"""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._lib.mpb_host_step_bytes(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def module_meteo(self):
        self._ck(self._lib.mpb_module_meteo(self._h))

    def run_modules(self, t: float, mask: int):
        self._ck(self._lib.mpb_run_modules(self._h, float(t), int(mask)))

    def module_timesteps(self, t: float):
        self._ck(self._lib.mpb_module_timesteps(self._h, float(t)))

    def module_position(self):
        self._ck(self._lib.mpb_module_position(self._h))

    def module_advect(self):
        self._ck(self._lib.mpb_module_advect(self._h))

    def module_diff_turb(self):
        self._ck(self._lib.mpb_module_diff_turb(self._h))

    def module_diff_meso(self):
        self._ck(self._lib.mpb_module_diff_meso(self._h))

    def module_sedi(self):
        self._ck(self._lib.mpb_module_sedi(self._h))

    def module_sort(self):
        self._ck(self._lib.mpb_module_sort(self._h))

    def module_mixing(self, t: float):
        self._ck(self._lib.mpb_module_mixing(self._h, float(t)))

    def mixing_accumulate_all(self, t: float):
        """box index per parcel + this rank's contributions of ALL mixed quantities to the box records"""
        self._ck(self._lib.mpb_mixing_accumulate_all(self._h, float(t)))

    def mixing_apply_all(self):
        self._ck(self._lib.mpb_mixing_apply_all(self._h))

    @property
    def mixing_rec_len(self) -> int:
        return int(self._lib.mpb_mixing_rec_len(self._h))

    # -- ranks that exchange through peer memory ----------------------------------------------------
    def peer_init(self, rank: int, nranks: int, mix_bytes: int, grid_bytes: int) -> bytes:
        """allocate this rank's exchange area; returns the handle the other ranks attach"""
        h = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ck(self._lib.mpb_peer_init(self._h, int(rank), int(nranks), int(mix_bytes), int(grid_bytes), h))
        return h.raw

    def peer_attach(self, handles):
        """handles of all ranks in rank order (other processes' areas are opened through CUDA IPC)"""
        blob = b"".join(handles)
        self._ck(self._lib.mpb_peer_attach(self._h, C.c_char_p(blob)))

    def peer_attach_local(self, engines):
        """contexts of THIS process (rank order): their areas are attached directly"""
        n = len(engines)
        areas = (C.c_void_p * n)(*[e._lib.mpb_peer_area(e._h) for e in engines])
        devs = (C.c_int * n)(*[e.device for e in engines])
        self._ck(self._lib.mpb_peer_attach_local(self._h, areas, devs))

    def peer_barrier(self):
        self._ck(self._lib.mpb_peer_barrier(self._h))

    def grid_reduce(self):
        self._ck(self._lib.mpb_grid_reduce(self._h))

    @property
    def mixing_nbox(self) -> int:
        return int(self._lib.mpb_mixing_nbox(self._h))

    def grid_accumulate(self, nx, ny, nz, lon0, lon1, lat0, lat1, z0, z1, t0, t1):
        g = _GridStruct(nx, ny, nz, 0, lon0, lon1, lat0, lat1, z0, z1, t0, t1)
        self._ck(self._lib.mpb_grid_accumulate(self._h, C.byref(g)))
        self._grid_nbox = nx * ny * nz

    def grid_fetch(self):
        nb = self._grid_nbox
        cnt = np.empty(nb, np.int32)
        s = np.empty((self.nq, nb), np.float64)
        sq = np.empty((self.nq, nb), np.float64)
        self._ck(self._lib.mpb_grid_fetch(self._h, _ptr(cnt), _ptr(s), _ptr(sq)))
        return cnt, s, sq

    # -- introspection --------------------------------------------------------------------------
    def device_ptr(self, name: str) -> int:
        p = self._lib.mpb_device_ptr(self._h, name.encode())
        if not p:
            raise MpbError(f"no device array named {name!r}")
        return int(p)

    @property
    def launch_count(self) -> int:
        return int(self._lib.mpb_launch_count(self._h))

    @property
    def met_bytes(self) -> int:
        b = C.c_int64()
        self._ck(self._lib.mpb_met_bytes(self._h, C.byref(b)))
        return int(b.value)


class Team:
    """Several devices behind one host thread (``mpb_team_*``): the whole parcel set is cut into contiguous index ranges, one
    per device; the calls mirror :class:`Engine`.  ``devices`` may name the same device twice (two contexts on one GPU:
    how the multi-device logic is exercised on a single-GPU box)."""

    def __init__(self, devices: Sequence[int], np_max: int, nq: int = 0, strict: bool = False):
        self._lib = load_library(strict)
        self._h = C.c_void_p()
        self.nq, self.np_max, self.devices = int(nq), int(np_max), [int(d) for d in devices]
        devs = (C.c_int * len(self.devices))(*self.devices)
        rc = self._lib.mpb_team_create(C.byref(self._h), len(self.devices), devs, int(np_max), int(nq))
        if rc:
            self._h = C.c_void_p()
            raise MpbError(self._lib.mpb_last_error().decode())

    def _ck(self, rc: int):
        if rc:
            raise MpbError(self._lib.mpb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.mpb_team_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return int(self._lib.mpb_team_size(self._h))

    def set_ctl(self, ctl: Ctl):
        s = ctl.to_struct()
        self._ck(self._lib.mpb_team_set_ctl(self._h, C.byref(s)))
        self.ctl = ctl

    def set_clim_tropo(self, time, lat, tropo):
        time, lat, tropo = (np.ascontiguousarray(a, np.float64) for a in (time, lat, tropo))
        self._ck(self._lib.mpb_team_set_clim_tropo(self._h, time.size, lat.size, _ptr(time), _ptr(lat), _ptr(tropo)))

    def set_met(self, slot: int, met: Met):
        v = met.view()
        self._ck(self._lib.mpb_team_set_met(self._h, int(slot), C.byref(v)))

    def swap_met(self):
        self._ck(self._lib.mpb_team_swap_met(self._h))

    def set_atm(self, time, p, lon, lat, q: Optional[np.ndarray] = None):
        arrs = [np.ascontiguousarray(a, np.float64) for a in (time, p, lon, lat)]
        n = arrs[0].size
        stride = 0
        if self.nq:
            q = np.ascontiguousarray(q, np.float64)
            if q.ndim != 2 or q.shape[0] != self.nq or q.shape[1] < n:
                raise ValueError("q must be [nq][>=np] float64")
            stride = q.strides[0] // 8
        self._keep = (arrs, q)
        self._ck(self._lib.mpb_team_set_atm(self._h, n, *[_ptr(a) for a in arrs], _ptr(q) if self.nq else None, stride))

    @property
    def np(self) -> int:
        return int(self._lib.mpb_team_get_np(self._h))

    def get_atm(self):
        n = self.np
        out = {k: np.empty(n, np.float64) for k in ("time", "p", "lon", "lat")}
        out["q"] = np.empty((self.nq, n), np.float64)
        self._ck(self._lib.mpb_team_get_atm(self._h, _ptr(out["time"]), _ptr(out["p"]), _ptr(out["lon"]), _ptr(out["lat"]),
                                            _ptr(out["q"]) if self.nq else None, n))
        return out

    def set_uvwp(self, uvwp):
        uvwp = np.ascontiguousarray(uvwp, np.float32)
        self._ck(self._lib.mpb_team_set_uvwp(self._h, _ptr(uvwp)))
        self.sync()

    def get_uvwp(self) -> np.ndarray:
        a = np.empty((self.np, 3), np.float32)
        self._ck(self._lib.mpb_team_get_uvwp(self._h, _ptr(a)))
        return a

    def get_dt(self) -> np.ndarray:
        a = np.empty(self.np, np.float64)
        self._ck(self._lib.mpb_team_get_dt(self._h, _ptr(a)))
        return a

    @property
    def rng_ctr(self) -> int:
        return int(self._lib.mpb_team_get_rng_ctr(self._h))

    @rng_ctr.setter
    def rng_ctr(self, v: int):
        self._ck(self._lib.mpb_team_set_rng_ctr(self._h, int(v)))

    def run_timestep(self, t: float):
        self._ck(self._lib.mpb_team_run_timestep(self._h, float(t)))

    def run_modules(self, t: float, mask: int):
        self._ck(self._lib.mpb_team_run_modules(self._h, float(t), int(mask)))

    def sync(self):
        self._ck(self._lib.mpb_team_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.mpb_team_launch_count(self._h))

    def grid_accumulate(self, nx, ny, nz, lon0, lon1, lat0, lat1, z0, z1, t0, t1):
        g = _GridStruct(nx, ny, nz, 0, lon0, lon1, lat0, lat1, z0, z1, t0, t1)
        self._ck(self._lib.mpb_team_grid_accumulate(self._h, C.byref(g)))
        self._grid_nbox = nx * ny * nz

    def grid_fetch(self):
        nb = self._grid_nbox
        cnt = np.empty(nb, np.int32)
        s = np.empty((self.nq, nb), np.float64)
        sq = np.empty((self.nq, nb), np.float64)
        self._ck(self._lib.mpb_team_grid_fetch(self._h, _ptr(cnt), _ptr(s), _ptr(sq)))
        return cnt, s, sq
