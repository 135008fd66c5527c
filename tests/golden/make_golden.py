#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE.

Run in the development container only (needs /root/reference and oracle/_ref built by
oracle/build_ref.sh).  The GPU box has neither; there the tests read the .npz files this writes.

Fixtures
  sedi_kat.npz     the 144 known-answer values of tests/tools_test/data.ref/sedi.tab
  dt_test.npz      tests/dt_test: met (ERA-Interim, read + pre-processed by the reference, cropped in
                   latitude/level to the rows the parcels touch), first NP_KEEP parcels of atm_split.tab, the
                   reference's binary results after each of the 7 steps, and the reference's own shipped
                   goldens (data.ref/atm_pl_*.tab, %g text) for the same parcels
  coord_test.npz   tests/coord_test (Cartesian UTM met, 3 hourly levels, 13 snapshots), same idea
  synth_full.npz   a small synthetic case (north->south latitudes, RK4 + turbulent + mesoscale diffusion +
                   sedimentation + mixing) with the reference's binary results.  (No module_sort here: the reference's
                   gsl_sort_index is not stable, so with diffusion on its result depends on the tie order.)
"""
import glob
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import ctypes as C  # noqa: E402

from mptrac_b200 import synth  # noqa: E402
from mptrac_b200.host import Ctl, Met  # noqa: E402
from oracle.oracle import Oracle, Parcels, Reference  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
NP_KEEP = 2000


def read_tab(path):
    return np.loadtxt(path, comments="#", ndmin=2)


def ref_read_met(ref, path, slot):
    return ref.read_met(path, slot, Met)


def ref_read_atm(ref, path):
    return ref.read_atm(path)  # time, p, lon, lat


def crop_met(m: Met, iy0, iy1, iz0, iz1) -> Met:
    sl = (slice(None), slice(iy0, iy1), slice(iz0, iz1))
    return Met(time=m.time, lon=m.lon, lat=m.lat[iy0:iy1], p=m.p[iz0:iz1], u=m.u[sl], v=m.v[sl], w=m.w[sl], t=m.t[sl],
               ps=m.ps[:, iy0:iy1], pbl=m.pbl[:, iy0:iy1], coord_type=m.coord_type)


def met_arrays(prefix, m: Met):
    return {f"{prefix}_{k}": getattr(m, k) for k in ("lon", "lat", "p", "u", "v", "w", "t", "ps", "pbl")} | {
        f"{prefix}_time": np.float64(m.time), f"{prefix}_coord_type": np.int32(m.coord_type)}


def snapshots(a: Parcels):
    return np.stack([a.time, a.p, a.lon, a.lat]).copy()


def Z(p):
    return 7.0 * np.log(1013.25 / p)


def check_ascii(tag, snap, tab, n):
    """snap [4][n] binary reference result vs. the %g text the reference shipped (cols time, z, lon, lat)."""
    err_t = np.max(np.abs(snap[0][:n] - tab[:n, 0]))
    err_z = np.max(np.abs(Z(snap[1][:n]) - tab[:n, 1]) / np.maximum(np.abs(tab[:n, 1]), 1e-30))
    err_x = np.max(np.abs(snap[2][:n] - tab[:n, 2]) / np.maximum(np.abs(tab[:n, 2]), 1e-30))
    err_y = np.max(np.abs(snap[3][:n] - tab[:n, 3]) / np.maximum(np.abs(tab[:n, 3]), 1e-30))
    print(f"  {tag}: vs shipped golden  dt={err_t:.3g} z-rel={err_z:.3g} x-rel={err_x:.3g} y-rel={err_y:.3g}")
    assert err_t < 0.006 and max(err_z, err_x, err_y) < 1e-5, "harness run does not reproduce the shipped golden"


def gen_sedi(ref):
    vals = {}
    rows = []
    for line in open(REF / "tests/tools_test/data.ref/sedi.tab"):
        if "=" not in line:
            continue
        k, v = line.split("=")
        vals[k.strip()] = float(v.split()[0])
        if k.strip() == "Re":
            rows.append([vals["p"], vals["T"], vals["r_p"], vals["rho_p"], vals["v_s"]])
    rows = np.array(rows)
    assert rows.shape == (144, 5)
    exact = np.array([ref.sedi(*r[:4]) for r in rows])
    assert np.max(np.abs(exact - rows[:, 4]) / rows[:, 4]) < 1e-5
    np.savez_compressed(OUT / "sedi_kat.npz", p=rows[:, 0], T=rows[:, 1], rp=rows[:, 2], rhop=rows[:, 3],
                        vs_text=rows[:, 4], vs_exact=exact)
    print("sedi_kat.npz: 144 known answers")


def run_steps(ref, ctl, atm, t_start, nsteps):
    out = []
    ref.ctr = 0
    for s in range(nsteps):
        ref.run("timestep", ctl, atm, t=t_start + s * ctl.dt_mod, nsteps=1)
        out.append(snapshots(atm))
    return np.stack(out)


def gen_dt_test(ref):
    print("dt_test:")
    t0 = 360547200.0
    nq = ref.read_ctl(["t", "u", "v", "w"], "DT_MOD 10.0 DIFFUSION 1 DT_MET 86400.0 T_STOP 360547260")
    assert nq == 4
    m0 = ref_read_met(ref, REF / "tests/data/ei_2011_06_05_00.nc", 0)
    m1 = ref_read_met(ref, REF / "tests/data/ei_2011_06_06_00.nc", 1)
    tm, p, lon, lat = ref_read_atm(ref, REF / "tests/dt_test/data.ref/atm_split.tab")
    n = tm.size
    ctl = Ctl(nq=4, advect=2, diffusion=1, dt_mod=10.0, dt_met=86400.0, t_start=t0, t_stop=t0 + 60.0)
    atm = Parcels(tm, p, lon, lat, np.zeros((4, n)))
    full = run_steps(ref, ctl, atm, t0, 7)
    tabs = sorted(glob.glob(str(REF / "tests/dt_test/data.ref/atm_pl_*.tab")))
    assert len(tabs) == 7
    shipped = np.stack([read_tab(f)[:, :4] for f in tabs])
    shipped_q = np.stack([read_tab(f)[:, 4:8] for f in tabs])     # the quantities t, u, v, w module_meteo wrote
    for s in range(7):
        check_ascii(Path(tabs[s]).name, full[s], shipped[s], n)

    # crop: keep the latitude rows / levels the parcels can touch (+ margin); longitudes stay global
    lat_lo, lat_hi = full[:, 3].min() - 7, full[:, 3].max() + 7
    p_lo, p_hi = full[:, 1].min() * 0.7, full[:, 1].max() * 1.4
    iy = np.where((m0.lat >= lat_lo) & (m0.lat <= lat_hi))[0]
    iz = np.where((m0.p >= p_lo) & (m0.p <= p_hi))[0]
    c0, c1 = (crop_met(m, iy[0], iy[-1] + 1, iz[0], iz[-1] + 1) for m in (m0, m1))
    ref.set_met(c0, c1)
    k = NP_KEEP
    sub = Parcels(tm[:k], p[:k], lon[:k], lat[:k], np.zeros((4, k)))
    # random numbers are addressed by parcel index, so a prefix of the parcels sees the same numbers as long as
    # the counter advances as if all n were present: emulate that by stepping the counter by hand
    res, res_q = [], []
    cm = Ctl(**{**ctl.__dict__, "met_dt_out": 0.1, "qnt_meteo": ref.qnt_meteo})
    ref.ctr = 0
    for s in range(7):
        base = ref.ctr
        a = sub  # in place
        # two module_rng calls per step (diff_turb, diff_meso), each 3n+1 counters
        ref.ctr = base
        c = Ctl(**{**ctl.__dict__})
        ref.run("timesteps", c, a, t=t0 + s * 10.0)
        ref.run("position", c, a)
        ref.run("advect", c, a)
        ref.ctr = base
        ref.run("diff_turb", c, a)
        ref.ctr = base + 3 * n + 1
        ref.run("diff_meso", c, a)
        ref.ctr = base + 2 * (3 * n + 1)
        ref.run("position", c, a)
        ref.run("meteo", cm, a)          # t, u, v, w at the new positions (module_meteo, every step: MET_DT_OUT 0.1)
        res.append(snapshots(a))
        res_q.append(a.q.copy())
    res = np.stack(res)
    assert np.array_equal(res, full[:, :, :k]), "cropped / prefix run differs from the full reference run"
    tt, tl, tr = ref.clim_tropo()
    np.savez_compressed(OUT / "dt_test.npz", **met_arrays("m0", c0), **met_arrays("m1", c1),
                        time=tm[:k], p=p[:k], lon=lon[:k], lat=lat[:k], np_total=np.int64(n), t_start=t0,
                        ref_binary=res, ref_shipped_text=shipped[:, :k, :],
                        ref_binary_q=np.stack(res_q), ref_shipped_q=shipped_q[:, :k, :],
                        tropo_time=tt, tropo_lat=tl, tropo=tr)
    print(f"  dt_test.npz: grid {c0.u.shape}, {k} of {n} parcels, 7 steps, bit-identical to the full run")


def gen_coord_test(ref):
    print("coord_test:")
    t0 = 799372800.0
    nq = ref.read_ctl(["t", "u", "v", "w"], "TRACER_CHEM 0 DIFFUSION 1 DT_MET 3600.0 T_STOP 799380000 MET_CAPE 0 DT_MOD 600 "
                      "MET_COORD_TYPE 1 MET_UTM_REF_LON 11.5692782 MET_UTM_REF_LAT 48.1507476")
    assert nq == 4
    mets = [ref_read_met(ref, REF / f"tests/data/era5_utm32_2025_05_01_{h:02d}.nc", 0) for h in range(3)]
    tabs = sorted(glob.glob(str(REF / "tests/coord_test/data.ref/atm_2025_05_01_*.tab")))
    assert len(tabs) == 13
    shipped = np.stack([read_tab(f)[:, :4] for f in tabs])
    shipped_q = np.stack([read_tab(f)[:, 4:8] for f in tabs])     # the quantities t, u, v, w module_meteo wrote
    # the t0 snapshot is the initial state (written after the dt = 0 step); read it through the reference reader
    tm, p, lon, lat = ref_read_atm(ref, REF / "tests/coord_test/data.ref/atm_2025_05_01_00_00_00.tab")
    n = tm.size
    print(f"  note: initial state taken from the %g snapshot at t0 ({n} parcels): later snapshots agree with the "
          f"shipped text only to the precision that rounding of the start positions allows")
    ctl = Ctl(nq=4, advect=2, diffusion=1, dt_mod=600.0, dt_met=3600.0, t_start=t0, t_stop=t0 + 7200.0, met_coord_type=1,
              met_utm_ref_lat=48.1507476)
    atm = Parcels(tm, p, lon, lat, np.zeros((4, n)))
    out = []
    ref.ctr = 0
    for s in range(13):
        t = t0 + s * 600.0
        lvl = min(int((t - t0 - 1e-9) // 3600.0), 1) if s > 0 else 0   # mptrac_get_met: swap when t > met1->time
        ref.set_met(mets[lvl], mets[lvl + 1])
        ref.run("timestep", ctl, atm, t=t, nsteps=1)
        out.append(snapshots(atm))
    out = np.stack(out)
    dev = [np.max(np.abs(out[s, 2] - shipped[s, :, 2])) for s in range(13)]
    devy = [np.max(np.abs(out[s, 3] - shipped[s, :, 3])) for s in range(13)]
    print("  max |x - shipped| per snapshot [m]:", " ".join(f"{d:.3g}" for d in dev))
    print("  max |y - shipped| per snapshot [m]:", " ".join(f"{d:.3g}" for d in devy))
    tt, tl, tr = ref.clim_tropo()
    d = {}
    for i, m in enumerate(mets):
        d.update(met_arrays(f"m{i}", m))
    np.savez_compressed(OUT / "coord_test.npz", **d, time=tm, p=p, lon=lon, lat=lat, t_start=t0, ref_binary=out,
                        ref_shipped_text=shipped, tropo_time=tt, tropo_lat=tl, tropo=tr)
    print(f"  coord_test.npz: grid {mets[0].u.shape}, {n} parcels, 13 steps")


def gen_synth(ref):
    print("synth_full:")
    nq = ref.read_ctl(["rp", "rhop", "m"], "")
    assert nq == 3 and ref.qnt["rp"] == 0 and ref.qnt["rhop"] == 1 and ref.qnt["m"] == 2
    m0, m1 = synth.make_met_pair(24, 13, 20, t0=0.0, dt_met=21600.0, lat_descending=True)
    n = 1500
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, zmin=0.1, zmax=45.0, seed=7)
    rng = np.random.default_rng(3)
    q = np.stack([np.full(n, 2.5), np.full(n, 1800.0), rng.uniform(0.0, 1.0, n)])
    tt, tl, tr = ref.clim_tropo()
    ctl = Ctl(nq=3, qnt_rp=0, qnt_rhop=1, advect=4, diffusion=1, t_start=0.0, t_stop=86400.0, dt_mod=600.0, dt_met=21600.0,
              turb_dz_trop=0.5, turb_dz_pbl=1.0, turb_dx_strat=20.0, turb_pbl_trans=0.2, mixing_trop=0.3, mixing_strat=0.1,
              mixing_dt=1200.0, mix_qnt=[2], mixing_nx=18, mixing_ny=9, mixing_nz=12)
    ref.set_met(m0, m1)
    atm = Parcels(tm, p, lon, lat, q)
    out, outq, outu = [], [], []
    ref.ctr = 0
    for s in range(8):
        ref.run("timestep", ctl, atm, t=s * 600.0, nsteps=1)
        out.append(snapshots(atm)); outq.append(atm.q.copy()); outu.append(atm.uvwp.copy())
    np.savez_compressed(OUT / "synth_full.npz", time=tm, p=p, lon=lon, lat=lat, q=q, ref_binary=np.stack(out),
                        ref_q=np.stack(outq), ref_uvwp=np.stack(outu), ref_ctr=np.uint64(ref.ctr),
                        tropo_time=tt, tropo_lat=tl, tropo=tr, ctl=np.array(repr(ctl.__dict__)))
    # the oracle must agree bit for bit right here, or the fixture is not worth committing
    orc = Oracle()
    b = Parcels(tm, p, lon, lat, q)
    for s in range(8):
        orc.run("timestep", ctl, (tt, tl, tr), m0, m1, b, t=s * 600.0, nsteps=1)
        same_pos = np.array_equal(snapshots(b), out[s])
        print(f"  step {s}: oracle == reference: positions {same_pos}, q {np.array_equal(b.q, outq[s])}, "
              f"uvwp {np.array_equal(b.uvwp, outu[s])}")
    print(f"  synth_full.npz: grid {m0.u.shape}, {n} parcels, 8 steps")


if __name__ == "__main__":
    ref = Reference()
    only = set(sys.argv[1:])
    for name, fn in (("sedi", gen_sedi), ("dt_test", gen_dt_test), ("coord_test", gen_coord_test), ("synth", gen_synth)):
        if not only or name in only:
            fn(ref)
