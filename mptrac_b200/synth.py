"""Synthetic inputs of the shapes BASELINE.json names: ERA5-shaped met grids and parcel clouds.

Pure data generation (numpy): no model physics lives here.  The met fields are smooth analytic
functions (a zonal jet with planetary waves, a meridional overturning, a standard-atmosphere
temperature profile, undulating surface / boundary-layer pressures) plus a deterministic small-scale
hash perturbation so that the 16-point wind standard deviation of the mesoscale module is not zero.
Longitudes run 0..360 with the periodic wrap column appended (what read_met_periodic produces,
src/mptrac.c:11714-11771); latitude order is selectable because ERA files run north->south.
"""
from __future__ import annotations

import numpy as np

from .host import Met

H0 = 7.0
P0 = 1013.25


def _hash_noise(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape, dtype=np.float32)


def make_met(nlon=360, nlat=181, nlev=60, time=0.0, phase=0.0, ztop=60.0, lat_descending=False, seed=1,
             noise=0.5) -> Met:
    """One met time level on a (nlon+1) x nlat x nlev grid; `phase` shifts the wave pattern (radians)."""
    lon = np.linspace(0.0, 360.0, nlon + 1)
    lat = np.linspace(-90.0, 90.0, nlat)
    if lat_descending:
        lat = lat[::-1].copy()
    z = np.linspace(0.0, ztop, nlev)
    p = P0 * np.exp(-z / H0)
    lam = np.deg2rad(lon)[:, None, None]
    phi = np.deg2rad(lat)[None, :, None]
    zz = z[None, None, :]
    jet = np.exp(-((zz - 11.0) / 6.0) ** 2)
    u = (10.0 + 35.0 * jet) * np.cos(phi) * (1.0 + 0.3 * np.sin(3 * lam + phase)) + 5.0 * np.sin(2 * phi)
    v = 8.0 * np.cos(phi) * np.sin(2 * lam + phase) * (0.3 + jet)
    w = 0.02 * np.sin(lam + phase) * np.cos(2 * phi) * np.sin(np.pi * zz / ztop) * (p[None, None, :] / P0 + 0.05)
    t = np.where(zz < 11.0, 288.15 - 6.5 * zz, 216.65 + 1.0 * np.maximum(zz - 20.0, 0.0)) + 3.0 * np.cos(lam) * np.cos(phi)
    u, v, w, t = (np.broadcast_to(a, (nlon + 1, nlat, nlev)).astype(np.float32) for a in (u, v, w, t))
    if noise > 0:
        u = u + noise * _hash_noise(u.shape, seed)
        v = v + noise * _hash_noise(v.shape, seed + 1)
        w = w + 1e-3 * noise * _hash_noise(w.shape, seed + 2)
    lam2, phi2 = lam[:, :, 0], phi[:, :, 0]
    ps = 985.0 + 25.0 * np.sin(2 * lam2 + phase) * np.cos(phi2) - 60.0 * np.exp(-((phi2 - 0.6) / 0.3) ** 2) * (1 + np.cos(lam2))
    pbl = ps - 60.0 - 40.0 * np.cos(phi2) * (1.0 + 0.5 * np.cos(lam2 + phase))
    ps = np.broadcast_to(ps, (nlon + 1, nlat)).astype(np.float32)
    pbl = np.broadcast_to(pbl, (nlon + 1, nlat)).astype(np.float32)
    out = [np.array(a) for a in (u, v, w, t)]
    for a in out:            # periodic wrap column = column 0
        a[-1] = a[0]
    ps, pbl = np.array(ps), np.array(pbl)
    ps[-1], pbl[-1] = ps[0], pbl[0]
    return Met(time=time, lon=lon, lat=lat, p=p, u=out[0], v=out[1], w=out[2], t=out[3], ps=ps, pbl=pbl, coord_type=0)


def make_met_pair(nlon=360, nlat=181, nlev=60, t0=0.0, dt_met=21600.0, **kw):
    m0 = make_met(nlon, nlat, nlev, time=t0, phase=0.0, seed=1, **kw)
    m1 = make_met(nlon, nlat, nlev, time=t0 + dt_met, phase=0.35, seed=11, **kw)
    return m0, m1


def make_parcels(n, t0=0.0, seed=123, zmin=1.0, zmax=30.0):
    """Parcels uniform on the sphere and in log-pressure height: time, p, lon, lat."""
    rng = np.random.default_rng(seed)
    lon = rng.uniform(-180.0, 180.0, n)
    lat = np.rad2deg(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    z = rng.uniform(zmin, zmax, n)
    p = P0 * np.exp(-z / H0)
    time = np.full(n, float(t0))
    return time, p, lon, lat


def make_clim_tropo():
    """A smooth stand-in for the tropopause climatology table ([12 months][73 latitudes], hPa)."""
    time = (np.arange(12) + 0.5) * (365.25 * 86400.0 / 12.0)
    lat = np.linspace(-90.0, 90.0, 73)
    season = np.cos(2 * np.pi * (time[:, None] / (365.25 * 86400.0)))
    tropo = 300.0 - 200.0 * np.exp(-(lat[None, :] / 35.0) ** 2) + 15.0 * season * np.sin(np.deg2rad(lat))[None, :]
    return time, lat, np.ascontiguousarray(tropo)


def add_model_levels(met: Met, npl=None, seed=7) -> Met:
    """Model-level fields for ADVECT_VERT_COORD 1 / 2 / 3 on the grid of ``met``: level pressures that undulate with the
    surface (hybrid-coordinate like, decreasing with the level index), winds and omega sampled from the pressure-level
    fields plus small-scale structure, a zeta coordinate that increases with height and its tendency."""
    nx, ny, nz = met.u.shape
    npl = npl or nz
    rng = np.random.default_rng(seed)
    lam = np.deg2rad(met.lon)[:, None, None]
    phi = np.deg2rad(met.lat)[None, :, None]
    k = np.arange(npl)[None, None, :]
    zlev = 60.0 * (k + 0.3) / npl                                   # km
    bump = 1.0 + 0.04 * np.sin(2 * lam + 0.3 * met.time / 21600.0) * np.cos(phi) * np.exp(-zlev / 8.0)
    pl = (P0 * np.exp(-zlev / H0) * bump).astype(np.float32)
    jet = np.exp(-((zlev - 11.0) / 6.0) ** 2)
    ul = ((10.0 + 35.0 * jet) * np.cos(phi) * (1.0 + 0.3 * np.sin(3 * lam)) + 5.0 * np.sin(2 * phi)).astype(np.float32)
    vl = (8.0 * np.cos(phi) * np.sin(2 * lam) * (0.3 + jet)).astype(np.float32)
    wl = (0.02 * np.sin(lam) * np.cos(2 * phi) * np.sin(np.pi * zlev / 60.0) * (pl / P0 + 0.05)).astype(np.float32)
    ul = ul + 0.5 * rng.standard_normal(ul.shape, dtype=np.float32)
    vl = vl + 0.5 * rng.standard_normal(vl.shape, dtype=np.float32)
    zetal = (250.0 + 30.0 * zlev * (1.0 + 0.02 * np.cos(lam) * np.cos(phi))).astype(np.float32)      # K, increasing upward
    zeta_dotl = (2e-4 * np.cos(2 * lam) * np.cos(phi) * np.sin(np.pi * zlev / 60.0)
                 * np.ones_like(zetal)).astype(np.float32)                                           # K / s
    out = []
    for a in (pl, ul, vl, wl, zetal, zeta_dotl):
        a = np.ascontiguousarray(np.broadcast_to(a, (nx, ny, npl))).copy()
        a[-1] = a[0]                     # periodic wrap column
        out.append(a)
    from dataclasses import replace
    return replace(met, pl=out[0], ul=out[1], vl=out[2], wl=out[3], zetal=out[4], zeta_dotl=out[5])


def add_meteo_fields(met: Met, seed=11, with_gaps=True) -> Met:
    """The further fields module_meteo interpolates (``Met.extra``): smooth synthetic geopotential height, potential
    vorticity, water vapour, ozone, cloud water contents and cover on the 3-D grid; surface, tropopause, cloud and
    convective 2-D fields.  ``with_gaps`` puts NaNs into a few 2-D fields the way the reference's diagnostics do (no
    cloud, no free-convection level), which exercises the nearest-neighbour rule of the 2-D interpolation."""
    from dataclasses import replace
    from .host import MET_X2, MET_X3
    nx, ny, nz = met.u.shape
    rng = np.random.default_rng(seed + int(met.time) % 977)
    lam = np.deg2rad(met.lon)[:, None, None]
    phi = np.deg2rad(met.lat)[None, :, None]
    z = (H0 * np.log(P0 / met.p))[None, None, :]
    wave = np.sin(2 * lam + 0.2 * met.time / 21600.0) * np.cos(phi)
    extra = {
        "z": z * (1.0 + 0.01 * wave) + 0.0 * lam,
        "pv": 0.3 * np.sign(np.sin(phi)) * np.exp(z / 7.0) * (1.0 + 0.2 * wave) * (np.abs(np.sin(phi)) + 0.05),
        "h2o": 0.02 * np.exp(-z / 2.2) * (1.0 + 0.5 * wave) + 3e-6,
        "o3": 8e-6 * np.exp(-((z - 32.0) / 9.0) ** 2) * (1.0 + 0.1 * wave) + 2e-8,
        "lwc": 2e-5 * np.exp(-((z - 2.0) / 1.5) ** 2) * np.maximum(wave, 0.0),
        "rwc": 1e-5 * np.exp(-((z - 1.5) / 1.0) ** 2) * np.maximum(wave, 0.0),
        "iwc": 1e-5 * np.exp(-((z - 9.0) / 2.0) ** 2) * np.maximum(-wave, 0.0),
        "swc": 5e-6 * np.exp(-((z - 6.0) / 2.0) ** 2) * np.maximum(-wave, 0.0),
        "cc": np.clip(0.5 * np.exp(-((z - 4.0) / 3.0) ** 2) * (1.0 + wave), 0.0, 1.0),
    }
    lam2, phi2 = lam[:, :, 0], phi[:, :, 0]
    w2 = np.sin(3 * lam2 + 0.1 * met.time / 21600.0) * np.cos(phi2)
    two = {
        "ts": 288.0 - 40.0 * np.sin(phi2) ** 2 + 3.0 * w2, "zs": 0.5 * (1.0 + w2) * np.cos(phi2) ** 2,
        "us": 5.0 * np.cos(phi2) + 2.0 * w2, "vs": 2.0 * w2, "ess": 0.1 * w2, "nss": -0.05 * w2, "shf": 120.0 * w2 - 10.0,
        "lsm": (w2 > 0.2).astype(np.float64), "sst": 290.0 - 25.0 * np.sin(phi2) ** 2 + w2,
        "pt": 100.0 + 200.0 * np.sin(phi2) ** 2 + 10.0 * w2, "tt": 200.0 + 15.0 * np.sin(phi2) ** 2 + w2,
        "zt": 17.0 - 9.0 * np.sin(phi2) ** 2 + 0.3 * w2, "h2ot": 4e-6 * (1.0 + 0.2 * w2),
        "pct": 300.0 + 100.0 * w2, "pcb": 850.0 + 50.0 * w2, "cl": 0.2 * (1.0 + w2),
        "plcl": 900.0 + 30.0 * w2, "plfc": 800.0 + 50.0 * w2, "pel": 250.0 + 60.0 * w2,
        "cape": 500.0 * np.maximum(w2, 0.0), "cin": 50.0 * np.maximum(-w2, 0.0), "o3c": 300.0 + 60.0 * np.sin(phi2) ** 2 + 10.0 * w2,
    }
    for k in two:
        two[k] = two[k] + 0.0 * lam2
    if with_gaps:
        for k in ("pct", "pcb", "plfc", "pel"):
            a = np.array(np.broadcast_to(two[k], (nx, ny)), dtype=np.float64)
            a[rng.uniform(size=a.shape) < 0.15] = np.nan
            two[k] = a
    extra.update(two)
    out = {}
    for k, a in extra.items():
        shape = (nx, ny, nz) if k in MET_X3 else (nx, ny)
        a = np.array(np.broadcast_to(a, shape), dtype=np.float32)
        a[-1] = a[0]                     # periodic wrap column
        out[k] = a
    assert set(out) == set(MET_X2) | set(MET_X3)
    return replace(met, extra=out)


MET_BIN_2D = ("ps", "ts", "zs", "us", "vs", "ess", "nss", "shf", "lsm", "sst", "pbl", "pt", "tt", "zt", "h2ot", "pct", "pcb", "cl",
              "plcl", "plfc", "pel", "cape", "cin", "o3c")
MET_BIN_3D = ("z", "t", "u", "v", "w", "pv", "h2o", "o3", "lwc", "rwc", "iwc", "swc", "cc")


def write_met_bin(path, met: Met):
    """Write ``met`` in the reference's uncompressed binary met format (MET_TYPE 1, version 104: write_met_bin,
    src/mptrac.c:14204 ff.): type, version, time, nx, ny, np, the three axes, 24 surface fields [nx][ny], 13 level fields
    [nx][ny][np], the final flag 999.  Fields the Met does not carry (``Met.extra``) are written as zeros."""
    nx, ny, nz = met.u.shape
    extra = getattr(met, "extra", None) or {}

    def field(name, shape):
        a = getattr(met, name, None) if name in ("u", "v", "w", "t", "ps", "pbl") else extra.get(name)
        return np.zeros(shape, np.float32) if a is None else np.ascontiguousarray(a, np.float32).reshape(shape)

    with open(path, "wb") as f:
        f.write(np.array([1, 104], np.int32).tobytes())
        f.write(np.array([met.time], np.float64).tobytes())
        f.write(np.array([nx, ny, nz], np.int32).tobytes())
        for ax in (met.lon, met.lat, met.p):
            f.write(np.ascontiguousarray(ax, np.float64).tobytes())
        for name in MET_BIN_2D:
            f.write(field(name, (nx, ny)).tobytes())
        for name in MET_BIN_3D:
            f.write(field(name, (nx, ny, nz)).tobytes())
        f.write(np.array([999], np.int32).tobytes())
