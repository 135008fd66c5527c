"""The decision logic of physics.cuh `box_level` (production device build), restated in numpy float32 arithmetic: wherever
the single-precision altitude is allowed to decide the level of a parcel, its answer must be the one of the reference's
double-precision sequence Z(p) = H0 log(P0 / p), (int)((z - z0) / dz) (src/mptrac.h:2243, src/mptrac.c:5214).  The GPU test
`test_binning_at_the_cell_faces` checks the kernel; this one checks the margins on 20 M pressures, including every face
of several grids approached to 1e-15 km, without a GPU."""
import numpy as np
import pytest

F = np.float32
Z_MARGIN = F(1e-3)


def fast_level(p, z0, z1, nz):
    """-2 = the fast path does not decide; otherwise the level (or -1 = outside) it returns"""
    rdz = 1.0 / ((z1 - z0) / nz)
    with np.errstate(all="ignore"):
        zf = F(7.0) * np.log(F(1013.25) / p.astype(F)).astype(F)
    lo, hi = F(z0), F(z1)
    m = Z_MARGIN + F(1e-6) * (abs(lo) + abs(hi))
    out = np.full(p.shape, -2, np.int64)
    with np.errstate(all="ignore"):
        return _decide(out, zf, lo, hi, m, rdz, nz)


def _decide(out, zf, lo, hi, m, rdz, nz):
    sane = np.abs(zf) < F(1e3)
    inside = sane & (zf > lo + m) & (zf < hi - m)
    t = ((zf - lo) * F(rdz)).astype(F)
    k = np.floor(t)
    mt = (m * F(rdz) + F(4e-7) * t).astype(F)
    sure = inside & (t - k > mt) & (k + F(1.0) - t > mt)
    out[sure] = np.where(k[sure] < F(nz), k[sure].astype(np.int64), -1)
    outside = sane & ~inside & ((zf < lo - m) | (zf > hi + m))
    out[outside] = -1
    return out


def exact_level(p, z0, z1, nz):
    with np.errstate(all="ignore"):
        z = 7.0 * np.log(1013.25 / p)
    dz = (z1 - z0) / nz
    with np.errstate(invalid="ignore"):
        iz = np.floor((z - z0) / dz)
    bad = ~(z >= z0) | (z >= z1) | ~(iz < nz)
    return np.where(bad, -1, iz).astype(np.int64)


@pytest.mark.parametrize("z0,z1,nz", [(-5.0, 85.0, 1), (-5.0, 85.0, 90), (0.0, 60.0, 240), (-2.0, 38.0, 40), (10.0, 10.5, 1000)])
def test_the_single_precision_altitude_only_decides_what_it_can(z0, z1, nz):
    rng = np.random.default_rng(17)
    mag = np.concatenate([[0.0], 10.0 ** np.arange(-15.0, -1.0, 0.25)])
    eps = np.concatenate([-mag[::-1], mag])
    faces = z0 + (z1 - z0) / nz * np.arange(nz + 1)
    near = (1013.25 * np.exp(-(faces[:, None] + eps[None, :]) / 7.0)).ravel()
    p = np.concatenate([near, rng.uniform(1e-4, 1100.0, 2_000_000), 1013.25 * np.exp(-rng.uniform(z0 - 3, z1 + 3, 2_000_000) / 7.0),
                        [0.0, -1.0, np.inf, np.nan, 1e-300, 1e300, 1013.25]])
    fast, exact = fast_level(p, z0, z1, nz), exact_level(p, z0, z1, nz)
    decided = fast != -2
    assert np.array_equal(fast[decided], exact[decided])
    # and it decides nearly always: the exact sequence is the exception (coarse grids) or at least the minority (1 m boxes)
    share = decided[near.size:-7].mean()
    assert share > (0.99 if (z1 - z0) / nz > 0.2 else 0.0), share
