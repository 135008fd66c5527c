"""The dispatcher of the device path (mpb_run_modules) as a plan: which kernels a model step launches, in which order, for
a control structure -- checked against the reference's dispatcher order (src/mptrac.c:7851-8001) without a GPU.
The planner is pure host logic inside libmptrac_b200.so (mpb_plan_modules); the executor walks the same plan."""
import pytest

from mptrac_b200 import Ctl
from mptrac_b200.host import (MOD_ADVECT, MOD_ALL, MOD_DIFF_MESO, MOD_DIFF_TURB, MOD_METEO, MOD_MIXING, MOD_POSITION0, MOD_POSITION1,
                              MOD_SEDI, MOD_SORT, MOD_TIMESTEPS, plan_modules)

BASE = dict(t_start=0.0, t_stop=1e9, dt_mod=300.0, dt_met=21600.0)
TS, PRE, POST, STORE = 0x01, 0x02, 0x40, 0x80        # step-kernel module bits
TURB, MESO, SEDI = 1, 2, 4


def step(advect, phys, mod):
    return f"step(advect={advect},phys=0x{phys:x},mod=0x{mod:x})"


def test_plain_step_is_one_fused_launch():
    c = Ctl(advect=4, **BASE)
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST)
    c = Ctl(advect=2, diffusion=1, turb_dz_trop=0.5, turb_mesox=0.16, nq=2, qnt_rp=0, qnt_rhop=1, **BASE)
    assert plan_modules(c, 300.0) == step(2, TURB | MESO | SEDI, TS | PRE | POST)
    c = Ctl(advect=0, **BASE)
    assert plan_modules(c, 300.0) == step(0, 0, TS | PRE | POST)


def test_sort_step_computes_dt_before_the_permutation():
    c = Ctl(advect=4, sort_dt=7200.0, **BASE)
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST)
    # on a sort step dt is computed in the sort's key pass (the reference computes it before it permutes the parcels)
    assert plan_modules(c, 7200.0) == " ".join([f"sort(mod=0x{TS | STORE:x})", step(4, 0, PRE | POST)])


def test_meteo_and_mixing_follow_the_step():
    c = Ctl(advect=4, nq=2, met_dt_out=0.1, qnt_meteo={"t": 0}, mixing_trop=1e-3, mixing_strat=1e-6, mixing_dt=600.0, mix_qnt=[1], **BASE)
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST) + " meteo"
    assert plan_modules(c, 600.0) == step(4, 0, TS | PRE | POST) + " meteo mixing"
    c = Ctl(advect=4, nq=1, met_dt_out=3600.0, qnt_meteo={"t": 0}, **BASE)
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST)
    assert plan_modules(c, 3600.0) == step(4, 0, TS | PRE | POST) + " meteo"
    c = Ctl(advect=4, nq=1, met_dt_out=0.1, **BASE)      # no module_meteo quantity asked for: nothing to launch
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST)


def test_hybrid_segments_of_the_shim():
    """the masks mptrac_shim.c passes when a CPU module sits between device segments: dt goes through cache_t::dt"""
    c = Ctl(advect=4, diffusion=1, turb_dz_trop=0.5, turb_mesox=0.16, nq=2, qnt_rp=0, qnt_rhop=1, met_dt_out=0.1, qnt_meteo={"t": 0},
            **BASE)
    seg = MOD_TIMESTEPS | MOD_SORT | MOD_POSITION0 | MOD_ADVECT | MOD_DIFF_TURB
    assert plan_modules(c, 300.0, seg) == step(4, TURB, TS | STORE | PRE)
    assert plan_modules(c, 300.0, MOD_DIFF_MESO) == step(0, MESO, 0)
    assert plan_modules(c, 300.0, MOD_DIFF_MESO | MOD_SEDI | MOD_POSITION1 | MOD_METEO) == step(0, MESO | SEDI, POST) + " meteo"
    assert plan_modules(c, 300.0, MOD_MIXING) == ""
    assert plan_modules(c, 300.0, MOD_ALL & ~MOD_MIXING) == step(4, TURB | MESO | SEDI, TS | PRE | POST) + " meteo"


def levels(modules):
    return f"advect_levels(mod=0x{modules:x})" if modules else "advect_levels"


def test_model_level_advection_takes_the_plain_segments_around_it_in():
    """the advection on model levels is its own kernel; segments around it that only carry timesteps / position checks are
    folded into its launch, segments with diffusion modules are not"""
    c = Ctl(advect=4, advect_vert_coord=2, diffusion=1, turb_dz_trop=0.5, **BASE)       # (TURB_MESOX / Z default to 0.16: on)
    assert plan_modules(c, 300.0) == " ".join([levels(TS | STORE | PRE), step(0, TURB | MESO, POST)])
    assert plan_modules(c, 0.0) == " ".join([levels(TS | STORE | PRE), step(0, TURB | MESO, POST)])
    c = Ctl(advect=2, advect_vert_coord=1, nq=1, qnt_zeta=0, **BASE)
    assert plan_modules(c, 0.0) == " ".join(["advect_init", levels(TS | STORE | PRE | POST)])
    assert plan_modules(c, 300.0) == levels(TS | STORE | PRE | POST)
    # a cell sort computes dt in its key pass (the reference computes dt before it permutes the parcels)
    c = Ctl(advect=4, advect_vert_coord=2, diffusion=0, sort_dt=300.0, **BASE)
    assert plan_modules(c, 300.0) == " ".join([f"sort(mod=0x{TS | STORE:x})", levels(PRE | POST)])
    # the shim's per-module masks (module timers) leave the advection alone
    assert plan_modules(c, 300.0, MOD_ADVECT) == "advect_levels"


def test_modules_with_their_own_kernel_split_the_fused_step_in_the_reference_order():
    c = Ctl(advect=4, diffusion=1, turb_pbl_scheme=1, turb_dz_trop=0.5, turb_mesox=0.16, nq=6, qnt_rp=0, qnt_rhop=1, conv_cape=100.0,
            isosurf=2, tdec_trop=86400.0, tdec_strat=864000.0, qnt_m=2, qnt_loss_rate=3, qnt_aoa=4, bound_lat0=-90.0, bound_lat1=90.0,
            bound_p0=1e10, bound_p1=-1e10, met_dt_out=0.1, qnt_meteo={"t": 5}, mixing_trop=1e-3, mixing_strat=1e-6, mixing_dt=300.0,
            mix_qnt=[2], **BASE)
    assert plan_modules(c, 0.0) == " ".join([
        "isosurf_init", step(4, TURB, TS | STORE | PRE), "diff_pbl", step(0, MESO, 0), "convection", step(0, SEDI, 0), "isosurf",
        step(0, 0, POST), "meteo", "bound_cond", "decay", "mixing", "bound_cond"])
    assert plan_modules(c, 300.0).startswith(step(4, TURB, TS | STORE | PRE) + " diff_pbl")
    c = Ctl(advect=4, conv_mix_pbl=1, conv_dt=3600.0, **BASE)
    assert plan_modules(c, 300.0) == step(4, 0, TS | PRE | POST)
    assert plan_modules(c, 3600.0) == " ".join([step(4, 0, TS | STORE | PRE), "convection", step(0, 0, POST)])
    c = Ctl(advect=4, nq=1, qnt_loss_rate=0, **BASE)      # the total loss rate is reset even without module_decay
    assert plan_modules(c, 300.0) == step(4, 0, TS | STORE | PRE | POST) + " decay"


@pytest.mark.parametrize("isosurf", [1, 4])
def test_isosurf_init_only_for_the_modes_that_compute_it(isosurf):
    c = Ctl(advect=4, isosurf=isosurf, **BASE)
    want = [step(4, 0, TS | STORE | PRE), "isosurf", step(0, 0, POST)]
    assert plan_modules(c, 300.0) == " ".join(want)
    assert plan_modules(c, 0.0) == " ".join((["isosurf_init"] if isosurf != 4 else []) + want)


def test_chem_grid_follows_mixing():
    c = Ctl(advect=4, nq=2, qnt_m=0, qnt_Cx=1, molmass=64.0, chemgrid=1, mixing_trop=1e-3, mixing_strat=1e-6, mixing_dt=300.0, mix_qnt=[0],
            bound_lat0=-90.0, bound_lat1=90.0, bound_p0=1e10, bound_p1=-1e10, **BASE)
    assert plan_modules(c, 300.0) == " ".join([step(4, 0, TS | STORE | PRE | POST), "bound_cond", "mixing", "chem_grid", "bound_cond"])
