#!/usr/bin/env python
"""Benchmark of the MPTRAC time-step path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c4g|c5|c2ml]

Workloads: c2 (default) = BASELINE configs[1]; c3 = configs[2]; c4 = one GPU's share of configs[3], c4g the same with
the gridded output reduced over the ranks every step; c5 = one GPU's share of configs[4] (cell sort + mixing all-reduce
every step); c2ml = configs[1] on model levels (ADVECT_VERT_COORD 2).  N > 1 runs under torchrun, one rank per GPU.

One "step" = one model time step (mptrac_run_timestep restricted to the path) over every parcel of the
workload.  Default workload = BASELINE.json configs[1] ("c2"): 1 M parcels, 1 deg x 1 deg x 60-level synthetic
ERA5-shaped met (361 x 181 x 60 after the periodic wrap column), RK4 advection only, fp64.

Printed JSON (one line, rank 0):
  value         particle-steps/s with parcels + met resident in HBM, CUDA-event timed per step, L2 flushed
                between steps (a 256 MiB device memset outside the timed events)
  e2e           the same metric through the C ABI with HOST (pinned) parcel buffers: every step uploads
                time/p/lon/lat, runs the step and downloads them again inside the timed region
  roofline      algorithmic bytes (64 B per parcel-step + the u,v,w grids of both time levels once per
                step) / average step-kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own OpenMP code (oracle/_ref harness) or the C port (oracle/) on this host's cores,
                on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (parcels per GPU, nlon, nlat, nlev, ctl overrides, algorithmic bytes per parcel-step excluding met, met fields)
    "c2": dict(np=1_000_000, grid=(360, 181, 60), ctl=dict(advect=4, diffusion=0, sort_dt=7200.0), state_bytes=64, met_fields=3,
               desc="1M parcels, 1x1 deg x 60 levels (361x181x60), RK4 advection only, SORT_DT 7200 (cell sort every 24 steps, timed)"),
    "c3": dict(np=10_000_000, grid=(720, 361, 137), ctl=dict(advect=4, diffusion=1, turb_dx_pbl=0, turb_dx_trop=0,
                                                              turb_dz_strat=0, qnt_rp=0, qnt_rhop=1, nq=2, sort_dt=3600.0),
               state_bytes=104, met_fields=4,
               desc="10M parcels, 0.5x0.5 deg x 137 levels (721x361x137), RK4 + mesoscale diffusion + sedimentation"),
    # one GPU's share of BASELINE configs[3] (100 M parcels over 8 GPUs): dense -- 6 parcels per grid cell
    "c4": dict(np=12_500_000, grid=(360, 181, 60), ctl=dict(advect=4, diffusion=1, sort_dt=3600.0), state_bytes=88, met_fields=3,
               desc="12.5M parcels per GPU (100M / 8), 1x1 deg x 60 levels, RK4 + turbulent + mesoscale diffusion"),
    # BASELINE configs[3] as stated: the c4 share plus the gridded output EVERY step (write_grid's binning on the device,
    # src/mptrac.c:13840-13872, reference default grid 360 x 180 x 1): local partial boxes -> one NCCL sum-reduction to
    # rank 0 -> rank 0 reads count / sum / sum of squares on the host
    "c4g": dict(np=12_500_000, grid=(360, 181, 60), ctl=dict(advect=4, diffusion=1, sort_dt=3600.0, nq=1), state_bytes=104, met_fields=3,
                q_init=1.0, out_grid=dict(nx=360, ny=180, nz=1, lon0=-180.0, lon1=180.0, lat0=-90.0, lat1=90.0, z0=-5.0, z1=85.0),
                label="configs[3]", kernel="step_kernel (+ grid binning, NCCL reduce)",
                desc="12.5M parcels per GPU (100M / 8), 1x1 deg x 60 levels, RK4 + turbulent + mesoscale diffusion, gridded output "
                     "(360x180x1 boxes) reduced over the ranks every step"),
    # BASELINE configs[4]: 10 M parcels over 8 GPUs, inter-parcel mixing and the cell sort EVERY step (MIXING_DT = SORT_DT =
    # DT_MOD); reference default mixing grid 360 x 180 x 90 = 5.8 M boxes: accumulate -> NCCL all-reduce of 70 MB -> relax
    "c5": dict(np=1_250_000, grid=(360, 181, 60), ctl=dict(advect=4, diffusion=0, sort_dt=300.0, nq=1, mixing_trop=1e-3, mixing_strat=1e-6,
                                                             mixing_dt=300.0, mix_qnt=[0]), state_bytes=112, met_fields=3,
               q_init="uniform", mixing=True, label="configs[4]", kernel="step_kernel (+ sort, mixing kernels, NCCL all-reduce)",
               desc="1.25M parcels per GPU (10M / 8), 1x1 deg x 60 levels, RK4, cell sort and inter-parcel mixing (360x180x90 boxes, "
                    "all-reduced over the ranks) every step"),
    # configs[1] on model levels: omega on 60 model levels, the reference's trac_test "ml" setting.
    # ONE launch per step (timesteps + position + model-level advection + position, folded by the plan); per parcel-step it
    # reads time, lon, lat, p and writes them (64 B), stores dt (8 B) and reads / writes the 2-byte level hint: 76 B
    "c2ml": dict(np=1_000_000, grid=(360, 181, 60), levels=60, ctl=dict(advect=4, advect_vert_coord=2, diffusion=0, sort_dt=7200.0),
                 state_bytes=76, met_fields=4, kernel="advect_levels_kernel (timesteps and position checks folded in)",
                 label="configs[1] on model levels",
                 roofline_note="Not HBM-bound: the model-level lookup is four dependent load rounds per Runge-Kutta stage (column "
                               "searches on both time levels, then the 8 records) at 16 resident warps/SM, ~800 instructions per stage; "
                               "ncu r02p under profiles/, DESIGN.md 7",
                 desc="1M parcels, 1x1 deg x 60 model levels, RK4 advection with omega on model levels (ADVECT_VERT_COORD 2)"),
    # probes, not BASELINE configurations: the module mixes of c3 and c4 exchanged between their grids (to tell an effect of
    # the instruction mix from an effect of the grid size when a build flag helps one of the two)
    "x3t": dict(np=10_000_000, grid=(720, 361, 137), ctl=dict(advect=4, diffusion=1, sort_dt=3600.0), state_bytes=88, met_fields=3,
                label="probe", desc="probe: c3's parcels and grid with c4's modules (RK4 + turbulent + mesoscale diffusion)"),
    "x4s": dict(np=12_500_000, grid=(360, 181, 60), ctl=dict(advect=4, diffusion=1, turb_dx_pbl=0, turb_dx_trop=0, turb_dz_strat=0,
                                                              qnt_rp=0, qnt_rhop=1, nq=2, sort_dt=3600.0), state_bytes=104, met_fields=4,
                label="probe", desc="probe: c4's parcels and grid with c3's modules (RK4 + mesoscale diffusion + sedimentation)"),
}
DT_MOD = 300.0
DT_MET = 21600.0


def workload_label(name):
    return WORKLOADS[name].get("label") or f"configs[{dict(c2=1, c3=2, c4=3)[name]}]"


T0 = time.perf_counter()


def _log(msg):
    """Progress marker on stderr (a stalled run then shows where it stopped)."""
    print(f"[bench {time.perf_counter() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def host_cores() -> int:
    """Cores this process may actually use: the affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = Path(f).read_text().split()
            if f.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = float(txt[0])
                per = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def cpu_threads() -> int:
    return int(os.environ.get("MPB_CPU_THREADS", "0")) or host_cores()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return json.loads(f.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(wl, rank, world):
    from mptrac_b200 import Ctl, synth
    nlon, nlat, nlev = wl["grid"]
    m0, m1 = synth.make_met_pair(nlon, nlat, nlev, t0=0.0, dt_met=DT_MET)
    if wl.get("levels"):
        m0, m1 = synth.add_model_levels(m0, npl=wl["levels"]), synth.add_model_levels(m1, npl=wl["levels"])
    if os.environ.get("MPB_BENCH_LON_AXIS") == "-180":
        # probe (not a BASELINE setting: the reference's `wind` tool writes 0..360): a met grid labelled -180..180, on which the
        # reference's sort key orders ALL parcels by column instead of clamping the western half into column 0
        m0.lon = m0.lon - 180.0
        m1.lon = m1.lon - 180.0
    n = wl["np"]
    tm, p, lon, lat = synth.make_parcels(n, t0=0.0, seed=123 + rank)
    kw = dict(nq=0, t_start=0.0, t_stop=1e9, dt_mod=DT_MOD, dt_met=DT_MET)
    kw.update(wl["ctl"])
    ctl = Ctl(**kw)
    q = None
    if ctl.nq and wl.get("q_init") is not None:
        q = np.full((ctl.nq, n), 1.0) if wl["q_init"] != "uniform" else np.random.default_rng(7 + rank).uniform(0.0, 1.0, (ctl.nq, n))
    elif ctl.nq:
        q = np.zeros((ctl.nq, n))
        q[0], q[1] = 1.0, 1500.0   # rp = 1 micron, rhop = 1500 kg/m3
    return ctl, m0, m1, (tm, p, lon, lat, q)


def bind_to_gpu_numa_node(local):
    """Several ranks share the host: keep this rank's threads -- and with them the pinned parcel buffers it is about to
    allocate (first touch) -- on the NUMA node its GPU hangs off, so that host <-> device copies do not cross the socket
    interconnect.  Returns a short description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{dev}/numa_node").read_text())
        if node < 0:
            return f"GPU {dev}: no NUMA information"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"GPU {dev} on NUMA node {node}: none of its CPUs is available to this process"
        os.sched_setaffinity(0, cpus)
        return f"GPU {dev} on NUMA node {node}: rank bound to {len(cpus)} CPUs of that node"
    except Exception as exc:      # affinity is an optimisation, never a requirement
        return f"not bound ({exc!r})"


def exchange_transport():
    return os.environ.get("MPB_BENCH_EXCHANGE", "peers")


def make_steps(eng, wl, ctl, dev, world):
    """(full step, the same without its exchange, attached) for a workload"""
    from mptrac_b200 import dist as mdist
    from mptrac_b200.host import MOD_ALL, MOD_MIXING
    attached = False
    if world > 1 and (wl.get("mixing") or wl.get("out_grid")) and exchange_transport() == "peers":
        g = wl.get("out_grid")
        attached = mdist.attach_peers(eng, ctl, g["nx"] * g["ny"] * g["nz"] if g else 0)
    if wl.get("mixing"):
        def transport_only(t):
            eng.run_modules(t, MOD_ALL & ~MOD_MIXING)
        if attached or world == 1:
            full = eng.run_timestep              # module_mixing inside: accumulate -> barrier -> apply
        else:
            def full(t):
                eng.run_modules(t, MOD_ALL & ~MOD_MIXING)
                mdist.mixing_step(eng, t, dev)
    elif wl.get("out_grid"):
        transport_only = eng.run_timestep

        def full(t):
            eng.run_timestep(t)
            # rank 0 holds the summed boxes on the host after this call
            mdist.grid_output(eng, dict(wl["out_grid"], t0=t - 0.5 * DT_MOD, t1=t + 0.5 * DT_MOD), dev, attached=attached)
    else:
        full = transport_only = eng.run_timestep
    return full, transport_only, attached


def exchange_record(name, rank, world, local, K=12, W=3):
    """One of the exchange workloads (configs[3] with its gridded output = c4g, configs[4] with its mixing = c5) on the
    ranks of this run: ms per step with and without the exchange step, max over ranks, CUDA events, L2 flushed."""
    import torch
    import torch.distributed as dist
    from mptrac_b200 import Engine, synth
    wl = WORKLOADS[name]
    ctl, m0, m1, (tm, p, lon, lat, q) = build_inputs(wl, rank, world)
    n = wl["np"]
    eng = Engine(n, nq=ctl.nq, device=local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_ctl(ctl)
    eng.set_clim_tropo(*synth.make_clim_tropo())
    eng.set_met(0, m0)
    eng.set_met(1, m1)
    eng.set_shard(rank * n, world * n)
    eng.set_atm(tm, p, lon, lat, q)
    dev = torch.device("cuda", local)
    full, transport_only, attached = make_steps(eng, wl, ctl, dev, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t = [0.0]

    def timed(fn):
        for _ in range(W):
            t[0] += DT_MOD
            fn(t[0])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for k in range(K):
            t[0] += DT_MOD
            flush.zero_()
            ev[k][0].record(stream)
            fn(t[0])
            ev[k][1].record(stream)
        torch.cuda.synchronize()
        ms = float(sum(a.elapsed_time(b) for a, b in ev)) / K
        if world > 1:
            x = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            ms = float(x.item())
        return ms

    spinup = int(round(wl["ctl"].get("sort_dt", 0.0) / DT_MOD))
    for _ in range(spinup):
        t[0] += DT_MOD
        full(t[0])
    ms_full = timed(full)
    ms_plain = timed(transport_only)
    eng.sync()
    if world > 1:
        dist.barrier()
    nmix = len([i for i in ctl.mix_qnt if i >= 0]) if wl.get("mixing") else 0
    if wl.get("mixing"):
        nbox = ctl.mixing_nx * ctl.mixing_ny * ctl.mixing_nz
        per_rank = (8 * (2 * nmix + 1) * n * (world - 1) // world if attached else      # {box, q ..} out, box means back
                    2 * 8 * (nmix + 1) * nbox * (world - 1) // world)
    else:
        g = wl["out_grid"]
        per_rank = g["nx"] * g["ny"] * g["nz"] * (16 * max(ctl.nq, 1) + 4) if world > 1 else 0
    eng.close()
    del flush
    return {"workload": f"BASELINE {workload_label(name)}: {wl['desc']}", "parcels_per_gpu": n, "steps": K,
            "ms_per_step": ms_full, "ms_transport_only": ms_plain, "ms_exchange": ms_full - ms_plain,
            "exchange_share": (ms_full - ms_plain) / ms_full if ms_full > 0 else None,
            "value": float(n) * world / (ms_full * 1e-3), "unit": "particle-steps/s",
            "transport": ("none (one rank)" if world == 1 else
                          "peer memory: contributions routed to the owner of each box and answers routed back as coalesced stores over NVLink, "
                          "local atomics, flag barriers (own kernels)" if attached
                          else "NCCL: one all-reduce of the dense box records" if wl.get("mixing") else "NCCL all-reduce of the output boxes"),
            "exchanged_bytes_per_rank_per_step": int(per_rank)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mptrac_b200 import Engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # libraries write to file descriptor 1 behind Python's back (NCCL prints its version line there): keep the real
    # stdout for the one JSON line and send everything else to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    wl = WORKLOADS[args.workload]
    if os.environ.get("MPB_BENCH_SORT_DT"):      # tuning aid: cell-sort cadence of the workload [s]
        wl = dict(wl, ctl=dict(wl["ctl"], sort_dt=float(os.environ["MPB_BENCH_SORT_DT"])))
    ctl, m0, m1, (tm, p, lon, lat, q) = build_inputs(wl, rank, world)
    n = wl["np"]

    _log(f"inputs ready ({args.workload}, rank {rank}/{world})")
    eng = Engine(n, nq=ctl.nq, device=local)
    stream = torch.cuda.Stream()          # a real (non-default) stream: the engine launches on it, the events time it
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_ctl(ctl)
    eng.set_clim_tropo(*synth.make_clim_tropo())
    eng.set_met(0, m0)
    eng.set_met(1, m1)
    eng.set_shard(rank * n, world * n)

    # pinned host buffers = what a host driver (trac) owns
    # (time, p, lon, lat as consecutive rows of one block, the way they sit in the reference's atm_t)
    host = torch.from_numpy(np.stack([tm, p, lon, lat])).pin_memory()
    hq = torch.from_numpy(q).pin_memory() if q is not None else None
    hn = {k: host[i].numpy() for i, k in enumerate(("time", "p", "lon", "lat"))}
    hqn = hq.numpy() if hq is not None else None

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Model:
        """Model clock of the run.  Like the reference driver (mptrac_get_met, src/mptrac.c:6489-6499) it rolls the two
        met levels forward when the clock reaches the later one -- swap + upload of a new level, alternating the two
        synthetic fields -- so that every timed step interpolates INSIDE the met interval however long the run is."""
        def __init__(self):
            self.t, self.t1, self.rolls, self.lev = 0.0, DT_MET, 0, [m0, m1]

        def next_t(self):
            self.t += DT_MOD
            if self.t > self.t1:
                from dataclasses import replace
                self.t1 += DT_MET
                eng.swap_met()
                eng.set_met(1, replace(self.lev[self.rolls % 2], time=self.t1))
                self.rolls += 1
            return self.t

    model = Model()

    # one model step.  Workloads with an exchange: by default the ranks are attached through peer memory and
    # the engine does the exchange itself (contributions routed to the owner of each box over NVLink + flag barriers);
    # MPB_BENCH_EXCHANGE=nccl sums the dense box records with ONE all-reduce on the stream the engine launches on
    dev = torch.device("cuda", local)
    step, _, attached = make_steps(eng, wl, ctl, dev, world)
    exchange = bool(wl.get("mixing") or wl.get("out_grid"))

    def host_step(t):
        if exchange:   # host buffers in, step with its exchange, host buffers out
            eng.set_atm(hn["time"], hn["p"], hn["lon"], hn["lat"], hqn)
            step(t)
            eng.get_atm({"time": hn["time"], "p": hn["p"], "lon": hn["lon"], "lat": hn["lat"], "q": hqn})
        else:
            eng.run_timestep_host(t, hn["time"], hn["p"], hn["lon"], hn["lat"], hqn)

    # ---------------- device-resident timing: per-step events, L2 flushed between steps ----------------
    _log("engine ready; device-resident timing")
    eng.set_atm(hn["time"], hn["p"], hn["lon"], hn["lat"], hqn)
    spinup = int(round(wl["ctl"].get("sort_dt", 0.0) / DT_MOD))    # reach the first cell sort: steady state of a long run
    for _ in range(spinup + W):
        step(model.next_t())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches = 0
    barrier()
    with ClockSampler(local) as clk:
        wall0 = time.perf_counter()
        for k in range(K):
            t_next = model.next_t()          # (a met roll, every 72 steps, happens here: outside the step's events)
            flush.zero_()
            l0 = eng.launch_count
            ev[k][0].record(stream)
            step(t_next)
            ev[k][1].record(stream)
            launches += eng.launch_count - l0     # kernels of ours inside the timed events
        barrier()
        wall1 = time.perf_counter()
        # the timed region lasts only milliseconds; keep the identical load running (untimed) for ~1 s so that the
        # 100 ms nvidia-smi sampler sees the clocks this kernel runs at
        t_end = time.perf_counter() + (0.0 if os.environ.get("MPB_BENCH_NO_SUSTAIN") else max(0.0, 1.0 - (wall1 - wall0)))
        while time.perf_counter() < t_end:
            for _ in range(24):
                step(model.next_t())
            torch.cuda.synchronize()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())

    # ---------------- back-to-back (no flush), one bracket ----------------
    _log("back-to-back timing")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rolls0 = model.rolls
    barrier()
    e0.record(stream)
    for k in range(K):
        step(model.next_t())
    e1.record(stream)
    barrier()
    b2b_ms = e0.elapsed_time(e1)
    b2b_rolls = model.rolls - rolls0

    # ---------------- end to end through the C ABI with host buffers ----------------
    # every step: parcels go host -> device, are stepped and come back (mpb_run_timestep_host pipelines the three
    # chunk by chunk); the host arrays are what the caller reads after each step
    _log("end-to-end timing")
    eng.get_atm({"time": hn["time"], "p": hn["p"], "lon": hn["lon"], "lat": hn["lat"], "q": hqn})
    for _ in range(max(W, 3)):
        host_step(model.next_t())
    barrier()
    e0.record(stream)
    for k in range(K):
        host_step(model.next_t())
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # what the host link gives on this box (one 64 MiB pinned copy per direction, then both at once): the bound of e2e
    pc = {}
    try:
        hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        hb2 = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        db, db2 = torch.empty_like(hb, device="cuda"), torch.empty_like(hb, device="cuda")
        s2 = torch.cuda.Stream()

        def timed(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            s2.synchronize()
            b.record(stream)
            torch.cuda.synchronize()
            return (64 << 20) * reps / (a.elapsed_time(b) * 1e-3) / 1e9

        def both():
            db.copy_(hb, non_blocking=True)
            with torch.cuda.stream(s2):
                hb2.copy_(db2, non_blocking=True)

        pc = {"h2d_gbs": round(timed(lambda: db.copy_(hb, non_blocking=True)), 1),
              "d2h_gbs": round(timed(lambda: hb.copy_(db, non_blocking=True)), 1),
              "both_gbs_per_direction": round(timed(both), 1), "numa": numa}
        if world > 1:
            # what the box gives when ALL ranks copy both ways at once: the ceiling of the host-resident (e2e) number
            dist.barrier()
            mine = timed(both, reps=10)
            tsum = torch.tensor([mine], dtype=torch.float64, device="cuda")
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            pc["all_ranks_at_once_gbs_per_direction"] = {"this_rank": round(mine, 1), "sum_over_ranks": round(float(tsum.item()), 1)}
        del hb, hb2, db, db2
    except Exception as exc:   # context only
        pc = {"error": repr(exc)}
    # what the last host-resident step moved across the link, as counted by the engine: p, lon, lat both ways (+ rp, rhop in when
    # sedimentation is on); time[] only when the parcels do not all carry the same time (mpb_host_step_bytes)
    h2d_bytes, d2h_bytes = (32 * n, 32 * n) if exchange else eng.host_step_bytes
    if exchange:                                       # set_atm / get_atm move every quantity both ways
        h2d_bytes = d2h_bytes = (32 + 8 * ctl.nq) * n
        if wl.get("out_grid") and rank == 0:
            g = wl["out_grid"]
            d2h_bytes += g["nx"] * g["ny"] * g["nz"] * (4 + 16 * max(ctl.nq, 1))   # count, sum, sum of squares on rank 0
    checksum = float(hn["lat"][:: max(1, n // 1024)].sum())   # the result is read on the host

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms, b2b_ms, e2e_ms = allmax(total_ms), allmax(b2b_ms), allmax(e2e_ms)
    units = float(n) * world * K
    value = units / (total_ms * 1e-3)

    # ---------------- roofline of the dominant (only) kernel ----------------
    nx, ny, nz = wl["grid"][0] + 1, wl["grid"][1], wl["grid"][2]
    met_bytes = wl["met_fields"] * 2 * nx * ny * nz * 4
    algo_bytes = wl["state_bytes"] * n + met_bytes
    kern_ms = total_ms / K
    peak, peak_src = peaks()
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic = traffic_src = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            traffic, traffic_src = tj.get(args.workload), tj.get("_source_" + args.workload)
        except Exception:
            traffic = None

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE {workload_label(args.workload)}: {wl['desc']}", "parcels_per_gpu": n,
                   "dt_mod_s": DT_MOD, "l2": "flushed between steps (256 MiB memset outside the per-step events)",
                   "parallelism": f"parcels sharded contiguously over {world} GPU(s), " + (
                       "no data-path collective" if not exchange or world == 1 else
                       "box records exchanged through peer memory once per step (own kernels over NVLink)" if attached else
                       "one NCCL all-reduce of the box records per step" if wl.get("mixing") else
                       "one NCCL reduce of the output-grid boxes per step")},
        "back_to_back": {"value": units / (b2b_ms * 1e-3), "ms_per_step": b2b_ms / K, "met_rolls_inside": b2b_rolls,
                         "note": "no L2 flush, one event bracket around K steps"},
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / K, "api": ("mpb_set_atm + step with its exchange + mpb_get_atm (pinned host arrays, every step)" if exchange else
                        "mpb_run_timestep_host (pinned host arrays in, same arrays out, every step; all parcels carry one time, so "
                        "time[] is filled on the host instead of crossing the link)"),
                "host_checksum": checksum, "host_link": pc},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": (f"from the ncu capture {traffic_src} (not re-measured in this run)" if traffic_src else None),
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel": wl.get("kernel", "step_kernel"), "peak_source": peak_src,
                     "note": wl.get("roofline_note") or (
                         "not HBM-bound: DRAM traffic is below the algorithmic bytes; the kernel is limited by latency at 16 "
                         "resident warps/SM (about 100 live fp64 registers per parcel) and by the 192 f32->f64 conversions per "
                         "RK4 step the reference's arithmetic needs (XU pipe 45 % busy: floor 47 us per 1M parcels, i.e. frac "
                         "0.51 at best); two re-designs were built and measured in round 2 -- one lane per coordinate (196 us) "
                         "and a TMA-staged met window (148 us) against 110-114 us of this kernel; ncu evidence under "
                         "profiles/, analysis in DESIGN.md 3.1 and 3.3")},
        "clocks": clk.summary(),
    }
    if world > 1:
        dist.barrier()       # (attached ranks: nobody frees its exchange area while a peer may still read it)
    eng.close()
    del flush
    # The workloads of BASELINE configs[3] / configs[4], whose step has an exchange between the ranks, measured next to the
    # headline on the same ranks (the headline workload itself shards without any exchange)
    if args.workload == "c2" and not args.no_exchange:
        line["exchange"] = {}

        def give_up():
            # a rank that failed on its own would leave the others waiting in a collective for ever: after the deadline every
            # rank leaves; rank 0 still prints the headline it has measured
            if rank == 0:
                line["exchange"]["error"] = "the exchange workloads did not finish within their deadline"
                os.write(real_stdout, (json.dumps(line) + "\n").encode())
            os._exit(0)
        watchdog = threading.Timer(float(os.environ.get("MPB_BENCH_EXCHANGE_DEADLINE", "240")), give_up)
        watchdog.daemon = True
        watchdog.start()
        for name in ("c5", "c4g"):
            _log(f"exchange workload {name}")
            try:
                line["exchange"][name] = exchange_record(name, rank, world, local)
            except Exception as exc:      # must never take the headline down with it
                line["exchange"][name] = {"error": repr(exc)[:300]}
        watchdog.cancel()
    if rank == 0:
        # CPU arm last, in its own process with a hard deadline: it can never take the GPU numbers down with it
        if world == 1 and not args.no_cpu:
            _log("cpu_baseline leg")
            line["cpu_baseline"] = cpu_baseline_subprocess(args.workload, args.cpu_budget)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    _log("done")
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(workload, budget_s=20.0, steps=None):
    """The reference's own CPU code (oracle/_ref) -- or the C port when _ref is absent -- on a bounded sample."""
    from oracle.oracle import Oracle, Parcels, Reference, reference_available
    from mptrac_b200 import synth
    wl = WORKLOADS[workload]
    ctl, m0, m1, (tm, p, lon, lat, q) = build_inputs(wl, 0, 1)
    n_sample = min(wl["np"], 1_000_000)
    sl = slice(0, n_sample)
    atm = Parcels(tm[sl], p[sl], lon[sl], lat[sl], None if q is None else q[:, sl])
    cores = cpu_threads()
    if reference_available():
        kind = "reference"
        ref = Reference()
        names = (["rp", "rhop"] if ctl.qnt_rp >= 0 else ["m"])[:ctl.nq]
        ref.read_ctl(names, "")
        ref.set_met(m0, m1)
        run = lambda t, k: ref.run("timestep", ctl, atm, t=t, nsteps=k)   # noqa: E731
    else:
        kind = "port"
        orc = Oracle()
        clim = synth.make_clim_tropo()
        run = lambda t, k: orc.run("timestep", ctl, clim, m0, m1, atm, t=t, nsteps=k)   # noqa: E731
    run(0.0, 1)                 # dt = 0 step: touches everything once
    t0 = time.perf_counter(); run(DT_MOD, 1); one = time.perf_counter() - t0
    k = steps or int(max(2, min(100, budget_s / max(one, 1e-3))))
    t0 = time.perf_counter(); run(2 * DT_MOD, k); el = time.perf_counter() - t0
    return {"value": n_sample * k / el, "unit": "particle-steps/s", "cores": cores, "kind": kind,
            "sample": f"{n_sample} parcels x {k} steps of the same workload ({el:.1f} s, OMP threads = {cores})"}


def cpu_baseline_subprocess(workload, budget_s):
    """Run the CPU arm in a fresh interpreter (no CUDA context, no torch thread pools next to the OpenMP team) with a
    hard deadline; the whole process group is killed when it is exceeded."""
    import signal
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", workload, "--cpu-budget", str(budget_s),
           "--steps", "0", "--warmup", "1"]
    deadline = max(120.0, 8 * budget_s) if workload == "c2" else max(600.0, 20 * budget_s)
    try:
        env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
        pr = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True, env=env)
        try:
            out, err = pr.communicate(timeout=deadline)
        except subprocess.TimeoutExpired:
            os.killpg(pr.pid, signal.SIGKILL)
            try:
                out, err = pr.communicate(timeout=10)
            except subprocess.TimeoutExpired:
                out, err = "", "unresponsive after SIGKILL"
            return {"error": f"cpu arm exceeded {deadline:.0f} s", "stderr": (err or "")[-300:]}
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (err or out)[-300:]}
    except Exception as exc:  # the CPU arm must never take the GPU number down with it
        return {"error": repr(exc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the OpenMP runtime reads this when the reference library is loaded (below).  torchrun exports OMP_NUM_THREADS=1 to
    # its workers, which is not what "all the host threads it can use" means: the count is ours (MPB_CPU_THREADS or
    # every core this process may run on)
    os.environ["OMP_NUM_THREADS"] = str(cpu_threads())
    _log(f"reference arm on {os.environ['OMP_NUM_THREADS']} OpenMP threads")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    # the reference logs through C stdio: keep file descriptor 1 clean for the one JSON line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.warmup > 0:
            cpu_baseline(args.workload, steps=1)
        # each "step" of this arm is a bounded sample: one model step over min(np, 1M) parcels on all host cores
        base = cpu_baseline(args.workload, budget_s=args.cpu_budget, steps=(max(1, min(args.steps, 20)) if args.steps > 0 else None))
    finally:
        sys.stdout.flush()
        import ctypes
        ctypes.CDLL(None).fflush(None)      # what the reference left in C stdio buffers goes to stderr, too
        os.dup2(saved, 1)
        os.close(saved)
    v = base["value"]
    n_sample = min(wl["np"], 1_000_000)
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_sample / v * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE {workload_label(args.workload)}: {wl['desc']}",
                       "note": "reference CPU arm: host cores only, rank 0 only"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-exchange", action="store_true", help="skip the c5 / c4g exchange sub-records of the default workload")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
