#!/usr/bin/env python
"""bench.py -- headline benchmark of the SPH step pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): particle-updates/s (and neighbour-pairs/s) of SPH::step.
A "step" is one full pass of the hot path (binning, sort, density+EOS, force +
integration + wall collision) over every particle of a synthetic dam-break /
box-drop scene (SURVEY 8(d) scene rule).  N=1: the 16M-particle dam-break the
north-star target is quoted on.  N>1: 16M per GPU box-drop, z-slab decomposition
(weak scaling).

`value`   : device-resident throughput (inputs already in HBM), CUDA events on the
            stream the kernels run on, max over ranks.
`e2e`     : the same metric through sphb200_step_host with PINNED HOST buffers:
            H2D of positions/velocities/masses + step + D2H of new positions /
            velocities inside the timed region, every step.
`roofline`: dominant kernel (force+integrate sweep), algorithmic bytes / its
            CUDA-event duration vs the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference physics compiled in
            place (oracle/_ref, timing build, 1 thread -- the reference has no
            active parallel region, sph.cpp:215-282) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from oracle import scenes  # noqa: E402  (input synthesis only)

ALG_BYTES_STEP = 240.0     # algorithmic bytes per particle-step (whole pipeline)
ALG_BYTES_FORCE = 72.0     # force+integrate+collide sweep: R(16+16+4) + W(16+16) + R4
NU = 40.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (host clock) -- the timed region."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.05)]
        if not rows:      # region shorter than one sampling period: nearest samples
            rows = [r for (_, r) in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_spec(name):
    cfg = scenes.CONFIGS[name]
    nx, ny, nz = cfg["sites"]
    d = scenes.lattice_spacing(0.1, NU)
    origin = [v * 0.2 for v in cfg["origin_vox"]]
    return cfg, nx, ny, nz, d, origin


# ----------------------------------------------------------------------------
def run_ours(args):
    import torch
    import smoothed_particle_hydrodynamics_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload or ("dambreak_16m" if world == 1 else "boxdrop_16m")
    cfg, nx, ny, nz, d, origin = workload_spec(name)
    n = nx * ny * nz
    sp = scenes.scene_params(nu=NU)
    if world > 1:
        raise SystemExit("multi-GPU slab mode: see bench_slab (not wired in this build)")
    p = S.default_params(particle_count=n, grid=cfg["grid"], examine_count=96, neighbor_mode=S.FULL,
                         use_uniform_gravity=1, use_wall_collision=1, rho0=sp["rho0"], stiffness=sp["stiffness"],
                         viscosity=sp["viscosity"], central_mass=0.0, gravity=sp["gravity"],
                         time_step=sp["time_step"])
    sph = S.SPH(p, device=local, init_scene=False)
    stream = torch.cuda.Stream()
    sph.set_stream(stream.cuda_stream)

    # synthetic scene in pinned host memory (the e2e leg copies from / to it every step)
    pos_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    vel_h = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
    mass_h = torch.ones((n,), dtype=torch.float32).pin_memory()
    S.scene_lattice(nx, ny, nz, d, origin, out=pos_h.numpy())
    sph.upload_ptr(pos_h.data_ptr(), vel_h.data_ptr(), mass_h.data_ptr())
    sph.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    sph.step_n(args.warmup)
    barrier()
    l0 = sph.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pairs = 0
    barrier()
    t_begin = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        sph.step_n(args.steps)
        e1.record(stream)
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    launches = sph.launch_count() - l0
    pairs_last, nmax, nmin = sph.neighbor_stats()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n * world * args.steps / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events inside the library, same stream) --
    sph.set_params(enable_timers=1)
    phase = np.zeros(6)
    reps = min(args.steps, 5)
    for _ in range(reps):
        sph.step_n(1)
        phase += np.array(sph.timings_ms())
    phase /= reps
    sph.set_params(enable_timers=0)
    force_ms = float(phase[4])
    hbm, peak_kind = measured_peaks()
    achieved = ALG_BYTES_FORCE * n / (force_ms * 1e-3) / 1e9 if force_ms > 0 else 0.0

    # ---- end to end through the host-buffer call ----------------------------
    pos_h.numpy()[...] = 0
    S.scene_lattice(nx, ny, nz, d, origin, out=pos_h.numpy())
    vel_h.zero_()
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        sph.step_host_ptr(pos_h.data_ptr(), vel_h.data_ptr(), mass_h.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sph.step_host_ptr(pos_h.data_ptr(), vel_h.data_ptr(), mass_h.data_ptr())
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_value = n * world * e2e_steps / e2e_s

    out = {
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %dx%dx%d jittered lattice = %d particles/GPU, h=0.1, ~%.0f neighbours, "
                               "voxel grid %s, FULL neighbour mode, gravity+walls on" % (
                                   name, nx, ny, nz, n, NU, "x".join(map(str, cfg["grid"]))),
                   "particles": n * world, "neighbor_pairs_per_sec": pairs_last * world / (ms_per_step * 1e-3),
                   "mean_neighbors": pairs_last / n, "l2": "inputs (>= 512 MB state) larger than the 126 MB L2",
                   "step_alg_bytes_per_particle": ALG_BYTES_STEP,
                   "step_hbm_frac": ALG_BYTES_STEP * n / (ms_per_step * 1e-3) / 1e9 / hbm,
                   "phase_ms": {"bin_sort_gather": float(phase[0]), "density_eos": float(phase[2]),
                                "force_integrate": force_ms, "reduce": float(phase[5])}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 24 * n,
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_force_tiled (force+integrate+collide sweep)", "achieved": achieved,
                     "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_kind": peak_kind,
                     "alg_bytes_per_particle": ALG_BYTES_FORCE, "kernel_ms": force_ms},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference_sample(steps=10, warmup=1)
    if rank == 0:
        print(json.dumps(out))
    sph.close()


# ----------------------------------------------------------------------------
def cpu_reference_sample(steps, warmup, sample="dambreak_1m"):
    """The reference's own computeDensity / computeAcceleration / integrate (+ the
    harness all-within-h search and the dead wall code) from oracle/_ref, timing
    build, 1 thread, on a 1/16 sample of the 16M workload."""
    from oracle import refharness
    kind = "reference" if refharness.available("timing") else "port"
    cfg, nx, ny, nz, d, origin = workload_spec(sample)
    n = nx * ny * nz
    sp = scenes.scene_params(nu=NU)
    pos = scenes.lattice_scene(nx, ny, nz, d, origin)
    vel = np.zeros((n, 3), np.float32)
    E = 96
    t_steps = []
    pairs = 0
    if kind == "reference":
        r = refharness.RefSPH("timing")
        r.resize(n, *cfg["grid"], E)
        r.set_params(rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0,
                     gravity=sp["gravity"], time_step=sp["time_step"])
        r.set_state(pos, vel, np.ones(n, np.float32))
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            r.step_phased(True, True, True)
            dt = time.perf_counter() - t0
            if s >= warmup:
                t_steps.append(dt)
        pairs = r.neighbor_stats()[0]
        phases = (r.phase_ns() / 1e6).tolist()
    else:
        from oracle.port import FULL, OracleSPH
        o = OracleSPH(n=n, grid=cfg["grid"], examine=E, init_scene=False, rho0=sp["rho0"], stiffness=sp["stiffness"],
                      viscosity=sp["viscosity"], central_mass=0.0, gravity=sp["gravity"], time_step=sp["time_step"])
        o.set_state(pos, vel)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            o.step(FULL, True, True)
            dt = time.perf_counter() - t0
            if s >= warmup:
                t_steps.append(dt)
        pairs = int(o.count.sum())
        phases = None
    t = float(np.mean(t_steps))
    return {"value": n / t, "unit": "particle-updates/s", "cores": 1, "kind": kind,
            "sample": "%s: %dx%dx%d lattice = %d particles (1/16 of the 16M workload, same spacing and parameters), "
                      "%d timed steps of the FULL-mode harness step" % (sample, nx, ny, nz, n, steps),
            "ms_per_step": t * 1e3, "neighbor_pairs_per_sec": pairs / t, "phase_ms_last": phases,
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = "dambreak_1m" if args.steps + args.warmup <= 40 else "dambreak_128k"
    cb = cpu_reference_sample(args.steps, args.warmup, sample)
    name = args.workload or ("dambreak_16m" if world == 1 else "boxdrop_16m")
    out = {"impl": "reference", "metric": "particle-updates/sec", "value": cb["value"], "unit": "particle-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "sampled_as": cb["sample"]},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
