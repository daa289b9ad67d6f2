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
`roofline`: dominant kernel (decided live: force or density sweep), algorithmic bytes / its
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
ALG_BYTES_DENSITY = 20.0   # density+EOS sweep: R16 + W4
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 16.7M particles, from the committed
# `ncu --set full` capture (profiles/r01_ncu_top_kernels_16m.csv)
NCU_TRAFFIC_16M = {"density": 0.870e9 + 1.554e9, "force": 2.010e9 + 0.852e9}
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
    # sparser lattices (nu < 40) are taller than the configured box: grow it (SURVEY config 5)
    need = [int(np.ceil(o / 0.2 + s * float(d) / 0.2)) + 2 for o, s in zip(origin, (nx, ny, nz))]
    cfg = dict(cfg, grid=tuple(max(g, m) for g, m in zip(cfg["grid"], need)))
    return cfg, nx, ny, nz, d, origin


# ----------------------------------------------------------------------------
def column_scene(world, rank, strong, sites_xy=(256, 128), planes=512):
    """Multi-GPU scene: ONE continuous jittered lattice column along z, 256x128 sites
    in x,y and 512 z-sites per GPU (weak scaling, SURVEY config 4: 16.7M per GPU) or 512
    z-sites in total (strong scaling, config 3), lifted 16 voxels off the floor and
    centred in x.  The box is cut into z-slabs on voxel layers so that every rank
    owns the same number of lattice planes (+-1).  Each rank generates only its own
    particles (counter-based jitter) and keeps those whose voxel layer it owns."""
    import smoothed_particle_hydrodynamics_b200 as S
    nx, ny = sites_xy
    nz = planes if strong else planes * world
    d = scenes.lattice_spacing(0.1, NU)
    vox = 0.2
    oz = 3
    origin = (49 * vox, 16 * vox, oz * vox)
    extent = nz * float(d) / vox                       # column height in voxel layers
    gz = int(np.ceil(oz + extent)) + 4
    bounds = [int(round(oz + extent * r / world)) for r in range(1, world)]
    layers = S.slab_layers(gz, world, bounds)
    z0, z1 = layers[rank]
    # lattice planes that can reach this slab (jitter is +-0.1 d): one plane of margin
    zpos = lambda iz: (origin[2] + (iz + 0.5) * float(d)) / vox
    planes = [iz for iz in range(nz) if z0 - 1 <= zpos(iz) < z1 + 1]
    first, last = (planes[0], planes[-1] + 1) if planes else (0, 0)
    first_id, count = first * nx * ny, (last - first) * nx * ny
    pos = np.empty((count, 3), np.float32)
    if count:
        S.scene_lattice(nx, ny, nz, d, origin, first_id=first_id, count=count, out=pos)
    inv2h = np.float32(1.0) / (np.float32(0.1) * np.float32(2.0))
    vz = S.voxel_layer(pos[:, 2], inv2h, gz)
    own = (vz >= z0) & (vz < z1)
    gids = (np.arange(count, dtype=np.int64) + first_id)[own].astype(np.uint32)
    return dict(grid=(160, 64, gz), layers=layers, pos=pos[own], gids=gids, total=nx * ny * nz,
                sites=(nx, ny, nz))


def run_ours(args):
    import torch
    import smoothed_particle_hydrodynamics_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = scenes.scene_params(nu=NU)
    common = dict(examine_count=96, neighbor_mode=S.FULL, use_uniform_gravity=1, use_wall_collision=1,
                  rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0,
                  gravity=sp["gravity"], time_step=sp["time_step"], kernel_variant=args.kernel_variant)
    strong = args.scaling == "strong"
    force_slab = world == 1 and args.force_slab
    if world == 1 and not force_slab:
        name = args.workload or "dambreak_16m"
        cfg, nx, ny, nz, d, origin = workload_spec(name)
        n = n_total = nx * ny * nz
        grid = cfg["grid"]
        sph = S.SPH(S.default_params(particle_count=n, grid=grid, **common), device=local, init_scene=False)
        capacity = n
        workload = ("%s: %dx%dx%d jittered lattice = %d particles, h=0.1, lattice spacing for ~%.0f neighbours "
                    "(continuum), voxel grid %s, FULL neighbour mode, gravity+walls on"
                    % (name, nx, ny, nz, n, NU, "x".join(map(str, grid))))
    else:
        sc = column_scene(world, rank, strong)
        n = sc["gids"].size
        n_total = sc["total"]
        grid = sc["grid"]
        z0, z1 = sc["layers"][rank]
        capacity = int(n * 1.12) + 400000
        nccl_id = None
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(S.SlabSPH.unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            nccl_id = bytes(idt.cpu().numpy().tobytes())
        sph = S.SlabSPH(S.default_params(particle_count=capacity, grid=grid, **common), rank, world, z0, z1,
                        nccl_id=nccl_id, device=local)
        name = "column_%s_%dgpu" % ("strong16m" if strong else "weak16m_per_gpu", world)
        workload = ("%s: continuous %dx%dx%d jittered lattice column = %d particles (%s), z-slabs of voxel layers "
                    "%s, one ghost voxel layer + migration per step over NCCL, h=0.1, voxel grid %s, FULL mode, "
                    "gravity+walls on" % (name, *sc["sites"], n_total,
                                          "16.7M per GPU" if not strong else "16.7M in total", sc["layers"],
                                          "x".join(map(str, grid))))
    stream = torch.cuda.Stream()
    sph.set_stream(stream.cuda_stream)

    # synthetic scene in pinned host memory (the e2e leg copies from / to it every step)
    pos_h = torch.empty((capacity, 3), dtype=torch.float32).pin_memory()
    vel_h = torch.zeros((capacity, 3), dtype=torch.float32).pin_memory()
    mass_h = torch.ones((capacity,), dtype=torch.float32).pin_memory()
    gid_h = torch.zeros((capacity,), dtype=torch.int32).pin_memory()

    slab = world > 1 or force_slab

    def load_scene():
        vel_h.zero_()
        if not slab:
            S.scene_lattice(nx, ny, nz, d, origin, out=pos_h.numpy())
            sph.upload_ptr(pos_h.data_ptr(), vel_h.data_ptr(), mass_h.data_ptr())
        else:
            pos_h.numpy()[:n] = sc["pos"]
            gid_h.numpy()[:n] = sc["gids"].view(np.int32)
            sph._check(sph._lib.sphb200_upload_slab(sph._h, n, pos_h.data_ptr(), vel_h.data_ptr(),
                                                    mass_h.data_ptr(), gid_h.data_ptr()))
    load_scene()
    sph.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident throughput ----------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()            # every rank has its scene on the device before the first exchange
    sph.step_n(args.warmup)
    barrier()
    l0 = sph.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        sph.step_n(args.steps)
        e1.record(stream)
    barrier()
    t_end = time.time()
    ms = allmax(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, t_end)
    launches = sph.launch_count() - l0
    pairs_last = allsum(sph.neighbor_stats()[0])
    if slab:
        sph.status()
    ms_per_step = ms / args.steps
    value = n_total * args.steps / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events inside the library, same stream) --
    sph.set_params(enable_timers=1)
    phase = np.zeros(6)
    reps = min(args.steps, 5)
    for _ in range(reps):
        sph.step_n(1)
        phase += np.array(sph.timings_ms())
    phase /= reps
    sph.set_params(enable_timers=0)
    dens_ms, force_ms = float(phase[2]), float(phase[4])
    hbm, peak_kind = measured_peaks()
    n_dev = n
    # the dominant kernel of the step, decided live: the force sweep or the density sweep
    if force_ms >= dens_ms:
        dom = {"key": "force", "ms": force_ms, "bytes": ALG_BYTES_FORCE,
               "kernel": "k_force_stream (pressure + viscosity + integrate + walls, hit-mask stream driven)",
               "note": "bound by L1 wavefronts of the scattered 16-byte neighbour gathers (ncu: "
                       "l1tex__data_pipe_lsu_wavefronts 86%, DRAM 14% busy), not by HBM: DESIGN.md section 5"}
    else:
        dom = {"key": "density", "ms": dens_ms, "bytes": ALG_BYTES_DENSITY,
               "kernel": "k_density_tiled (density + EOS + hit-mask stream sweep, packed FP32)",
               "note": "FP32 pipe / issue bound (ncu: FMA pipe 61% of active cycles at 2 cycles per packed "
                       "instruction, DRAM 13% busy), not HBM bound: DESIGN.md section 5"}
    achieved = dom["bytes"] * n_dev / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0

    # ---- end to end through host buffers -------------------------------------
    # every step: H2D of positions / velocities / masses (/ ids) from pinned memory,
    # the step, D2H of the new positions and velocities into pinned memory
    load_scene()
    e2e_steps = max(1, min(args.steps, 10))

    def e2e_step():
        if not slab:
            sph.step_host_ptr(pos_h.data_ptr(), vel_h.data_ptr(), mass_h.data_ptr())
        else:
            sph._check(sph._lib.sphb200_upload_slab(sph._h, n, pos_h.data_ptr(), vel_h.data_ptr(),
                                                    mass_h.data_ptr(), gid_h.data_ptr()))
            sph.step_n(1)
            sph._check(sph._lib.sphb200_download(sph._h, S.Field.POSITION, pos_h.data_ptr(), pos_h.numel() * 4))
            sph._check(sph._lib.sphb200_download(sph._h, S.Field.VELOCITY, vel_h.data_ptr(), vel_h.numel() * 4))
    for _ in range(2):
        e2e_step()
    if slab:
        load_scene()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()      # (the id buffer is an input only: nothing writes it, nothing to refill)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = n_total * e2e_steps / e2e_s
    h2d = (32 if slab else 28) * n
    d2h = 24 * (capacity if slab else n)

    out = {
        "metric": "particle-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong and world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "particles": n_total,
                   "neighbor_pairs_per_sec": pairs_last / (ms_per_step * 1e-3),
                   "mean_neighbors": pairs_last / n_total,
                   "l2": "inputs (>= 512 MB of state per GPU) larger than the 126 MB L2",
                   "halo": ("peer puts by the force sweep (CUDA IPC / NVLink)" if sph.put_mode() else
                            "grouped ncclSend/ncclRecv") if slab and world > 1 else None,
                   "step_alg_bytes_per_particle": ALG_BYTES_STEP,
                   "step_hbm_frac": ALG_BYTES_STEP * n_dev / (ms_per_step * 1e-3) / 1e9 / hbm,
                   "phase_ms_rank0": {"exchange_bin_sort_gather": float(phase[0]), "density_eos": dens_ms,
                                      "force_integrate": force_ms, "reduce": float(phase[5])}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom["kernel"],
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": NCU_TRAFFIC_16M[dom["key"]] if (n_dev == 16777216 and NU == 40.0) else None,
                     "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_ncu_top_kernels_16m.csv)",
                     "peak_kind": peak_kind, "alg_bytes_per_particle": dom["bytes"], "kernel_ms": dom["ms"],
                     "other_kernel_ms": {"density": dens_ms, "force": force_ms},
                     "note": dom["note"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference_sample(steps=10, warmup=1)
    if rank == 0:
        emit(out)
    barrier()            # peer-put halos: no rank frees its receive buffers while a neighbour may still write
    sph.close()
    if world > 1:
        dist.destroy_process_group()


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
    E = 96 if NU <= 40.0 else int(NU * 1.6) + 32
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
    emit(out)


def _claim_stdout():
    """Libraries (NCCL prints its version banner) must not pollute the one JSON line
    rank 0 prints: fd 1 is pointed at stderr for the run and the saved descriptor is
    used for the result."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


RESULT_OUT = None


def emit(obj):
    RESULT_OUT.write(json.dumps(obj) + "\n")
    RESULT_OUT.flush()


def main():
    global RESULT_OUT
    RESULT_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nu", type=float, default=40.0,
                    help="lattice spacing for this many neighbours in the continuum limit (config 5 sweep: 30/60/120)")
    ap.add_argument("--kernel-variant", type=int, default=0,
                    help="A/B: 0 = tiled density + flat force sweep (default), 3 = force sweep tiled in shared memory, 1 = untiled")
    ap.add_argument("--force-slab", action="store_true", help="N=1: run the slab code path as a single slab")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: 16.7M particles per GPU (weak, default) or 16.7M in total (strong)")
    args = ap.parse_args()
    global NU
    NU = float(args.nu)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
