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

ALG_BYTES_STEP = 240.0     # algorithmic bytes per particle-step (whole pipeline)
ALG_BYTES_FORCE = 72.0     # force+integrate+collide sweep: R(16+16+4) + W(16+16) + R4
ALG_BYTES_DENSITY = 20.0   # density+EOS sweep: R16 + W4
NU = 40.0
# per-launch counters of the two sweeps from the committed `ncu --set full` capture of this very
# command (tools/ncu_summarize.py writes the file from the raw ncu csv): DRAM bytes, FP32 operations,
# what binds.  Read, not typed in: a stale or missing file yields traffic = null.
KERNEL_METRICS = os.path.join(ROOT, "profiles", "kernel_metrics.json")


def kernel_metrics(key, particles, nu):
    try:
        with open(KERNEL_METRICS) as f:
            m = json.load(f)
        if int(m["particles"]) != int(particles) or float(m["nu"]) != float(nu):
            return None
        return dict(m["kernels"][key], source=m.get("source"))
    except (OSError, KeyError, ValueError):
        return None


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


def scene_kwargs(p):
    """The physical parameters of a configured scene (sphb200_scene_config) as create() keywords."""
    return dict(examine_count=p.examine_count, neighbor_mode=p.neighbor_mode, use_uniform_gravity=p.use_uniform_gravity,
                use_wall_collision=p.use_wall_collision, rho0=p.rho0, stiffness=p.stiffness, viscosity=p.viscosity,
                central_mass=p.central_mass, gravity=tuple(p.gravity), time_step=p.time_step)


# ----------------------------------------------------------------------------
def column_scene(S, world, rank, strong, sites_xy=(256, 128), planes=512):
    """Multi-GPU scene: ONE continuous jittered lattice column along z, 256x128 sites
    in x,y and 512 z-sites per GPU (weak scaling, SURVEY config 4: 16.7M per GPU) or 512
    z-sites in total (strong scaling, config 3), lifted off the floor and centred in x like
    the box-drop scene.  The box is cut into z-slabs on voxel layers so that every rank
    owns the same number of lattice planes (+-1).  Each rank generates only its own
    particles (counter-based jitter) and keeps those whose voxel layer it owns."""
    p, lat = S.scene_config("boxdrop_16m", NU)
    nx, ny = sites_xy
    nz = planes if strong else planes * world
    d = lat.spacing
    vox = 0.2
    oz = 3
    origin = (lat.origin[0], lat.origin[1], oz * vox)
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
    return dict(grid=(p.grid_x, p.grid_y, gz), layers=layers, pos=pos[own], gids=gids, total=nx * ny * nz,
                sites=(nx, ny, nz), params=p)


def id_checksums(gids):
    """(count, sum, xor) of a rank's owned global ids -- combined over ranks by sum / sum / xor."""
    g = np.asarray(gids).astype(np.uint64)
    return int(g.size), int(g.sum()), (int(np.bitwise_xor.reduce(g)) if g.size else 0)


def id_checksums_ok(count, idsum, idxor, n_total):
    """Every id 0 .. n_total-1 exactly once <=> (necessary) count, sum and xor equal their closed forms."""
    xor_ref = [n_total - 1, 1, n_total, 0][(n_total - 1) % 4]
    return {"owned_total": int(count), "expected": int(n_total), "id_sum_ok": idsum == n_total * (n_total - 1) // 2,
            "id_xor_ok": idxor == xor_ref}


class Dist:
    """torch.distributed plumbing (bootstrap, barriers, max / sum over ranks); a no-op for one rank."""

    def __init__(self, torch, world, local):
        self.torch, self.world, self.dist = torch, world, None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _all(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._all(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x):
        return self._all(x, self.dist.ReduceOp.SUM if self.dist else None)

    def nccl_id(self, S, rank):
        if not self.dist:
            return None
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(self.torch.frombuffer(bytearray(S.SlabSPH.unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


class Job:
    """One scene on this rank's GPU: the context, its pinned host buffers and the measurements."""

    def __init__(self, S, torch, D, rank, world, local, args, slab, strong=False, name=None):
        self.S, self.torch, self.D, self.rank, self.world, self.slab = S, torch, D, rank, world, slab
        kv = dict(kernel_variant=args.kernel_variant)
        if not slab:
            name = name or "dambreak_16m"
            p, lat = S.scene_config(name, NU)
            self.lat = lat
            self.n = self.n_total = self.capacity = p.particle_count
            grid = (p.grid_x, p.grid_y, p.grid_z)
            self.sph = S.SPH(S.default_params(particle_count=self.n, grid=grid, **scene_kwargs(p), **kv), device=local,
                             init_scene=False)
            self.workload = ("%s: %dx%dx%d jittered lattice = %d particles, h=0.1, lattice spacing for ~%.0f "
                             "neighbours (continuum), voxel grid %s, FULL neighbour mode, gravity+walls on"
                             % (name, lat.nx, lat.ny, lat.nz, self.n, NU, "x".join(map(str, grid))))
        else:
            sc = self.sc = column_scene(S, world, rank, strong)
            self.n, self.n_total = sc["gids"].size, sc["total"]
            z0, z1 = sc["layers"][rank]
            # slots: the owned particles, two ghost voxel layers and migration head-room (every kernel
            # runs over the slots, so the margin is kept small: 6 % + 300 K)
            self.capacity = int(self.n * 1.06) + 300000
            self.sph = S.SlabSPH(S.default_params(particle_count=self.capacity, grid=sc["grid"],
                                                  **scene_kwargs(sc["params"]), **kv),
                                 rank, world, z0, z1, nccl_id=D.nccl_id(S, rank), device=local)
            name = "column_%s_%dgpu" % ("strong16m" if strong else "weak16m_per_gpu", world)
            self.workload = ("%s: continuous %dx%dx%d jittered lattice column = %d particles (%s), z-slabs of voxel "
                             "layers %s, one ghost voxel layer + migration per step (%s), h=0.1, voxel grid %s, FULL "
                             "mode, gravity+walls on"
                             % (name, *sc["sites"], self.n_total, "16.7M in total" if strong else "16.7M per GPU",
                                sc["layers"], "HALO", "x".join(map(str, sc["grid"]))))
        self.name = name
        self.stream = torch.cuda.Stream()
        self.sph.set_stream(self.stream.cuda_stream)
        cap = self.capacity
        # synthetic scene in pinned host memory (the e2e leg copies from / to it every step)
        self.pos_h = torch.empty((cap, 3), dtype=torch.float32).pin_memory()
        self.vel_h = torch.zeros((cap, 3), dtype=torch.float32).pin_memory()
        self.mass_h = torch.ones((cap,), dtype=torch.float32).pin_memory()
        self.gid_h = torch.zeros((cap,), dtype=torch.int32).pin_memory()
        self.load_scene()
        self.sph.synchronize()
        if slab:
            halo = ("peer puts by the force sweep (CUDA IPC / NVLink)" if self.sph.put_mode() else
                    "grouped ncclSend/ncclRecv") if world > 1 else "single slab, no neighbour"
            self.halo = halo
            self.workload = self.workload.replace("HALO", halo)
        else:
            self.halo = None

    def load_scene(self):
        S, sph = self.S, self.sph
        self.vel_h.zero_()
        if not self.slab:
            S.scene_generate(self.lat, out=self.pos_h.numpy())
            sph.upload_ptr(self.pos_h.data_ptr(), self.vel_h.data_ptr(), self.mass_h.data_ptr())
        else:
            self.pos_h.numpy()[:self.n] = self.sc["pos"]
            self.gid_h.numpy()[:self.n] = self.sc["gids"].view(np.int32)
            self.upload_slab()

    def upload_slab(self):
        sph = self.sph
        sph._check(sph._lib.sphb200_upload_slab(sph._h, self.n, self.pos_h.data_ptr(), self.vel_h.data_ptr(),
                                                self.mass_h.data_ptr(), self.gid_h.data_ptr()))

    def timed(self, steps, warmup):
        """`steps` device-resident steps after `warmup`: CUDA events on the stream the kernels run on,
        barrier + synchronize on both sides, max over ranks."""
        torch, D, sph = self.torch, self.D, self.sph
        D.barrier()            # every rank has its scene on the device before the first exchange
        sph.step_n(warmup)
        D.barrier()
        l0 = sph.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        t_begin = time.time()
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            sph.step_n(steps)
            e1.record(self.stream)
        D.barrier()
        t_end = time.time()
        ms = D.max(e0.elapsed_time(e1))
        launches = sph.launch_count() - l0
        pairs = D.sum(sph.neighbor_stats()[0])
        if self.slab:
            sph.status()
        return dict(ms=ms, ms_per_step=ms / steps, value=self.n_total * steps / (ms * 1e-3), launches=launches,
                    pairs=pairs, t_begin=t_begin, t_end=t_end)

    def phases(self, reps):
        """Per-kernel durations: CUDA events inside the library, on the same stream."""
        sph = self.sph
        sph.set_params(enable_timers=1)
        phase = np.zeros(6)
        for _ in range(reps):
            sph.step_n(1)
            phase += np.array(sph.timings_ms())
        sph.set_params(enable_timers=0)
        return phase / reps

    def integrity(self):
        """Slab runs check their own result: every particle is owned by exactly one rank (count, sum
        and xor of the global ids against the closed forms for ids 0 .. n_total-1), no exchange error,
        and the state is finite.  The neighbour total is compared by the caller."""
        sph, D = self.sph, self.D
        sph.status()
        pos, gids = sph.download_slab(self.S.Field.POSITION)
        count, idsum, idxor = id_checksums(gids)
        # (sums stay below 2^53: 134M ids sum to 9.0e15)
        count, idsum = int(D.sum(float(count))), int(D.sum(float(idsum)))
        if D.dist:
            t = self.torch.tensor([idxor], device="cuda", dtype=self.torch.int64)
            parts = [self.torch.zeros_like(t) for _ in range(self.world)]
            D.dist.all_gather(parts, t)
            idxor = 0
            for q in parts:
                idxor ^= int(q.item())
        out = id_checksums_ok(count, idsum, idxor, self.n_total)
        out["state_finite"] = D.sum(float(np.isfinite(pos).all())) == self.world
        out["slab_status"] = "ok"
        out["ok"] = bool(out["owned_total"] == out["expected"] and out["id_sum_ok"] and out["id_xor_ok"]
                         and out["state_finite"])
        return out

    def e2e(self, steps):
        """End to end through host buffers: every step H2D of positions / velocities / masses (/ ids) from
        pinned memory, the step, D2H of the new positions and velocities into pinned memory."""
        S, sph, D = self.S, self.sph, self.D
        self.load_scene()

        def one():
            if not self.slab:
                sph.step_host_ptr(self.pos_h.data_ptr(), self.vel_h.data_ptr(), self.mass_h.data_ptr())
            else:
                self.upload_slab()
                sph.step_n(1)
                sph._check(sph._lib.sphb200_download(sph._h, S.Field.POSITION, self.pos_h.data_ptr(),
                                                     self.pos_h.numel() * 4))
                sph._check(sph._lib.sphb200_download(sph._h, S.Field.VELOCITY, self.vel_h.data_ptr(),
                                                     self.vel_h.numel() * 4))
        for _ in range(2):
            one()
        if self.slab:
            self.load_scene()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()      # (the id buffer is an input only: nothing writes it, nothing to refill)
        D.barrier()
        sec = D.max(time.perf_counter() - t0)
        h2d = (32 if self.slab else 28) * self.n
        d2h = 24 * (self.capacity if self.slab else self.n)
        return {"value": self.n_total * steps / sec, "unit": "particle-updates/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": steps}

    def close(self):
        self.D.barrier()     # peer-put halos: no rank frees its receive buffers while a neighbour may still write
        self.sph.close()


def roofline_block(job, phase, clocks, hbm, peak_kind):
    """The dominant kernel of the step, decided live, against (1) the contract figure -- algorithmic
    bytes / CUDA-event duration vs the measured HBM copy bandwidth -- and (2) what actually binds it:
    the FP32 pipe (148 SMs x 128 lanes x 2 flop x clock) with the FP32 operation count of the committed
    ncu capture, and that capture's DRAM traffic."""
    dens_ms, force_ms = float(phase[2]), float(phase[4])
    # the two sweeps take the same time to within run-to-run noise; a near-tie (3 %) goes to the force sweep, the
    # one that carries more of the step's algorithmic bytes (both are listed under "sweeps")
    if force_ms >= 0.97 * dens_ms:
        dom = {"key": "force", "ms": force_ms, "bytes": ALG_BYTES_FORCE, "bound": "l1",
               "kernel": "k_force_stream (pressure + viscosity + integrate + walls, hit-mask stream driven)",
               "note": "bound by the L1 data pipe: two scattered 16-byte neighbour gathers per pair, ~13 distinct "
                       "128-byte lines per warp-wide load (ncu: l1tex__data_pipe_lsu_wavefronts); HBM and the FP32 "
                       "pipe are far from saturated: DESIGN.md section 5"}
    else:
        dom = {"key": "density", "ms": dens_ms, "bytes": ALG_BYTES_DENSITY, "bound": "fp32",
               "kernel": "k_density_persist (density + EOS + hit-mask stream sweep, packed FP32, x-trimmed runs)",
               "note": "bound by the FP32 pipe / instruction issue (ncu: FMA pipe cycles, packed f32x2 instructions "
                       "hold the pipe two cycles), not by HBM: DESIGN.md section 5"}
    achieved = dom["bytes"] * job.n / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    km = kernel_metrics(dom["key"], job.n, NU) if not job.slab else None
    clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
    peak_tf = 148 * 128 * 2 * clk / 1e12
    fp32 = None
    if km and dom["ms"] > 0 and km.get("fma_pipe_pct") is not None:
        # ncu's op counters do not see packed f32x2 instructions, so the FP32 figure is the FMA pipe's
        # busy fraction of the capture, rescaled by captured / live kernel duration
        frac = km["fma_pipe_pct"] / 100.0 * km["ncu_ms"] / dom["ms"]
        fp32 = {"bound": "fp32", "achieved_tflops": frac * peak_tf, "peak_tflops": peak_tf, "frac": frac,
                "fma_pipe_pct_ncu": km.get("fma_pipe_pct"), "l1_lsu_wavefront_pct_ncu": km.get("lsu_wavefront_pct"),
                "issue_active_pct_ncu": km.get("issue_pct"), "dram_pct_ncu": km.get("dram_pct"),
                "inst_executed_ncu": km.get("inst_executed"), "ncu_kernel_ms": km["ncu_ms"],
                "peak_formula": "148 SMs x 128 lanes x 2 flop x %.0f MHz" % (clk / 1e6)}
    sweeps = [{"kernel": name, "kernel_ms": ms, "alg_bytes_per_particle": b, "achieved": b * job.n / (ms * 1e-3) / 1e9,
               "frac": b * job.n / (ms * 1e-3) / 1e9 / hbm}
              for name, ms, b in (("k_tile_list + k_density_persist", dens_ms, ALG_BYTES_DENSITY), ("k_force_stream", force_ms, ALG_BYTES_FORCE),
                                  ("k_cell_keys + scan + k_scatter + k_rank_gather", float(phase[0]),
                                   ALG_BYTES_STEP - ALG_BYTES_DENSITY - ALG_BYTES_FORCE)) if ms > 0]
    return {"bound": dom["bound"], "contract_bound": "hbm", "kernel": dom["kernel"], "sweeps": sweeps,
            "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": km["dram_bytes"] if km else None,
            "traffic_unit": "bytes per launch: ncu dram__bytes_read.sum + dram__bytes_write.sum (%s)"
                            % (km["source"] if km else "no capture committed for this workload"),
            "peak_kind": peak_kind, "alg_bytes_per_particle": dom["bytes"], "kernel_ms": dom["ms"],
            "other_kernel_ms": {"density": dens_ms, "force": force_ms},
            "secondary": fp32, "note": dom["note"]}


def bind_to_gpu_numa(torch, local):
    """Pinned host buffers and the submitting thread on the NUMA node of this rank's GPU (N ranks share
    the host's memory controllers during the e2e leg)."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


def run_ours(args):
    import torch
    import smoothed_particle_hydrodynamics_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local)
    D = Dist(torch, world, local)
    strong = args.scaling == "strong"
    slab = world > 1 or args.force_slab
    job = Job(S, torch, D, rank, world, local, args, slab, strong=strong, name=args.workload)

    # ---- device-resident throughput ----------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    t = job.timed(args.steps, args.warmup)
    clocks = sampler.stop(t["t_begin"], t["t_end"])
    phase = job.phases(min(args.steps, 5))
    hbm, peak_kind = measured_peaks()
    config = {"workload": job.workload, "particles": job.n_total,
              "neighbor_pairs_per_sec": t["pairs"] / (t["ms_per_step"] * 1e-3),
              "mean_neighbors": t["pairs"] / job.n_total,
              "l2": "inputs (>= 512 MB of state per GPU) larger than the 126 MB L2",
              "halo": job.halo, "numa_node": numa,
              "step_alg_bytes_per_particle": ALG_BYTES_STEP,
              "step_hbm_frac": ALG_BYTES_STEP * job.n / (t["ms_per_step"] * 1e-3) / 1e9 / hbm,
              "phase_ms_rank0": {"exchange_bin_sort_gather": float(phase[0]), "density_eos": float(phase[2]),
                                 "force_integrate": float(phase[4]), "reduce": float(phase[5])}}
    if slab:
        config["integrity"] = job.integrity()
    e2e = job.e2e(max(1, min(args.steps, 10)))
    out = {
        "metric": "particle-updates/sec", "value": t["value"], "unit": "particle-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong and world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(t["launches"]),
        "roofline": roofline_block(job, phase, clocks, hbm, peak_kind),
    }
    job.close()

    # ---- the other north-star configurations, measured in the same run ---------------------
    # (each is its own scene and context; K and W as above; skipped with --no-extras)
    if not args.no_extras:
        extras = {}
        if world == 1 and not slab:
            # the denominator of the weak-scaling curve on ITS workload: the 16.7M-per-GPU column scene
            # through the slab code path as a single slab (what every rank of an N-GPU run executes)
            j = Job(S, torch, D, rank, world, local, args, True)
            r = j.timed(args.steps, args.warmup)
            extras["weak_scaling_base"] = {"workload": j.workload, "value": r["value"], "ms_per_step": r["ms_per_step"],
                                           "mean_neighbors": r["pairs"] / j.n_total, "integrity": j.integrity(),
                                           "use": "efficiency_N = value_N / (N x this value) compares one workload"}
            j.close()
        if world > 1 and not strong:
            # config 3: the 16.7M scene cut into N slabs (strong scaling)
            j = Job(S, torch, D, rank, world, local, args, True, strong=True)
            r = j.timed(args.steps, args.warmup)
            extras["strong_scaling"] = {"workload": j.workload, "value": r["value"], "ms_per_step": r["ms_per_step"],
                                        "mean_neighbors": r["pairs"] / j.n_total, "integrity": j.integrity()}
            j.close()
        # config 5: neighbour-density sweep at 16.7M particles (in total): 1 GPU, or N slabs
        global NU
        nu0, sweep = NU, []
        for nu in ([30.0, 60.0, 120.0] if nu0 == 40.0 else []):
            NU = nu
            j = Job(S, torch, D, rank, world, local, args, world > 1, strong=True)
            r = j.timed(max(2, args.steps // 4), 3)
            sweep.append({"nu": nu, "mean_neighbors": r["pairs"] / j.n_total, "ms_per_step": r["ms_per_step"],
                          "value": r["value"], "neighbor_pairs_per_sec": r["pairs"] / (r["ms_per_step"] * 1e-3)})
            j.close()
        NU = nu0
        if sweep:
            extras["neighbor_sweep_16m"] = sweep
        out["config"]["extras"] = extras
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference_sample(steps=10, warmup=1)
    if rank == 0:
        emit(out)
    D.barrier()
    D.close()


# ----------------------------------------------------------------------------
def cpu_reference_sample(steps, warmup, sample="dambreak_1m"):
    """The reference's own computeDensity / computeAcceleration / integrate (+ the
    harness all-within-h search and the dead wall code) from oracle/_ref, timing
    build, 1 thread, on a bounded sample of the 16M workload (same lattice spacing and
    parameters; the scene tables of the checker, oracle/scenes.py, generate it)."""
    from oracle import refharness, scenes
    kind = "reference" if refharness.available("timing") else "port"
    cfg = scenes.CONFIGS[sample]
    nx, ny, nz = cfg["sites"]
    d = scenes.lattice_spacing(0.1, NU)
    origin = [v * 0.2 for v in cfg["origin_vox"]]
    need = [int(np.ceil(o / 0.2 + s * float(d) / 0.2)) + 2 for o, s in zip(origin, (nx, ny, nz))]
    grid = tuple(max(g, m) for g, m in zip(cfg["grid"], need))
    n = nx * ny * nz
    sp = scenes.scene_params(nu=NU)
    pos = scenes.lattice_scene(nx, ny, nz, d, origin)
    vel = np.zeros((n, 3), np.float32)
    E = 96 if NU <= 40.0 else int(NU * 1.6) + 32
    t_steps = []
    pairs = 0
    if kind == "reference":
        r = refharness.RefSPH("timing")
        r.resize(n, *grid, E)
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
        o = OracleSPH(n=n, grid=grid, examine=E, init_scene=False, rho0=sp["rho0"], stiffness=sp["stiffness"],
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
    ratio = 16777216 // n
    return {"value": n / t, "unit": "particle-updates/s", "cores": 1, "kind": kind,
            "sample": "%s: %dx%dx%d lattice = %d particles (1/%d of the 16.7M workload, same spacing and parameters), "
                      "%d timed steps of the FULL-mode harness step" % (sample, nx, ny, nz, n, ratio, steps),
            "sample_ratio": ratio, "ms_per_step": t * 1e3, "neighbor_pairs_per_sec": pairs / t,
            "phase_ms_last": phases,
            "phase_names": ["voxelize", "findNeighbors (harness all-within-h search, FULL mode)", "computeDensity",
                            "computePressure", "computeAcceleration", "integrate + walls"],
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the reference does ~1 us per particle-step on one core (FULL harness step): the largest sample of
    # the 16.7M scene whose K + W steps end within about two minutes -- the true scene when K + W is small
    total = max(1, args.steps + args.warmup)
    sample = "dambreak_128k"
    for name, n in (("dambreak_16m", 16777216), ("dambreak_4m", 4194304), ("dambreak_1m", 1048576)):
        if total * n * 1.05e-6 <= 130.0:
            sample = name
            break
    cb = cpu_reference_sample(args.steps, args.warmup, sample)
    name = args.workload or ("dambreak_16m" if world == 1 else "boxdrop_16m")
    out = {"impl": "reference", "metric": "particle-updates/sec", "value": cb["value"], "unit": "particle-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "sampled_as": cb["sample"], "sample_ratio": cb["sample_ratio"],
                      "note": "rate per particle; the CPU step is linear in the particle count, so the sample is "
                              "if anything kind to the CPU (smaller working set)"},
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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records of the other north-star configurations (weak-scaling base, strong scaling, neighbour sweep)")
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
