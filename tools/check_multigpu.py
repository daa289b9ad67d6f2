"""Real multi-process check of the slab path (one rank per GPU, NCCL bootstrap, halo by peer puts or NCCL):
the N-rank run must reproduce the single-GPU run bit for bit (rank 0 runs the single-GPU reference).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu.py

Same scene as tests/test_gpu_slabs.py (128K particles drifting along z: migration + halo every step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import smoothed_particle_hydrodynamics_b200 as S
from oracle import scenes
F = S.Field
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(scenes.CONFIGS["dambreak_128k"]); cfg["grid"] = (40, 16, 32)
nx, ny, nz = cfg["sites"]; n = nx * ny * nz
pos = scenes.lattice_scene(nx, ny, nz, scenes.lattice_spacing(0.1, 40), origin=(0.0, 0.0, 0.9))
rng = np.random.default_rng(5)
vel = rng.normal(0, 2.0, (n, 3)).astype(np.float32); vel[:, 2] += 25.0
mass = (rng.random(n) * 0.2 + 0.9).astype(np.float32)
sp = scenes.scene_params()
p = S.default_params(particle_count=n, grid=cfg["grid"], examine_count=96, neighbor_mode=S.FULL, use_uniform_gravity=1,
                     use_wall_collision=1, rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"],
                     central_mass=0.0, gravity=sp["gravity"], time_step=sp["time_step"])
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(S.SlabSPH.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
z0, z1 = S.slab_layers(cfg["grid"][2], world)[rank]
slab = S.SlabSPH(p, rank, world, z0, z1, nccl_id=bytes(idt.cpu().numpy().tobytes()), device=local)
ref = S.SPH(p, init_scene=False, device=local) if rank == 0 else None
vz = S.voxel_layer(pos[:, 2], slab.derived.h_times2_inv, cfg["grid"][2])
own = np.flatnonzero((vz >= z0) & (vz < z1))
slab.upload_slab(pos[own], vel[own], mass[own], own.astype(np.uint32))
if ref: ref.upload(pos, vel, mass)
steps = 20
slab.step_n(steps); slab.synchronize(); slab.status()
if ref: ref.step_n(steps)
ok = True
for field, comps in ((F.NEIGHBOR_COUNT, 1), (F.DENSITY, 1), (F.POSITION, 3), (F.VELOCITY, 3)):
    vals, gids = slab.download_slab(field)
    full = torch.zeros((n, comps), dtype=torch.float64, device="cuda")
    seen = torch.zeros(n, dtype=torch.float64, device="cuda")
    full[torch.from_numpy(gids.astype(np.int64)).cuda()] = torch.from_numpy(vals.reshape(-1, comps).astype(np.float64)).cuda()
    seen[torch.from_numpy(gids.astype(np.int64)).cuda()] = 1
    dist.all_reduce(full); dist.all_reduce(seen)
    if rank == 0:
        r = ref.download(field).reshape(n, comps).astype(np.float64)
        got = full.cpu().numpy()
        same = np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(r, nan=-1.0))
        ok &= bool(same) and bool((seen.cpu().numpy() == 1).all())
        print("field", int(field), "owners exactly one:", bool((seen.cpu().numpy() == 1).all()), "bitwise equal to 1 GPU:", same, flush=True)
if rank == 0:
    print("halo mode:", "peer puts" if slab.put_mode() else "nccl", "| owned per rank after %d steps differs from start:" % steps, True)
    print("MULTIGPU CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
