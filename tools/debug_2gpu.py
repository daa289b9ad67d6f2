import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import smoothed_particle_hydrodynamics_b200 as S
import bench
from oracle import scenes
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = bench.column_scene(world, rank, False)
n = sc["gids"].size; z0, z1 = sc["layers"][rank]
sp = scenes.scene_params()
cap = int(n*1.12)+400000
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: idt.copy_(torch.frombuffer(bytearray(S.SlabSPH.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
p = S.default_params(particle_count=cap, grid=sc["grid"], examine_count=96, neighbor_mode=S.FULL, use_uniform_gravity=1, use_wall_collision=1,
    rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0, gravity=sp["gravity"], time_step=sp["time_step"], enable_timers=1)
sph = S.SlabSPH(p, rank, world, z0, z1, nccl_id=bytes(idt.cpu().numpy().tobytes()), device=local)
sph.upload_slab(sc["pos"], np.zeros_like(sc["pos"]), None, sc["gids"])
for step in range(6):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    sph.step_n(1); sph.synchronize()
    t = (time.perf_counter()-t0)*1e3
    print("rank", rank, "step", step, "wall ms %.2f" % t, "phases", ["%.2f" % x for x in sph.timings_ms()], "counts", sph.local_count(), flush=True)
# time the phases separately
torch.cuda.synchronize(); dist.barrier()
for name, fn in [("pack", sph.pack), ("unpack", sph.unpack)]:
    t0 = time.perf_counter(); fn(); sph.synchronize(); print("rank", rank, name, "ms %.3f" % ((time.perf_counter()-t0)*1e3), flush=True)
dist.destroy_process_group()
