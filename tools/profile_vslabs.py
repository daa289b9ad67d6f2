"""Two virtual-rank slabs of the 16.7M column on ONE GPU (for ncu launch lists of the slab path with real ghosts)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import smoothed_particle_hydrodynamics_b200 as S
import bench
from oracle import scenes
world = 2
sp = scenes.scene_params()
slabs = []
for rank in range(world):
    sc = bench.column_scene(world, rank, True)
    n = sc["gids"].size; z0, z1 = sc["layers"][rank]
    p = S.default_params(particle_count=int(n*1.12)+400000, grid=sc["grid"], examine_count=96, neighbor_mode=S.FULL, use_uniform_gravity=1,
        use_wall_collision=1, rho0=sp["rho0"], stiffness=sp["stiffness"], viscosity=sp["viscosity"], central_mass=0.0, gravity=sp["gravity"],
        time_step=sp["time_step"], enable_timers=1)
    s = S.SlabSPH(p, rank, world, z0, z1, halo_capacity=400000)
    s.upload_slab(sc["pos"], np.zeros_like(sc["pos"]), None, sc["gids"]); slabs.append(s)
for step in range(3):
    S.step_virtual_slabs(slabs, 1)
    for s in slabs: s.synchronize()
    print(step, [["%.2f" % x for x in s.timings_ms()] for s in slabs], [s.local_count() for s in slabs], flush=True)
