import sys, time
sys.path.insert(0, '.')
import numpy as np
import smoothed_particle_hydrodynamics_b200 as S
sph = S.SPH(); sph.set_params(enable_timers=1)
acc = np.zeros(6)
for i in range(200):
    sph.step_n(1); acc += np.array(sph.timings_ms())
print("phase ms (voxelize, find, density, pressure, accel, integrate):", (acc/200).round(4), "sum", (acc/200).sum().round(4))
sph.set_params(enable_timers=0); sph.synchronize()
t0=time.perf_counter(); sph.step_n(500); sph.synchronize(); print("us/step async loop:", (time.perf_counter()-t0)/500*1e6)
