"""BASELINE config 1: the reference's default scene (32,768 particles, seeded sphere), 1001 steps in
REFERENCE_SAMPLED mode: device-resident loop, the C++ headless driver (reference log files), and the compiled
reference (oracle/_ref timing build) on one host core."""
import os, subprocess, sys, time
sys.path.insert(0, '.')
import numpy as np
import smoothed_particle_hydrodynamics_b200 as S
sph = S.SPH()
sph.step_n(10); sph.synchronize()
t0 = time.perf_counter(); sph.step_n(1001); sph.synchronize(); t1 = time.perf_counter()
tot, mx, mn = sph.neighbor_stats()
print("GPU device-resident: 1001 steps %.3f s  (%.1f us/step, %.3e particle-updates/s), last-step neighbours %d" % (t1 - t0, (t1 - t0) / 1001 * 1e6, 32768 * 1001 / (t1 - t0), tot))
exe = "smoothed_particle_hydrodynamics_b200/sph_headless"
t0 = time.perf_counter(); r = subprocess.run([exe, "1000", "/tmp/sph_out"], capture_output=True, text=True); t1 = time.perf_counter()
print("GPU headless driver (./sph r equivalent, 1001 steps + 4 log files + per-step position readback): %.3f s  rc=%d" % (t1 - t0, r.returncode))
print(open("/tmp/sph_out/energy.txt").read().splitlines()[-1])
from oracle import refharness
if refharness.available("timing"):
    ref = refharness.RefSPH("timing")
    t0 = time.perf_counter()
    for _ in range(100): ref.step()
    t1 = time.perf_counter()
    print("reference CPU (timing build, 1 core): %.2f ms/step -> 1001 steps = %.1f s" % ((t1 - t0) * 10, (t1 - t0) * 10.01))
