"""The C++ facade (host/sph.cpp: the reference's `class SPH` interface over the
C ABI) through the headless driver -- the reference's `./sph r` (main.cpp:23-28 ->
SPH::run, sph.cpp:149-187): step count, the four log files and their formats,
energies against the reference's own log (golden fixture)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "smoothed_particle_hydrodynamics_b200", "sph_headless")


def test_headless_run_writes_the_reference_logs(tmp_path, golden_default):
    if not os.path.exists(EXE):
        pytest.fail("sph_headless is not built (python -m smoothed_particle_hydrodynamics_b200.build)")
    out = tmp_path / "out"
    r = subprocess.run([EXE, "19", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Directory created" in r.stdout
    energy = (out / "energy.txt").read_text().splitlines()
    assert energy[0] == "Step, Kinetic Energy, Potential Energy, Total Energy"
    rows = np.array([[float(x) for x in line.split(",")] for line in energy[1:]])
    assert rows.shape == (20, 4)                        # totalSteps + 1 steps, like the reference loop
    assert np.array_equal(rows[:, 0], np.arange(20))
    ref = golden_default["energy_20"].astype(np.float64)   # the reference's own energies, 20 steps
    # first steps tightly (6 printed digits); later steps loosely (chaotic trajectories)
    np.testing.assert_allclose(rows[:3, 1:3], ref[:3], rtol=2e-5)
    np.testing.assert_allclose(rows[:, 1:3], ref, rtol=2e-3)
    np.testing.assert_allclose(rows[:, 3], rows[:, 1] + rows[:, 2], rtol=1e-4)
    assert energy[1].startswith("0, 4.69595e+06, -8.37892e+06")   # the reference's step-0 line
    timing = (out / "timing.txt").read_text().splitlines()
    assert timing[0] == ("Step, Voxelize, Find Neighbors, Compute Density, Compute Pressure, "
                         "Compute Acceleration, Integrate")
    assert len(timing) == 21 and all(len(t.split(",")) == 7 for t in timing[1:])
    am = (out / "angularmomentum.txt").read_text().splitlines()
    assert am[0] == "Step, Angular Momentum" and am[1] == "0, 0"
    nb = (out / "neighbors.txt").read_text().splitlines()
    assert len(nb) == 20
    avg, mx, mn = [int(x) for x in nb[0].split(",")]
    cnt = golden_default["nbr_count_1"].astype(np.int64)
    assert (avg, mx, mn) == (int(cnt.sum()) // cnt.size, int(cnt.max()), min(34, int(cnt.min())))


def test_facade_readback_path():
    """getGrid()[c].count() == CELL_COUNT for every voxel, the position mirror is
    current after step(), step() with the throttled position readback costs < 1.2x a step without
    (1 M particles), and the GUI's gravity row is live (host/facade_check.cpp prints the numbers)."""
    exe = os.path.join(ROOT, "smoothed_particle_hydrodynamics_b200", "facade_check")
    if not os.path.exists(exe):
        pytest.fail("facade_check is not built (python -m smoothed_particle_hydrodynamics_b200.build)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("mirror_positions", "grid_counts", "grid_members", "readback_cost", "gravity_row_live"):
        assert "PASS " + name in r.stdout, r.stdout


def test_bench_line_contract():
    """bench.py prints ONE json line with the keys the driver reads (metric / value / e2e / roofline / clocks /
    gpu_launches / cpu_baseline), on the small dam-break so that the test stays short."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "dambreak_1m", "--steps", "5",
                        "--warmup", "3", "--no-extras", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["metric"] == "particle-updates/sec" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 1e8 and abs(d["value"] - 1048576 / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
    assert d["gpu_launches"] >= 5 * 9
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 28 * 1048576 and e["d2h_bytes_per_step"] == 24 * 1048576
    rf = d["roofline"]
    assert rf["bound"] in ("fp32", "l1") and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1 and rf["peak"] > 1000
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and len(rf["sweeps"]) == 3
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"] and "dambreak_1m" in d["config"]["workload"] and "model" not in d["config"]
