"""Synthetic scenes for the throughput configs (SURVEY 8(d) "scene rule").

TEST/BENCH INPUT SYNTHESIS in numpy -- an independent restatement of the
product's host-side generator (sphb200_scene_lattice in csrc/sph_scene.cpp);
tests/test_scenes.py checks the two are bit-identical.

Lattice: nx*ny*nz sites of spacing d, first site d/2 from `origin`, particle id
= (z*ny + y)*nx + x, jitter uniform in +-0.1 d per axis from a counter-based
integer hash of (seed, id, axis).  All arithmetic is float32 with one rounding
per operation so that C++ and numpy agree bit for bit.
"""
import math

import numpy as np


def lattice_spacing(h, nu):
    """d = h * (4 pi / (3 nu))^(1/3): nu = target mean neighbour count."""
    return np.float32(float(h) * (4.0 * math.pi / (3.0 * float(nu))) ** (1.0 / 3.0))


def _hash01(idx_u32, seed):
    """murmur3 finaliser of (id*3+axis) xor seed*golden; top 24 bits -> [0,1)."""
    x = idx_u32.astype(np.uint32) ^ np.uint32((int(seed) * 0x9E3779B9) & 0xFFFFFFFF)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x85EBCA6B)).astype(np.uint32)
    x ^= x >> np.uint32(13)
    x = (x * np.uint32(0xC2B2AE35)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def lattice_scene(nx, ny, nz, spacing, origin=(0.0, 0.0, 0.0), seed=42, first_id=0, count=None):
    """Returns pos[n,3] float32 for ids first_id .. first_id+count-1."""
    total = nx * ny * nz
    if count is None:
        count = total - first_id
    d = np.float32(spacing)
    amp = np.float32(0.1) * d
    ids = np.arange(first_id, first_id + count, dtype=np.int64)
    ix = (ids % nx).astype(np.float32)
    iy = ((ids // nx) % ny).astype(np.float32)
    iz = (ids // (nx * ny)).astype(np.float32)
    pos = np.empty((count, 3), np.float32)
    with np.errstate(over="ignore"):
        for axis, site in enumerate((ix, iy, iz)):
            r = _hash01((ids * 3 + axis).astype(np.uint32), seed)
            jit = (np.float32(2.0) * r - np.float32(1.0)) * amp
            base = (site + np.float32(0.5)) * d
            pos[:, axis] = (np.float32(origin[axis]) + base) + jit
    return pos


# name -> (sites, voxel grid, origin in voxels) at nu = 40
CONFIGS = {
    "dambreak_1m": dict(sites=(128, 64, 128), grid=(80, 32, 32), origin_vox=(0, 0, 0)),
    "dambreak_16m": dict(sites=(256, 128, 512), grid=(160, 64, 128), origin_vox=(0, 0, 0)),
    # a quarter of the 16M block in z: the reference arm's sample when 16M would not end in minutes
    "dambreak_4m": dict(sites=(256, 128, 128), grid=(160, 64, 32), origin_vox=(0, 0, 0)),
    # per-GPU slab of the 128M box-drop: lifted 16 voxels, centred in x
    "boxdrop_16m": dict(sites=(256, 128, 512), grid=(160, 64, 128), origin_vox=(49, 16, 3)),
    # small cases for parity tests
    "dambreak_16k": dict(sites=(32, 16, 32), grid=(20, 8, 8), origin_vox=(0, 0, 0)),
    "dambreak_128k": dict(sites=(64, 32, 64), grid=(40, 16, 16), origin_vox=(0, 0, 0)),
}


def scene_params(h=0.1, nu=40.0):
    """Physical parameters of the throughput scenes (stated with every result):
    uniform gravity (0,-9.8,0), no central mass, walls on, rest density of the
    lattice; stiffness / viscosity are the reference defaults."""
    d = lattice_spacing(h, nu)
    rho0 = np.float32(1.0) / (d * d * d)
    return dict(h=np.float32(h), rho0=float(rho0), stiffness=0.001, viscosity=0.01,
                central_mass=0.0, gravity=(0.0, -9.8, 0.0), time_step=0.001)
