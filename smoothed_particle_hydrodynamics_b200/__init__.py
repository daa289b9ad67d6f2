"""B200-native SPH step pipeline (drop-in for the hot path of
DanielaCourel/smoothed_particle_hydrodynamics, reference src/sph.cpp:190-304).

The product is the C-ABI shared library ``libsphb200.so`` (include/sphb200.h:
hand-written CUDA for sm_100a) and the C++ ``SPH`` facade in ``host/``.  This
package is the ctypes binding used by tests, bench.py and the smoke test; it
contains no compute and no CPU fallback -- without the built library or without
a CUDA device every entry point raises.
"""
from .binding import (  # noqa: F401
    FULL, SAMPLED, SCENES, SPH, Field, SlabSPH, SphDerived, SphError, SphParams, SphSceneLattice, default_params, derive,
    lib, lib_path, scene_config, scene_generate, scene_lattice, scene_sphere, slab_layers, step_virtual_slabs,
    voxel_layer,
)
