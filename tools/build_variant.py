"""Builds an A/B variant of libsphb200.so with extra -D flags on sph_full.cu:

    python tools/build_variant.py NAME -DSPH_TBY=4 -DSPH_TILE_THREADS=256 ...

-> smoothed_particle_hydrodynamics_b200/variants/libsphb200_NAME.so; select it with
SPHB200_LIB=<path> (binding.lib_path).  The other objects come from the normal build."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smoothed_particle_hydrodynamics_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
B.build()
out_dir = os.path.join(B.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "sph_full_%s.o" % name)
subprocess.check_call([B.NVCC] + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, "sph_full.cu"), "-o", obj])
objs = [os.path.join(B.OBJ, os.path.splitext(s)[0] + ".o") for s in B.CU_SOURCES + B.CPP_SOURCES if s != "sph_full.cu"]
lib = os.path.join(out_dir, "libsphb200_%s.so" % name)
subprocess.check_call([B.NVCC, "-shared", "-o", lib, obj] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                     "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"])
print(lib)
