"""Build a variant of the library with extra -D flags on ONE source file (kernel A/B experiments):
    python scripts/build_variant.py tcfast gx_fused.cu -DGX_F2_TILE_FAST=1 -DGX_F2_TC12=2
writes giwaxsim_b200/_variants/libgiwaxs_b200_<name>.so; run with GIWAXS_B200_LIB=<that path>."""
import os, subprocess, sys
sys.path.insert(0, ".")
from giwaxsim_b200 import build as b
name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
out_dir = os.path.join(b.HERE, "_variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, src.replace(".cu", "_%s.o" % name))
subprocess.check_call([b.NVCC] + [f for f in b.FLAGS if f not in ("-Xptxas", "-v")] + flags +
                      ["-c", os.path.join(b.CSRC, src), "-o", obj])
objs = [os.path.join(b.OBJ, b._obj_name(s)) for s in b.SOURCES if s != src] + [obj]
lib = os.path.join(out_dir, "libgiwaxs_b200_%s.so" % name)
subprocess.check_call([b.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                               "-Xcompiler", "-fPIC", "-ldl"])
print(lib)
