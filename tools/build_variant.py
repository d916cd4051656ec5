"""Developer tool: build compas_tno_b200/variants/libtno_b200_<name>.so with extra nvcc flags on tno_onchip.cu (the other
objects come from the regular build), for A/B timing on the GPU box with TNO_LIB_VARIANT=<name>.
usage: python tools/build_variant.py name [-DTNO_PF=2 ...]"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from compas_tno_b200 import _build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
_build.build()
objdir = os.path.join(ROOT, "build", "obj")
outdir = os.path.join(ROOT, "compas_tno_b200", "variants")
os.makedirs(outdir, exist_ok=True)
obj = os.path.join(objdir, "tno_onchip_%s.o" % name)
flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-O3"]
subprocess.check_call([_build._nvcc()] + flags + extra + ["-c", os.path.join(_build.CSRC, "tno_onchip.cu"), "-o", obj])
others = [os.path.join(objdir, os.path.splitext(s)[0] + ".o") for s in _build.SOURCES if s != "tno_onchip.cu"]
out = os.path.join(outdir, "libtno_b200_%s.so" % name)
subprocess.check_call([_build._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, obj] + others)
print(out)
