"""Build a library variant with extra -D flags for one source file:  python tools/build_variant.py TAG FILE.cu -DX=1 ...
-> tools/_variants/libcfx_TAG.so (the other objects come from the in-tree build's object cache)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmm_chargeflux_b200 import _build
tag, fname, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
_build.build_cuda()
objdir = os.path.join(_build.PKG, "_obj")
out = os.path.join(_build.ROOT, "tools", "_variants")
os.makedirs(out, exist_ok=True)
obj = os.path.join(out, "%s_%s.o" % (fname[:-3], tag))
cflags = [f for f in _build.NVCC_FLAGS if f not in ("-shared", "-ldl", "-Xlinker", "--exclude-libs=ALL")]
subprocess.run(["nvcc"] + cflags + flags + ["-c", "-o", obj, os.path.join(_build.CSRC, fname)], check=True)
objs = [obj if o == fname[:-3] + ".o" else os.path.join(objdir, o) for o in sorted(os.listdir(objdir)) if o.endswith(".o")]
subprocess.run(["nvcc"] + _build.NVCC_FLAGS + ["-o", os.path.join(out, "libcfx_%s.so" % tag)] + objs, check=True)
print("built", tag)
