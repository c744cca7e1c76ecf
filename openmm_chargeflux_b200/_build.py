"""Build recipes for the native code (run by __graft_entry__.build()).

* libcfx_b200.so       -- the product: hand-written CUDA for sm_100a + the C ABI (nvcc, in-tree).
* libOpenMMCoulB200.so -- the OpenMM plugin adapter (g++), needs the plugin's API headers from
                          /root/reference and the OpenMM stand-in; skipped when those are absent.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcfx_b200.so")
PLUGIN_LIB = os.path.join(PKG, "plugin", "libOpenMMCoulB200.so")
REF = os.environ.get("CFX_REFERENCE_DIR", "/root/reference")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared",
              # keep the statically linked CUDA runtime private: plugin loaders dlopen with RTLD_GLOBAL
              "-Xlinker", "--exclude-libs=ALL",
              "-ldl"]                                   # NCCL is dlopen'ed (comm.cu): no link-time dependency
if os.environ.get("CFX_PTXAS_V"):
    NVCC_FLAGS += ["-Xptxas", "-v"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    """Each .cu is compiled to its own object (in parallel, rebuilt only when it or a header changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    names = ("api.cu", "flux.cu", "kspace.cu", "kspace_tc.cu", "direct.cu", "md.cu", "comm.cu", "platform.cu")
    headers = [os.path.join(CSRC, "cfx_internal.cuh"), os.path.join(CSRC, "ptx_sm100.cuh"), os.path.join(ROOT, "include", "cfx_b200.h")]
    objdir = os.path.join(PKG, "_obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-ldl")]
    compile_flags = [f for i, f in enumerate(compile_flags) if f not in ("-Xlinker", "--exclude-libs=ALL")]
    jobs, objects = [], []
    for n in names:
        src, obj = os.path.join(CSRC, n), os.path.join(objdir, n[:-3] + ".o")
        objects.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([nvcc] + compile_flags + ["-c", "-o", obj, src])
    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([nvcc] + NVCC_FLAGS + ["-o", LIB] + objects)
    return LIB


def build_plugin(force=False, verbose=False):
    """The KernelFactory-registered adapter; needs the plugin's own openmmapi headers."""
    api_inc = os.path.join(REF, "openmmapi", "include")
    # the OpenMM stand-in and the plugin's unchanged API library (built by oracle/Makefile from the reference tree):
    # what a real installation provides as libOpenMM.so / libOpenMMCoul.so
    api_out = os.path.join(ROOT, "shim", "_build")
    if not os.path.isdir(api_inc) or not os.path.exists(os.path.join(api_out, "libOpenMMCoul.so")):
        return PLUGIN_LIB if os.path.exists(PLUGIN_LIB) else None
    src_dir = os.path.join(PKG, "plugin")
    sources = [os.path.join(src_dir, f) for f in ("B200CoulKernels.cpp", "B200CudaCoulKernels.cpp", "B200CoulKernelFactory.cpp",
                                                      "CoulForceProxy.cpp")]
    deps = sources + [os.path.join(src_dir, "B200CoulKernels.h"), os.path.join(src_dir, "B200CudaCoulKernels.h"),
                      os.path.join(src_dir, "B200CoulKernelFactory.h"), os.path.join(src_dir, "CoulForceProxy.h"),
                      os.path.join(ROOT, "include", "cfx_b200.h"), os.path.join(api_out, "libOpenMMShim.so")]
    if not force and not _newer(PLUGIN_LIB, deps):
        return PLUGIN_LIB
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-I" + os.path.join(ROOT, "shim"), "-I" + api_inc,
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda_home, "include"), "-o", PLUGIN_LIB] + sources + [
           "-L" + api_out, "-lOpenMMCoul", "-lOpenMMCudaShim", "-lOpenMMShim", "-L" + PKG, "-lcfx_b200",
           "-Wl,-rpath,$ORIGIN/../../shim/_build", "-Wl,-rpath,$ORIGIN/.."]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return PLUGIN_LIB
