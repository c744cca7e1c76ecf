"""GPU: the zero-copy adapter for OpenMM's CUDA platform.

`B200CudaCalcCoulForceKernel` (openmm_chargeflux_b200/plugin/B200CudaCoulKernels.cpp) replaces the reference's
`CudaCalcCoulForceKernel` (platforms/cuda/src/CudaCoulKernels.cpp): it is registered on the platform named "CUDA"
through the plugin loader contract, finds the platform's CudaContext as the reference's factory does, and works on the
platform's own device buffers. OpenMM is absent, so the platform is the stand-in of shim/openmm/cuda/ driven by
shim/cuda_harness.cpp: real4 posq in a SHUFFLED atom order + atomIndex, padded arrays, the 64-bit fixed-point force
buffer, the energy buffer, the platform's stream -- with the plugin's unmodified CoulForce / CoulForceImpl in between.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import E_RTOL, F_RTOL, ROOT, rel_rms
from openmm_chargeflux_b200 import _abi, synthetic
from oracle import Oracle

HARNESS = os.path.join(ROOT, "shim", "_build", "libcfx_cudaharness.so")
PLUGIN = os.path.join(ROOT, "openmm_chargeflux_b200", "plugin", "libOpenMMCoulB200.so")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(HARNESS) and os.path.exists(PLUGIN)), reason="CUDA-platform harness / plugin adapter not built")]

_lib = None


def harness():
    global _lib
    if _lib is None:
        _lib = C.CDLL(HARNESS)
        _lib.cfxcu_last_error.restype = C.c_char_p
        _lib.cfxcu_destroy.argtypes = [C.c_void_p]
        _lib.cfxcu_execute.argtypes = [C.c_void_p, _abi.c_double_p, _abi.c_double_p, C.c_int, C.c_int, _abi.c_double_p, _abi.c_double_p]
        _lib.cfxcu_force_info_groups.argtypes = [C.c_void_p]
        _lib.cfxcu_force_info_group.argtypes = [C.c_void_p, C.c_int, _abi.c_int32_p, C.c_int]
        _lib.cfxcu_force_info_particles_identical.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.cfxcu_force_info_groups_identical.argtypes = [C.c_void_p, C.c_int, C.c_int]
        assert _lib.cfxcu_load_plugin(PLUGIN.encode()) == 0, _lib.cfxcu_last_error().decode()
    return _lib


class CudaPlatformContext:
    PRECISION = {"single": 0, "mixed": 1, "double": 2}

    def __init__(self, force, box, precision, seed=11):
        self.lib = harness()
        desc, self._keep = force.to_desc(np.asarray(box, dtype=np.float64))
        self.h = C.c_void_p()
        rc = self.lib.cfxcu_create(C.byref(desc), self.PRECISION[precision], 0, seed, C.byref(self.h))
        assert rc == 0, self.lib.cfxcu_last_error().decode()
        self.n = force.getNumParticles()

    def evaluate(self, pos, box, inc_f=True, inc_e=True):
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        e = np.zeros(_abi.E_COUNT)
        f = np.zeros(3 * self.n)
        rc = self.lib.cfxcu_execute(self.h, p.ctypes.data_as(_abi.c_double_p), b.ctypes.data_as(_abi.c_double_p), int(inc_f), int(inc_e),
                                    e.ctypes.data_as(_abi.c_double_p), f.ctypes.data_as(_abi.c_double_p))
        assert rc == 0, self.lib.cfxcu_last_error().decode()
        return e[_abi.E_TOTAL], f.reshape(-1, 3)

    def close(self):
        if self.h:
            self.lib.cfxcu_destroy(self.h)
            self.h = None


CASES = {
    "water216_pbc": lambda: synthetic.water_box(216, seed=1, cutoff=0.9),
    "c1_nonperiodic": lambda: synthetic.config("c1"),
    "c2_4k_water": lambda: synthetic.config("c2"),
    "c5_methanol_water": lambda: synthetic.config("c5"),
}


@pytest.mark.parametrize("precision", ["double", "mixed", "single"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_platform_adapter_matches_the_reference_platform(case, precision, build_native):
    pos, box, force = CASES[case]()
    eo, fo = Oracle(force, box).execute(pos, box, True, True)
    ctx = CudaPlatformContext(force, box, precision)
    for repeat in range(2):                                # the second call replays the cached CUDA graph
        e, f = ctx.evaluate(pos, box, True, True)
        # a single-precision platform hands over float positions (2e-7 nm rounding) and sums energies in a float buffer
        e_tol, f_tol = (E_RTOL, F_RTOL) if precision != "single" else (2e-5, 1e-4)
        assert abs(e - eo[4]) <= e_tol * max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max()), (repeat, e, eo[4])
        assert rel_rms(f, fo) <= f_tol, repeat
    # forces-only call: the adapter opts out of the partial energy OpenMM discards, forces are unchanged
    e2, f2 = ctx.evaluate(pos, box, True, False)
    assert rel_rms(f2, fo) <= (F_RTOL if precision != "single" else 1e-4)
    ctx.close()


def test_cuda_platform_atom_order_does_not_matter(build_native):
    pos, box, force = CASES["water216_pbc"]()
    results = []
    for seed in (1, 2):
        ctx = CudaPlatformContext(force, box, "double", seed=seed)
        results.append(ctx.evaluate(pos, box))
        ctx.close()
    assert results[0][0] == results[1][0] and np.array_equal(results[0][1], results[1][1])     # fixed-point sums: bit-identical


def test_force_info_covers_exclusions_and_flux_terms(build_native):
    """CudaForceInfo (CudaCoulKernels.cpp:20-47 lists exclusions only): here the flux bonds / angles / waters are
    particle groups too, and groups compare by kind and parameters."""
    pos, box, force = synthetic.config("c5")               # methanol/water: mixed bond / angle parameters
    ctx = CudaPlatformContext(force, box, "double")
    lib = ctx.lib
    nx, nb, na, nw = force.getNumExceptions(), force.getNumFluxBonds(), force.getNumFluxAngles(), force.getNumFluxWaters()
    assert lib.cfxcu_force_info_groups(ctx.h) == nx + nb + na + nw
    buf = np.zeros(4, np.int32)
    cnt = lib.cfxcu_force_info_group(ctx.h, 0, buf.ctypes.data_as(_abi.c_int32_p), 4)
    assert cnt == 2 and tuple(buf[:2]) == tuple(force.getExceptionParameters(0))
    cnt = lib.cfxcu_force_info_group(ctx.h, nx, buf.ctypes.data_as(_abi.c_int32_p), 4)
    assert cnt == 2 and tuple(buf[:2]) == tuple(force.getFluxBondParameters(0)[:2])
    cnt = lib.cfxcu_force_info_group(ctx.h, nx + nb, buf.ctypes.data_as(_abi.c_int32_p), 4)
    assert cnt == 3 and tuple(buf[:3]) == tuple(force.getFluxAngleParameters(0)[:3])
    assert lib.cfxcu_force_info_groups_identical(ctx.h, 0, 1) == 1                      # two exclusions
    assert lib.cfxcu_force_info_groups_identical(ctx.h, 0, nx) == 0                     # exclusion vs flux bond
    params = [force.getFluxBondParameters(k)[2:] for k in range(nb)]
    same = next(k for k in range(1, nb) if params[k] == params[0])
    diff = next(k for k in range(1, nb) if params[k] != params[0])
    assert lib.cfxcu_force_info_groups_identical(ctx.h, nx, nx + same) == 1
    assert lib.cfxcu_force_info_groups_identical(ctx.h, nx, nx + diff) == 0
    q = [force.getParticleParameters(k) for k in range(force.getNumParticles())]
    twin = next(k for k in range(1, len(q)) if q[k] == q[0])
    other = next(k for k in range(1, len(q)) if q[k] != q[0])
    assert lib.cfxcu_force_info_particles_identical(ctx.h, 0, twin) == 1
    assert lib.cfxcu_force_info_particles_identical(ctx.h, 0, other) == 0
    ctx.close()
