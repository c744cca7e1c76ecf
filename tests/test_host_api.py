"""CPU: the host-side mirror of the plugin API and the C-ABI library's surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from openmm_chargeflux_b200 import CoulForce, _abi, synthetic, runtime


def test_coulforce_defaults_and_accessors():
    f = CoulForce()
    # CoulForce.cpp:12-16
    assert f.getCutoffDistance() == 1.0 and f.getEwaldErrorTolerance() == 0.0001 and not f.usesPeriodicBoundaryConditions()
    f.addParticle(-0.8, 0.3, 0.6); f.addParticle(0.4, 0.1, 0.0); f.addParticle(0.4, 0.1, 0.0)
    f.setParticleParameters(1, 0.41, 0.11, 0.01)
    assert f.getNumParticles() == 3 and f.getParticleParameters(1) == (0.41, 0.11, 0.01)
    f.addException(0, 1); f.addException(0, 2)
    assert f.getNumExceptions() == 2 and f.getExceptionParameters(1) == (0, 2)
    f.addFluxBond(0, 1, 2.0, 0.1); f.addFluxAngle(1, 0, 2, 0.08, 1.8); f.addFluxWater(0, 1, 2, 1, 2, 3, 4, 5)
    assert f.getNumFluxBonds() == 1 and f.getFluxBondParameters(0) == (0, 1, 2.0, 0.1)
    assert f.getNumFluxAngles() == 1 and f.getFluxAngleParameters(0) == (1, 0, 2, 0.08, 1.8)
    assert f.getNumFluxWaters() == 1 and f.getFluxWaterParameters(0) == (0, 1, 2, 1, 2, 3, 4, 5)
    f.setCutoffDistance(0.9); f.setEwaldErrorTolerance(1e-5); f.setUsesPeriodicBoundaryConditions(True)
    d, keep = f.to_desc(np.diag([3.0, 3.0, 3.0]))
    assert d.num_particles == 3 and d.num_flux_waters == 1 and d.use_pbc == 1 and d.cutoff == 0.9
    assert list(d.default_box) == [3, 0, 0, 0, 3, 0, 0, 0, 3]
    assert d.flux_water_params[4] == 5.0 and d.exception_pairs[3] == 2


def test_synthetic_configs_have_the_surveyed_sizes():
    pos, box, f = synthetic.config("c1")
    assert len(pos) == 192 and f.getNumFluxBonds() == 128 and f.getNumFluxAngles() == 64 and f.getNumExceptions() == 192
    assert not f.usesPeriodicBoundaryConditions()
    pos, box, f = synthetic.config("c2")
    assert len(pos) == 4095 and abs(box[0, 0] - 3.4435) < 1e-3 and f.getEwaldErrorTolerance() == 1e-4
    pos, box, f = synthetic.config("c5")
    assert len(pos) == 1500 and f.getNumExceptions() == 100 * 12 + 300 * 3
    assert f.getNumFluxBonds() == 500 and f.getNumFluxAngles() == 700 and f.getNumFluxWaters() == 300


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cfx_b200.h")).read()
    declared = set(re.findall(r"\b(cfx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTED_SYMBOLS)
    lib = runtime.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    # compile-free check: sizes the C compiler would produce on x86-64
    assert ctypes.sizeof(_abi.SystemDesc) == 216
    assert ctypes.sizeof(_abi.Options) == 32
    assert ctypes.sizeof(_abi.EwaldParams) == 32
    assert ctypes.sizeof(_abi.Stats) == 48


def test_no_cpu_fallback():
    lib = runtime.load_library()
    if lib.cfx_device_count() > 0:
        pytest.skip("a GPU is present")
    pos, box, f = synthetic.water_box(8, 1)
    with pytest.raises(runtime.CfxError, match="no CPU fallback"):
        runtime.CoulContext(f, box)
