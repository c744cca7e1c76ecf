"""Serialization of a CoulForce (the reference has no proxy, so nothing of its own to compare with).

* the Python mirror round-trips its own XML exactly and the restatement evaluates the copy bit-identically;
* plugin/CoulForceProxy.cpp, reached through OpenMM's XmlSerializer entry points (stand-in in shim/), writes the same bytes
  as the mirror, reads them back, and the reference's own Reference-platform kernel evaluates the deserialized force
  bit-identically to the original (element order = add* order = Jacobian row order)."""
import os

import numpy as np
import pytest

from openmm_chargeflux_b200 import synthetic
from openmm_chargeflux_b200.force import CoulForce
from oracle import Oracle
from oracle.binding import ReferenceBuild, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "openmm_chargeflux_b200", "plugin", "libOpenMMCoulB200.so")


def cases():
    yield "water", synthetic.water_box(27, seed=3, cutoff=0.45, ewald_tol=1e-4)
    yield "methanol", synthetic.methanol_water(6, 20, seed=5, cutoff=0.45, ewald_tol=1e-5)
    yield "nopbc", synthetic.config("c1")
    pos, box, f = synthetic.water_box(27, seed=4, cutoff=0.45, ewald_tol=1e-4)
    g = CoulForce()                                   # flux-water terms, a force group, awkward doubles
    for i in range(f.getNumParticles()):
        q, s, e = f.getParticleParameters(i)
        g.addParticle(q * (1 + 1e-16 * i) + 1e-300, s / 3.0, e * np.pi)
    for i in range(f.getNumExceptions()):
        g.addException(*f.getExceptionParameters(i))
    for w in range(f.getNumParticles() // 3):
        g.addFluxWater(3 * w, 3 * w + 1, 3 * w + 2, 0.1 / 3, -0.2 / 7, 0.05, 0.09572, 0.15139)
    g.setCutoffDistance(0.45); g.setEwaldErrorTolerance(1e-4); g.setUsesPeriodicBoundaryConditions(True); g.setForceGroup(7)
    yield "fluxwater", (pos, box, g)


def same_force(a, b):
    return (a._charges == b._charges and a._ljparams == b._ljparams and a._exclusions == b._exclusions
            and a._fbond_idx == b._fbond_idx and a._fbond_params == b._fbond_params
            and a._fangle_idx == b._fangle_idx and a._fangle_params == b._fangle_params
            and a._fwater_idx == b._fwater_idx and a._fwater_params == b._fwater_params
            and a._cutoff == b._cutoff and a._ewald_tol == b._ewald_tol and a._pbc == b._pbc and a._force_group == b._force_group)


@pytest.mark.parametrize("name,case", list(cases()), ids=[c[0] for c in cases()])
def test_python_mirror_round_trip_is_exact(name, case):
    pos, box, force = case
    text = force.to_xml()
    copy = CoulForce.from_xml(text)
    assert same_force(force, copy)                    # doubles survive %.17g
    assert copy.to_xml() == text
    e0, f0 = Oracle(force, box).execute(pos, box)
    e1, f1 = Oracle(copy, box).execute(pos, box)
    assert np.array_equal(e0, e1) and np.array_equal(f0, f1)


def test_from_xml_rejects_other_types_and_versions():
    text = CoulForce().to_xml()
    with pytest.raises(ValueError):
        CoulForce.from_xml(text.replace('type="CoulForce"', 'type="NonbondedForce"'))
    with pytest.raises(ValueError):
        CoulForce.from_xml(text.replace('version="1"', 'version="2"'))


needs_plugin = pytest.mark.skipif(not (reference_available() and os.path.exists(PLUGIN)),
                                  reason="needs oracle/_ref and the plugin adapter (built where /root/reference exists)")


@needs_plugin
@pytest.mark.parametrize("name,case", list(cases()), ids=[c[0] for c in cases()])
def test_proxy_writes_the_mirror_bytes_and_restores_a_bit_identical_force(name, case):
    pos, box, force = case
    # the proxy is registered when the plugin library is loaded (what OpenMM's plugin loader does)
    original = ReferenceBuild(force, box, platform="Reference")
    original._check(original.lib.cfxref_load_plugin(PLUGIN.encode()))
    text = original.to_xml()
    # (the harness builds its CoulForce from the C parameter block, which carries no force group)
    assert text == force.to_xml().replace('forceGroup="%d"' % force.getForceGroup(), 'forceGroup="0"')
    restored = ReferenceBuild.from_xml(force.to_xml(), box, platform="Reference")       # the mirror's bytes, group included
    assert restored.n == force.getNumParticles()
    for flags in ((True, True), (True, False), (False, True)):
        e0, f0 = original.execute(pos, box, *flags)
        e1, f1 = restored.execute(pos, box, *flags)
        assert e0[4] == e1[4] and np.array_equal(f0, f1)
    if force.getNumFluxBonds() + force.getNumFluxAngles() + force.getNumFluxWaters() > 0:
        for a, b in zip(original.jacobian(), restored.jacobian()):
            assert np.array_equal(a, b)
    # what the proxy writes for the force it restored is again the mirror's text
    assert restored.to_xml() == force.to_xml()
    assert same_force(CoulForce.from_xml(restored.to_xml()), force)


@needs_plugin
def test_proxy_rejects_unknown_versions_and_types():
    pos, box, force = synthetic.config("c1")
    ctx = ReferenceBuild(force, box, platform="Reference")
    ctx._check(ctx.lib.cfxref_load_plugin(PLUGIN.encode()))
    with pytest.raises(RuntimeError, match="Unsupported version"):
        ReferenceBuild.from_xml(force.to_xml().replace('version="1"', 'version="3"'), box)
    with pytest.raises(RuntimeError, match="no serialization proxy"):
        ReferenceBuild.from_xml(force.to_xml().replace('type="CoulForce"', 'type="Other"'), box)
