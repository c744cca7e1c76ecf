"""CPU, build container only: the oracle against the reference's own sources executed live
(oracle/_ref, compiled from /root/reference against the OpenMM stand-in)."""
import numpy as np
import pytest

from conftest import rel_rms
from openmm_chargeflux_b200 import synthetic
from oracle import Oracle, ReferenceBuild, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built (no /root/reference on this box)")


@pytest.mark.parametrize("seed,flux,periodic", [(11, "bond+angle", True), (12, "water", True), (13, "none", True),
                                                 (14, "bond+angle", False), (15, "water", False)])
def test_oracle_is_bit_identical_to_reference_build(seed, flux, periodic, build_native):
    pos, box, force = synthetic.water_box(125, seed=seed, periodic=periodic, cutoff=0.75, ewald_tol=3e-5, flux=flux)
    o, r = Oracle(force, box), ReferenceBuild(force, box)
    for inc_f, inc_e in ((True, True), (True, False), (False, True), (False, False)):
        eo, fo = o.execute(pos, box, inc_f, inc_e)
        er, fr = r.execute(pos, box, inc_f, inc_e)
        assert eo[4] == er[4]
        assert np.array_equal(fo, fr)
    assert np.array_equal(o.charges(), r.charges())
    assert np.array_equal(o.jacobian()[2], r.jacobian()[2])
    if periodic:
        assert o.ewald_params() == r.ewald_params()
        assert np.array_equal(o.neighbor_pairs(), r.neighbor_pairs())


def test_forces_are_added_to_existing_forces(build_native):
    pos, box, force = synthetic.water_box(27, seed=3, cutoff=0.45, ewald_tol=1e-4)
    o, r = Oracle(force, box), ReferenceBuild(force, box)
    base = np.random.default_rng(0).normal(size=pos.shape)
    _, fo = o.execute(pos, box, forces_in=base)
    _, fr = r.execute(pos, box, forces_in=base)
    _, f0 = o.execute(pos, box)
    assert np.array_equal(fo, fr)
    assert rel_rms(fo - base, f0) < 1e-12


def test_methanol_water_matches(build_native):
    pos, box, force = synthetic.methanol_water(20, 60, seed=8, cutoff=0.7, ewald_tol=1e-5)
    o, r = Oracle(force, box), ReferenceBuild(force, box)
    eo, fo = o.execute(pos, box)
    er, fr = r.execute(pos, box)
    assert eo[4] == er[4] and np.array_equal(fo, fr)
