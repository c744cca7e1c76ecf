"""CPU: the oracle restatement against the golden vectors produced by the reference's own sources."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, golden_case, rel_rms
from oracle import Oracle


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_generator_is_deterministic(name, build_native):
    data, pos, box, force = golden_case(name)
    assert np.array_equal(pos, data["positions"])
    assert np.array_equal(box, data["box"])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name, build_native):
    data, pos, box, force = golden_case(name)
    o = Oracle(force, box)
    for inc_f in (1, 0):
        for inc_e in (1, 0):
            e, f = o.execute(pos, box, bool(inc_f), bool(inc_e))
            e_ref = float(data["energy_f%d_e%d" % (inc_f, inc_e)])
            f_ref = data["forces_f%d_e%d" % (inc_f, inc_e)]
            assert abs(e[4] - e_ref) <= 1e-12 * max(1.0, abs(e_ref)), (name, inc_f, inc_e)
            assert rel_rms(f, f_ref) <= 1e-12, (name, inc_f, inc_e)
    o.execute(pos, box, True, True)
    assert np.abs(o.charges() - data["charges"]).max() <= 1e-15
    dq, dx, val = o.jacobian()
    assert np.array_equal(dq, data["jac_dq"]) and np.array_equal(dx, data["jac_dx"])
    if len(val):
        assert np.abs(val - data["jac_val"]).max() <= 1e-12
    if force.usesPeriodicBoundaryConditions():
        alpha, kmax, nk = o.ewald_params()
        assert alpha == float(data["alpha"]) and tuple(kmax) == tuple(data["kmax"]) and nk == int(data["num_kvectors"])
        assert np.array_equal(o.neighbor_pairs(), data["pairs"])      # neighbour list: bit-exact
