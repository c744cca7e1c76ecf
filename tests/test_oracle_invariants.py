"""CPU: known answers and invariants that pin the oracle independently of the reference
(the reference ships no tests)."""
import numpy as np
import pytest

from openmm_chargeflux_b200 import synthetic, _abi
from openmm_chargeflux_b200.force import CoulForce
from oracle import Oracle


def _fd_forces(o, pos, box, atoms, h=1e-5):
    out = np.zeros((len(atoms), 3))
    for k, a in enumerate(atoms):
        for c in range(3):
            p = pos.copy(); p[a, c] += h
            ep = o.execute(p, box, False, True)[0][4]
            p[a, c] -= 2 * h
            em = o.execute(p, box, False, True)[0][4]
            out[k, c] = -(ep - em) / (2 * h)
    return out


@pytest.mark.parametrize("flux", ["bond+angle", "water"])
def test_finite_difference_of_chain_rule_periodic(flux, build_native):
    pos, box, force = synthetic.water_box(27, seed=5, cutoff=0.45, ewald_tol=1e-6, flux=flux)
    o = Oracle(force, box)
    _, f = o.execute(pos, box)
    atoms = [0, 1, 2, 40, 41, 80]
    fd = _fd_forces(o, pos, box, atoms)
    assert np.abs(fd - f[atoms]).max() <= 2e-5 * np.abs(f).max()


def test_finite_difference_methanol_water(build_native):
    pos, box, force = synthetic.methanol_water(8, 24, seed=5, cutoff=0.5, ewald_tol=1e-6)
    o = Oracle(force, box)
    _, f = o.execute(pos, box)
    atoms = [0, 4, 5, 9, 48, 49, 50]
    fd = _fd_forces(o, pos, box, atoms)
    assert np.abs(fd - f[atoms]).max() <= 2e-5 * np.abs(f).max()


def test_finite_difference_nonperiodic(build_native):
    pos, box, force = synthetic.water_box(20, seed=9, periodic=False)
    o = Oracle(force, box)
    _, f = o.execute(pos, box)
    atoms = [0, 1, 2, 30]
    fd = _fd_forces(o, pos, box, atoms)
    assert np.abs(fd - f[atoms]).max() <= 2e-6 * np.abs(f).max()


def test_charge_conservation_and_jacobian_column_sums(build_native):
    pos, box, force = synthetic.methanol_water(10, 30, seed=1, cutoff=0.55, ewald_tol=1e-4)
    o = Oracle(force, box)
    _, f = o.execute(pos, box)
    q = o.charges()
    q0 = np.array([force.getParticleParameters(i)[0] for i in range(force.getNumParticles())])
    assert abs(q.sum() - q0.sum()) < 1e-12
    dq, dx, val = o.jacobian()
    # translation invariance of q(x): for every charge, the rows over all dx sum to zero
    colsum = np.zeros((force.getNumParticles(), 3))
    np.add.at(colsum, dq, val)
    assert np.abs(colsum).max() < 1e-10
    assert np.abs(f.sum(axis=0)).max() < 1e-7 * np.abs(f).max() * len(f)


def test_madelung_constant_of_rock_salt(build_native):
    a = 0.564
    pos, box, force = synthetic.rock_salt(cells=3, a=a, ewald_tol=1e-7)
    o = Oracle(force, box)
    e, f = o.execute(pos, box)
    n_pairs = len(pos) / 2
    madelung = -e[4] / n_pairs * (a / 2) / _abi.ONE_4PI_EPS0
    assert abs(madelung - 1.747564594633) < 2e-6
    assert np.abs(f).max() < 1e-8


def test_ewald_sum_is_independent_of_splitting(build_native):
    pos, box, force = synthetic.water_box(64, seed=3, cutoff=0.6, ewald_tol=1e-6)
    # the cutoff stays fixed (the LJ part is truncated there); the tolerance moves alpha and kmax
    e1 = Oracle(force, box).execute(pos, box)[0]
    force.setEwaldErrorTolerance(1e-7)
    e2 = Oracle(force, box).execute(pos, box)[0]
    assert abs(e1[0] - e2[0]) > 1e3                       # the split really changed
    assert abs(e1[4] - e2[4]) < 1e-5 * abs(e1[4])


def test_periodic_limit_reproduces_nonperiodic_branch(build_native):
    # neutral cluster in a huge box, no LJ: Ewald energy -> plain Coulomb as L grows
    # (the leading image correction is dipole-dipole, ~ 2 pi p^2 / (3 V))
    rng = np.random.default_rng(4)
    pos = rng.uniform(0, 0.4, size=(6, 3)) + 4.0
    q = rng.normal(size=6); q -= q.mean()
    f_np, f_p = CoulForce(), CoulForce()
    for f in (f_np, f_p):
        f._bulk(q, np.zeros(6), np.zeros(6))
    f_p.setUsesPeriodicBoundaryConditions(True)
    f_p.setCutoffDistance(2.0)
    f_p.setEwaldErrorTolerance(1e-6)
    box = np.diag([8.0] * 3)
    e_np, g_np = Oracle(f_np, box).execute(pos, box)
    e_p, g_p = Oracle(f_p, box).execute(pos, box)
    dip = (q[:, None] * pos).sum(0)
    corr = 2 * np.pi / (3 * 8.0 ** 3) * (dip @ dip) * _abi.ONE_4PI_EPS0
    assert abs(e_p[4] - corr - e_np[4]) < 2e-3 * abs(e_np[4]) + 1e-3


def test_kmax_and_alpha_follow_the_reference_rule(build_native):
    # values quoted in SURVEY.md section 8 for the benchmark boxes
    for n_w, tol, kmax, nk in ((1365, 1e-4, 11, 4630), (10922, 1e-5, 27, 74438)):
        box_len = (n_w / synthetic.WATER_DENSITY) ** (1 / 3)
        f = CoulForce(); f.addParticle(0, 0, 0)
        f.setUsesPeriodicBoundaryConditions(True); f.setEwaldErrorTolerance(tol)
        alpha, k, n = Oracle(f, np.diag([box_len] * 3)).ewald_params()
        assert k == (kmax,) * 3 and n == nk
        assert abs(alpha - np.sqrt(-np.log(2 * tol))) < 1e-12


def test_parallel_k_slabs_reproduce_the_single_threaded_oracle(build_native):
    """oracle/slabs.py (used for the full-size parity pins): slab sum == one-thread evaluation to rounding."""
    from oracle.slabs import execute_parallel
    pos, box, force = synthetic.water_box(216, seed=3, cutoff=0.9, ewald_tol=1e-5)
    o = Oracle(force, box)
    e, f = o.execute(pos, box, True, True)
    e2, f2, d2, kmax = execute_parallel(force, box, pos, workers=3)
    assert kmax == o.ewald_params()[1]
    assert np.abs(e - e2).max() <= 1e-12 * np.abs(e[:4]).max()
    assert np.sqrt(((f - f2) ** 2).sum() / (f ** 2).sum()) <= 1e-13
    assert np.sqrt(((o.dedq() - d2) ** 2).sum() / (d2 ** 2).sum()) <= 1e-13
