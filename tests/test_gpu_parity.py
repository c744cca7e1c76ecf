"""GPU: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Tolerances are those of BASELINE.json's north_star: neighbour and exclusion lists bit-exact, energies
1e-6 relative, forces 1e-5 relative RMS (mixed precision)."""
import numpy as np
import pytest

from conftest import E_RTOL, E_RTOL_DISCARDED, F_RTOL, GOLDEN_NAMES, golden_case, rel_rms
from openmm_chargeflux_b200 import _abi, runtime, synthetic
from openmm_chargeflux_b200.force import CoulForce
from oracle import Oracle

pytestmark = pytest.mark.gpu


def _compare(pos, box, force, check_pairs=True, flags=((True, True), (True, False), (False, True), (False, False))):
    o = Oracle(force, box)
    ctx = runtime.CoulContext(force, box)
    for inc_f, inc_e in flags:
        eo, fo = o.execute(pos, box, inc_f, inc_e)
        e, f, comps = ctx.evaluate(pos, inc_f, inc_e)
        # With includeEnergy=False the reference still returns self + direct + exclusion (a value OpenMM
        # discards); it is reproduced with FP32 pair terms, hence the looser bound in that mode.
        rtol = E_RTOL if inc_e else E_RTOL_DISCARDED
        scale = max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max())
        assert abs(e - eo[4]) <= rtol * abs(eo[4]) or abs(e - eo[4]) <= 1e-9 * np.abs(eo[:4]).max(), (inc_f, inc_e, e, eo)
        assert np.abs(comps[:4] - eo[:4]).max() <= rtol * scale
        if inc_f:
            assert rel_rms(f, fo) <= F_RTOL, (inc_f, inc_e)
        else:
            # without includeForces only the self-term chain rule reaches the forces (reference quirk)
            assert np.abs(f - fo).max() <= 1e-9 * max(1.0, np.abs(fo).max())
    eo, fo = o.execute(pos, box)
    ctx.evaluate(pos)
    assert np.abs(ctx.kernel.charges() - o.charges()).max() <= 1e-14
    assert rel_rms(ctx.kernel.dedq(), o.dedq()) <= F_RTOL
    dq, dx, val = ctx.kernel.jacobian()
    odq, odx, oval = o.jacobian()
    assert np.array_equal(dq, odq) and np.array_equal(dx, odx)
    if len(oval):
        assert np.abs(val - oval).max() <= 1e-12 * max(1.0, np.abs(oval).max())
    if force.usesPeriodicBoundaryConditions() and check_pairs:
        assert np.array_equal(ctx.kernel.neighbor_pairs(), o.neighbor_pairs())
        assert ctx.kernel.stats().pairs_in_cutoff == len(o.neighbor_pairs())
        assert ctx.kernel.ewald_params() == o.ewald_params()
    return ctx, o


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_vectors(name, build_native):
    data, pos, box, force = golden_case(name)
    ctx = runtime.CoulContext(force, box)
    for inc_f in (1, 0):
        for inc_e in (1, 0):
            e, f, _ = ctx.evaluate(pos, bool(inc_f), bool(inc_e))
            e_ref = float(data["energy_f%d_e%d" % (inc_f, inc_e)])
            f_ref = data["forces_f%d_e%d" % (inc_f, inc_e)]
            assert abs(e - e_ref) <= (E_RTOL if inc_e else E_RTOL_DISCARDED) * abs(e_ref)
            if inc_f and np.abs(f_ref).max() > 1e-6:
                assert rel_rms(f, f_ref) <= F_RTOL
            elif inc_f:
                assert np.abs(f - f_ref).max() <= 1e-2          # rock salt: forces vanish by symmetry
    ctx.evaluate(pos)
    assert np.abs(ctx.kernel.charges() - data["charges"]).max() <= 1e-14
    if len(data["jac_val"]):
        assert np.abs(ctx.kernel.jacobian()[2] - data["jac_val"]).max() <= 1e-12 * max(1.0, np.abs(data["jac_val"]).max())
    if force.usesPeriodicBoundaryConditions():
        assert np.array_equal(ctx.kernel.neighbor_pairs(), data["pairs"])          # bit-exact neighbour list


def test_config1_64_waters_nonperiodic(build_native):
    _compare(*synthetic.config("c1"))


def test_config2_4k_water_box(build_native):
    _compare(*synthetic.config("c2"), flags=((True, True), (True, False)))


def test_config5_methanol_water(build_native):
    _compare(*synthetic.config("c5"))


@pytest.mark.parametrize("flux", ["bond+angle", "water", "none"])
def test_small_periodic_boxes(flux, build_native):
    _compare(*synthetic.water_box(125, seed=21, cutoff=0.75, ewald_tol=3e-5, flux=flux))


def test_exclusion_lists_match_reference_sets(build_native):
    pos, box, force = synthetic.methanol_water(10, 30, seed=1, cutoff=0.55, ewald_tol=1e-4)
    force.addException(3, 0)       # duplicate of an existing pair, reversed
    force.addException(7, 7)       # self pair: never evaluated by the reference (p1 < p2 test)
    ctx, o = _compare(pos, box, force)
    ptr, cols = ctx.kernel.exclusions()
    sets = [set() for _ in range(force.getNumParticles())]
    for k in range(force.getNumExceptions()):
        a, b = force.getExceptionParameters(k)
        sets[a].add(b); sets[b].add(a)
    for i, s in enumerate(sets):
        assert list(cols[ptr[i]:ptr[i + 1]]) == sorted(s)


def test_edge_cases(build_native):
    # no flux terms, no exclusions, atoms far outside the box, non-cubic box, tiny system
    rng = np.random.default_rng(5)
    n = 37
    box = np.diag([2.3, 2.9, 2.1])
    pos = rng.uniform(-7, 9, size=(n, 3))
    q = rng.normal(size=n); q -= q.mean()
    f = CoulForce()
    f._bulk(q, rng.uniform(0.2, 0.3, n), rng.uniform(0.0, 0.5, n))
    f.setUsesPeriodicBoundaryConditions(True); f.setCutoffDistance(1.0); f.setEwaldErrorTolerance(1e-5)
    _compare(pos, box, f)
    # two particles
    f2 = CoulForce(); f2.addParticle(1.0, 0.3, 0.2); f2.addParticle(-1.0, 0.3, 0.2)
    f2.setUsesPeriodicBoundaryConditions(True); f2.setCutoffDistance(0.9)
    _compare(np.array([[0.1, 0.2, 0.3], [0.5, 0.6, 0.2]]), np.diag([2.0, 2.0, 2.0]), f2)
    # empty system
    f0 = CoulForce(); f0.setUsesPeriodicBoundaryConditions(True)
    ctx = runtime.CoulContext(f0, np.diag([3.0] * 3))
    assert ctx.evaluate(np.zeros((0, 3)))[0] == 0.0


def test_errors_are_reported_not_swallowed(build_native):
    pos, box, force = synthetic.water_box(27, seed=5, cutoff=0.45)
    ctx = runtime.CoulContext(force, box)
    with pytest.raises(runtime.CfxError, match="rectangular"):
        ctx.kernel.execute(pos, np.array([[1.5, 0, 0], [0.2, 1.5, 0], [0, 0, 1.5]]))
    with pytest.raises(runtime.CfxError, match="twice the cutoff"):
        ctx.kernel.execute(pos, np.diag([0.8, 2.0, 2.0]))
    with pytest.raises(runtime.CfxError):
        ctx.kernel.execute(pos[:-1], box)
    bad = CoulForce(); bad.addParticle(0, 0, 0); bad.addFluxBond(0, 5, 1.0, 0.1)
    with pytest.raises(runtime.CfxError, match="out of range"):
        runtime.CoulContext(bad, box)


def test_forces_accumulate_and_box_can_change(build_native):
    pos, box, force = synthetic.water_box(64, seed=3, cutoff=0.6, ewald_tol=1e-5)
    ctx = runtime.CoulContext(force, box)
    base = np.random.default_rng(0).normal(size=pos.shape)
    forces = base.copy()
    ctx.kernel.execute(pos, box, forces)
    _, f0, _ = ctx.evaluate(pos)
    assert rel_rms(forces - base, f0) < 1e-12
    # a different box at execute time: kmax/alpha stay those of the default box (reference semantics)
    box2 = box * 1.07
    o = Oracle(force, box)
    eo, fo = o.execute(pos * 1.07, box2)
    ctx.box = box2
    e, f, _ = ctx.evaluate(pos * 1.07)
    assert abs(e - eo[4]) <= E_RTOL * abs(eo[4]) and rel_rms(f, fo) <= F_RTOL


def test_kmax_can_follow_the_box(build_native):
    """CFX_OPT_KMAX_FOLLOWS_BOX: after a box change the handle behaves like one created for that box
    (the reference's estimator, ReferenceCoulKernels.cpp:403-420, re-applied); both flag sets, there and back."""
    pos, box, force = synthetic.water_box(216, seed=1, cutoff=0.9, ewald_tol=1e-4)
    ctx = runtime.CoulContext(force, box, kmax_follows_box=True)
    fixed = runtime.CoulContext(force, box)
    k0 = ctx.kernel.ewald_params()[1]
    for scale in (1.0, 1.4, 0.97, 1.0):
        b = box * scale
        o = Oracle(force, b)                                   # kmax from this box
        ctx.box = b
        for inc_e in (True, False, True):
            e, f, _ = ctx.evaluate(pos * scale, True, inc_e)
            eo, fo = o.execute(pos * scale, b, True, inc_e)
            assert ctx.kernel.ewald_params() == o.ewald_params()
            # (the stretched boxes have small totals next to their components: same scale as the other small-box tests)
            assert abs(e - eo[4]) <= (E_RTOL if inc_e else E_RTOL_DISCARDED) * max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max()), (scale, inc_e)
            assert rel_rms(f, fo) <= F_RTOL, (scale, inc_e)
    assert Oracle(force, box * 1.4).ewald_params()[1] != k0     # the sweep did change kmax
    # default: the reference's behaviour, kmax of the default box whatever the box of the call
    fixed.box = box * 1.4
    fixed.evaluate(pos * 1.4)
    assert fixed.kernel.ewald_params()[1] == k0


def test_results_are_bitwise_reproducible(build_native):
    pos, box, force = synthetic.config("c2")
    ctx = runtime.CoulContext(force, box)
    e1, f1, c1 = ctx.evaluate(pos)
    e2, f2, c2 = ctx.evaluate(pos)
    ctx2 = runtime.CoulContext(force, box)
    e3, f3, c3 = ctx2.evaluate(pos)
    assert e1 == e2 == e3 and np.array_equal(f1, f2) and np.array_equal(f1, f3)


def test_finite_difference_of_gpu_energy(build_native):
    # chain rule on the GPU: central differences of the GPU energy vs the GPU forces
    pos, box, force = synthetic.methanol_water(8, 24, seed=5, cutoff=0.5, ewald_tol=1e-6)
    ctx = runtime.CoulContext(force, box)
    _, f, _ = ctx.evaluate(pos)
    h = 2e-4
    for a in (0, 4, 5, 50):
        for c in range(3):
            p = pos.copy(); p[a, c] += h
            ep = ctx.evaluate(p, False, True)[0]
            p[a, c] -= 2 * h
            em = ctx.evaluate(p, False, True)[0]
            assert abs(-(ep - em) / (2 * h) - f[a, c]) <= 2e-3 * np.abs(f).max()


def test_update_parameters_equals_a_fresh_handle(build_native):
    """cfx_update_parameters: new charges / LJ / flux parameters on a live handle give exactly what a
    handle created with them gives, and a topology change is refused."""
    pos, box, f0 = synthetic.water_box(216, seed=3, cutoff=0.9, ewald_tol=1e-5)
    f1 = CoulForce()                                            # same topology, other parameter values
    for i in range(f0.getNumParticles()):
        q, s, e = f0.getParticleParameters(i)
        f1.addParticle(0.9 * q, 1.05 * s, 1.2 * e)
    for x in range(f0.getNumExceptions()):
        f1.addException(*f0.getExceptionParameters(x))
    for t in range(f0.getNumFluxBonds()):
        p1, p2, k, b = f0.getFluxBondParameters(t)
        f1.addFluxBond(p1, p2, 1.3 * k, b + 0.001)
    for t in range(f0.getNumFluxAngles()):
        p1, p2, p3, k, th = f0.getFluxAngleParameters(t)
        f1.addFluxAngle(p1, p2, p3, 0.7 * k, th - 0.01)
    f1.setUsesPeriodicBoundaryConditions(True)
    f1.setCutoffDistance(f0.getCutoffDistance())
    f1.setEwaldErrorTolerance(f0.getEwaldErrorTolerance())
    live = runtime.CoulContext(f0, box)
    e_old, f_old, _ = live.evaluate(pos)
    live.kernel.copyParametersToContext(box, f1)
    e_new, f_new, c_new = live.evaluate(pos)
    fresh = runtime.CoulContext(f1, box)
    e_ref, f_ref, c_ref = fresh.evaluate(pos)
    assert e_new == e_ref and np.array_equal(f_new, f_ref) and np.array_equal(c_new, c_ref)
    assert abs(e_new - e_old) > 1e-3 * abs(e_old)
    o = Oracle(f1, box)
    eo, fo = o.execute(pos, box)
    assert abs(e_new - eo[4]) <= E_RTOL * max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max()) and rel_rms(f_new, fo) <= F_RTOL
    _, _, smaller = synthetic.water_box(125, seed=3, cutoff=0.9, ewald_tol=1e-5)
    with pytest.raises(runtime.CfxError):
        live.kernel.copyParametersToContext(box, smaller)


def test_sharded_handle_needs_a_periodic_system(build_native):
    """The non-periodic all-pairs branch is not partitioned: a sharded handle for it would count every pair on every rank."""
    pos, box, force = synthetic.config("c1")
    k = runtime.CalcCoulForceKernel(shard_rank=0, shard_count=2)
    with pytest.raises(runtime.CfxError, match="periodic"):
        k.initialize(box, force)


def test_failed_create_reports_and_leaves_the_library_usable(build_native):
    pos, box, force = synthetic.water_box(64, seed=5, cutoff=0.6)
    k = runtime.CalcCoulForceKernel(device=99)
    with pytest.raises(runtime.CfxError, match="device"):
        k.initialize(box, force)
    ctx = runtime.CoulContext(force, box)                 # the failed create released everything it had allocated
    assert np.isfinite(ctx.evaluate(pos)[0])


def test_pinned_caller_buffers_give_the_staged_result(build_native):
    """CFX_OPT_PIN_CALLER_BUFFERS: arrays passed repeatedly are page-locked in place (DMA-read positions, forces
    accumulated into the caller's array by the GPU); results are identical to the default staging path."""
    pos, box, force = synthetic.water_box(216, seed=9, cutoff=0.9)
    staged = runtime.CalcCoulForceKernel()
    staged.initialize(box, force)
    pinned = runtime.CalcCoulForceKernel(pin_caller_buffers=True)
    pinned.initialize(box, force)
    f_ref = np.zeros_like(pos)
    e_ref = staged.execute(pos, box, f_ref)
    f_pin = np.zeros_like(pos)
    for call in range(4):                                  # registered from the second sighting on
        f_pin[:] = 1.0                                     # forces are ADDED to what the array holds
        e = pinned.execute(pos, box, f_pin)
        assert e == e_ref and np.array_equal(f_pin - 1.0, (f_ref + 1.0) - 1.0), call
    other = pos + 0.0                                      # a different array: unregisters the first one
    f2 = np.zeros_like(pos)
    assert pinned.execute(other, box, f2) == e_ref and np.array_equal(f2, f_ref)
    pinned.close(); staged.close()


def test_slab_with_vacuum_clusters_stretched_over_empty_cells(build_native):
    """A liquid slab in a box three times as long: the last atoms of a z-ordered column and the first atoms of the next
    one form i-clusters stretched over the empty cells, whose stencil wraps onto itself. The fast pair kernel lists them
    for the generic (min-image) one; neighbour list and forces must not notice."""
    pos, box, force = synthetic.water_box(400, seed=21, cutoff=0.6, ewald_tol=1e-4)
    box = np.diag([box[0, 0], box[1, 1], 3.0 * box[2, 2]])
    o = Oracle(force, box)
    ctx = runtime.CoulContext(force, box)
    assert min(ctx.kernel.stats().cells) >= 7                       # large enough for the fast kernel
    for inc_e in (True, False):
        eo, fo = o.execute(pos, box, True, inc_e)
        e, f, comps = ctx.evaluate(pos, True, inc_e)
        assert abs(e - eo[4]) <= (E_RTOL if inc_e else E_RTOL_DISCARDED) * max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max())
        assert rel_rms(f, fo) <= F_RTOL
    assert np.array_equal(ctx.kernel.neighbor_pairs(), o.neighbor_pairs())
