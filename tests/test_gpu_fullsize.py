"""GPU: size-independent properties at BASELINE.json's full 32k-atom configuration, plus the parts
of the oracle that finish in seconds at that size (everything except the explicit k-sum)."""
import os

import numpy as np
import pytest

from conftest import E_RTOL, F_RTOL, GOLDEN_DIR, rel_rms
from openmm_chargeflux_b200 import _abi, runtime, synthetic
from oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3(build_native):
    pos, box, force = synthetic.config("c3")
    ctx = runtime.CoulContext(force, box)
    e, f, comps = ctx.evaluate(pos)
    return pos, box, force, ctx, e, f, comps


def test_c3_sizes(c3):
    pos, box, force, ctx, e, f, comps = c3
    alpha, kmax, nk = ctx.kernel.ewald_params()
    assert len(pos) == 32766 and kmax == (27, 27, 27) and nk == 74438


def test_c3_direct_self_exclusion_and_neighbour_list_against_oracle(c3):
    pos, box, force, ctx, e, f, comps = c3
    o = Oracle(force, box)
    o.set_kx_range(0, 0)                       # skip the explicit k-sum (minutes on the CPU at this size)
    eo, fo = o.execute(pos, box)
    for k in (_abi.E_SELF, _abi.E_DIRECT, _abi.E_EXCL):
        assert abs(comps[k] - eo[k]) <= E_RTOL * abs(e), k
    assert ctx.kernel.stats().pairs_in_cutoff == len(o.neighbor_pairs())
    assert np.array_equal(ctx.kernel.neighbor_pairs(), o.neighbor_pairs())
    assert np.abs(ctx.kernel.charges() - o.charges()).max() <= 1e-14


def test_c3_reciprocal_space_against_oracle_slab_by_linearity(c3):
    """E_recip and the reciprocal forces are sums over k: a second handle whose default box is three
    times smaller gets kmax=(9,9,9) from the reference rule, i.e. a sub-block of the same k lattice,
    which the CPU oracle can evaluate in seconds on the same positions and box."""
    pos, box, force, ctx, e, f, comps = c3
    small_default = box / 3.0
    o = Oracle(force, small_default)
    assert o.ewald_params()[1] == (11, 11, 11) or o.ewald_params()[1][0] < 27
    eo, fo = o.execute(pos, box)
    k = runtime.CalcCoulForceKernel()
    k.initialize(small_default, force)
    forces = np.zeros_like(pos)
    c = np.zeros(_abi.E_COUNT)
    k.execute(pos, box, forces, components=c)
    assert abs(c[_abi.E_RECIP] - eo[_abi.E_RECIP]) <= E_RTOL * abs(eo[4])
    assert abs(c[_abi.E_TOTAL] - eo[4]) <= E_RTOL * abs(eo[4])
    assert rel_rms(forces, fo) <= F_RTOL


def test_c3_momentum_and_charge_conservation(c3):
    pos, box, force, ctx, e, f, comps = c3
    assert np.abs(f.sum(axis=0)).max() <= 1e-6 * np.abs(f).max() * np.sqrt(len(f))
    q = ctx.kernel.charges()
    assert abs(q.sum()) < 1e-9
    assert abs(comps[:4].sum() - e) < 1e-9 * abs(e)


def test_c3_periodic_translation_invariance(c3):
    pos, box, force, ctx, e, f, comps = c3
    shift = np.random.default_rng(1).integers(-3, 4, size=(len(pos) // 3, 1, 3)) * np.diag(box)[None, None, :]
    pos2 = (pos.reshape(-1, 3, 3) + shift).reshape(-1, 3)      # whole molecules moved by lattice vectors
    e2, f2, _ = ctx.evaluate(pos2)
    assert abs(e2 - e) <= E_RTOL * abs(e)
    assert rel_rms(f2, f) <= F_RTOL


def test_c3_reproducible(c3):
    pos, box, force, ctx, e, f, comps = c3
    e2, f2, _ = ctx.evaluate(pos)
    assert e2 == e and np.array_equal(f2, f)


# ---------------------------------------------------------------------------------------------------------
# Full k lattice at the benchmark sizes. tests/golden/{c3,c4}_fullk.npz hold the oracle's energies and forces
# with its explicit k-sum spread over the host cores (tests/golden/make_golden_fullsize.py, oracle/slabs.py).
# ---------------------------------------------------------------------------------------------------------
def _fullk(name):
    path = os.path.join(GOLDEN_DIR, name + "_fullk.npz")
    if not os.path.exists(path):
        pytest.skip(name + "_fullk.npz not generated")
    data = np.load(path)
    pos, box, force = synthetic.config(name)
    assert abs(float((pos * np.arange(1, 4)[None, :]).sum()) - float(data["pos_checksum"])) <= 1e-9 * abs(float(data["pos_checksum"]))
    return data, pos, box, force


def _check_fullk(ctx, data, pos, include_energy):
    e, f, comps = ctx.evaluate(pos, True, include_energy)
    eg = data["energy"]
    if include_energy:
        assert abs(e - eg[4]) <= E_RTOL * abs(eg[4]), (e, eg)
        assert np.abs(comps[:4] - eg[:4]).max() <= E_RTOL * abs(eg[4])
    else:
        # the reference returns self + direct + exclusion when includeEnergy is false (no reciprocal term)
        partial = eg[0] + eg[2] + eg[3]
        assert abs(e - partial) <= 5e-5 * abs(partial)
    assert rel_rms(f, data["forces_f32"].astype(np.float64)) <= F_RTOL
    probe = data["probe"]
    assert rel_rms(f[probe], data["forces_probe"]) <= F_RTOL
    assert rel_rms(ctx.kernel.dedq()[probe], data["dedq_probe"]) <= F_RTOL
    assert ctx.kernel.stats().pairs_in_cutoff == int(data["pairs_in_cutoff"])


@pytest.mark.parametrize("include_energy", [True, False])
def test_c3_full_kmax27_against_the_oracle(c3, include_energy):
    """C3 at its real kmax = (27,27,27), 74,438 k-vectors: energy 1e-6, forces 1e-5 relative RMS. The forces-only call
    is the per-MD-step call of the benchmark (tensor-core structure factors + tensor-core gather)."""
    data, pos, box, force = _fullk("c3")
    ctx = c3[3]
    assert tuple(data["kmax"]) == ctx.kernel.ewald_params()[1] == (27, 27, 27)
    _check_fullk(ctx, data, pos, include_energy)


def test_c3_full_kmax27_live_oracle_over_the_host_cores(c3):
    """The same comparison against the oracle run HERE, its k-sum in parallel slabs (about a minute of CPU time)."""
    from oracle.slabs import execute_parallel, host_cores
    if host_cores() < 4:
        pytest.skip("needs a few host cores")
    pos, box, force, ctx, e, f, comps = c3
    eo, fo, dedq, kmax = execute_parallel(force, box, pos)
    assert kmax == (27, 27, 27)
    assert abs(e - eo[4]) <= E_RTOL * abs(eo[4])
    assert np.abs(comps[:4] - eo[:4]).max() <= E_RTOL * abs(eo[4])
    assert rel_rms(f, fo) <= F_RTOL
    e2, f2, _ = ctx.evaluate(pos, True, False)
    assert rel_rms(f2, fo) <= F_RTOL
    assert rel_rms(ctx.kernel.dedq(), dedq) <= F_RTOL


@pytest.mark.parametrize("include_energy", [True, False])
def test_c4_262k_atoms_kmax55_against_the_oracle(build_native, include_energy):
    """C4 (262,143 atoms, kmax 55, 647,514 k-vectors): the <1,64> gather and <128,1> structure-factor variants."""
    data, pos, box, force = _fullk("c4")
    ctx = runtime.CoulContext(force, box)
    assert tuple(data["kmax"]) == ctx.kernel.ewald_params()[1] == (55, 55, 55)
    _check_fullk(ctx, data, pos, include_energy)
    ctx.kernel.close()
