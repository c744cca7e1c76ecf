"""Direct-space candidate lists (csrc/direct.cu): built for cutoff + skin, reused until an atom has moved skin/2, with the
in-cutoff test exact at every evaluation (reference: the neighbour list ReferenceCoulKernels.cpp:559-565 rebuilds each call)."""
import os

import numpy as np
import pytest

from openmm_chargeflux_b200 import runtime, synthetic
from oracle import Oracle

pytestmark = pytest.mark.gpu


def _box(n_waters=2744, seed=11):
    # 8,232 atoms, L = 4.35 nm: 7 cells per axis with the default skin -> the list (fast) path
    return synthetic.water_box(n_waters, seed=seed, cutoff=1.0, ewald_tol=1e-4)


def _relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def test_lists_are_reused_and_rebuilt_on_the_displacement_trigger():
    pos, box, force = _box()
    ctx = runtime.CoulContext(force, box)                          # default skin 0.1 nm
    fresh = runtime.CoulContext(force, box, list_skin=0.0)          # rebuilds at every evaluation
    oracle = Oracle(force, box)
    rng = np.random.default_rng(5)
    step = rng.normal(scale=0.002, size=pos.shape)                  # fastest atom ~0.009 nm per step: skin/2 after ~6 steps
    builds = []
    p = pos.copy()
    for it in range(12):
        e, f, _ = ctx.evaluate(p)
        e2, f2, comps = fresh.evaluate(p)
        builds.append(ctx.kernel.stats().pair_list_builds)
        # same neighbour set (bit-exact, against the oracle at a few steps), same physics as a handle without history
        assert ctx.kernel.stats().pairs_in_cutoff == fresh.kernel.stats().pairs_in_cutoff
        # (the total of this box is ~80 kJ/mol out of components of 1e5: compare at the scale of the components; the
        # energies are summed as 2^-24 fixed point per warp, so two summation orders differ by ~1e-6 kJ/mol)
        assert abs(e - e2) <= 1e-10 * np.abs(comps[:4]).max(), (it, e, e2)
        assert _relrms(f, f2) < 4e-6, (it, _relrms(f, f2))      # two FP32 summation orders, each ~1e-6 from the oracle
        if it in (0, 5, 11):
            eo, fo = oracle.execute(p, box)
            assert np.array_equal(ctx.kernel.neighbor_pairs(), oracle.neighbor_pairs())
            assert abs(e - eo[4]) <= 1e-6 * abs(eo[4]) and _relrms(f, fo) < 1e-5
        p = p + step
    assert builds[0] == 1 and builds[1] == 1 and builds[3] == 1     # reused while nothing has moved skin/2
    assert 2 <= builds[-1] <= 3                                     # rebuilt when the trigger fired
    assert fresh.kernel.stats().pair_list_builds >= 12
    # same positions again: bitwise reproducible (no rebuild in between)
    ea, fa, _ = ctx.evaluate(p)
    eb, fb, _ = ctx.evaluate(p)
    assert ea == eb and np.array_equal(fa, fb)
    ctx.kernel.close(); fresh.kernel.close()


def test_jump_by_a_box_length_is_not_a_displacement():
    pos, box, force = _box(seed=3)
    ctx = runtime.CoulContext(force, box)
    e, f, _ = ctx.evaluate(pos)
    shifted = pos.copy()
    shifted[::7] += np.diag(box) * np.array([1.0, -2.0, 1.0])       # callers may re-wrap molecules between calls
    e2, f2, _ = ctx.evaluate(shifted)
    assert ctx.kernel.stats().pair_list_builds == 1
    assert abs(e - e2) <= 1e-9 * abs(e) and _relrms(f2, f) < 1e-6
    ctx.kernel.close()


def test_list_overflow_falls_back_to_the_generic_kernel_and_grows(monkeypatch):
    pos, box, force = _box(seed=4)
    ref = runtime.CoulContext(force, box)
    e0, f0, _ = ref.evaluate(pos)
    monkeypatch.setenv("CFX_LIST_CAP", "256")                       # far too small: every cluster overflows
    ctx = runtime.CoulContext(force, box)
    monkeypatch.delenv("CFX_LIST_CAP")
    for _ in range(5):                                              # 256 -> 512 -> 1024 -> 2048 -> 4096
        e, f, _ = ctx.evaluate(pos)
        assert abs(e - e0) <= 1e-9 * abs(e0) and _relrms(f, f0) < 2e-6
        assert ctx.kernel.stats().pairs_in_cutoff == ref.kernel.stats().pairs_in_cutoff
    assert np.array_equal(f, f0) and e == e0                        # lists large enough again: identical to the reference handle
    ctx.kernel.close(); ref.kernel.close()


@pytest.mark.parametrize("world", [3])
def test_sharded_handles_on_the_list_path_sum_to_the_unsharded_result(world):
    """Every rank lists only its own slab of i-clusters; the slabs' fixed-point sums reproduce the unsharded evaluation,
    before and after the atoms have moved (ranks emulated as consecutive launches on one GPU)."""
    import torch
    pos, box, force = _box(seed=6)
    n = len(pos)
    whole = runtime.CoulContext(force, box)
    kernels = [runtime.CalcCoulForceKernel(shard_rank=r, shard_count=world) for r in range(world)]
    for k in kernels:
        k.initialize(box, force)
    npad = kernels[0].padded_num_particles()
    stream = torch.cuda.Stream()
    rng = np.random.default_rng(9)
    p = pos.copy()
    for it in range(3):
        e, f, comps = whole.evaluate(p)
        d_pos = torch.tensor(p.reshape(-1), device="cuda")
        d_force = torch.zeros(3 * npad, dtype=torch.int64, device="cuda")
        d_energy = torch.zeros(8, dtype=torch.float64, device="cuda")
        with torch.cuda.stream(stream):
            for k in kernels:
                k.execute_device(d_pos.data_ptr(), box, d_force.data_ptr(), 0, d_energy.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        fs = d_force.cpu().numpy().reshape(3, npad)[:, :n].T / 4294967296.0
        es = d_energy.cpu().numpy()
        assert np.abs(es[:4] - comps[:4]).max() <= 1e-9 * np.abs(comps[:4]).max()
        assert _relrms(fs, f) <= 2e-6
        p = p + rng.normal(scale=0.01, size=p.shape)                # it = 1, 2: past skin/2 for some atoms -> rebuild
    assert all(k.stats().pair_list_builds >= 2 for k in kernels)
    for k in kernels:
        k.close()
    whole.kernel.close()
