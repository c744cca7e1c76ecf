"""GPU: the tensor-core reciprocal-space kernels (kspace_tc.cu) against the FP32 CUDA-core kernels and the oracle.

The tensor kernels replace the FP32 ones whenever the geometry allows (|nz| <= 32 for the structure factors of a
forces-only evaluation, 2*Kz <= 112 and Ky >= 8 for the gather); CFX_KSPACE_S / CFX_KSPACE_GATHER = fp32 pin the
FP32 kernels at plan time, which is how the two paths are compared here."""
import os

import numpy as np
import pytest

from conftest import F_RTOL, rel_rms
from openmm_chargeflux_b200 import runtime, synthetic
from oracle import Oracle

pytestmark = pytest.mark.gpu


def _context(force, box, fp32):
    keys = ("CFX_KSPACE_S", "CFX_KSPACE_GATHER")
    old = {k: os.environ.get(k) for k in keys}
    try:
        for k in keys:
            if fp32:
                os.environ[k] = "fp32"
            else:
                os.environ.pop(k, None)
        return runtime.CoulContext(force, box)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_tensor_and_fp32_kspace_paths_agree_on_the_4k_box(build_native):
    pos, box, force = synthetic.config("c2")
    tens, fp32 = _context(force, box, False), _context(force, box, True)
    for inc_e in (False, True):
        e_t, f_t, c_t = tens.evaluate(pos, True, inc_e)
        e_f, f_f, c_f = fp32.evaluate(pos, True, inc_e)
        assert rel_rms(f_t, f_f) <= 3e-6, inc_e
        assert rel_rms(tens.kernel.dedq(), fp32.kernel.dedq()) <= 3e-6
        if inc_e:
            # energy evaluations use the FP32 (round-to-nearest) structure-factor kernel on both paths
            assert abs(e_t - e_f) <= 1e-9 * np.abs(c_f[:4]).max()


@pytest.mark.parametrize("cutoff,flux,kz_range", [(0.24, "bond+angle", (29, 32)), (0.24, "water", (29, 32)), (0.23, "bond+angle", (33, 56))])
def test_large_kmax_variants_of_the_tensor_kernels_against_the_oracle(build_native, cutoff, flux, kz_range):
    """Short cutoff => large alpha => |nz| up to ~30: the gather runs its one-atom-tile / 64-column variant
    (K padded to 64 > 56) and the structure factors still fit one N = 64 MMA; |nz| > 32 selects their N = 128 variant."""
    pos, box, force = synthetic.water_box(216, seed=11, cutoff=cutoff, ewald_tol=1e-5, flux=flux)
    o = Oracle(force, box)
    ctx = _context(force, box, False)
    kmax = ctx.kernel.ewald_params()[1]
    assert kz_range[0] <= kmax[2] <= kz_range[1], kmax
    eo, fo = o.execute(pos, box, True, True)
    e, f, comps = ctx.evaluate(pos, True, True)
    # alpha = 13.7/nm makes the components (2.4e5 kJ/mol) cancel to 6e2: the FP32 reciprocal sums are good to 1e-7 of
    # the component scale, which is what is checked here; the headline tolerances are tested on the physical boxes
    assert abs(e - eo[4]) <= 1e-7 * np.abs(eo[:4]).max()
    assert rel_rms(f, fo) <= F_RTOL
    eo, fo = o.execute(pos, box, True, False)
    e, f, comps = ctx.evaluate(pos, True, False)                  # forces only: tensor-core structure factors too
    assert rel_rms(f, fo) <= F_RTOL
    assert rel_rms(ctx.kernel.dedq(), o.dedq()) <= F_RTOL


def test_tf32_peak_measurement_is_plausible(build_native):
    tf = runtime.measure_tf32_peak(iters=5000)
    assert 300.0 < tf < 2500.0, tf


def test_i8_peak_measurement_is_plausible(build_native):
    tops = runtime.measure_i8_peak(iters=4000)
    assert 2000.0 < tops < 5000.0          # B200 dense INT8 nominal 4.5 POP/s


def test_one_large_charge_switches_the_forces_only_call_to_four_digit_planes(build_native):
    """The integer structure-factor kernel scales its fixed point by the largest |q|: when one charge dwarfs the others
    (api.cu: forceDigitsFor), the forces-only call must use the four-digit variant to keep the small charges' bits. A
    +-6 e ion pair in water (10x the typical charge): forces-only and energy calls against the oracle."""
    from openmm_chargeflux_b200.force import CoulForce
    pos, box, f = synthetic.water_box(512, seed=9, cutoff=0.9, ewald_tol=1e-5)
    g = CoulForce()
    for i in range(f.getNumParticles()):
        q, s, e = f.getParticleParameters(i)
        g.addParticle(6.0 if i == 0 else (-6.0 if i == 300 else q), s, e)
    for i in range(f.getNumExceptions()):
        g.addException(*f.getExceptionParameters(i))
    for i in range(f.getNumFluxBonds()):
        g.addFluxBond(*f.getFluxBondParameters(i))
    for i in range(f.getNumFluxAngles()):
        g.addFluxAngle(*f.getFluxAngleParameters(i))
    g.setCutoffDistance(f.getCutoffDistance()); g.setEwaldErrorTolerance(f.getEwaldErrorTolerance())
    g.setUsesPeriodicBoundaryConditions(True)
    o = Oracle(g, box)
    ctx = runtime.CoulContext(g, box)
    for inc_e in (False, True):
        e, frc, _ = ctx.evaluate(pos, True, inc_e)
        eo, fo = o.execute(pos, box, True, inc_e)
        assert rel_rms(frc, fo) <= F_RTOL, inc_e
        if inc_e:
            assert abs(e - eo[4]) <= 1e-6 * max(abs(eo[4]), 1e-3 * np.abs(eo[:4]).max())
    assert rel_rms(ctx.kernel.dedq(), o.dedq()) <= F_RTOL


def test_integer_structure_factors_do_not_depend_on_the_pipeline_depth(build_native, monkeypatch):
    """Integer sums are exact, so nothing about the kernel's schedule may show in the result: the same evaluation with
    2/2, 3/2 and the default row / operand ring depths must agree to the last bit (a race in the TMA -> formers -> MMA
    pipeline would not)."""
    pos, box, force = synthetic.config("c2")
    results = []
    for rs, ob in ((None, None), ("2", "2"), ("3", "2")):
        for name, val in (("CFX_SI_ROW_STAGES", rs), ("CFX_SI_OP_STAGES", ob)):
            if val is None:
                monkeypatch.delenv(name, raising=False)
            else:
                monkeypatch.setenv(name, val)
        ctx = runtime.CoulContext(force, box)
        out = []
        for inc_e in (True, False, True):
            e, f, comps = ctx.evaluate(pos, True, inc_e)
            out.append((e, f.copy(), comps.copy()))
        assert out[0][0] == out[2][0] and np.array_equal(out[0][1], out[2][1])      # repeatable within a handle
        results.append(out)
        ctx.kernel.close()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
