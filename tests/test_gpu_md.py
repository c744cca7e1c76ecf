"""GPU: the MD harness (bonded forces + velocity Verlet around the CalcCoulForce kernel).
Energy conservation in NVE is a whole-path check: it fails unless the charge-flux chain-rule forces are
the exact gradient of the energy the kernel reports."""
import numpy as np
import pytest

from openmm_chargeflux_b200 import md

pytestmark = pytest.mark.gpu


def test_bonded_forces_are_the_gradient_of_the_bonded_energy(build_native):
    sim, pos = md.flexible_water_simulation(27, seed=5, cutoff=0.45, ewald_tol=1e-5)
    e0 = sim.energies()
    # total force from a finite difference of the total potential energy (bonded + CoulForce)
    h = 2e-4          # the FP32-accumulated reciprocal energy carries ~1e-4 kJ/mol of rounding noise
    p, _ = sim.get_state()
    for atom, c in ((0, 0), (1, 2), (41, 1)):
        q = p.copy(); q[atom, c] += h
        sim.set_state(q); ep = sim.energies()
        q[atom, c] -= 2 * h
        sim.set_state(q); em = sim.energies()
        fd = -((ep["bonded"] + ep["coulomb"]) - (em["bonded"] + em["coulomb"])) / (2 * h)
        # the analytic force: one tiny step of velocity Verlet from rest gives v = F dt / m
        sim.set_state(p, np.zeros_like(p))
        dt = 1e-7
        sim.step(1, dt)
        _, v = sim.get_state()
        f = v[atom, c] * sim.masses[atom] / dt
        assert abs(f - fd) <= 1e-3 * max(abs(fd), 1e3)
    assert np.isfinite(e0["total"])


def test_nve_energy_conservation_small_box(build_native):
    sim, pos = md.flexible_water_simulation(216, seed=1, cutoff=0.9, ewald_tol=1e-5)
    sim.minimize(300, 0.002)
    p, _ = sim.get_state()
    sim.set_state(p, sim.maxwell_boltzmann(300.0, seed=3))
    e0 = sim.energies()
    sim.step(400, 0.00025)
    e1 = sim.energies()
    drift = abs(e1["total"] - e0["total"])
    assert drift <= 2e-3 * e0["kinetic"], (e0, e1)
    assert abs(e1["kinetic"] - e0["kinetic"]) > 1e-3 * e0["kinetic"]      # something actually moved
