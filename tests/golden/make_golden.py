"""Generate the golden fixtures in tests/golden/ from the plugin's OWN sources.

Runs in the build container only (needs oracle/_ref/libcfx_ref.so, i.e. /root/reference compiled
against the OpenMM stand-in by oracle/Makefile). Each fixture stores the inputs (positions, box and
the generator call that rebuilds the CoulForce) and the outputs of the reference's
ReferenceCalcCoulForceKernel: total energy and forces for all four (includeForces, includeEnergy)
combinations, the charges q(x), the Jacobian rows and the neighbour list.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from openmm_chargeflux_b200 import synthetic  # noqa: E402
from oracle import ReferenceBuild  # noqa: E402

CASES = {
    # name: (generator, kwargs)
    "c1_water64_nopbc": ("water_box", dict(n_waters=64, seed=64, periodic=False)),
    "water216_pbc": ("water_box", dict(n_waters=216, seed=1, periodic=True, cutoff=0.9, ewald_tol=1e-4)),
    "fluxwater216_pbc": ("water_box", dict(n_waters=216, seed=2, periodic=True, cutoff=0.9, ewald_tol=1e-5, flux="water")),
    "water400_rect": ("water_box", dict(n_waters=400, seed=7, periodic=True, cutoff=0.8, ewald_tol=1e-4)),
    "methanol_water_small": ("methanol_water", dict(n_methanol=30, n_water=90, seed=5, cutoff=0.8, ewald_tol=1e-5)),
    "rock_salt": ("rock_salt", dict(cells=3)),
}


def build(name):
    gen, kw = CASES[name]
    pos, box, force = getattr(synthetic, gen)(**kw)
    if name == "water400_rect":
        # non-cubic box + atoms displaced by whole box vectors (unwrapped input)
        box = np.diag([box[0, 0] * 1.10, box[1, 1] * 0.95, box[2, 2] * 1.02])
        rng = np.random.Generator(np.random.PCG64(99))
        shift = rng.integers(-2, 3, size=(len(pos) // 3, 1, 3)) * np.diag(box)[None, None, :]
        pos = (pos.reshape(-1, 3, 3) + shift).reshape(-1, 3)
    return pos, box, force


def main():
    for name in CASES:
        pos, box, force = build(name)
        ref = ReferenceBuild(force, box)
        out = dict(positions=pos, box=box)
        for inc_f in (1, 0):
            for inc_e in (1, 0):
                e, f = ref.execute(pos, box, bool(inc_f), bool(inc_e))
                out["energy_f%d_e%d" % (inc_f, inc_e)] = e[4]
                out["forces_f%d_e%d" % (inc_f, inc_e)] = f
        ref.execute(pos, box, True, True)
        out["charges"] = ref.charges()
        dq, dx, val = ref.jacobian()
        out["jac_dq"], out["jac_dx"], out["jac_val"] = dq, dx, val
        if force.usesPeriodicBoundaryConditions():
            alpha, kmax, nk = ref.ewald_params()
            out["alpha"], out["kmax"], out["num_kvectors"] = alpha, np.asarray(kmax), nk
            out["pairs"] = ref.neighbor_pairs()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-24s N=%5d  E=%.10g  %.0f KB" % (name, len(pos), out["energy_f1_e1"], os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
