"""Golden vectors at BASELINE.json's FULL benchmark sizes: C3 (32,766 atoms, kmax 27) and C4 (262,143 atoms, kmax 55).

The single-threaded reference needs ~90 s (C3) / ~1.7 h (C4) per evaluation, so the fixtures are produced by the
oracle restatement (pinned bit-identical to the plugin's own sources by tests/test_oracle_vs_ref.py) with its
explicit k-sum spread over the host cores (oracle/slabs.py). Inputs are NOT stored: they are rebuilt from
openmm_chargeflux_b200.synthetic.config(name) (seeded), a checksum of the positions guards against drift.

Stored per case: the five energy components (float64), forces as float32 (relative rounding 6e-8, three orders
below the 1e-5 bar they are compared at) plus float64 forces and dE/dq of 2,048 probe atoms, float64 column sums,
and the in-cutoff pair count.

    python tests/golden/make_golden_fullsize.py c3 c4
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from openmm_chargeflux_b200 import synthetic  # noqa: E402
from oracle import Oracle  # noqa: E402
from oracle.slabs import execute_parallel  # noqa: E402


def probe_atoms(n):
    return np.sort(np.random.Generator(np.random.PCG64(2024)).choice(n, size=min(2048, n), replace=False))


def main(names):
    for name in names:
        pos, box, force = synthetic.config(name)
        t = time.time()
        e, f, dedq, kmax = execute_parallel(force, box, pos)
        o = Oracle(force, box)
        o.set_kx_range(0, 0)
        o.execute(pos, box, True, True)
        npairs = o.stats().pairs_in_cutoff
        probe = probe_atoms(len(pos))
        out = dict(energy=e, kmax=np.asarray(kmax), forces_f32=f.astype(np.float32), probe=probe, forces_probe=f[probe],
                   dedq_probe=dedq[probe], force_sum=f.sum(axis=0), force_sq_sum=float((f ** 2).sum()),
                   dedq_f32=dedq.astype(np.float32), pos_checksum=float((pos * np.arange(1, 4)[None, :]).sum()),
                   pairs_in_cutoff=np.int64(npairs))
        path = os.path.join(HERE, name + "_fullk.npz")
        np.savez_compressed(path, **out)
        print("%s N=%d kmax=%s E=%.10g pairs=%d  %.0f s  %.0f KB" % (name, len(pos), kmax, e[4], npairs, time.time() - t,
                                                                     os.path.getsize(path) / 1024), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c3"])
