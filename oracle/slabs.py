"""The CPU oracle's explicit k-sum spread over the host cores. TEST INFRASTRUCTURE, NOT PRODUCT.

The reference's reciprocal loops (ReferenceCoulKernels.cpp:513-556) are a plain sum over the half-space
k-vectors, outermost index nkx. ``Oracle.set_kx_range(lo, hi)`` restricts one evaluation to the slab
lo <= nkx < hi, so the full sum is assembled from independent slab evaluations, one per worker process:

    total = R(no k) + sum_slabs [R(slab) - R(no k)]

(every slab evaluation repeats the direct/self/exclusion terms, which cancel in the difference; the chain
rule is linear in dE/dq, so it distributes over the slabs as well). Energies and forces agree with the
single-threaded evaluation to the order of the floating-point sum (~1e-15 relative): good for the 1e-6 /
1e-5 parity bars at the full benchmark sizes, where one thread would need minutes (C3) to hours (C4).
"""
import multiprocessing as mp
import os

import numpy as np

from openmm_chargeflux_b200 import _abi

_JOB = {}


def _run_slab(rng):
    from oracle import Oracle
    lo, hi = rng
    o = Oracle(_JOB["force"], _JOB["box"])
    o.set_kx_range(lo, hi)
    e, f = o.execute(_JOB["pos"], _JOB["box"], _JOB["inc_f"], True)
    d = o.dedq()
    o.close()
    return lo, e, f, d


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def execute_parallel(force, box, pos, include_forces=True, workers=None):
    """One includeEnergy=True evaluation of the oracle with the k-sum spread over `workers` processes.
    Returns (energy[5], forces[N,3], dedq[N], kmax)."""
    from oracle import Oracle
    box = np.asarray(box, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    probe = Oracle(force, box)
    kmax = probe.ewald_params()[1]
    probe.close()
    workers = workers or host_cores()
    # one slab per nkx (nkx = 0 carries half the work of the others): a pool balances them
    slabs = [(0, 0)] + [(nx, nx + 1) for nx in range(kmax[0])]
    _JOB.update(force=force, box=box, pos=pos, inc_f=bool(include_forces))
    ctx = mp.get_context("fork")
    with ctx.Pool(min(workers, len(slabs))) as pool:
        results = pool.map(_run_slab, slabs, chunksize=1)
    _JOB.clear()
    base = results[0]
    energy = base[1].copy()
    forces = base[2].copy()
    dedq = base[3].copy()
    for lo, e, f, d in results[1:]:
        energy[_abi.E_RECIP] += e[_abi.E_RECIP]
        forces += f - base[2]
        dedq += d - base[3]
    energy[_abi.E_TOTAL] = energy[_abi.E_SELF] + energy[_abi.E_RECIP] + energy[_abi.E_DIRECT] + energy[_abi.E_EXCL]
    return energy, forces, dedq, kmax
