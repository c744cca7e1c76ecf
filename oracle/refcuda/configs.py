"""Emit the -D flags the reference's host code would pass to its runtime compiler
(platforms/cuda/src/CudaCoulKernels.cpp:377-389,466-506) for one benchmark configuration.
BASELINE TOOLING, NOT PRODUCT."""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from openmm_chargeflux_b200 import synthetic  # noqa: E402


def kmax_rule(length, alpha, tol):
    k = 1
    while 0.05 * math.sqrt(length * alpha) * k * math.exp(-(k * math.pi / (length * alpha)) ** 2) > tol:
        k += 1
    return k + 1 if k % 2 == 0 else k


def defines(name):
    pos, box, f = synthetic.config(name)
    n = f.getNumParticles()
    padded = (n + 31) // 32 * 32
    tol, rc = f.getEwaldErrorTolerance(), f.getCutoffDistance()
    alpha = math.sqrt(-math.log(2 * tol)) / rc
    k = [kmax_rule(box[d][d], alpha, tol) for d in range(3)]
    nb, na, nw = f.getNumFluxBonds(), f.getNumFluxAngles(), f.getNumFluxWaters()
    d = {
        "NUM_ATOMS": n, "PADDED_NUM_ATOMS": padded, "NUM_BLOCKS": padded // 32, "THREAD_BLOCK_SIZE": 64,
        "EWALDFORCEBLOCK": 32, "TILE_SIZE": 32, "NUM_TILES_WITH_EXCLUSIONS": padded // 32,
        "FIRST_EXCLUSION_TILE": 0, "LAST_EXCLUSION_TILE": padded // 32,
        "USE_PERIODIC": 1, "USE_CUTOFF": 1, "USE_SYMMETRIC": 1, "INCLUDE_FORCES": 1, "INCLUDE_ENERGY": 1,
        "CUTOFF": "%.9ff" % rc, "EWALD_ALPHA": "%.9ff" % alpha,
        "TWO_OVER_SQRT_PI": "%.9ff" % (2 / math.sqrt(math.pi)), "ONE_OVER_SQRT_PI": "%.9ff" % (1 / math.sqrt(math.pi)),
        "KMAX_X": k[0], "KMAX_Y": k[1], "KMAX_Z": k[2], "KSIZEX": 2 * k[0] - 1, "KSIZEY": 2 * k[1] - 1,
        "KSIZEZ": 2 * k[2] - 1, "KSIZEYZ": (2 * k[1] - 1) * (2 * k[2] - 1),
        "TOTALK": (2 * k[0] - 1) * (2 * k[1] - 1) * (2 * k[2] - 1),
        "EXP_COEFFICIENT": "%.9ef" % (-1.0 / (4 * alpha * alpha)), "ONE_4PI_EPS0": "138.935456f",
        "NUM_FLUX_BONDS": nb, "NUM_FLUX_ANGLES": na, "NUM_FLUX_WATERS": nw, "BSHIFT": 4 * nb, "BASHIFT": 4 * nb + 9 * na,
        "NUM_DQDX_PAIRS": 4 * nb + 9 * na + 9 * nw,
    }
    return " ".join("-D%s=%s" % kv for kv in d.items())


if __name__ == "__main__":
    print(defines(sys.argv[1]))
