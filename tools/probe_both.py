"""Two energy+forces and two forces-only evaluations of C3 through the C ABI without graphs: the command ncu wraps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmm_chargeflux_b200 import runtime, synthetic
pos, box, force = synthetic.config("c3")
k = runtime.CalcCoulForceKernel(use_graph=False)
k.initialize(box, force)
f = np.zeros_like(pos)
for inc_e in (True, True, False, False):
    f[:] = 0
    e = k.execute(pos, box, f, True, inc_e)
print("E", e, "pairs", k.stats().pairs_in_cutoff)
