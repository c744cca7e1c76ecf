import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmm_chargeflux_b200 import runtime, synthetic
pos, box, force = synthetic.config("c3")
k = runtime.CalcCoulForceKernel(use_graph=False)
k.initialize(box, force)
f = np.zeros_like(pos)
for _ in range(3):
    e = k.execute(pos, box, f, False, True)
print("E", e)
