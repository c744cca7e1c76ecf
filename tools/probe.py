"""A few evaluations of one benchmark configuration through the C ABI: the command ncu wraps.

    python tools/probe.py [c3] [forces|energy] [evaluations]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmm_chargeflux_b200 import runtime, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
energy = len(sys.argv) > 2 and sys.argv[2] == "energy"
count = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pos, box, force = synthetic.config(name)
if os.environ.get("CFX_PROBE_POSITIONS"):
    pos = np.load(os.environ["CFX_PROBE_POSITIONS"])
k = runtime.CalcCoulForceKernel(use_graph=False)
k.initialize(box, force)
f = np.zeros_like(pos)
for _ in range(count):
    f[:] = 0
    e = k.execute(pos, box, f, True, energy)
st = k.stats()
print(name, "E", e, "|F|rms", float(np.sqrt((f ** 2).mean())), "pairs", st.pairs_in_cutoff, "distance tests", st.pair_candidates)
