import sys, numpy as np, torch
sys.path.insert(0,'.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, f = synthetic.config('c3')
ctx = runtime.CoulContext(f, box)
for i in range(3):
    e, frc, comps = ctx.evaluate(pos)
print(e)
