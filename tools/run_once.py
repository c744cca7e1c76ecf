import sys, numpy as np, torch
sys.path.insert(0,'.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, f = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'c3')
ctx = runtime.CoulContext(f, box)
for i in range(2):
    e, frc, comps = ctx.evaluate(pos, True, False)      # the per-MD-step call: tensor-core k-space kernels
for i in range(2):
    e, frc, comps = ctx.evaluate(pos, True, True)       # energy + forces: FP32 structure factors, FP64 pair energies
print(e)
