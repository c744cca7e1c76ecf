"""GPU box: structure-factor kernel time at C3 for both calls (integer tensor-core kernel) under the current CFX_SI_* settings."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, force = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'c3')
dpos = torch.tensor(pos.reshape(-1), device='cuda')
k = runtime.CalcCoulForceKernel()
k.initialize(box, force)
out = []
for inc_e in (True, False):
    kt = k.time_kernels(dpos.data_ptr(), box, 20, True, inc_e)
    out.append("E=%d S %.4f step %.4f" % (inc_e, kt['structure_factor'], k.time_device(dpos.data_ptr(), box, 30, True, inc_e)))
f = np.zeros_like(pos)
print("  ".join(out), " E %.8f" % k.execute(pos, box, f, True, True))
