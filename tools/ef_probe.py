"""GPU box: per-kernel times of the energy+forces and forces-only calls at C3 (or argv[1]) and the energy components."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
pos, box, force = synthetic.config(name)
dpos = torch.tensor(pos.reshape(-1), device='cuda')
k = runtime.CalcCoulForceKernel()
k.initialize(box, force)
for inc_e in (True, False):
    kt = k.time_kernels(dpos.data_ptr(), box, 20, True, inc_e)
    print("includeEnergy=%d  device step %.4f ms  sum %.4f  %s" % (inc_e, k.time_device(dpos.data_ptr(), box, 50, True, inc_e),
          sum(kt.values()), {a: round(b, 4) for a, b in kt.items()}))
f = np.zeros_like(pos)
e = k.execute(pos, box, f, True, True)
print("E", repr(e), "components", k.energy_components() if hasattr(k, "energy_components") else "")
