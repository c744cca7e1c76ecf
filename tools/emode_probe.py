"""GPU box: per-kernel times of the forces-only call with and without the discarded partial energy, and the MD step."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime, md
pos, box, force = synthetic.config('c3')
dpos = torch.tensor(pos.reshape(-1), device='cuda')
for skip in (False, True):
    k = runtime.CalcCoulForceKernel(skip_discarded_energy=skip)
    k.initialize(box, force)
    kt = k.time_kernels(dpos.data_ptr(), box, 20, True, False)
    print("skip_discarded_energy=%s  device step %.4f ms  pairs %.4f  sum %.4f" % (skip, k.time_device(dpos.data_ptr(), box, 50, True, False), kt['direct_pairs'], sum(kt.values())))
    k.close()
sim, p0 = md.flexible_water_simulation(10922, 32768, cutoff=1.0, ewald_tol=1e-5)
sim.minimize(200, 0.002)
p, _ = sim.get_state()
sim.set_state(p, sim.maxwell_boltzmann(300.0, seed=7))
sim.step(50, 0.0005)
print("MD ms/step", sim.step(1000, 0.0005)/1000)
p, _ = sim.get_state()
sim.close()
dpos2 = torch.tensor(p.reshape(-1), device='cuda')
k = runtime.CalcCoulForceKernel(skip_discarded_energy=True)
k.initialize(box, force)
kt = k.time_kernels(dpos2.data_ptr(), box, 20, True, False)
print("thermalised positions: device step %.4f ms  kernels %s" % (k.time_device(dpos2.data_ptr(), box, 50, True, False), {a: round(b, 4) for a, b in kt.items()}))
