"""GPU box: thermalise a benchmark box with the MD harness (minimise, 300 K, 2 ps) and save the positions."""
import sys
import numpy as np
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import md, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/thermal_%s.npy" % name
cfg = synthetic.CONFIGS[name]
sim, p0 = md.flexible_water_simulation(cfg["n_waters"], cfg["seed"], cutoff=cfg["cutoff"], ewald_tol=cfg["ewald_tol"])
sim.minimize(200, 0.002)
p, _ = sim.get_state()
sim.set_state(p, sim.maxwell_boltzmann(300.0, seed=7))
sim.step(int(sys.argv[3]) if len(sys.argv) > 3 else 4000, 0.0005)
p, _ = sim.get_state()
np.save(out, p)
print("saved", out, p.shape, "energies", sim.energies())
