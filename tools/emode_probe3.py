import sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime, md
pos, box, force = synthetic.config('c3')
d = torch.tensor(pos.reshape(-1), device='cuda')
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def report(label):
    k = runtime.CalcCoulForceKernel(skip_discarded_energy=True)
    k.initialize(box, force)
    kt = k.time_kernels(d.data_ptr(), box, 20, True, False)
    print("%-34s pairs %.4f S %.4f gather %.4f | %s" % (label, kt['direct_pairs'], kt['structure_factor'], kt['kspace_gather'], smi()), flush=True)
    return k
k = report("fresh")
ms = k.time_device(d.data_ptr(), box, 6000, True, False)
print("6000 back-to-back evaluations: %.4f ms each | %s" % (ms, smi()))
k.close()
report("after 6000 evaluations").close()
sim, p0 = md.flexible_water_simulation(10922, 32768, cutoff=1.0, ewald_tol=1e-5)
sim.minimize(50, 0.002)
report("after md minimize(50)").close()
print("MD 1000 steps: %.4f ms/step | %s" % (sim.step(1000, 0.0005)/1000, smi()))
report("after md 1000 steps (sim alive)").close()
sim.close()
report("after sim.close()").close()
time.sleep(3)
report("after 3 s idle").close()
