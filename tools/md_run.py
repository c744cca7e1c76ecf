import sys, json, time, numpy as np
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import md
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 10922
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dt = 0.0005
sim, pos = md.flexible_water_simulation(nw, seed=32768 if nw == 10922 else 1, cutoff=1.0, ewald_tol=1e-5)
e_start = sim.energies(); print("initial", e_start)
t = time.time(); sim.minimize(400, 0.002); print("minimize s", time.time() - t, sim.energies())
p, _ = sim.get_state()
sim.set_state(p, sim.maxwell_boltzmann(300.0, seed=7))
sim.step(200, dt)                                    # settle
log = []
e0 = sim.energies(); log.append(e0); print("t=0", e0)
total_ms = 0.0
chunk = steps // 10
for k in range(10):
    total_ms += sim.step(chunk, dt)
    e = sim.energies(); log.append(e)
    print("step", (k + 1) * chunk, {a: round(b, 3) for a, b in e.items()}, "drift", e["total"] - e0["total"])
ms_step = total_ms / (chunk * 10)
n = 3 * nw
out = {"atoms": n, "steps": chunk * 10, "dt_fs": dt * 1000, "ms_per_step": ms_step, "steps_per_s": 1e3 / ms_step,
       "ns_per_day": 1e3 / ms_step * dt * 1e-3 * 86400, "e_total_start": e0["total"], "e_total_end": log[-1]["total"],
       "drift_kj_mol": log[-1]["total"] - e0["total"], "kinetic": e0["kinetic"],
       "drift_over_kinetic": (log[-1]["total"] - e0["total"]) / e0["kinetic"],
       "max_fluctuation_over_kinetic": max(abs(l["total"] - e0["total"]) for l in log) / e0["kinetic"]}
print(json.dumps(out))
