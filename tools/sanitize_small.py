import sys, numpy as np
sys.path.insert(0,'.')
from openmm_chargeflux_b200 import synthetic, runtime, md
for name,(pos,box,f) in (("c1", synthetic.config("c1")), ("w125", synthetic.water_box(125, seed=21, cutoff=0.75, ewald_tol=3e-5)),
                         ("mw", synthetic.methanol_water(10, 30, seed=1, cutoff=0.55, ewald_tol=1e-4))):
    ctx = runtime.CoulContext(f, box)
    for fl in ((True,True),(True,False),(False,True)):
        e, frc, comps = ctx.evaluate(pos, *fl)
    if f.usesPeriodicBoundaryConditions():
        print(name, e, len(ctx.kernel.neighbor_pairs()))
    else: print(name, e)
sim, pos = md.flexible_water_simulation(64, seed=3, cutoff=0.6, ewald_tol=1e-5)
sim.minimize(5, 0.002); sim.step(5, 0.0005); print("md", sim.energies())
