import sys, time, numpy as np, torch
sys.path.insert(0,'.')
from openmm_chargeflux_b200 import synthetic, runtime
t=time.time(); pos, box, f = synthetic.config('c4'); print("gen", time.time()-t)
ctx = runtime.CoulContext(f, box)
print(ctx.kernel.ewald_params())
t=time.time(); e, frc, comps = ctx.evaluate(pos); print("first eval", time.time()-t, comps)
print("sumF", frc.sum(0), "Fmax", np.abs(frc).max(), "finite", np.isfinite(frc).all())
dpos = torch.tensor(pos, device='cuda')
print("ms/eval E+F", ctx.kernel.time_device(dpos.data_ptr(), box, 5), "F only", ctx.kernel.time_device(dpos.data_ptr(), box, 5, True, False))
print("E+F", ctx.kernel.time_kernels(dpos.data_ptr(), box, 3))
print("F only", ctx.kernel.time_kernels(dpos.data_ptr(), box, 3, True, False))
s = ctx.kernel.stats(); print("pairs", s.pairs_in_cutoff, "cand", s.pair_candidates, "cells", tuple(s.cells))
