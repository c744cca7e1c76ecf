import sys, time, numpy as np
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime, _abi
from oracle import Oracle

def relrms(a, b):
    return np.sqrt(((a-b)**2).sum()/ (b**2).sum())

def check(name, pos, box, f, pairs=True):
    o = Oracle(f, box)
    t=time.time(); eo, fo = o.execute(pos, box); to=time.time()-t
    ctx = runtime.CoulContext(f, box)
    t=time.time(); e, frc, comps = ctx.evaluate(pos); tg=time.time()-t
    t=time.time(); e, frc, comps = ctx.evaluate(pos); tg2=time.time()-t
    print(f"== {name}: N={len(pos)} ewald={ctx.kernel.ewald_params()} oracle {to:.2f}s gpu first {tg*1e3:.1f}ms second {tg2*1e3:.2f}ms")
    print("   oracle E", eo)
    print("   gpu    E", comps)
    print("   E rel err total", abs(e-eo[4])/abs(eo[4]), " comps", [abs(comps[k]-eo[k])/max(1e-30,abs(eo[4])) for k in range(4)])
    print("   F rel RMS", relrms(frc, fo), " max abs", np.abs(frc-fo).max(), " Frms", np.sqrt((fo**2).mean()))
    print("   q max err", np.abs(ctx.kernel.charges()-o.charges()).max(), " dedq rel", relrms(ctx.kernel.dedq(), o.dedq()))
    jo = o.jacobian(); jg = ctx.kernel.jacobian()
    if len(jo[0]): print("   jac idx eq", np.array_equal(jo[0], jg[0]) and np.array_equal(jo[1], jg[1]), " val max err", np.abs(jo[2]-jg[2]).max())
    if f.usesPeriodicBoundaryConditions() and pairs:
        po = o.neighbor_pairs(); pg = ctx.kernel.neighbor_pairs()
        print("   pairs oracle", len(po), "gpu", len(pg), "identical", np.array_equal(po, pg))
        s = ctx.kernel.stats(); print("   stats pairs", s.pairs_in_cutoff, "cand", s.pair_candidates, "launches", s.kernel_launches, "cells", tuple(s.cells))
    # flags
    for (incF, incE) in ((True, False), (False, True), (False, False)):
        eo2, fo2 = o.execute(pos, box, incF, incE)
        e2, f2, c2 = ctx.evaluate(pos, incF, incE)
        print(f"   flags F={incF} E={incE}: E err {abs(e2-eo2[4])/max(1e-30,abs(eo2[4])):.2e}  F relrms {relrms(f2, fo2) if (fo2**2).sum()>0 else np.abs(f2).max():.2e}")
    sys.stdout.flush()

which = sys.argv[1:] or ['c1','small','c5','c2']
print("fp32 peak", runtime.measure_fp32_peak())
if 'c1' in which:
    check("C1 64 waters nonPBC", *synthetic.config('c1'))
if 'small' in which:
    check("216 waters PBC rc=0.9", *synthetic.water_box(216, seed=1, ewald_tol=1e-4, cutoff=0.9))
    check("216 flux-water PBC", *synthetic.water_box(216, seed=2, ewald_tol=1e-5, cutoff=0.9, flux='water'))
    check("rock salt", *synthetic.rock_salt(cells=3))
if 'c5' in which:
    check("C5 methanol/water", *synthetic.config('c5'))
if 'c2' in which:
    check("C2 4k water", *synthetic.config('c2'))
if 'c3' in which:
    pos, box, f = synthetic.config('c3')
    ctx = runtime.CoulContext(f, box)
    for i in range(3):
        t=time.time(); e, frc, comps = ctx.evaluate(pos); print("C3 eval", (time.time()-t)*1e3, "ms", comps)
    import torch
    dpos = torch.tensor(pos, device='cuda')
    print("C3 device ms/eval", ctx.kernel.time_device(dpos.data_ptr(), box, 20))
    print("C3 kernels", ctx.kernel.time_kernels(dpos.data_ptr(), box, 5))
    s = ctx.kernel.stats(); print("   stats pairs", s.pairs_in_cutoff, "cand", s.pair_candidates, "launches", s.kernel_launches, "cells", tuple(s.cells))
