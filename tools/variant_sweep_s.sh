#!/bin/bash
# GPU box: time the FP32 structure-factor kernel (energy call) of each library variant at C3 and check the energy.
cp openmm_chargeflux_b200/libcfx_b200.so /tmp/libcfx_keep.so
for tag in "$@"; do
    cp tools/_variants/libcfx_$tag.so openmm_chargeflux_b200/libcfx_b200.so
    python - <<PY
import sys, numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, force = synthetic.config('c3')
ctx = runtime.CoulContext(force, box)
g = np.load('tests/golden/c3_fullk.npz')
e, f, comps = ctx.evaluate(pos, True, True)
dpos = torch.tensor(pos.reshape(-1), device='cuda')
kte = ctx.kernel.time_kernels(dpos.data_ptr(), box, 20, True, True)
mse = ctx.kernel.time_device(dpos.data_ptr(), box, 50, True, True)
print("$tag: E %.8f golden %.8f rel %.2e  step E+F %.4f ms  S %.4f gather %.4f" % (e, float(g['energy'][4]) if g['energy'].ndim else float(g['energy']), abs(e - float(np.ravel(g['energy'])[-1]))/abs(float(np.ravel(g['energy'])[-1])), mse, kte['structure_factor'], kte['kspace_gather']))
PY
done
cp /tmp/libcfx_keep.so openmm_chargeflux_b200/libcfx_b200.so
