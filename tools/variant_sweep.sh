#!/bin/bash
# GPU box: time the pair kernel of each prebuilt library variant (tools/_variants/libcfx_<tag>.so) at C3.
# usage: tools/variant_sweep.sh tag1 tag2 ...   (restores the in-tree library afterwards)
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
ref = g['forces_f32'].astype(np.float64)
rr = float(np.sqrt(((f - ref) ** 2).sum() / (ref ** 2).sum()))
dpos = torch.tensor(pos.reshape(-1), device='cuda')
kt = ctx.kernel.time_kernels(dpos.data_ptr(), box, 20, True, False)
kte = ctx.kernel.time_kernels(dpos.data_ptr(), box, 20, True, True)
kto = ctx.kernel.time_kernels(dpos.data_ptr(), box, 20, False, True)
ms = ctx.kernel.time_device(dpos.data_ptr(), box, 50, True, False)
mse = ctx.kernel.time_device(dpos.data_ptr(), box, 50, True, True)
print("$tag: F rel-RMS vs golden %.2e  E %.8f (%s)  pairs %d  step F %.4f E+F %.4f ms  pairs: F %.4f  E+F %.4f  E-only %.4f" % (
    rr, e, [k for k in g.files if 'ener' in k.lower()][:2], ctx.kernel.stats().pairs_in_cutoff, ms, mse, kt['direct_pairs'], kte['direct_pairs'], kto['direct_pairs']))
PY
done
cp /tmp/libcfx_keep.so openmm_chargeflux_b200/libcfx_b200.so
