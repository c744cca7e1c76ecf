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
from conftest import golden_case
pos, box, force = synthetic.config('c3')
ctx = runtime.CoulContext(force, box)
e, f, comps = ctx.evaluate(pos, True, False)
g = np.load('tests/golden/c3_fullk.npz')
ref = g['forces_f32'].astype(np.float64)
rr = float(np.sqrt(((f - ref) ** 2).sum() / (ref ** 2).sum()))
dpos = torch.tensor(pos.reshape(-1), device='cuda')
kt = ctx.kernel.time_kernels(dpos.data_ptr(), box, 20, True, False)
kte = ctx.kernel.time_kernels(dpos.data_ptr(), box, 10, True, True)
ms = ctx.kernel.time_device(dpos.data_ptr(), box, 50, True, False)
print("$tag: F rel-RMS vs golden %.2e  pairs %d  step %.4f ms  cell_build %.4f direct_pairs %.4f (energy call %.4f)  sum %.4f" % (
    rr, ctx.kernel.stats().pairs_in_cutoff, ms, kt['cell_build'], kt['direct_pairs'], kte['direct_pairs'], sum(kt.values())))
PY
done
cp /tmp/libcfx_keep.so openmm_chargeflux_b200/libcfx_b200.so
