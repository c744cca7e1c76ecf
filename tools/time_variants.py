import sys, numpy as np, torch
sys.path.insert(0,'.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, f = synthetic.config(sys.argv[1] if len(sys.argv)>1 else 'c3')
ctx = runtime.CoulContext(f, box)
dpos = torch.tensor(pos, device='cuda')
for incF, incE in ((True, True), (True, False), (False, True)):
    ms = ctx.kernel.time_device(dpos.data_ptr(), box, 20, incF, incE)
    print(f"forces={incF} energy={incE}: {ms:.4f} ms/eval")
print('E+F kernels', ctx.kernel.time_kernels(dpos.data_ptr(), box, 5))
print('F-only kernels', ctx.kernel.time_kernels(dpos.data_ptr(), box, 5, True, False))
