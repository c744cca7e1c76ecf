import sys
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, force = synthetic.config('c3')
th = np.load('gpurun_out/thermal_c3.npy')
for label, p in (("lattice", pos), ("thermal", th)):
    d = torch.tensor(p.reshape(-1), device='cuda')
    for skip in (False, True):
        for inc_e in (False, True):
            k = runtime.CalcCoulForceKernel(skip_discarded_energy=skip)
            k.initialize(box, force)
            kt = k.time_kernels(d.data_ptr(), box, 20, True, inc_e)
            ms = k.time_device(d.data_ptr(), box, 50, True, inc_e)
            st = k.stats()
            print("%s skip=%d energy=%d: step %.4f ms pairs-kernel %.4f cell %.4f excl %.4f | pairs %d tests %d" % (label, skip, inc_e, ms, kt['direct_pairs'], kt['cell_build'], kt['exclusion_corr'], st.pairs_in_cutoff, st.pair_candidates))
            k.close()
