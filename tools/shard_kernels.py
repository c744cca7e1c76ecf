"""Per-kernel times of ONE shard (rank r of world w) of the C3 evaluation on one GPU: what a rank of the multi-GPU run executes."""
import sys, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pos, box, f = synthetic.config(sys.argv[2] if len(sys.argv) > 2 else 'c3')
dpos = torch.tensor(pos, device='cuda')
for rank in (0, world - 1):
    k = runtime.CalcCoulForceKernel(shard_rank=rank, shard_count=world)
    k.initialize(box, f)
    for inc_e in (False, True):
        kt = k.time_kernels(dpos.data_ptr(), box, 10, True, inc_e)
        print("rank", rank, "of", world, "energy", inc_e, "sum %.4f ms" % sum(kt.values()), {a: round(b, 4) for a, b in kt.items()})
        print("   device time per eval (graph): %.4f ms" % k.time_device(dpos.data_ptr(), box, 20, True, inc_e))
