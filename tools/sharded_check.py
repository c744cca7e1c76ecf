"""torchrun --nproc-per-node N tools/sharded_check.py [c2|c3]: the NCCL-sharded evaluation against the single-GPU one."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
from openmm_chargeflux_b200.parallel import ShardedCoulContext
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.pop("NCCL_DEBUG", None)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, box, force = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else "c2")
ctx = ShardedCoulContext(force, box, rank=rank, world=world, device=local)
for flags in ((True, True), (True, False)):
    e, f, comps = ctx.evaluate(pos, *flags)
    if rank == 0:
        one = runtime.CoulContext(force, box)
        e1, f1, c1 = one.evaluate(pos, *flags)
        rel = np.sqrt(((f - f1) ** 2).sum() / (f1 ** 2).sum())
        print("world", world, "flags", flags, "E", e, "single", e1, "dE/|E|", abs(e - e1) / abs(e1), "F rel RMS", rel, "comps", comps - c1)
        assert rel < 5e-6 and abs(e - e1) <= 2e-6 * max(abs(e1), 1e-3 * np.abs(c1[:4]).max())
dist.barrier()
dist.destroy_process_group()
