import os, sys
print("rank", os.environ.get("RANK"), "NCCL_DEBUG before:", os.environ.get("NCCL_DEBUG"), "FILE:", os.environ.get("NCCL_DEBUG_FILE"), file=sys.stderr, flush=True)
mode = sys.argv[1]
os.environ["NCCL_DEBUG"] = "INFO"
if mode == "stderr":
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
elif mode == "fd2":
    os.environ["NCCL_DEBUG_FILE"] = "/proc/self/fd/2"
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(4, device="cuda"); dist.all_reduce(x); torch.cuda.synchronize()
print("rank", os.environ.get("RANK"), "sum", x[0].item(), flush=True)
dist.destroy_process_group()
