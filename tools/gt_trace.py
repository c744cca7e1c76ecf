"""Phase timeline of the tensor-core gather kernel (CFX_GT_TRACE=1): per-CTA globaltimer stamps."""
import os, sys, ctypes, numpy as np, torch
os.environ["CFX_GT_TRACE"] = "1"
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
pos, box, f = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'c3')
ctx = runtime.CoulContext(f, box)
dpos = torch.tensor(pos, device='cuda')
for _ in range(3):
    ms = ctx.kernel.time_device(dpos.data_ptr(), box, 1, True, False)
lib = runtime.load_library()
buf = np.zeros((148, 32), dtype=np.uint64)
lib.cfx_debug_gather_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
rc = lib.cfx_debug_gather_trace(ctx.kernel._h, buf.ctypes.data)
print("rc", rc)
t = buf.astype(np.int64)
t0 = t[:, 0].min()
names = {0: "setup done (barriers, TMEM alloc)",
         1: "epi: group 0 begin", 2: "epi: g0 flush done", 3: "epi: g0 barrier 1", 4: "epi: g0 phase operand loads + tcgen05.st issued",
         5: "epi: g0 operand visible, arrive", 6: "epi: g0 barrier 2 (Ey columns ready)", 20: "MMA: first phase operand seen",
         8: "epi: group 1 begin", 9: "epi: g1 flush done", 10: "epi: g1 barrier 1", 11: "epi: g1 operand stores issued",
         12: "epi: g1 operand visible, arrive", 13: "epi: g1 barrier 2", 21: "MMA: second phase operand seen",
         24: "MMA: commit tile seq 18", 26: "epi: accumulator seq 18 seen", 27: "epi: slot of seq 18 released", 14: "epi: math of seq 18 done",
         19: "epi: accumulator seq 19 seen", 15: "epi: math of seq 19 done",
         28: "MMA: wait operand planes of unit 10", 29: "MMA: operand planes of unit 10 seen",
         22: "MMA: wait slot for seq 20", 23: "MMA: slot for seq 20 free", 25: "MMA: commit tile seq 20", 7: "epi: accumulator seq 20 seen",
         16: "epi: unit loop done", 17: "epi: final flush done", 18: "all roles done"}
for cta in (0, 73):
    print("CTA", cta)
    for k in sorted(names, key=lambda k: t[cta, k] if t[cta, k] else 1 << 62):
        if t[cta, k]:
            print(f"   {names[k]:28s} {(t[cta, k] - t0)/1000.0:8.2f} us")
