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
names = {0: "setup done", 1: "g0 begin", 2: "g0 flush done", 3: "g0 barrier1", 4: "g0 A loads+st issued", 5: "g0 A visible+arrive", 6: "g0 barrier2 (Ey ready)",
         20: "MMA first aFull", 21: "MMA second aFull", 8: "epi math done seq18", 9: "epi math done seq19", 19: "epi dFull seq19 seen", 7: "epi dFull seq20 seen", 10: "split u10 begin wait raw", 11: "split u10 raw arrived", 12: "split u10 op buffer free", 13: "split u10 loop done", 14: "split u10 fenced+arrived", 15: "split u10 group barrier",
         24: "MMA commit seq18", 26: "epi dFull seq18 seen", 27: "epi arrive dEmpty seq18", 22: "MMA begin wait dEmpty seq20", 23: "MMA dEmpty seq20 seen", 25: "MMA commit seq20",
         28: "MMA begin wait coefFull unit10", 29: "MMA coefFull unit10 seen", 30: "TMA begin wait empty for unit10", 31: "TMA issue load unit10",
         16: "epilogue loop done", 17: "final flush done", 18: "all done"}
for cta in (0, 73):
    print("CTA", cta)
    for k in sorted(names, key=lambda k: t[cta, k] if t[cta, k] else 1 << 62):
        if t[cta, k]:
            print(f"   {names[k]:28s} {(t[cta, k] - t0)/1000.0:8.2f} us")
