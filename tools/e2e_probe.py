"""GPU box: where the end-to-end (host buffers) call spends its time beyond the device step at C3.
   python tools/e2e_probe.py [steps]      (CFX_HOST_COPY_KERNELS=0: copy nodes instead of load/store-through kernels)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from openmm_chargeflux_b200 import synthetic, runtime
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
pos, box, force = synthetic.config('c3')
frames = synthetic.ballistic_frames(pos, 33, dt_ps=0.5e-3)
flush = torch.empty(384 << 20, dtype=torch.uint8, device='cuda')

def wall(fn, pre=None):
    ts = []
    for i in range(steps + 3):
        if pre: pre(i)
        flush.zero_(); torch.cuda.synchronize()
        t = time.perf_counter(); fn(); dt = time.perf_counter() - t
        if i >= 3: ts.append(dt)
    return 1e3*float(np.mean(ts)), 1e3*float(np.min(ts))

ref_f = None
for mode, pin in (("0", True), ("1", True), ("0", False), ("1", False)):
    os.environ["CFX_HOST_COPY_KERNELS"] = mode
    k = runtime.CalcCoulForceKernel(pin_caller_buffers=pin)
    k.initialize(box, force)
    ph, fh = pos.copy(), np.zeros_like(pos)
    def move(i): ph[:] = frames[synthetic.ping_pong(i, 33)]
    for inc_e in (True, False):
        def full(): fh[:] = 0.0; k.execute(ph, box, fh, True, inc_e)
        def nozero(): k.execute(ph, box, fh, True, inc_e)
        def noforce(): k.execute(ph, box, None, True, inc_e)
        print("kernels=%s pin=%d includeEnergy=%d  zero+execute %.4f (min %.4f)  execute %.4f (min %.4f)  execute without force array %.4f (min %.4f) ms"
              % ((mode, pin, inc_e) + wall(full, move) + wall(nozero, move) + wall(noforce, move)), flush=True)
    # reference check of the in-place update
    fh[:] = 1.0; e = k.execute(ph, box, fh, True, True)
    if ref_f is None: ref_f = fh.copy()
    print("  sum check: mean(f) %.6f (expect ~1.0: forces sum to zero)  E %.8f  max |f - first variant| %.3e" % (fh.mean(), e, np.abs(fh - ref_f).max()))
    k.close()

# device-resident pieces
k = runtime.CalcCoulForceKernel()
k.initialize(box, force)
dpos = torch.tensor(pos.reshape(-1), device='cuda')
npad = k.padded_num_particles()
dbuf = torch.zeros(3*npad + 8, dtype=torch.int64, device='cuda')
st = torch.cuda.Stream()
for inc_e in (True, False):
    def dev(): k.execute_shard(dpos.data_ptr(), box, dbuf.data_ptr(), st.cuda_stream, True, inc_e); st.synchronize()
    print("device-resident call + sync (host wall) includeEnergy=%d  %.4f (min %.4f) ms;  device step by events %.4f ms"
          % ((inc_e,) + wall(dev) + (k.time_device(dpos.data_ptr(), box, 30, True, inc_e),)), flush=True)
z = np.zeros_like(pos)
t = time.perf_counter()
for _ in range(100): z[:] = 0.0
print("numpy zero of the force array %.4f ms" % (10*(time.perf_counter() - t)))
hp = torch.empty(pos.size, dtype=torch.float64).pin_memory(); dp = torch.empty(pos.size, dtype=torch.float64, device='cuda')
def h2d(): dp.copy_(hp, non_blocking=True); torch.cuda.synchronize()
def d2h(): hp.copy_(dp, non_blocking=True); torch.cuda.synchronize()
print("pinned H2D of one vector %.4f ms, D2H %.4f ms (host wall, with sync)" % (wall(h2d)[0], wall(d2h)[0]))
