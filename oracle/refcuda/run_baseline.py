"""Time the plugin's EXISTING CUDA kernels (the "repo's existing CUDA platform" baseline of north_star) on
this GPU. BASELINE TOOLING, NOT PRODUCT.

The kernels are the reference's own (.cubin files built from /root/reference/platforms/cuda/src/kernels by
oracle/refcuda/Makefile, with an OpenMM-equivalent preamble). They are launched in the order and with the
launch geometry of the reference's host code (platforms/cuda/src/CudaCoulKernels.cpp:527-661: block 32 for the
two reciprocal-space kernels, OpenMM's default 64 elsewhere, grid capped at 4 blocks per SM as OpenMM's
executeKernel does), on the same synthetic box the new implementation is benchmarked on.

`computeNonbonded` needs OpenMM's tile neighbour list (CudaNonbondedUtilities), which does not exist outside
OpenMM, so it is NOT launched: the reported time is a LOWER bound on the existing platform's time
(8 of its 9 launches; the two reciprocal kernels are > 95 % of it).

    python oracle/refcuda/run_baseline.py [c3|c2] [iters]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "oracle", "_ref")


class Float4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


def available(cfg):
    return all(os.path.exists(os.path.join(OUT, "refcuda_%s_%s.cubin" % (cfg, k))) for k in ("pbc", "flux"))


def run(cfg="c3", iters=3, check=True):
    import torch
    from cuda.bindings import driver as cu
    from openmm_chargeflux_b200 import synthetic

    def ok(res):
        if res[0] != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError("CUDA driver error %s" % (res[0],))
        return res[1] if len(res) == 2 else res[1:]

    pos, box, f = synthetic.config(cfg)
    n = f.getNumParticles()
    padded = (n + 31) // 32 * 32
    dev = torch.device("cuda")
    torch.zeros(1, device=dev)                        # primary context
    mods = {}
    for k in ("pbc", "flux"):
        data = open(os.path.join(OUT, "refcuda_%s_%s.cubin" % (cfg, k)), "rb").read()
        mods[k] = ok(cu.cuModuleLoadData(data))
    fn = lambda m, name: ok(cu.cuModuleGetFunction(mods[m], name.encode()))

    L = np.diag(box)
    wrapped = pos - np.floor(pos / L) * L
    q0 = np.array([f.getParticleParameters(i)[0] for i in range(n)])
    lj = np.array(f._ljparams).reshape(n, 2)
    posq = torch.zeros(padded, 4, dtype=torch.float32, device=dev)
    posq[:n, :3] = torch.tensor(wrapped, dtype=torch.float32)
    params = torch.tensor(np.stack([q0, lj[:, 0] / 2, 2 * np.sqrt(lj[:, 1]), np.zeros(n)], 1), dtype=torch.float32, device=dev)
    atom_index = torch.arange(padded, dtype=torch.int32, device=dev)
    index_atom = torch.zeros(padded, dtype=torch.int32, device=dev)
    nb, na, nw = f.getNumFluxBonds(), f.getNumFluxAngles(), f.getNumFluxWaters()
    cf_idx = np.zeros((max(nb + na, 1), 4), np.int32)
    cf_par = np.zeros((max(nb + na, 1), 2), np.float32)
    cf_idx[:nb, :2] = np.array(f._fbond_idx).reshape(-1, 2)
    cf_par[:nb] = np.array(f._fbond_params).reshape(-1, 2)
    cf_idx[nb:nb + na, :3] = np.array(f._fangle_idx).reshape(-1, 3)
    cf_par[nb:nb + na] = np.array(f._fangle_params).reshape(-1, 2)
    wat_idx = np.zeros((max(nw, 1), 4), np.int32)
    wat_par = np.zeros(max(5 * nw, 1), np.float32)
    if nw:
        wat_idx[:, :3] = np.array(f._fwater_idx).reshape(-1, 3)
        wat_par[:] = np.array(f._fwater_params)
    dq_idx, dx_idx = [], []
    for t in range(nb):
        p = f._fbond_idx[2 * t:2 * t + 2]
        dq_idx += [p[0], p[0], p[1], p[1]]; dx_idx += [p[0], p[1], p[0], p[1]]
    for t in range(na):
        p = f._fangle_idx[3 * t:3 * t + 3]
        for a in range(3):
            for b in range(3):
                dq_idx.append(p[a]); dx_idx.append(p[b])
    for t in range(nw):
        p = f._fwater_idx[3 * t:3 * t + 3]
        for a in range(3):
            for b in range(3):
                dq_idx.append(p[a]); dx_idx.append(p[b])
    npairs = len(dq_idx)
    ex = np.array(f._exclusions, np.int32).reshape(-1, 2)
    t = lambda a, dt: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    d_cf_idx, d_cf_par = t(cf_idx, torch.int32), t(cf_par, torch.float32)
    d_wat_idx, d_wat_par = t(wat_idx, torch.int32), t(wat_par, torch.float32)
    d_dq_idx, d_dx_idx = t(np.array(dq_idx, np.int32), torch.int32), t(np.array(dx_idx, np.int32), torch.int32)
    d_ex0, d_ex1 = t(ex[:, 0], torch.int32), t(ex[:, 1], torch.int32)
    dedq = torch.zeros(n, dtype=torch.float32, device=dev)
    dqdx_val = torch.zeros(max(npairs, 1), 4, dtype=torch.float32, device=dev)
    force = torch.zeros(3 * padded, dtype=torch.int64, device=dev)
    energy = torch.zeros(1 << 17, dtype=torch.float64, device=dev)
    defs = dict(kv.split("=") for kv in __import__("subprocess").run(
        [sys.executable, os.path.join(HERE, "configs.py"), cfg], capture_output=True, text=True).stdout.replace("-D", "").split())
    totalk = int(defs["TOTALK"])
    cos_sin = torch.zeros(totalk, 2, dtype=torch.float32, device=dev)
    box4 = Float4(L[0], L[1], L[2], 0.0)
    inv4 = Float4(1 / L[0], 1 / L[1], 1 / L[2], 0.0)
    vx, vy, vz = Float4(L[0], 0, 0, 0), Float4(0, L[1], 0, 0), Float4(0, 0, L[2], 0)
    n_ex = C.c_int(len(ex))

    sm = torch.cuda.get_device_properties(0).multi_processor_count
    max_blocks = 4 * sm                                  # OpenMM: numThreadBlocksPerComputeUnit = 4

    def launch(func, args, work, block=64):
        keep = []
        for a in args:
            if isinstance(a, torch.Tensor):
                keep.append(C.c_void_p(a.data_ptr()))
            else:
                keep.append(a)
        ptrs = (C.c_void_p * len(keep))(*[C.cast(C.pointer(k), C.c_void_p) for k in keep])
        grid = max(1, min((work + block - 1) // block, max_blocks))
        ok(cu.cuLaunchKernel(func, grid, 1, 1, block, 1, 1, 0, torch.cuda.current_stream().cuda_stream, C.addressof(ptrs), 0))
        return keep, ptrs

    k_index, k_copy, k_real = fn("pbc", "genIndexAtom"), fn("flux", "copyCharge"), fn("flux", "calcRealCharge")
    k_self, k_rec_e, k_rec_f = fn("pbc", "computeEwaldSelfEner"), fn("pbc", "computeEwaldRecEner"), fn("pbc", "computeEwaldRecForce")
    k_excl, k_mult = fn("pbc", "computeExclusion"), fn("flux", "multdQdX")

    steps = [
        ("genIndexAtom", k_index, [atom_index, index_atom], n, 64),
        ("copyCharge", k_copy, [posq, dedq, dqdx_val, params, index_atom], n + npairs, 64),
        ("calcRealCharge", k_real, [dqdx_val, posq, d_cf_idx, d_cf_par, d_wat_idx, d_wat_par, index_atom, box4, inv4, vx, vy, vz],
         nb + na + nw, 64),
        ("computeEwaldSelfEner", k_self, [energy, dedq, posq, atom_index], n, 64),
        ("computeEwaldRecEner", k_rec_e, [energy, posq, atom_index, cos_sin, box4, inv4], totalk, 32),
        ("computeEwaldRecForce", k_rec_f, [force, dedq, posq, atom_index, cos_sin, box4, inv4], n, 32),
        ("computeExclusion", k_excl, [force, energy, dedq, posq, atom_index, index_atom, params, d_ex0, d_ex1, n_ex, box4, inv4, vx, vy, vz],
         len(ex), 64),
        ("multdQdX", k_mult, [force, dedq, index_atom, d_dq_idx, d_dx_idx, dqdx_val], npairs, 64),
    ]
    times = {name: [] for name, *_ in steps}
    total = []
    for it in range(iters + 1):
        force.zero_(); energy.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(steps) + 1)]
        evs[0].record()
        keep = []
        for i, (name, func, args, work, block) in enumerate(steps):
            keep.append(launch(func, args, work, block))
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it == 0:
            continue                                           # warm-up
        for i, (name, *_r) in enumerate(steps):
            times[name].append(evs[i].elapsed_time(evs[i + 1]))
        total.append(evs[0].elapsed_time(evs[-1]))
    out = {"config": cfg, "atoms": n, "totalk": totalk, "ms_per_eval_lower_bound": float(np.mean(total)),
           "evals_per_s_upper_bound": 1e3 / float(np.mean(total)),
           "kernels_ms": {k: float(np.mean(v)) for k, v in times.items()},
           "note": "computeNonbonded not launched (needs OpenMM's tile neighbour list): lower bound on the existing platform's time"}
    if check:
        # sanity: the reciprocal + self + exclusion energy the reference kernels produced
        out["energy_sum_kernels"] = float(energy.sum().item())
    return out


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    print(json.dumps(run(cfg, iters)))
