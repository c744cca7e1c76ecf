"""Time the plugin's EXISTING CUDA kernels (the "repo's existing CUDA platform" baseline of north_star) on
this GPU. BASELINE TOOLING, NOT PRODUCT.

The kernels are the reference's own (.cubin files built from /root/reference/platforms/cuda/src/kernels by
oracle/refcuda/Makefile, with an OpenMM-equivalent preamble). They are launched in the order and with the
launch geometry of the reference's host code (platforms/cuda/src/CudaCoulKernels.cpp:527-661: block 32 for the
two reciprocal-space kernels, OpenMM's default 64 elsewhere, grid capped at 4 blocks per SM as OpenMM's
executeKernel does), on the same synthetic box the new implementation is benchmarked on.

`computeNonbonded` reads OpenMM's tile neighbour list (CudaNonbondedUtilities). OpenMM does not exist here, so the harness
builds that list itself (`tile_lists`, numpy, outside the timed region -- OpenMM rebuilds it on the GPU only when atoms
have moved, so its cost is not part of a step either): atoms sorted into compact blocks of 32, block bounding boxes,
the diagonal tiles as "exclusion tiles" (the plugin compiles the kernel WITHOUT USE_EXCLUSIONS,
CudaCoulKernels.cpp:480: every pair of a tile is evaluated, excluded pairs are subtracted by computeExclusion), and for
every block x the atoms of blocks y > x within the cutoff of x's bounding box, 32 per tile entry. All nine launches
of the existing platform are then timed.

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


def spatial_order(wrapped, L, cell=0.7):
    """Molecules (3 consecutive atoms) sorted by the cell of their first atom, z fastest: compact blocks of 32 atoms."""
    n = len(wrapped)
    nc = np.maximum(1, np.floor(L / cell).astype(int))
    first = wrapped[0::3][: (n + 2) // 3]
    c = np.minimum((first / L * nc).astype(int), nc - 1)
    key = (c[:, 0] * nc[1] + c[:, 1]) * nc[2] + c[:, 2]
    mol = np.argsort(key, kind="stable")
    order = (3 * mol[:, None] + np.arange(3)[None, :]).reshape(-1)
    return order[order < n].astype(np.int32)


def tile_lists(p, L, rc, padded):
    """OpenMM's tile neighbour list for atoms p (platform order): per block of 32 its bounding box (centre, half size),
    and tile entries (x, 32 atoms of blocks y > x within rc of x's box), CudaNonbondedUtilities findBlockBounds /
    findBlocksWithInteractions restated in numpy."""
    n = len(p)
    nb = padded // 32
    centre = np.zeros((nb, 4), np.float32)
    size = np.zeros((nb, 4), np.float32)
    for b in range(nb):
        a = p[32 * b:min(32 * b + 32, n)]
        if len(a) == 0:
            continue
        rel = a - a[0]
        rel -= np.round(rel / L) * L                       # one periodic copy, around the block's first atom
        q = a[0] + rel
        lo, hi = q.min(0), q.max(0)
        centre[b, :3] = 0.5 * (lo + hi)
        size[b, :3] = 0.5 * (hi - lo)
    c64, s64 = centre[:, :3].astype(np.float64), size[:, :3].astype(np.float64)
    tiles, inter = [], []
    pad = np.full(32, padded, np.int64)
    nreal = (n + 31) // 32
    for x in range(nreal):
        ys = np.arange(x + 1, nreal)
        if len(ys) == 0:
            break
        dc = c64[ys] - c64[x]
        dc -= np.round(dc / L) * L
        gap = np.maximum(0.0, np.abs(dc) - s64[ys] - s64[x])
        near = ys[(gap ** 2).sum(1) < rc * rc]
        if len(near) == 0:
            continue
        atoms = (32 * near[:, None] + np.arange(32)[None, :]).reshape(-1)
        atoms = atoms[atoms < n]
        d = p[atoms] - c64[x]
        d -= np.round(d / L) * L
        g = np.maximum(0.0, np.abs(d) - s64[x])
        atoms = atoms[(g ** 2).sum(1) < rc * rc]
        for k in range(0, len(atoms), 32):
            chunk = pad.copy()
            chunk[:len(atoms[k:k + 32])] = atoms[k:k + 32]
            tiles.append(x)
            inter.append(chunk)
    return {"tiles": np.array(tiles, np.int32), "interacting_atoms": np.array(inter, np.int64).reshape(-1),
            "block_center": centre, "block_size": size}


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
    rc = f.getCutoffDistance()
    # the platform's atom order: molecules (3 consecutive atoms) sorted into compact spatial blocks, as OpenMM's
    # CudaContext::reorderAtoms does (there along a Hilbert curve; here by 0.7 nm cells, z fastest)
    order = spatial_order(wrapped, L)
    tl = tile_lists(wrapped[order], L, rc, padded)
    q0 = np.array([f.getParticleParameters(i)[0] for i in range(n)])
    lj = np.array(f._ljparams).reshape(n, 2)
    posq = torch.zeros(padded, 4, dtype=torch.float32, device=dev)
    posq[:n, :3] = torch.tensor(wrapped[order], dtype=torch.float32)
    params = torch.tensor(np.stack([q0, lj[:, 0] / 2, 2 * np.sqrt(lj[:, 1]), np.zeros(n)], 1), dtype=torch.float32, device=dev)
    ai = np.arange(padded, dtype=np.int32)
    ai[:n] = order                                         # atomIndex[platform slot] = user index (padding: identity)
    atom_index = torch.tensor(ai, dtype=torch.int32, device=dev)
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

    k_nb = None

    def launch(func, args, work, block=64):
        keep = []
        for a in args:
            if isinstance(a, torch.Tensor):
                keep.append(C.c_void_p(a.data_ptr()))
            else:
                keep.append(a)
        ptrs = (C.c_void_p * len(keep))(*[C.cast(C.pointer(k), C.c_void_p) for k in keep])
        grid = max(1, min((work + block - 1) // block, max_blocks))
        if func is k_nb:
            grid = max_blocks                                # CudaNonbondedUtilities: numForceThreadBlocks = 4 per SM
        ok(cu.cuLaunchKernel(func, grid, 1, 1, block, 1, 1, 0, torch.cuda.current_stream().cuda_stream, C.addressof(ptrs), 0))
        return keep, ptrs

    k_index, k_copy, k_real = fn("pbc", "genIndexAtom"), fn("flux", "copyCharge"), fn("flux", "calcRealCharge")
    k_self, k_rec_e, k_rec_f = fn("pbc", "computeEwaldSelfEner"), fn("pbc", "computeEwaldRecEner"), fn("pbc", "computeEwaldRecForce")
    k_excl, k_mult = fn("pbc", "computeExclusion"), fn("flux", "multdQdX")
    k_nb = fn("pbc", "computeNonbonded")          # (assigned before the first launch: the closure above sees it)
    nblocks = padded // 32
    d_tiles = t(tl["tiles"], torch.int32)
    d_inter = t(tl["interacting_atoms"].astype(np.int64), torch.int64).to(torch.int32)       # uint32 values < 2^31
    d_count = t(np.array([len(tl["tiles"]), 0], np.int64), torch.int64).to(torch.int32)
    d_center = t(tl["block_center"], torch.float32)
    d_size = t(tl["block_size"], torch.float32)
    d_excl_tiles = t(np.stack([np.arange(nblocks), np.arange(nblocks)], 1).astype(np.int32), torch.int32)
    d_dummy = torch.zeros(64, dtype=torch.int32, device=dev)
    dedq_nb = torch.zeros(padded, dtype=torch.float32, device=dev)

    steps = [
        ("genIndexAtom", k_index, [atom_index, index_atom], n, 64),
        ("copyCharge", k_copy, [posq, dedq, dqdx_val, params, index_atom], n + npairs, 64),
        ("calcRealCharge", k_real, [dqdx_val, posq, d_cf_idx, d_cf_par, d_wat_idx, d_wat_par, index_atom, box4, inv4, vx, vy, vz],
         nb + na + nw, 64),
        ("computeEwaldSelfEner", k_self, [energy, dedq, posq, atom_index], n, 64),
        ("computeEwaldRecEner", k_rec_e, [energy, posq, atom_index, cos_sin, box4, inv4], totalk, 32),
        ("computeEwaldRecForce", k_rec_f, [force, dedq, posq, atom_index, cos_sin, box4, inv4], n, 32),
        ("computeNonbonded", k_nb, [force, energy, dedq_nb, posq, atom_index, params, d_dummy, d_excl_tiles, C.c_uint(0), C.c_ulonglong(0),
                                    d_tiles, d_count, box4, inv4, vx, vy, vz, C.c_uint(len(tl["tiles"]) + 1), d_center, d_size, d_inter,
                                    C.c_uint(0), d_dummy], 1 << 30, 64),
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
    out = {"config": cfg, "atoms": n, "totalk": totalk, "ms_per_eval": float(np.mean(total)),
           "evals_per_s": 1e3 / float(np.mean(total)),
           "kernels_ms": {k: float(np.mean(v)) for k, v in times.items()},
           "nonbonded_tiles": int(len(tl["tiles"])), "nonbonded_pairs_evaluated": int(len(tl["tiles"])) * 1024 + nblocks * 1024,
           "note": "all nine launches of the existing platform; computeNonbonded runs on a harness-built tile list (OpenMM's "
                   "own list builder is absent and is not part of a step: it only runs when atoms have moved)"}
    if check:
        # sanity: the energy the reference kernels produced, and the tile list: computeNonbonded + computeExclusion alone
        # must reproduce the new implementation's direct + excluded-pair energy (same charges, same pairs)
        out["energy_sum_kernels"] = float(energy.sum().item())
        energy.zero_()
        names = [s_[0] for s_ in steps]
        keep = [launch(*steps[names.index(k)][1:]) for k in ("computeNonbonded", "computeExclusion")]
        torch.cuda.synchronize()
        e_pairs = float(energy.sum().item())
        from openmm_chargeflux_b200 import runtime
        ctx = runtime.CoulContext(f, box)
        _, _, comps = ctx.evaluate(pos)
        ctx.kernel.close()
        out["check_direct_plus_exclusion"] = {"existing_kernels": e_pairs, "new_implementation": float(comps[2] + comps[3]),
                                              "rel": abs(e_pairs - float(comps[2] + comps[3])) / abs(float(comps[2] + comps[3]))}
    return out


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    print(json.dumps(run(cfg, iters)))
