"""The OpenMM plugin adapter (drop-in boundary) and the multi-GPU sharding logic.

CPU part: the plugin library exports the loader symbols; shard partition rule; a world_size-2 gloo run
of ShardedCoulContext with an oracle-backed stand-in for the CUDA kernel (each rank contributes its
k-slab, rank 0 the non-reciprocal terms) must reproduce the unsharded result after the all-reduce.
GPU part: the same unmodified CoulForce / CoulForceImpl objects evaluated by the reference's own
Reference-platform kernel and by the B200 kernel behind the plugin's KernelFactory boundary; and
sharded handles (emulated on one GPU) summing to the unsharded result."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import E_RTOL, F_RTOL, ROOT, rel_rms
from openmm_chargeflux_b200 import _abi, synthetic
from openmm_chargeflux_b200.parallel import FIXED_SCALE, ShardedCoulContext, shard_bounds
from oracle import Oracle, ReferenceBuild, reference_available

PLUGIN = os.path.join(ROOT, "openmm_chargeflux_b200", "plugin", "libOpenMMCoulB200.so")


def test_shard_bounds_partition_everything_once():
    for count in (0, 1, 7, 729, 4096):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(count, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.skipif(not os.path.exists(PLUGIN), reason="plugin adapter not built (needs the plugin's API headers)")
def test_plugin_exports_the_loader_contract():
    lib = ctypes.CDLL(PLUGIN)
    for sym in ("registerPlatforms", "registerKernelFactories", "registerCoulB200KernelFactories", "registerCoulSerializationProxies"):
        assert hasattr(lib, sym)


class _OracleShard:
    """Stand-in for the CUDA kernel in the gloo test: rank r evaluates its slab of nkx with the CPU
    oracle; rank 0 also owns the direct, self and exclusion terms (as the C library does)."""

    def __init__(self, force, box, rank, world):
        self.o = Oracle(force, box)
        self.n = force.getNumParticles()
        kx = self.o.ewald_params()[1][0]
        self.lo, self.hi = shard_bounds(kx, rank, world)
        self.rank = rank

    def padded_num_particles(self):
        return (self.n + 127) // 128 * 128

    def execute_device(self, d_pos, box, d_force, d_dedq, d_energy, stream, inc_f, inc_e):
        import torch
        npad = self.padded_num_particles()
        pos = np.ctypeslib.as_array(ctypes.cast(d_pos, ctypes.POINTER(ctypes.c_double)), shape=(3 * self.n,)).reshape(-1, 3)
        self.o.set_kx_range(self.lo, self.hi)
        e, f = self.o.execute(pos, box, inc_f, inc_e)
        if self.rank != 0:
            self.o.set_kx_range(0, 0)
            e0, f0 = self.o.execute(pos, box, inc_f, inc_e)
            e, f = e - e0, f - f0
        fx = np.ctypeslib.as_array(ctypes.cast(d_force, ctypes.POINTER(ctypes.c_int64)), shape=(3, npad))
        fx[:, :self.n] += np.rint(f.T * FIXED_SCALE).astype(np.int64)
        en = np.ctypeslib.as_array(ctypes.cast(d_energy, ctypes.POINTER(ctypes.c_double)), shape=(8,))
        en[:5] += e


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, box, force = synthetic.water_box(27, seed=5, cutoff=0.45, ewald_tol=1e-5)
    ctx = ShardedCoulContext(force, box, rank=rank, world=world, backend=_OracleShard(force, box, rank, world))
    e, f, comps = ctx.evaluate(pos)
    if rank == 0:
        np.savez(out, e=e, f=f, comps=comps)
    dist.destroy_process_group()


def test_sharded_context_allreduce_over_gloo(tmp_path, build_native):
    import torch.multiprocessing as mp
    out = str(tmp_path / "sharded.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    pos, box, force = synthetic.water_box(27, seed=5, cutoff=0.45, ewald_tol=1e-5)
    eo, fo = Oracle(force, box).execute(pos, box)
    assert abs(float(got["e"]) - eo[4]) <= 1e-9 * abs(eo[4])
    assert rel_rms(got["f"], fo) <= 1e-9
    assert np.abs(got["comps"][:4] - eo[:4]).max() <= 1e-9 * np.abs(eo[:4]).max()


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.skipif(not (reference_available() and os.path.exists(PLUGIN)), reason="needs oracle/_ref and the plugin adapter")
@pytest.mark.parametrize("case", ["water", "methanol", "nopbc"])
def test_plugin_is_a_drop_in_for_the_reference_kernel(case, build_native):
    if case == "water":
        pos, box, force = synthetic.water_box(216, seed=1, cutoff=0.9, ewald_tol=1e-4)
    elif case == "methanol":
        pos, box, force = synthetic.methanol_water(30, 90, seed=5, cutoff=0.8, ewald_tol=1e-5)
    else:
        pos, box, force = synthetic.config("c1")
    ref = ReferenceBuild(force, box, platform="Reference")
    b200 = ReferenceBuild(force, box, platform="B200", plugin=PLUGIN)
    base = np.random.default_rng(1).normal(size=pos.shape)
    for inc_f, inc_e in ((True, True), (True, False)):
        er, fr = ref.execute(pos, box, inc_f, inc_e, forces_in=base)
        eb, fb = b200.execute(pos, box, inc_f, inc_e, forces_in=base)
        if inc_e:
            assert abs(eb[4] - er[4]) <= E_RTOL * abs(er[4])
        assert rel_rms(fb - base, fr - base) <= F_RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_handles_sum_to_the_unsharded_result(world, build_native):
    import torch
    from openmm_chargeflux_b200 import runtime
    pos, box, force = synthetic.config("c2")
    n = len(pos)
    whole = runtime.CoulContext(force, box)
    e, f, comps = whole.evaluate(pos)
    d_pos = torch.tensor(pos.reshape(-1), device="cuda")
    kernels = [runtime.CalcCoulForceKernel(shard_rank=r, shard_count=world) for r in range(world)]
    for k in kernels:
        k.initialize(box, force)
    npad = kernels[0].padded_num_particles()
    d_force = torch.zeros(3 * npad, dtype=torch.int64, device="cuda")
    d_energy = torch.zeros(8, dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for k in kernels:                                   # ranks emulated as consecutive launches on one GPU
            k.execute_device(d_pos.data_ptr(), box, d_force.data_ptr(), 0, d_energy.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    fs = d_force.cpu().numpy().reshape(3, npad)[:, :n].T / FIXED_SCALE
    es = d_energy.cpu().numpy()
    # (a rank's FP32 structure-factor partial sums run over other atom splits than the unsharded handle's: the
    # reciprocal energy of the two agrees to FP32 summation noise, ~2e-8 of the total here, not to the last bit)
    assert abs(es[4] - e) <= 1e-7 * abs(e)
    assert np.abs(es[:4] - comps[:4]).max() <= 1e-8 * np.abs(comps[:4]).max()
    assert rel_rms(fs, f) <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [(True, True), (True, False)])
def test_execute_shard_fills_the_reduction_buffer_like_execute_device(flags, build_native):
    """cfx_execute_shard = zero + forces (2^32) + energies (2^24) in one int64 buffer: summing the buffers of the shard
    handles (what the all-reduce does) reproduces the unsharded evaluation."""
    import torch
    from openmm_chargeflux_b200 import runtime
    from openmm_chargeflux_b200.parallel import ENERGY_SCALE
    pos, box, force = synthetic.config("c2")
    n, world = len(pos), 3
    whole = runtime.CoulContext(force, box)
    e, f, comps = whole.evaluate(pos, *flags)
    d_pos = torch.tensor(pos.reshape(-1), device="cuda")
    kernels = [runtime.CalcCoulForceKernel(shard_rank=r, shard_count=world) for r in range(world)]
    for k in kernels:
        k.initialize(box, force)
    npad = kernels[0].padded_num_particles()
    stream = torch.cuda.Stream()
    total = torch.zeros(3 * npad + 8, dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        for k in kernels:
            d_reduce = torch.full((3 * npad + 8,), 12345, dtype=torch.int64, device="cuda")     # stale contents must be cleared
            for _ in range(2):                                                                    # second call replays the graph
                k.execute_shard(d_pos.data_ptr(), box, d_reduce.data_ptr(), stream.cuda_stream, *flags)
            total += d_reduce
    stream.synchronize()
    buf = total.cpu().numpy()
    fs = buf[:3 * npad].reshape(3, npad)[:, :n].T / FIXED_SCALE
    es = buf[3 * npad:3 * npad + 5] / ENERGY_SCALE
    # includeEnergy=False: the partial energy OpenMM discards is summed from FP32 pair terms, and a shard deals its
    # clusters' stencil columns over more CTAs than the whole evaluation does (different FP32 partial sums)
    assert abs(es[4] - e) <= (1e-7 if flags[1] else 1e-6) * max(abs(e), 1e-3 * np.abs(comps[:4]).max())
    assert abs(es[:4].sum() - es[4]) <= 1e-6
    assert rel_rms(fs, f) <= 2e-6


# ---------------------------------------------------------------------------------------------------------
# Multi-GPU inside the C ABI (comm.cu): single-process multi-device handle, NCCL owned by the library.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_device_handle_matches_the_single_gpu_evaluation(ndev, build_native):
    """cfx_multi_create(devices): one sharded handle + one NCCL rank per device in ONE process; cfx_multi_execute has
    the contract of cfx_execute (host positions in, energy + forces added to the caller's array out)."""
    import torch
    from openmm_chargeflux_b200 import runtime
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    pos, box, force = synthetic.config("c2")
    whole = runtime.CoulContext(force, box)
    multi = runtime.MultiGpuCoulKernel(box, force, list(range(ndev)))
    for flags in ((True, True), (True, False)):
        e, f, comps = whole.evaluate(pos, *flags)
        fm = np.full_like(pos, 0.5)
        cm = np.zeros(_abi.E_COUNT)
        em = multi.execute(pos, box, fm, *flags, components=cm)
        assert abs(em - e) <= (1e-7 if flags[1] else 1e-6) * max(abs(e), 1e-3 * np.abs(comps[:4]).max())
        assert rel_rms(fm - 0.5, f) <= 2e-6
    multi.close()


@pytest.mark.gpu
def test_comm_entry_points_fail_cleanly_without_a_communicator(build_native):
    import torch
    from openmm_chargeflux_b200 import runtime
    pos, box, force = synthetic.config("c2")
    k = runtime.CalcCoulForceKernel(shard_rank=0, shard_count=2)
    k.initialize(box, force)
    assert k.comm_size() == 0
    with pytest.raises(runtime.CfxError, match="communicator"):
        k.execute(pos, box, np.zeros_like(pos))
    d_pos = torch.tensor(pos.reshape(-1), device="cuda")
    d_red = torch.zeros(3 * k.padded_num_particles() + 8, dtype=torch.int64, device="cuda")
    with pytest.raises(runtime.CfxError, match="communicator"):
        k.execute_sharded(d_pos.data_ptr(), box, d_red.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert len(runtime.comm_unique_id()) == 128
