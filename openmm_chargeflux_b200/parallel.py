"""Multi-GPU evaluation: one process per GPU, k-vectors and direct-space i-tiles sharded, one
all-reduce of the fixed-point forces.

Every rank holds the whole system (positions, parameters: O(N), replicated) and a kernel handle
created with (shard_rank, shard_count): it computes S(k) for its block of (nx,|ny|) rows over ALL
atoms, scatters that block's forces/dE/dq to all atoms, evaluates the i-tiles of its spatial slab,
and applies the chain rule to its PARTIAL dE/dq (the chain rule is linear in dE/dq). Rank 0 adds the
self and excluded-pair terms. The partial results are int64 fixed point, so the sum over ranks is
exact and independent of the reduction order (bitwise reproducible for a given GPU count; different GPU
counts group the FP32 partial sums differently and agree to FP32 rounding, ~1e-6 relative).
The collective is ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np

FIXED_SCALE = 4294967296.0
ENERGY_SCALE = 16777216.0


def shard_bounds(count, rank, world):
    """Contiguous block [lo, hi) of `count` work units owned by `rank` (same rule as the C library)."""
    return (count * rank) // world, (count * (rank + 1)) // world


class ShardedCoulContext:
    """Evaluate one CoulForce on `world` GPUs. `backend` is an object with ``padded_num_particles()``
    and ``execute_device(...)`` (the CUDA kernel), injectable for the CPU (gloo) tests."""

    def __init__(self, force, box, rank=None, world=None, device=None, backend=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.box = np.asarray(box, dtype=np.float64)
        self.n = force.getNumParticles()
        if backend is None:
            from . import runtime
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            backend = runtime.CalcCoulForceKernel(device=self.device.index, shard_rank=self.rank, shard_count=self.world)
            backend.initialize(self.box, force)
        else:
            self.device = torch.device("cpu")
        self.kernel = backend
        self.npad = backend.padded_num_particles()
        self.d_pos = torch.zeros(3 * self.n, dtype=torch.float64, device=self.device)
        # one buffer = forces [3][Npad] followed by the four energy components (fixed point 2^24 as
        # int64 would lose nothing, but energies are kept as doubles in a second tiny buffer)
        # one reduction buffer: forces [3][Npad] followed by 8 slots for the energies as 2^24 fixed point,
        # so a step needs a single all-reduce
        self.d_buf = torch.zeros(3 * self.npad + 8, dtype=torch.int64, device=self.device)
        self.d_force = self.d_buf[:3 * self.npad]
        self.d_energy = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.h_pos = torch.zeros(3 * self.n, dtype=torch.float64).pin_memory() if self.device.type == "cuda" else None
        if self.device.type == "cuda":
            self._out_dev = torch.zeros(3 * self.n + 8, dtype=torch.float64, device=self.device)
            self._out_host = torch.zeros(3 * self.n + 8, dtype=torch.float64).pin_memory()
        # a dedicated (capturable) stream: the step is replayed as one CUDA graph on it
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def evaluate_device(self, include_forces=True, include_energy=True):
        """Positions already in ``self.d_pos``. Leaves the reduced fixed-point results in ``d_buf``
        (forces [3][Npad] at 2^32, then the energy components at 2^24): one library call (one CUDA graph that
        also zeroes the buffer) and, for world > 1, one all-reduce."""
        with self._on_stream():
            stream = self.stream.cuda_stream if self.stream is not None else 0
            if hasattr(self.kernel, "execute_shard"):
                self.kernel.execute_shard(self.d_pos.data_ptr(), self.box, self.d_buf.data_ptr(), stream, include_forces, include_energy)
            else:                                              # injected CPU backend of the gloo tests
                t = self.torch
                self.d_force.zero_()
                self.d_energy.zero_()
                self.kernel.execute_device(self.d_pos.data_ptr(), self.box, self.d_force.data_ptr(), 0, self.d_energy.data_ptr(),
                                           stream, include_forces, include_energy)
                self.d_buf[3 * self.npad:] = t.round(self.d_energy * ENERGY_SCALE).to(t.int64)
            if self.world > 1:
                self.dist.all_reduce(self.d_buf)

    def energies(self):
        """The five energy components [self, recip, direct, excl, total] of the last evaluation (host array)."""
        if self.stream is not None:
            self.stream.synchronize()
        return self.d_buf[3 * self.npad:3 * self.npad + 5].cpu().numpy().astype(np.float64) / ENERGY_SCALE

    def evaluate(self, positions, include_forces=True, include_energy=True):
        """Host positions in, (energy, forces[N,3], components[5]) out -- the end-to-end call."""
        torch = self.torch
        pos = torch.from_numpy(np.ascontiguousarray(positions, dtype=np.float64).reshape(-1))
        with self._on_stream():
            if self.h_pos is not None:
                self.h_pos.copy_(pos)
                self.d_pos.copy_(self.h_pos, non_blocking=True)
            else:
                self.d_pos.copy_(pos)
        self.evaluate_device(include_forces, include_energy)
        if self.stream is None:                                # CPU (gloo) test backend
            buf = self.d_buf.numpy()
            f = buf[:3 * self.npad].reshape(3, self.npad)[:, :self.n].T.astype(np.float64) / FIXED_SCALE
            e = buf[3 * self.npad:3 * self.npad + 5].astype(np.float64) / ENERGY_SCALE
            return float(e[4]), np.ascontiguousarray(f), e
        # fixed point -> double and [3][Npad] -> [N][3] on the GPU, one pinned D2H of forces + energies
        with self._on_stream():
            out = self._out_dev
            out[:3 * self.n] = (self.d_force.view(3, self.npad)[:, :self.n].to(torch.float64) * (1.0 / FIXED_SCALE)).t().reshape(-1)
            out[3 * self.n:] = self.d_buf[3 * self.npad:].to(torch.float64) * (1.0 / ENERGY_SCALE)
            self._out_host.copy_(out, non_blocking=True)
        self.stream.synchronize()
        host = self._out_host.numpy()
        e = host[3 * self.n:3 * self.n + 5].copy()
        return float(e[4]), host[:3 * self.n].reshape(self.n, 3).copy(), e
