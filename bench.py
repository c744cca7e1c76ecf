#!/usr/bin/env python
"""bench.py -- charge-flux Ewald force evaluations per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--impl reference]

One "step" = one CalcCoulForceKernel::execute(includeForces=true, includeEnergy=true), BASELINE.md's definition of a
force evaluation: charge-flux assembly, Ewald direct + explicit-k reciprocal + self + excluded-pair correction and the
dE/dq.dq/dx chain rule, energy and forces, on the synthetic flexible-water box named in `config.workload`. Both arms
(`--impl reference` too) use these flags. The forces-only call OpenMM makes once per MD time step (includeEnergy=false) is
measured the same way and reported at top level as `value_forces_only` / `ms_per_step_forces_only`; `ns_per_day` derives
from it.

  value     whole-job force-evals/s with positions resident in HBM, CUDA events around every step,
            L2 flushed between steps, max over ranks.
  e2e       the same through the reference-facing call with HOST buffers: pinned H2D of the positions
            and D2H of forces + energy inside the timed region.
  roofline  the longest kernel launch of the step: algorithmic FLOP / its CUDA-event duration against the peak of the unit
            it runs on, measured live on this GPU (MEASURED_PEAKS.json has neither a CUDA-core nor a TF32 / INT8 figure).
            `roofline.kernels` lists every large kernel with its own bound: pair passes against the FP32 FMA peak, the
            gather (tcgen05 kind::tf32, three-product split) against the TF32 peak, the structure factors (tcgen05 kind::i8
            on digit planes of fixed-point operands) against the INT8 peak -- each as algorithmic FLOP and as executed
            tensor operations.
  cpu_baseline  the plugin's Reference-platform kernel (oracle/_ref when present, else the oracle port)
            on one host core, bounded sample, extrapolated in the number of k-vectors.

`--impl reference` times that CPU implementation alone: one full evaluation at the real kmax, no extrapolation (the
reference platform is single-threaded; ~90 s at C3).
N > 1 (launched by torchrun): k-vectors and direct-space i-clusters sharded over the ranks, one NCCL all-reduce of the
fixed-point reduction buffer per step, issued by the library inside the step's CUDA graph (cfx_comm_init /
cfx_execute_sharded); total work fixed => "scaling": "strong". The line adds `parity_vs_single` (sharded against the same
evaluation on one GPU) and a `c4` block: the 262k-atom box at N GPUs and on one.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "charge-flux Ewald force-evals/sec (ns/day), 32k-atom water, 1/2/4/8 B200"
UNIT = "force-evals/s"
WORKLOADS = {
    "c2": "c2: 4,095-atom periodic flexible-water box, cutoff 1.0 nm, Ewald tol 1e-4, bond+angle charge flux",
    "c3": "c3: 32,766-atom periodic flexible-water box, cutoff 1.0 nm, Ewald tol 1e-5, bond+angle charge flux",
    "c4": "c4: 262,143-atom periodic flexible-water box, cutoff 1.0 nm, Ewald tol 1e-5, bond+angle charge flux",
}
TIMESTEP_FS = 0.5
N_FRAMES = 33                   # positions advance one frame per step (synthetic.ballistic_frames), walked back and forth
INCLUDE_ENERGY = True           # one force eval = one full execute, energy + forces
FLAGS_NOTE = "includeForces=1 includeEnergy=1 (BASELINE.md: one force eval = energy + forces)"


def ns_per_day(evals_per_s):
    # evals/s x fs/step x 1e-6 ns/fs x 86400 s/day (x 0.0432 at 0.5 fs)
    return evals_per_s * TIMESTEP_FS * 86.4e-3


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                mhz, mx = float(parts[1].split()[0]), float(parts[2].split()[0])
            except ValueError:
                continue
            smax.append(mx)
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(mhz)
                for name, val in zip(names, parts[5:9]):
                    if val == "Active":
                        reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[1].split()[0]) for r in self.rows[-3:]] if self.rows else [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU reference (oracle/ is only ever the thing MEASURED here in the cpu_baseline / --impl reference legs)
# --------------------------------------------------------------------------------------------------
def _cpu_factory():
    from oracle import Oracle, ReferenceBuild, oracle_available, reference_available
    if reference_available():
        return "reference", (lambda force, default_box: ReferenceBuild(force, default_box))
    if not oracle_available():
        import __graft_entry__ as g
        g.build_oracle()
    return "port", (lambda force, default_box: Oracle(force, default_box))


def _kcount(kmax):
    kx, ky, kz = kmax
    return (kz - 1) + (ky - 1) * (2 * kz - 1) + (kx - 1) * (2 * ky - 1) * (2 * kz - 1)


def _kmax_rule(length, alpha, tol):
    """The reference's kmax rule (ReferenceCoulKernels.cpp:32-35,403-420), used only to pick the default box."""
    k = 1
    while 0.05 * np.sqrt(length * alpha) * k * np.exp(-(k * np.pi / (length * alpha)) ** 2) > tol:
        k += 1
    return k + 1 if k % 2 == 0 else k


def cpu_reference_sample(pos, box, force, target_kmax, make):
    """Time one execute of the CPU reference whose k lattice is a sub-block (kmax <= target_kmax, fixed
    through the default box exactly as the reference's own rule does) of the full one.
    Returns (seconds, K_sample, kmax_sample, K_full)."""
    tol, rc = force.getEwaldErrorTolerance(), force.getCutoffDistance()
    alpha = np.sqrt(-np.log(2 * tol)) / rc
    nfull = _kcount([_kmax_rule(box[d, d], alpha, tol) for d in range(3)])
    div = 1.0
    while max(_kmax_rule(box[d, d] / div, alpha, tol) for d in range(3)) > target_kmax:
        div *= 1.05
    h = make(force, box / div)
    k = h.ewald_params()[1]
    t = time.perf_counter()
    h.execute(pos, box, True, INCLUDE_ENERGY)
    dt = time.perf_counter() - t
    return dt, _kcount(k), k, nfull


def cpu_baseline(pos, box, force, budget_s=25.0):
    kind, make = _cpu_factory()
    t_nonk, k0, _, nfull = cpu_reference_sample(pos, box, force, 1, make)          # kmax=(1,1,1): no k-vectors
    n = len(pos)
    # pick the sample so the k part takes ~budget: ~2 loops x (cos+sin) per (atom,k), ~45 ns each on one core
    per_ak = 9.0e-8
    target = 3
    for km in (5, 7, 9, 11, 13):
        ks = ((2 * km - 1) ** 3 - 1) // 2
        if ks * n * per_ak <= max(budget_s - 2 * t_nonk, 2.0):
            target = km
    t_s, ks, kmax_s, nfull = cpu_reference_sample(pos, box, force, target, make)
    t_full = t_nonk + max(t_s - t_nonk, 0.0) * nfull / max(ks, 1)
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "one execute with kmax=%s (%d of %d half-space k-vectors, fixed via the default box as the reference's own "
                      "rule does) = %.2f s, and one with no k-vectors = %.2f s (direct+self+exclusion+chain rule); "
                      "k part scaled by %d/%d; single thread (the reference platform has no threading)"
                      % (tuple(kmax_s), ks, nfull, t_s, t_nonk, nfull, ks),
            "seconds_per_eval_extrapolated": t_full}


def md_leg(args):
    """NVE MD of the same box: bonded forces + charge-flux Ewald forces + velocity Verlet, one CUDA graph per step."""
    try:
        from openmm_chargeflux_b200 import md, synthetic
        cfgs = synthetic.CONFIGS[args.workload]
        sim, pos = md.flexible_water_simulation(cfgs["n_waters"], cfgs["seed"], cutoff=cfgs["cutoff"], ewald_tol=cfgs["ewald_tol"])
        sim.minimize(200, 0.002)
        p, _ = sim.get_state()
        sim.set_state(p, sim.maxwell_boltzmann(300.0, seed=7))
        dt = TIMESTEP_FS * 1e-3
        sim.step(50, dt)
        e0 = sim.energies()
        ms = sim.step(args.md_steps, dt) / args.md_steps
        e1 = sim.energies()
        sim.close()
        return {"steps": args.md_steps, "dt_fs": TIMESTEP_FS, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "ns_per_day": ns_per_day(1e3 / ms),
                "energy_drift_over_kinetic": (e1["total"] - e0["total"]) / e0["kinetic"],
                "note": "forces-only evaluations; profiles/ holds the 10,000-step run"}
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def existing_cuda_baseline(workload, our_ms):
    """The reference's own platforms/cuda kernels (prebuilt cubins in oracle/_ref, see oracle/refcuda/) timed here."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "refcuda"))
        import run_baseline
        if not run_baseline.available(workload):
            return {"unavailable": "oracle/_ref/refcuda_%s_*.cubin not built (needs /root/reference at build time)" % workload}
        r = run_baseline.run(workload, iters=3)
        out = {"ms_per_eval": r["ms_per_eval"], "value": r["evals_per_s"], "unit": UNIT,
               "kernels_ms": {k: round(v, 4) for k, v in r["kernels_ms"].items()},
               "speedup": r["ms_per_eval"] / our_ms,
               "nonbonded_tiles": r.get("nonbonded_tiles"),
               "note": "all nine launches of the existing platform, reference launch geometry, mixed precision; computeNonbonded runs "
                       "on a tile neighbour list built by the harness (OpenMM's own builder is absent; it is not part of a step). "
                       "The existing platform ignores includeForces/includeEnergy (CudaCoulKernels.cpp:522-660 launches every kernel "
                       "on every call)"}
        if "check_direct_plus_exclusion" in r:
            out["check_direct_plus_exclusion"] = r["check_direct_plus_exclusion"]
        return out
    except Exception as e:                                   # a baseline must never break the bench line
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference_arm(args, pos, box, force, workload):
    """The reference's own CPU implementation (oracle/_ref: the plugin's unmodified platforms/reference sources; else the
    oracle port), ONE full, un-extrapolated evaluation at the real kmax, timed once: a C3 evaluation takes ~90 s on one
    core and the reference platform has no threading, so K + W repetitions would not fit a bench run. `steps` reports
    what was actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, make = _cpu_factory()
    h = make(force, box)
    alpha, kmax, nk = h.ewald_params()
    t = time.perf_counter()
    h.execute(pos, box, True, INCLUDE_ENERGY)
    t_full = time.perf_counter() - t
    value = 1.0 / t_full
    n = len(pos)
    sample = ("one full execute(includeForces=1, includeEnergy=1) at the real kmax=%s (%d half-space k-vectors), timed once, no "
              "extrapolation; requested --steps %d --warmup %d not repeated (%.0f s per evaluation); single thread: the reference "
              "platform has no threading" % (tuple(kmax), nk, args.steps, args.warmup, t_full))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 * t_full, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "atoms": n, "kvectors": int(nk), "flags": FLAGS_NOTE},
            "ns_per_day": ns_per_day(value),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# the CUDA path
# --------------------------------------------------------------------------------------------------
def algorithmic_flops(n_atoms, n_k, pairs, n_terms, n_rows, n_excl, energy):
    """FMA = 2 FLOP. Of the 80 FLOP of an in-cutoff pair 12 are its energy terms (E_coul 2, LJ energy 10):
    an energy+forces call runs them as a second, half-shell FP64 pass (`direct_pairs_energy`)."""
    return {"structure_factor": 4.0 * n_atoms * n_k, "kspace_gather": 8.0 * n_atoms * n_k,
            "direct_pairs": (68.0 if energy else 80.0) * pairs, "direct_pairs_energy": 12.0 * pairs,
            "total": 12.0 * n_atoms * n_k + 80.0 * pairs + 150.0 * n_terms + 6.0 * n_rows + 40.0 * n_excl + 4.0 * n_atoms}


def executed_tensor_ops(n_atoms, kmax, energy):
    """Operations the tensor-core k-space kernels execute per launch (padded tiles, split operands):
    name -> (ops, kind). Gather: tcgen05 kind::tf32, three products. Structure factors: tcgen05 kind::i8 on digit planes --
    per row tile and 32-atom stage the MMAs span 8 NN columns (3 digits, forces-only call) or 10 NN (4 digits, energy call)."""
    kx, ky, kz = kmax
    npad = (n_atoms + 255) // 256 * 256
    out = {}
    kp = (2 * kz + 7) // 8 * 8
    if kp <= 112 and ky >= 8:
        nt = 128 if kp <= 56 else 64
        signed = ky + (kx - 1) * (2 * ky - 1)
        cols = (signed + nt // 4 - 1) // (nt // 4) * nt
        out["kspace_gather"] = (3 * 2.0 * npad * kp * cols, "tf32")
    if kz <= 64 and os.environ.get("CFX_KSPACE_S", "") not in ("tf32", "fp32"):
        nn = 64 if kz <= 32 else 128
        rows_per_cta = 32 * (128 // nn)
        rows = (kx * ky + rows_per_cta - 1) // rows_per_cta * rows_per_cta
        out["structure_factor"] = (2.0 * 128 * (10 if energy else 8) * nn * npad * (rows // 32), "i8")
    return out


class Rank:
    """One rank of the bench: a (possibly sharded) kernel handle, device buffers, and the step it times."""

    def __init__(self, torch, dist, force, box, pos, rank, world, local, comm_id=None):
        from openmm_chargeflux_b200 import runtime
        self.torch, self.dist, self.world, self.box = torch, dist, world, box
        self.kernel = runtime.CalcCoulForceKernel(device=local, shard_rank=rank, shard_count=world, pin_caller_buffers=True)
        self.kernel.initialize(box, force)
        if world > 1:
            self.kernel.comm_init(comm_id)
        self.n, self.npad = len(pos), self.kernel.padded_num_particles()
        from openmm_chargeflux_b200 import synthetic
        self.frames_host = synthetic.ballistic_frames(pos, N_FRAMES, dt_ps=TIMESTEP_FS * 1e-3)
        self.d_frames = torch.tensor(self.frames_host.reshape(N_FRAMES, -1), dtype=torch.float64, device="cuda")
        self.d_pos = self.d_frames[0].clone()
        self.step_no = 0
        self.d_buf = torch.zeros(3 * self.npad + 8, dtype=torch.int64, device="cuda")
        self.stream = torch.cuda.Stream()

    def advance(self):
        """The integrator's part, outside the timed region: the atoms move to the next frame (device-to-device copy)."""
        from openmm_chargeflux_b200 import synthetic
        self.step_no += 1
        with self.torch.cuda.stream(self.stream):
            self.d_pos.copy_(self.d_frames[synthetic.ping_pong(self.step_no, N_FRAMES)])

    def rewind(self):
        self.step_no = 0
        with self.torch.cuda.stream(self.stream):
            self.d_pos.copy_(self.d_frames[0])

    def step(self, include_energy):
        """One evaluation on device-resident positions: the shard's CUDA graph, and for world > 1 the all-reduce inside it."""
        fn = self.kernel.execute_sharded if self.world > 1 else self.kernel.execute_shard
        fn(self.d_pos.data_ptr(), self.box, self.d_buf.data_ptr(), self.stream.cuda_stream, True, include_energy)

    def forces(self):
        self.stream.synchronize()
        return self.d_buf[:3 * self.npad].view(3, self.npad)[:, :self.n].t().to(self.torch.float64).cpu().numpy() / 4294967296.0

    def energies(self):
        self.stream.synchronize()
        return self.d_buf[3 * self.npad:3 * self.npad + 5].cpu().numpy().astype(np.float64) / 16777216.0


def broadcast_comm_id(dist, rank):
    from openmm_chargeflux_b200 import runtime
    box_ = [runtime.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box_, src=0)
    return box_[0]


def run_ours(args, pos, box, force, workload):
    import torch
    import torch.distributed as dist
    from openmm_chargeflux_b200 import runtime, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's INFO log (rank count, transports, NVLS) stays visible to whoever runs this: it is sent to stderr, so that
        # stdout holds the one JSON line
        # (the GPU boxes preset NCCL_DEBUG=VERSION: raise it to INFO unless the caller asked for more)
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = len(pos)
    comm_id = broadcast_comm_id(dist, rank) if world > 1 else None
    me = Rank(torch, dist, force, box, pos, rank, world, local, comm_id)
    torch.cuda.synchronize()
    flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(r, include_energy, steps):
        for _ in range(max(args.warmup, 3)):
            r.advance()
            r.step(include_energy)
        barrier()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        for i in range(steps):
            r.advance()
            with torch.cuda.stream(r.stream):
                flush.zero_()
                ev0[i].record(r.stream)
            r.step(include_energy)
            ev1[i].record(r.stream)
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    ms_f = timed(me, False, args.steps)                      # forces only: the per-MD-step call, reported beside the headline
    sampler = ClockSampler(local) if rank == 0 else None
    t_start = time.perf_counter()
    builds0 = me.kernel.stats().pair_list_builds
    ms_per_step = timed(me, INCLUDE_ENERGY, args.steps)      # headline: energy + forces
    st_ = me.kernel.stats()
    launches_per_eval, list_builds = st_.kernel_launches, st_.pair_list_builds - builds0
    me.rewind()
    me.step(INCLUDE_ENERGY)
    f_sharded, e_sharded = me.forces(), me.energies()

    # end to end through the reference-facing call: host positions in, host energy + forces out (every rank passes the same
    # positions and receives the whole result; sharded handles all-reduce inside the library call)
    # the caller's persistent arrays (pin_caller_buffers: they must outlive the handle), positions updated in place
    forces_host = np.zeros_like(pos)
    pos_host = pos.copy()

    def e2e(include_energy):
        times = []
        for i in range(args.steps + 3):
            pos_host[:] = me.frames_host[synthetic.ping_pong(i, N_FRAMES)]     # the integrator's part, outside the timed call
            flush.zero_()
            barrier()
            t = time.perf_counter()
            forces_host[:] = 0.0
            me.kernel.execute(pos_host, box, forces_host, True, include_energy)
            dt = time.perf_counter() - t
            if i >= 3:
                times.append(dt)
        tt = torch.tensor([float(np.mean(times))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_s = e2e(INCLUDE_ENERGY)
    clocks = sampler.stop(t_start, time.perf_counter()) if sampler else None     # timed loop + e2e loop, both under load
    e2e_f_s = e2e(False)
    # the host-buffer call returns what the device-resident call does (same positions: frame 0)
    pos_host[:] = me.frames_host[0]
    forces_host[:] = 0.0
    e_host = me.kernel.execute(pos_host, box, forces_host, True, INCLUDE_ENERGY)
    e2e_check = {"forces_rel_rms_vs_device_call": float(np.sqrt(((forces_host - f_sharded) ** 2).sum() / (f_sharded ** 2).sum())),
                 "energy_rel_vs_device_call": float(abs(e_host - e_sharded[4]) / abs(e_sharded[4]))}

    alpha, kmax, nk = me.kernel.ewald_params()
    line = None
    if rank == 0:
        value = 1e3 / ms_per_step
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 pair terms and tf32x3 gather, s8 digit planes / s32 sums for the structure factors (f64 energies and cross-CTA sums, int64 fixed-point forces)", "data": "synthetic",
                "config": {"workload": workload, "atoms": n, "kmax": list(kmax), "kvectors": int(nk), "alpha": alpha,
                           "flags": FLAGS_NOTE,
                           "positions": "the atoms move every step: thermal (300 K) straight-line motion, %.1f fs per step, %d frames "
                                        "walked back and forth (harder on the candidate lists than MD); lists built for cutoff + 0.1 nm "
                                        "and rebuilt when an atom has moved 0.05 nm: %d builds (re-sort + list) in the %d timed "
                                        "steps, inside the timed region" % (TIMESTEP_FS, N_FRAMES, list_builds, args.steps),
                           "l2": "384 MiB buffer written between timed steps (L2 flush)"
                                 + ("; working set < L2" if n < 100000 else "; phase tables (0.5 GB) stream from HBM"),
                           "parallelism": "k-vector rows + direct-space i-clusters sharded x%d, one NCCL all-reduce of the int64 "
                                          "reduction buffer inside the library's CUDA graph" % world if world > 1 else "single GPU"},
                "value_forces_only": 1e3 / ms_f, "ms_per_step_forces_only": ms_f,
                "forces_only_note": "includeForces=1 includeEnergy=0: the call OpenMM makes once per MD step (tcgen05 structure "
                                    "factors, FP32 pair terms); ns_per_day is derived from it",
                "ns_per_day": ns_per_day(1e3 / ms_f),
                "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 24 * n + 40,
                        "forces_only": {"value": 1.0 / e2e_f_s, "ns_per_day": ns_per_day(1.0 / e2e_f_s)},
                        "check": e2e_check},
                "gpu_launches": int(launches_per_eval) * args.steps,
                "clocks": clocks}

    if world > 1:
        # parity of the sharded evaluation against the same evaluation on ONE GPU (rank 0's), and the 256k-atom box
        par = {}
        if rank == 0:
            single = Rank(torch, dist, force, box, pos, 0, 1, local)
            single.step(INCLUDE_ENERGY)
            f1, e1 = single.forces(), single.energies()
            par = {"forces_rel_rms": float(np.sqrt(((f_sharded - f1) ** 2).sum() / (f1 ** 2).sum())),
                   "energy_rel": float(abs(e_sharded[4] - e1[4]) / abs(e1[4])), "energy_sharded": float(e_sharded[4]), "energy_single": float(e1[4])}
            single.kernel.close()
        barrier()
        c4 = None
        if not args.no_c4:
            pos4, box4, force4 = synthetic.config("c4")
            r4 = Rank(torch, dist, force4, box4, pos4, rank, world, local, broadcast_comm_id(dist, rank))
            ms4_f, ms4_ef = timed(r4, False, 10), timed(r4, True, 10)
            r4.rewind()
            r4.step(False)
            f4 = r4.forces()
            r4.kernel.close()
            if rank == 0:
                s4 = Rank(torch, dist, force4, box4, pos4, 0, 1, local)
                one_f = one_ef = None
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for k, inc_e in enumerate((False, True)):
                    for _ in range(3):
                        s4.advance()
                        s4.step(inc_e)
                    ev[2 * k].record(s4.stream)
                    for _ in range(5):
                        s4.advance()                         # (a 6 MB device copy inside the timed span: negligible at this size)
                        s4.step(inc_e)
                    ev[2 * k + 1].record(s4.stream)
                torch.cuda.synchronize()
                one_f, one_ef = ev[0].elapsed_time(ev[1]) / 5, ev[2].elapsed_time(ev[3]) / 5
                s4.rewind()
                s4.step(False)
                g = s4.forces()
                c4 = {"workload": WORKLOADS["c4"], "atoms": len(pos4), "n_gpus": world,
                      "ms_per_step_forces_only": ms4_f, "ms_per_step": ms4_ef,
                      "ms_per_step_forces_only_1gpu": one_f, "ms_per_step_1gpu": one_ef,
                      "speedup_forces_only": one_f / ms4_f, "speedup": one_ef / ms4_ef,
                      "parity_vs_single_forces_rel_rms": float(np.sqrt(((f4 - g) ** 2).sum() / (g ** 2).sum())),
                      "note": "1-GPU figures measured in this run on rank 0's GPU (no L2 flush at this size: the tables exceed L2)"}
                s4.kernel.close()
            barrier()
        if rank == 0:
            line["parity_vs_single"] = par
            line["c4"] = c4
            line["comm"] = {"nranks": me.kernel.comm_size(), "backend": "NCCL (dlopen'ed by libcfx_b200.so), int64 sum all-reduce of %d bytes"
                            % (8 * (3 * me.npad + 8))}

    if world == 1:
        # per-kernel durations (CUDA events on the launching stream) and the roofline of the dominant one
        tf_peak, _ = runtime.measure_fp32_peak(local, 5)
        tf32_peak = runtime.measure_tf32_peak(local)
        i8_peak = runtime.measure_i8_peak(local)
        kt = me.kernel.time_kernels(me.d_pos.data_ptr(), box, 10, True, INCLUDE_ENERGY)
        kt_f = me.kernel.time_kernels(me.d_pos.data_ptr(), box, 10, True, False)
        pairs = me.kernel.stats().pairs_in_cutoff
        terms = (force.getNumFluxBonds() + force.getNumFluxAngles() + force.getNumFluxWaters(),
                 4 * force.getNumFluxBonds() + 9 * force.getNumFluxAngles() + 9 * force.getNumFluxWaters(), force.getNumExceptions())
        fl = algorithmic_flops(n, nk, pairs, *terms, INCLUDE_ENERGY)
        try:
            traffic_db = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"))) if args.workload == "c3" else {}
        except (OSError, ValueError):
            traffic_db = {}

        def traffic_of(name):
            tr = traffic_db.get(name)
            return tr["dram_read_bytes"] + tr["dram_write_bytes"] if tr else None

        def kernel_rows(times, energy):
            ex = executed_tensor_ops(n, kmax, energy)
            flk = algorithmic_flops(n, nk, pairs, *terms, energy)
            total = sum(times.values())
            rows = []
            for name in ("direct_pairs", "direct_pairs_energy", "kspace_gather", "structure_factor"):
                if name not in times:
                    continue
                kind = ex[name][1] if name in ex else None
                peak, unit = {"tf32": (tf32_peak, "TFLOP/s"), "i8": (i8_peak, "TOP/s"), None: (tf_peak, "TFLOP/s")}[kind]
                a = flk[name] / (times[name] * 1e-3) / 1e12
                row = {"kernel": name, "bound": "tensor" if kind else "fp32", "ms": times[name], "share_of_step": times[name] / total,
                       "algorithmic_flop_per_launch": flk[name], "achieved": a, "peak": peak, "unit": unit, "frac": a / peak,
                       "traffic": traffic_of(name + ("@E" if energy else "@F"))}
                if kind:
                    row["tensor_kind"] = "tcgen05.mma kind::" + kind
                    row["executed_ops_per_launch"] = ex[name][0]
                    row["executed_tops"] = ex[name][0] / (times[name] * 1e-3) / 1e12
                    row["executed_frac"] = row["executed_tops"] / peak
                    row["fp32_equivalent_frac_of_fp32_peak"] = a / tf_peak
                if name == "direct_pairs_energy":
                    row["note"] = "half-shell FP64 pass (34 FP64 instructions per pair); fraction quoted against the FP32 peak as SURVEY.md 8d defines"
                rows.append(row)
            return rows

        rows = kernel_rows(kt, INCLUDE_ENERGY)
        top = max(rows, key=lambda r: r["ms"])
        line["roofline"] = {"bound": top["bound"], "kernel": top["kernel"], "achieved": top["achieved"], "peak": top["peak"],
                            "unit": top["unit"], "frac": top["frac"], "traffic": top["traffic"],
                            "traffic_note": "DRAM bytes per launch of this kernel from the committed ncu capture (profiles/); far below "
                                            "the algorithmic FLOP x 4 B: the kernel is instruction-issue bound, not HBM bound",
                            "peak_source": "NOT in MEASURED_PEAKS.json (it holds HBM and bf16 figures only): FP32 FMA microbenchmark run "
                                           "in this process (cfx_measure_fp32_peak; theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4) "
                                           "and tcgen05 kind::tf32 128x128x8 / kind::i8 128x256x32 microbenchmarks run in this process "
                                           "(cfx_measure_tf32_peak, cfx_measure_i8_peak)",
                            "kernel_ms": top["ms"], "kernel_share_of_step": top["share_of_step"],
                            "algorithmic_flop_per_launch": top["algorithmic_flop_per_launch"],
                            "kernels": rows,
                            "kernels_forces_only": kernel_rows(kt_f, False),
                            "whole_step": {"achieved": fl["total"] / (ms_per_step * 1e-3) / 1e12,
                                           "frac": fl["total"] / (ms_per_step * 1e-3) / 1e12 / tf_peak,
                                           "frac_forces_only": fl["total"] / (ms_f * 1e-3) / 1e12 / tf_peak,
                                           "algorithmic_flop": fl["total"],
                                           "note": "algorithmic FP32-equivalent FLOP of the whole evaluation / step time, against the "
                                                   "FP32 FMA peak; both reciprocal-space sums run on tensor cores (gather: kind::tf32 x 3, structure factors: kind::i8 digit planes)"}}
        # what one list build costs (a handle that rebuilds at every evaluation): re-sort + list kernel
        rb = runtime.CalcCoulForceKernel(device=local, list_skin=0.0)
        rb.initialize(box, force)
        kt_rb = rb.time_kernels(me.d_pos.data_ptr(), box, 10, True, False)
        rb.close()
        line["pair_list"] = {"skin_nm": 0.1, "builds_in_timed_steps": int(list_builds), "timed_steps": args.steps,
                             "build_ms": round(kt_rb.get("cell_build", 0.0) + kt_rb.get("pair_list", 0.0), 5),
                             "note": "build = re-sort into cells + candidate-list kernel, run inside the step's graph on the steps "
                                     "whose displacement check fires; the other steps launch the same kernels, which return at once"}
        line["kernels_ms"] = {k: round(v, 5) for k, v in kt.items()}
        line["kernels_ms_forces_only"] = {k: round(v, 5) for k, v in kt_f.items()}
        line["peaks"] = {"fp32_fma_tflops": tf_peak, "tf32_tcgen05_tflops": tf32_peak, "i8_tcgen05_tops": i8_peak}
        line["config"]["pairs_in_cutoff"] = int(pairs)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pos, box, force)
        if args.md_steps > 0:
            me.kernel.close()
            line["md_nve"] = md_leg(args)
        # second reported baseline of north_star: the plugin's EXISTING CUDA kernels on this GPU
        line["existing_cuda_baseline"] = existing_cuda_baseline(args.workload, ms_per_step)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--md-steps", type=int, default=2000, help="length of the NVE MD leg (0 = skip)")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the 262k-atom strong-scaling block")
    args = ap.parse_args()
    from openmm_chargeflux_b200 import synthetic
    pos, box, force = synthetic.config(args.workload)
    if args.impl == "reference":
        run_reference_arm(args, pos, box, force, WORKLOADS[args.workload])
    else:
        run_ours(args, pos, box, force, WORKLOADS[args.workload])


if __name__ == "__main__":
    main()
