#!/usr/bin/env python
"""bench.py -- charge-flux Ewald force evaluations per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--impl reference]

One "step" = one CalcCoulForceKernel::execute(includeForces=true, includeEnergy=false) -- the call OpenMM makes
once per MD time step: charge-flux assembly, Ewald direct + explicit-k reciprocal + self + excluded-pair
correction and the dE/dq.dq/dx chain rule, on the synthetic flexible-water box named in `config.workload`.
The energy+forces call (what a reporter / minimiser makes) is measured the same way and reported beside it as
`energy_and_forces`; both arms (`--impl reference` too) use the same flags.

  value     whole-job force-evals/s with positions resident in HBM, CUDA events around every step,
            L2 flushed between steps, max over ranks.
  e2e       the same through the reference-facing call with HOST buffers: pinned H2D of the positions
            and D2H of forces + energy inside the timed region.
  roofline  dominant kernel (the direct-space pair kernel): algorithmic FLOP / its CUDA-event duration, against
            the FP32 FMA peak measured live on this GPU (MEASURED_PEAKS.json has no CUDA-core figure).
            `roofline.kernels` lists every large kernel with its own bound: the reciprocal-space kernels run on
            the tensor cores (tcgen05 kind::tf32, three-product split) and are held against the TF32 peak measured
            live, both as algorithmic FP32-equivalent FLOP and as executed TF32 FLOP.
  cpu_baseline  the plugin's Reference-platform kernel (oracle/_ref when present, else the oracle port)
            on one host core, bounded sample, extrapolated in the number of k-vectors.

`--impl reference` times that CPU implementation alone (the reference platform is single-threaded).
N > 1 (launched by torchrun): k-vectors and direct-space i-tiles sharded over the ranks, one NCCL
all-reduce of the fixed-point forces per step; total work fixed => "scaling": "strong".
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
INCLUDE_ENERGY = False          # the headline step is the MD-step call: forces only


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
        return {"ms_per_eval_lower_bound": r["ms_per_eval_lower_bound"], "value_upper_bound": r["evals_per_s_upper_bound"], "unit": UNIT,
                "kernels_ms": {k: round(v, 4) for k, v in r["kernels_ms"].items()},
                "speedup_vs_lower_bound": r["ms_per_eval_lower_bound"] / our_ms,
                "note": "8 of the existing platform's 9 launches, reference launch geometry, mixed precision; computeNonbonded needs "
                        "OpenMM's tile list and is not launched, so this is a lower bound on its time. The existing platform "
                        "ignores includeForces/includeEnergy (CudaCoulKernels.cpp:522-660 launches every kernel on every call)"}
    except Exception as e:                                   # a baseline must never break the bench line
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference_arm(args, pos, box, force, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, make = _cpu_factory()
    total_calls = args.steps + args.warmup
    t_nonk, _, _, nfull = cpu_reference_sample(pos, box, force, 1, make)
    budget = max(150.0 / max(total_calls, 1) - t_nonk, 0.3)
    n = len(pos)
    target = 3
    for km in (5, 7, 9, 11):
        if (((2 * km - 1) ** 3 - 1) // 2) * n * 9.0e-8 <= budget:
            target = km
    times, ks_used, kmax_used = [], None, None
    for it in range(total_calls):
        t_s, ks, kmax_s, nfull = cpu_reference_sample(pos, box, force, target, make)
        if it >= args.warmup:
            times.append(t_nonk + max(t_s - t_nonk, 0.0) * nfull / max(ks, 1))
        ks_used, kmax_used = ks, kmax_s
    t_full = float(np.mean(times))
    value = 1.0 / t_full
    sample = ("each step: one execute of the CPU reference with kmax=%s (%d of %d k-vectors), k part scaled by %d/%d, "
              "non-k part %.2f s measured once; single thread" % (tuple(kmax_used), ks_used, nfull, nfull, ks_used, t_nonk))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_full, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "atoms": n, "kvectors": int(nfull), "flags": "includeForces=1 includeEnergy=0"},
            "ns_per_day": ns_per_day(value),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# the CUDA path
# --------------------------------------------------------------------------------------------------
def algorithmic_flops(n_atoms, n_k, pairs, n_terms, n_rows, n_excl):
    """FMA = 2 FLOP."""
    return {"structure_factor": 4.0 * n_atoms * n_k, "kspace_gather": 8.0 * n_atoms * n_k, "direct_pairs": 80.0 * pairs,
            "total": 12.0 * n_atoms * n_k + 80.0 * pairs + 150.0 * n_terms + 6.0 * n_rows + 40.0 * n_excl + 4.0 * n_atoms}


def executed_tensor_flops(n_atoms, kmax):
    """TF32 FLOP the tensor-core k-space kernels execute per launch: three products, padded tiles."""
    kx, ky, kz = kmax
    npad = (n_atoms + 255) // 256 * 256
    out = {}
    kp = (2 * kz + 7) // 8 * 8
    if kp <= 112 and ky >= 8:
        nt = 128 if kp <= 56 else 64
        signed = ky + (kx - 1) * (2 * ky - 1)
        cols = (signed + nt // 4 - 1) // (nt // 4) * nt
        out["kspace_gather"] = 3 * 2.0 * npad * kp * cols
    if kz <= 64:
        nn = 64 if kz <= 32 else 128
        rows_per_cta = 32 * (128 // nn)
        rows = (kx * ky + rows_per_cta - 1) // rows_per_cta * rows_per_cta
        out["structure_factor"] = 3 * 2.0 * (4 * rows) * nn * npad
    return out


def run_ours(args, pos, box, force, workload):
    import torch
    import torch.distributed as dist
    from openmm_chargeflux_b200 import runtime
    from openmm_chargeflux_b200.parallel import ShardedCoulContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # rank 0 must print exactly one JSON line: NCCL's banner / debug output (NCCL_DEBUG >= VERSION) goes to a file
        if "CFX_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["CFX_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/cfx_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = len(pos)
    ctx = ShardedCoulContext(force, box, rank=rank, world=world, device=local)
    ctx.d_pos.copy_(torch.from_numpy(pos.reshape(-1)))
    torch.cuda.synchronize()
    flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(include_energy):
        for _ in range(max(args.warmup, 3)):
            ctx.evaluate_device(True, include_energy)
        barrier()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        for i in range(args.steps):
            with torch.cuda.stream(ctx.stream):
                flush.zero_()
                ev0[i].record(ctx.stream)
            ctx.evaluate_device(True, include_energy)
            ev1[i].record(ctx.stream)
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / args.steps

    ms_ef = timed(True)                                       # energy + forces, reported beside the headline
    sampler = ClockSampler(local) if rank == 0 else None
    t_start = time.perf_counter()
    ms_per_step = timed(INCLUDE_ENERGY)
    launches_per_eval = ctx.kernel.stats().kernel_launches

    # end to end through the reference-facing call (host buffers in, host buffers out)
    e2e_times = []
    forces_host = np.zeros_like(pos)
    for i in range(args.steps + 3):
        flush.zero_()
        barrier()
        t = time.perf_counter()
        if world == 1:
            forces_host[:] = 0.0
            if i == 0:
                k1 = runtime.CalcCoulForceKernel(device=local)
                k1.initialize(box, force)
            k1.execute(pos, box, forces_host, True, INCLUDE_ENERGY)
        else:
            ctx.evaluate(pos, True, INCLUDE_ENERGY)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if i >= 3:
            e2e_times.append(dt)
    clocks = sampler.stop(t_start, time.perf_counter()) if sampler else None     # timed loop + e2e loop, both under load
    e2e_t = torch.tensor([float(np.mean(e2e_times))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    line = None
    if rank == 0:
        alpha, kmax, nk = ctx.kernel.ewald_params()
        st = ctx.kernel.stats()
        pairs = st.pairs_in_cutoff if world == 1 else None
        value = 1e3 / ms_per_step
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (f64 energies/accumulation, int64 fixed-point forces)", "data": "synthetic",
                "config": {"workload": workload, "atoms": n, "kmax": list(kmax), "kvectors": int(nk), "alpha": alpha,
                           "flags": "includeForces=1 includeEnergy=0 (the per-MD-step call)",
                           "l2": "384 MiB buffer written between timed steps (L2 flush); working set < L2",
                           "parallelism": "k-vector rows + direct-space i-tiles sharded x%d, NCCL all-reduce of int64 forces" % world
                           if world > 1 else "single GPU"},
                "ns_per_day": ns_per_day(value),
                "energy_and_forces": {"value": 1e3 / ms_ef, "unit": UNIT, "ms_per_step": ms_ef,
                                      "note": "includeEnergy=1: FP64 pair energies and the FP32 (round-to-nearest) structure-factor kernel"},
                "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 24 * n + 40,
                        "ns_per_day": ns_per_day(1.0 / e2e_s)},
                "gpu_launches": int(launches_per_eval) * args.steps,
                "clocks": clocks}
    if world == 1:
        # per-kernel durations (CUDA events on the launching stream) and the roofline of the dominant one
        tf_peak, _ = runtime.measure_fp32_peak(local, 5)
        tf32_peak = runtime.measure_tf32_peak(local)
        kt = ctx.kernel.time_kernels(ctx.d_pos.data_ptr(), box, 10, True, INCLUDE_ENERGY)
        kt_ef = ctx.kernel.time_kernels(ctx.d_pos.data_ptr(), box, 10, True, True)
        ctx.evaluate_device(True, INCLUDE_ENERGY)
        torch.cuda.synchronize()
        pairs = ctx.kernel.stats().pairs_in_cutoff
        fl = algorithmic_flops(n, nk, pairs, force.getNumFluxBonds() + force.getNumFluxAngles() + force.getNumFluxWaters(),
                               4 * force.getNumFluxBonds() + 9 * force.getNumFluxAngles() + 9 * force.getNumFluxWaters(),
                               force.getNumExceptions())
        ex = executed_tensor_flops(n, kmax)
        top = max(kt, key=kt.get)
        total_kernel_ms = sum(kt.values())
        achieved = fl.get(top, 0.0) / (kt[top] * 1e-3) / 1e12
        try:
            traffic_db = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json"))) if args.workload == "c3" else {}
        except (OSError, ValueError):
            traffic_db = {}

        def traffic_of(name):
            tr = traffic_db.get(name)
            return tr["dram_read_bytes"] + tr["dram_write_bytes"] if tr else None

        kernels = []
        for name in ("direct_pairs", "kspace_gather", "structure_factor"):
            if name not in kt:
                continue
            tensor = name in ex
            peak = tf32_peak if tensor else tf_peak
            a = fl[name] / (kt[name] * 1e-3) / 1e12
            row = {"kernel": name, "bound": "tensor" if tensor else "fp32", "ms": kt[name], "share_of_step": kt[name] / total_kernel_ms,
                   "algorithmic_flop_per_launch": fl[name], "achieved": a, "peak": peak, "unit": "TFLOP/s", "frac": a / peak,
                   "traffic": traffic_of(name)}
            if tensor:
                row["executed_tf32_flop_per_launch"] = ex[name]
                row["executed_tflops"] = ex[name] / (kt[name] * 1e-3) / 1e12
                row["executed_frac"] = row["executed_tflops"] / peak
                row["fp32_equivalent_frac_of_fp32_peak"] = a / tf_peak
            kernels.append(row)
        line["roofline"] = {"bound": "fp32", "kernel": top, "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                            "frac": achieved / tf_peak, "traffic": traffic_of(top),
                            "traffic_note": "DRAM bytes per launch of this kernel from the committed ncu capture (profiles/); far below "
                                            "the algorithmic FLOP x 4 B: the kernel is instruction-issue / SFU bound, not HBM bound",
                            "peak_source": "FP32 FMA microbenchmark run in this process (cfx_measure_fp32_peak); "
                                           "theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4. TF32 peak: tcgen05 kind::tf32 "
                                           "128x128x8 microbenchmark run in this process (cfx_measure_tf32_peak)",
                            "kernel_ms": kt[top], "kernel_share_of_step": kt[top] / total_kernel_ms,
                            "algorithmic_flop_per_launch": fl.get(top, 0.0),
                            "kernels": kernels,
                            "whole_step": {"achieved": fl["total"] / (ms_per_step * 1e-3) / 1e12,
                                           "frac": fl["total"] / (ms_per_step * 1e-3) / 1e12 / tf_peak,
                                           "algorithmic_flop": fl["total"],
                                           "note": "algorithmic FP32-equivalent FLOP of the whole evaluation / step time, against the "
                                                   "FP32 FMA peak; the reciprocal-space part runs on tensor cores"}}
        line["kernels_ms"] = {k: round(v, 5) for k, v in kt.items()}
        line["energy_and_forces"]["kernels_ms"] = {k: round(v, 5) for k, v in kt_ef.items()}
        line["peaks"] = {"fp32_fma_tflops": tf_peak, "tf32_tcgen05_tflops": tf32_peak}
        line["config"]["pairs_in_cutoff"] = int(pairs)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pos, box, force)
        if args.md_steps > 0:
            line["md_nve"] = md_leg(args)
        # second reported baseline of north_star: the plugin's EXISTING CUDA kernels on this GPU
        line["existing_cuda_baseline"] = existing_cuda_baseline(args.workload, ms_per_step)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
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
    args = ap.parse_args()
    from openmm_chargeflux_b200 import synthetic
    pos, box, force = synthetic.config(args.workload)
    if args.impl == "reference":
        run_reference_arm(args, pos, box, force, WORKLOADS[args.workload])
    else:
        run_ours(args, pos, box, force, WORKLOADS[args.workload])


if __name__ == "__main__":
    main()
