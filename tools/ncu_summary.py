"""Summarise an .ncu-rep (ncu --set full) launch by launch: duration, instructions, issue rate, pipes, stalls, DRAM / L2 bytes.
   python tools/ncu_summary.py file.ncu-rep [traffic.json [bench]]   -- the optional JSON gets {kernel: dram bytes} of the LAST launch
   of each name; with "bench" the keys are bench.py's kernel groups (capture of tools/probe_both.py: energy+forces calls first, then
   forces only; "<group>@E" = the energy+forces call, "<group>@F" = the forces-only call)"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}
traffic = {}
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    print("==", name)
    for c in cols:
        if c in d and d[c] != "":
            print("   %-88s %s %s" % (c, d[c], u.get(c, "")))
    try:
        rd = float(d["dram__bytes_read.sum"]) * scale.get(u["dram__bytes_read.sum"], 1.0)
        wr = float(d["dram__bytes_write.sum"]) * scale.get(u["dram__bytes_write.sum"], 1.0)
        traffic[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us": float(d["gpu__time_duration.sum"])}
        launches.append((name, traffic[name]))
    except (KeyError, ValueError):
        pass
if len(sys.argv) > 3 and sys.argv[3] == "bench":
    def group(name):
        if "pairKernel" in name:                       # pairKernel<FAST, FORCES, EMODE, EMIT>: <., 0, 2, .> is the FP64 energy pass
            args = name.split("<")[1].split(">")[0].replace(" ", "").split(",")
            return "direct_pairs_energy" if (args[1] == "0" and args[2] == "2") else "direct_pairs"
        if "gatherTensorKernel" in name or "gatherKernel" in name: return "kspace_gather"
        if "structureFactor" in name: return "structure_factor"
        return None
    # split the launch sequence into evaluations at every flux-term kernel (the first kernel of a call that is captured)
    evals, cur = [], None
    for name, t in launches:
        if "fluxTermKernel" in name or cur is None and group(name):
            cur = {}
            evals.append(cur)
        g = group(name)
        if g:
            acc = cur.setdefault(g, {"dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "duration_us": 0.0})
            for k in ("dram_read_bytes", "dram_write_bytes", "duration_us"): acc[k] += t[k]
    out = {}
    for ev in evals:                                   # later evaluations overwrite earlier ones (warm lists); @E = energy+forces call
        tag = "@E" if "direct_pairs_energy" in ev else "@F"
        for g, acc in ev.items():
            out[g + tag] = acc
    json.dump(out, open(sys.argv[2], "w"), indent=1)
elif len(sys.argv) > 2:
    json.dump(traffic, open(sys.argv[2], "w"), indent=1)
