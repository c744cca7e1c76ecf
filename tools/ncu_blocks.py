"""Summarise an .ncu-rep of one kernel: headline counters and a basic-block table (consecutive SASS instructions with the
same execution count) with instruction and stall-sample shares.   python tools/ncu_blocks.py file.ncu-rep [min_pct]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # which launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2 + kidx]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"]
for h, v in zip(hdr, r):
    if h in want or ("issue_stalled" in h and h.endswith("ratio") and float(v or 0) > 0.15):
        print("%-90s %s" % (h, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[min(2*kidx, len(starts) - 1)]:]        # every launch is exported twice
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
blocks, cur, k = [], None, 0
for r in rows[2:]:
    if r[0] == "Kernel Name":
        break
    try:
        n, s = int(r[iE]), int(r[iSm])
    except (ValueError, IndexError):
        continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if cur and cur["n"] == n:
        cur["cnt"] += 1; cur["smp"] += s; cur["ops"].append(op)
    else:
        cur = {"n": n, "cnt": 1, "smp": s, "ops": [op], "idx": k}; blocks.append(cur)
    k += 1
tot = sum(b["n"] * b["cnt"] for b in blocks); ts = sum(b["smp"] for b in blocks)
print("total warp instructions", tot, "samples", ts)
for b in blocks:
    w = b["n"] * b["cnt"]
    if w > tot * minpct / 100:
        print("idx %5d len %4d exec %8d  instr %5.1f%%  smp %5.1f%%  %s" % (b["idx"], b["cnt"], b["n"], 100 * w / tot, 100 * b["smp"] / ts,
              dict(collections.Counter(b["ops"]).most_common(7))))
