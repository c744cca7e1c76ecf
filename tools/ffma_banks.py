"""Offline register-bank check of the FFMA instructions of a kernel (no GPU needed).
For every FFMA count the distinct even and odd source registers that are not served by the operand reuse
cache (.reuse on the PREVIOUS use of the same slot is what matters, approximated here by the flag on this
instruction's predecessor in the same slot). rt = max(#even, #odd) per the two-bank register file model."""
import re, subprocess, sys, collections
lib, pattern = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs:
    name = f.split("\n", 1)[0]
    if pattern not in name:
        continue
    hist = collections.Counter(); total = 0
    prev_reuse = {}           # slot -> register kept in the reuse cache
    for line in f.split("\n"):
        m = re.search(r"\bFFMA\s+(R\d+|RZ),\s*(-?\|?R\d+\|?(?:\.reuse)?|-?RZ|[^,]+),\s*(-?\|?R\d+\|?(?:\.reuse)?|[^,]+),\s*(-?\|?R\d+\|?(?:\.reuse)?|[^;]+?)\s*;", line)
        if not m:
            if re.search(r"^\s*/\*[0-9a-f]{4}\*/\s+(?!FFMA)", line):
                prev_reuse = {}
            continue
        srcs = [m.group(2), m.group(3), m.group(4)]
        fresh = []
        new_reuse = {}
        for slot, s in enumerate(srcs):
            r = re.search(r"R(\d+)", s)
            if not r:
                continue
            reg = int(r.group(1))
            if prev_reuse.get(slot) == reg:
                pass                          # served by the reuse cache
            else:
                fresh.append(reg)
            if ".reuse" in s:
                new_reuse[slot] = reg
        prev_reuse = new_reuse
        ev = len({r for r in fresh if r % 2 == 0}); od = len({r for r in fresh if r % 2 == 1})
        hist[max(ev, od, 1)] += 1; total += 1
    print(name[:90]); print("   FFMA:", total, " issue cycles by bank model:", dict(sorted(hist.items())),
          " avg rt = %.3f" % (sum(k * v for k, v in hist.items()) / max(total, 1)))
