#!/usr/bin/env python3
"""Top stall sites of one kernel from an ncu report (source page, SASS view).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep edge_fwd [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first kernel instance only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(lines[start:end]))
hdr = rows[0]
iS, iN, iE = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
body = [(int(r[iS] or 0), int(r[iN] or 0), int(r[iE] or 0), idx, r[1].strip()) for idx, r in enumerate(rows[1:])]
tot = sum(b[0] for b in body)
print(f"total samples {tot}, instructions {len(body)}, warp-inst executed {sum(b[2] for b in body)}")
for s, ni, ex, idx, src in sorted(body, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  all={s:6d} notissued={ni:6d} exec={ex:9d}  #{idx:5d}  {src[:90]}")
