#!/usr/bin/env python
"""Per-source-line cost of the first kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_source_lines.py rep.ncu-rep [top N] -> prints lines sorted by stall samples and by warp-instructions executed"""
import csv, io, subprocess, sys
def num(x):
    try: return int(x)
    except ValueError: return 0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; cur_fn = None; take = True; recs = []; hdr = None; seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "Function Name":
        if cur_fn is None: cur_fn = r[1]
        take = (r[1] == cur_fn)
        continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur and take and r[0].isdigit():
        d = dict(zip(hdr, r))
        recs.append((cur, int(r[0]), r[1].strip()[:90], num(d["# Samples"]), num(d["Instructions Executed"]), d))
ts = sum(x[3] for x in recs); ti = sum(x[4] for x in recs)
print(f"total samples {ts}, total warp-instructions {ti}")
print("--- by samples")
for f, ln, src, s, i, d in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{f}:{ln:5d} {100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins  long_sb={d.get('stall_long_sb','')} wait={d.get('stall_wait','')} noinst={d.get('stall_no_inst','')} short={d.get('stall_short_sb','')} | {src}")
print("--- by instructions")
for f, ln, src, s, i, d in sorted(recs, key=lambda x: -x[4])[:top]:
    print(f"{f}:{ln:5d} {100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins | {src}")
# --- grouped by function region of physics.cuh (line ranges), for a quick "where do the instructions go"
import os
if os.environ.get("REGIONS"):
    groups = []
    for g in os.environ["REGIONS"].split(","):
        name, a, b = g.split(":"); groups.append((name, int(a), int(b)))
    acc = {n: [0, 0] for n, _, _ in groups}; acc["other"] = [0, 0]
    for f, ln, src, s, i, d in recs:
        key = "other"
        if f == "physics.cuh":
            for n, a, b in groups:
                if a <= ln <= b: key = n; break
        elif f == "engine.cu": key = "engine.cu"
        acc.setdefault(key, [0, 0]); acc[key][0] += s; acc[key][1] += i
    print("--- groups")
    for k, (s, i) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:24s} {100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins")
