#!/usr/bin/env python
"""SASS instruction histogram of the kernels that matter, per library flavour (no GPU needed: cuobjdump on the built .so).
usage: scripts/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
LIBS = {"production (libmptrac_b200.so)": ROOT / "mptrac_b200/_lib/libmptrac_b200.so",
        "strict (libmptrac_b200_strict.so: -fmad=false, TMA bulk staging of the parcel stream)": ROOT / "mptrac_b200/_lib/libmptrac_b200_strict.so"}
KERNELS = ["step_kernelILi4ELj0E", "step_kernelILi4ELj3E", "step_kernelILi4ELj6E", "tile_step_kernelILi4ELj0E", "quad_step_kernelILi4ELi3E",
           "advect_levels_kernelILi4E", "mix_route_kernel", "mix_fold_kernel", "mix_answer_kernel", "peer_barrier_kernel"]
WATCH = ["LDG.E.ENL2.256", "LDG.E.128", "LDG.E.64", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDS.128", "F2F.F64.F32", "DFMA", "DADD", "DMUL",
         "SHFL", "RED", "ATOM", "ST.E", "MUFU"]
for name, lib in LIBS.items():
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    print(f"== {name}")
    for k in KERNELS:
        body = next((f for f in funcs if f.split("\n", 1)[0].find(k) >= 0), None)
        if body is None:
            continue
        ops = collections.Counter()
        for ln in body.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        sel = {w: sum(v for o, v in ops.items() if o.startswith(w)) for w in WATCH}
        sel = {w: v for w, v in sel.items() if v}
        print(f"  {body.split(chr(10), 1)[0].strip()[:80]}: {total} instructions; " + ", ".join(f"{w} {v}" for w, v in sel.items()))
