#!/usr/bin/env python
"""Compare the SASS of every kernel of two builds of libmptrac_b200.so (cuobjdump; no GPU needed): which kernels are
instruction-for-instruction identical, which changed, which are new.  Used to show that a change which was made without
GPU time left the verified and measured kernels untouched.
usage: scripts/sass_diff.py OLD.so NEW.so     (build an old commit with: git archive <rev> mptrac_b200/csrc include | tar -x -C DIR)"""
import re
import subprocess
import sys


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            funcs[name].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return funcs


a, b = sass(sys.argv[1]), sass(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
print(f"identical: {len(same)}   changed: {len(diff)}   removed: {len([k for k in a if k not in b])}   new: {len([k for k in b if k not in a])}")
for k in diff:
    print("  changed", k, len(a[k]), "->", len(b[k]), "instructions")
for k in b:
    if k not in a:
        print("  new    ", k, len(b[k]), "instructions")
sys.exit(1 if diff else 0)
