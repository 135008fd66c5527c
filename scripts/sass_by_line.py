import re, collections, subprocess, sys, os
lib, fun = os.path.abspath(sys.argv[1]), sys.argv[2]
import tempfile, os, glob
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {lib} > /dev/null && nvdisasm -g *.cubin > all.lines", shell=True, check=True)
txt = open(f"{d}/all.lines").read()
start = txt.index(f".text.{fun}:")
end = txt.find("\t.section", start)
body = txt[start:end]
cur = None
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter); allops = collections.Counter()
for ln in body.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)', ln)
    if m and cur:
        cnt[cur] += 1; ops[cur][m.group(2)] += 1; allops[m.group(2)] += 1
print("total", sum(cnt.values()))
print(dict(allops.most_common(25)))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{k[0]}:{k[1]:4d} {v:5d}  ", dict(ops[k].most_common(6)))
