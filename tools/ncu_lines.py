"""Per-source-line executed-instruction attribution for one kernel of an .ncu-rep (read on the CPU box).
usage: python tools/ncu_lines.py <rep> <kernel-substring> [so]"""
import csv, subprocess, os, tempfile, re, sys
from collections import Counter
rep, want = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else "ra_b200/csrc/libra_engine.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = [b for b in blocks if want in b["name"]][0]
h = blk["rows"][0]; ix = {k: i for i, k in enumerate(h)}
data = [r for r in blk["rows"][1:] if len(r) == len(h)]
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[ix['Instructions Executed']]) for r in data)
samp = sum(f(r[ix['# Samples']]) for r in data)
print(blk["name"][:90], "SASS lines", len(data), "warp instructions", tot)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
m = re.search(r"raft_(\w+)_kernel<\(int\)(\d+)(?:, \((int|bool)\)(\d+))?>", blk["name"])
kind, mm, ty, roles = m.group(1), m.group(2), m.group(3), m.group(4)
want = "raft_%s_kernelILi%sE" % (kind, mm) + ((("Lb%sE" if ty == "bool" else "Li%sE") % roles) if roles else "")
ns = re.search(r"(ra_narrow|ra_wide)::", blk["name"])          # the hot kernel exists once per index width
if ns: want = "%d%s%d%s" % (len(ns.group(1)), ns.group(1), len("raft_%s_kernel" % kind), want)
a2l = {}
for fn in os.listdir(tmp):
    if not (fn.endswith(".cubin") and fn.startswith("engine.")): continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, fn)], capture_output=True, text=True).stdout.split("\n")
    for sidx in [i for i, l in enumerate(txt) if l.startswith('.text.') and want in l]:
        curl = None
        for l in txt[sidx + 1:]:
            if l.startswith('.text.') or l.startswith('//--------------------- .text'): break
            mm_ = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if mm_: curl = (os.path.basename(mm_.group(1)), int(mm_.group(2))); continue
            mm_ = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
            if mm_: a2l[int(mm_.group(1), 16)] = curl
c = Counter(); cs = Counter(); base = None
for r in data:
    a = int(r[ix['Address']], 16)
    if base is None: base = a
    k = a2l.get(a - base, ('?', 0))
    c[k] += f(r[ix['Instructions Executed']]); cs[k] += f(r[ix['# Samples']])
for (fnm, l), v in c.most_common(int(os.environ.get("TOP", "40"))):
    try: text = open('ra_b200/csrc/' + fnm).read().splitlines()[l - 1].strip()[:88]
    except Exception: text = ''
    print("%5.1f%% inst %5.1f%% samp %-14s %5d  %s" % (100 * v / tot, 100 * cs[(fnm, l)] / max(1, samp), fnm, l, text))
