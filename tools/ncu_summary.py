"""Summarise an .ncu-rep (read on the CPU box): key metrics, stall mix, opcode mix, hot lines."""
import csv, re, subprocess, sys, os
from collections import Counter
rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "ra_b200/csrc/libra_engine.so"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, d = rows[0], rows[1], rows[2]
def g(k):
    return d[hdr.index(k)] if k in hdr else None
print("kernel:", g("Kernel Name"))
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
          'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers',
          'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active',
          'smsp__warps_active.avg.per_cycle_active', 'launch__grid_size', 'launch__block_size']:
    print("  %-62s %s [%s]" % (k, g(k), units[hdr.index(k)] if k in hdr else ""))
st = [(float(d[i] or 0), h) for i, h in enumerate(hdr) if 'stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
print("stall mix (warps stalled per issue):")
for v, h in sorted(st, reverse=True)[:8]:
    print("  %6.2f %s" % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}
data = rows[2:]
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[ix['Instructions Executed']]) for r in data); samp = sum(f(r[ix['# Samples']]) for r in data)
c = Counter(); cs = Counter()
for r in data:
    t = r[ix['Source']].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    op = op.split('.')[0]
    c[op] += f(r[ix['Instructions Executed']]); cs[op] += f(r[ix['# Samples']])
print("opcode mix (executed warp instructions %.0f, %d SASS lines):" % (tot, len(data)))
for op, v in c.most_common(14):
    print("  %-8s %5.1f%% inst  %5.1f%% samples" % (op, 100 * v / tot, 100 * cs[op] / samp))
# line attribution through nvdisasm -g
import tempfile
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
kname = g("Kernel Name").split("(")[0]
best = None
for fn in os.listdir(tmp):
    if fn.endswith(".cubin") and fn.startswith("engine."):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, fn)], capture_output=True, text=True).stdout.split("\n")
        secs = [i for i, l in enumerate(txt) if l.startswith('.text.') and 'raft_step' in l]
        mm = re.search(r"<(\d+)>", g("Kernel Name"))
        want = "ILi%sE" % mm.group(1) if mm else ""
        for sidx in secs:
            if want in txt[sidx]:
                best = (txt, sidx)
if best:
    txt, start = best
    a2l = {}; cur = None
    for l in txt[start + 1:]:
        if l.startswith('.text.') or l.startswith('//--------------------- .text'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m: a2l[int(m.group(1), 16)] = cur
    bl = Counter(); bs = Counter(); base = None
    for r in data:
        a = int(r[ix['Address']], 16)
        if base is None: base = a
        k = a2l.get(a - base)
        bl[k] += f(r[ix['Instructions Executed']]); bs[k] += f(r[ix['# Samples']])
    srcs = {}
    for fn in ['ra_b200/csrc/raft_step.cuh', 'ra_b200/csrc/engine.cu']:
        srcs[os.path.basename(fn)] = open(fn).read().split('\n')
    print("hot source lines (by stall samples):")
    for k, v in bs.most_common(22):
        line = srcs.get(k[0], [''] * 100000)[k[1] - 1].strip()[:80] if k and k[0] in srcs else ''
        print("  %5.1f%% samp %5.1f%% inst  %s:%s  %s" % (100 * v / samp, 100 * bl[k] / tot, k[0] if k else None, k[1] if k else None, line))
