"""per-source-line instruction / sample shares of an ncu report, aggregated over line ranges of dp_group.cuh
usage: python tools/ncu_lines.py report.ncu-rep"""
import csv, sys, subprocess, collections, re, os
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
res = []; cur = None; h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        h = r; ix = {c: i for i, c in enumerate(h) if c not in ('Source',)}; continue
    if r[0] != "" and h:
        try: res.append((cur, int(r[0]), float(r[ix['# Samples']]), float(r[ix['Instructions Executed']]), float(r[ix['Thread Instructions Executed']]) if 'Thread Instructions Executed' in ix else 0.0))
        except Exception: pass
base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'decision-making-and-path-planning_b200', 'csrc', 'dp_group.cuh')
# section markers: function definitions and phase comments
marks = []
for i, l in enumerate(open(base), 1):
    m = re.match(r'^(DG_FN|DG_NOINLINE|template <bool OFFS> DG_FN|template <bool OFFS>)\s*.*?(\w+)\(', l)
    if m: marks.append((i, m.group(2)))
    m = re.match(r'^\s*// ---- (P\w+)', l)
    if m: marks.append((i, m.group(1)))
marks.sort()
def sec(line):
    name = '?'
    for i, n in marks:
        if i <= line: name = n
        else: break
    return name
agg = collections.Counter(); smp = collections.Counter(); thr = collections.Counter()
for f, l, s, i, t in res:
    k = sec(l) if f == 'dp_group.cuh' else f
    agg[k] += i; smp[k] += s; thr[k] += t
ti = sum(agg.values()); ts = sum(smp.values())
print("total warp inst %.0f samples %.0f" % (ti, ts))
for k, v in agg.most_common(40):
    print("%-22s inst %5.1f%% (%9.0f, lanes %4.1f)  samples %5.1f%%" % (k, 100 * v / ti, v, thr[k] / max(v, 1), 100 * smp[k] / ts))
