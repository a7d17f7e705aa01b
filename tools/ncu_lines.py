"""per-kernel executed-instruction breakdown by source line from `ncu --page source --csv` (file saved beforehand)
usage: python tools/ncu_lines.py src.csv <kernel-substring> [top]"""
import collections
import csv
import sys

rows = csv.reader(open(sys.argv[1]))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = kern = None
ix = None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kern = r[1]; continue
    if r[0] == "Line No":
        ix = {c: i for i, c in enumerate(r) if c != "Source"}; continue
    if r[0] != "" and ix and kern and want in kern:
        try:
            a = agg.setdefault((cur, int(r[0])), [r[1].strip()[:100], 0.0, 0.0])
            a[1] += float(r[ix["Instructions Executed"]]); a[2] += float(r[ix["# Samples"]])
        except (ValueError, KeyError):
            pass
ti = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("kernel %s: %d warp instructions, %d samples" % (want, ti, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-14s %4d inst %5.2f%% (%6.0f/scene@4096) smp %4.1f%% | %s" % (k[0][:14], k[1], 100 * a[1] / ti, a[1] / 4096, 100 * a[2] / max(ts, 1), a[0]))
