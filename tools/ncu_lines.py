"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line (samples, instructions)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
secs = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and "Instructions Executed" in r]
s = secs[which]
e = secs[which + 1] - 2 if which + 1 < len(secs) else len(rows)
hdr = rows[s]
li = hdr.index("Line No"); sc = hdr.index("Source"); inst = hdr.index("Instructions Executed")
samp = hdr.index("# Samples"); tinst = hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0, 0, ""]); tot = [0, 0, 0]
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows[s + 1:e]:
    if len(r) <= tinst: continue
    try: ln = int(r[li])
    except Exception: continue
    a = agg[ln]
    a[0] += num(r[inst]); a[1] += num(r[samp]); a[2] += num(r[tinst]); a[3] = r[sc][:100]
    tot[0] += num(r[inst]); tot[1] += num(r[samp]); tot[2] += num(r[tinst])
print("section", which, rows[s - 1][1][:80], "| warp inst", tot[0], "samples", tot[1], "avg threads/inst %.2f" % (tot[2] / max(1, tot[0])))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print("%5d inst%%=%5.1f samp%%=%5.1f thr=%4.1f | %s" % (ln, 100 * a[0] / tot[0], 100 * a[1] / tot[1], a[2] / max(1, a[0]), a[3]))
