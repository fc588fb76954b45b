"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line (samples, instructions).

    python tools/ncu_lines.py dump.csv [launch index] [top N]
The dump holds, per profiled launch, one section per source FILE of the kernel; all sections of the chosen launch are merged
(lines are reported as file:line)."""
import csv, collections, os, sys
rows = list(csv.reader(open(sys.argv[1])))
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
secs = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and "Instructions Executed" in r]
# sections of one launch repeat the same file list: a launch starts where the first file path shows up again
files = [rows[s - 2][1] if s >= 2 and rows[s - 2] and rows[s - 2][0] == "File Path" else "?" for s in secs]
per_launch = files.index(files[0], 1) if files.count(files[0]) > 1 else len(files)
mine = list(range(launch * per_launch, min(len(secs), (launch + 1) * per_launch)))
agg = collections.defaultdict(lambda: [0, 0, 0, ""]); tot = [0, 0, 0]
def num(x):
    try: return int(x)
    except Exception: return 0
for k in mine:
    s = secs[k]
    e = secs[k + 1] - 2 if k + 1 < len(secs) else len(rows)
    hdr = rows[s]
    li = hdr.index("Line No"); sc = hdr.index("Source"); inst = hdr.index("Instructions Executed")
    samp = hdr.index("# Samples"); tinst = hdr.index("Thread Instructions Executed")
    fname = os.path.basename(files[k])
    for r in rows[s + 1:e]:
        if len(r) <= tinst: continue
        try: ln = int(r[li])
        except Exception: continue
        a = agg[(fname, ln)]
        a[0] += num(r[inst]); a[1] += num(r[samp]); a[2] += num(r[tinst]); a[3] = r[sc][:110]
        tot[0] += num(r[inst]); tot[1] += num(r[samp]); tot[2] += num(r[tinst])
print("launch", launch, rows[secs[mine[0]] - 1][1][:90], "| warp inst", tot[0], "samples", tot[1], "avg threads/inst %.2f" % (tot[2] / max(1, tot[0])))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("%-22s %5d inst%%=%5.1f samp%%=%5.1f thr=%4.1f | %s" % (f, ln, 100 * a[0] / tot[0], 100 * a[1] / max(1, tot[1]), a[2] / max(1, a[0]), a[3]))
