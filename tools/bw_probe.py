"""Measures the read bandwidth of L2 and HBM on this GPU with the library's probe kernel (ecbProbeReadBandwidth) and writes
profiles/measured_l2.json -- the denominator of the roofline's L2 fraction (bench.py reads it next to the driver-written
MEASURED_PEAKS.json).

    python tools/bw_probe.py [--write]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embree_compressed_b200 as ecb

MB = 1 << 20
rows = []
for ws, passes in ((8 * MB, 256), (16 * MB, 128), (32 * MB, 64), (48 * MB, 48), (64 * MB, 32), (96 * MB, 24), (128 * MB, 16), (256 * MB, 8), (1024 * MB, 2), (4096 * MB, 1)):
    g = ecb.probe_read_bandwidth(ws, passes)
    rows.append({"working_set_mb": ws // MB, "passes": passes, "read_gbs": round(g, 1)})
    print("%6d MB x %3d passes: %8.1f GB/s" % (ws // MB, passes, g), flush=True)
l2 = max(r["read_gbs"] for r in rows if r["working_set_mb"] <= 64)
hbm = rows[-1]["read_gbs"]
out = {"l2_read_gbs": l2, "hbm_read_gbs": hbm, "sweep": rows, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
       "how": "ecbProbeReadBandwidth: 148 x 8 CTAs x 256 threads, four 128-bit ld.global.cg loads in flight per thread, best of 5 launches "
              "after a warm-up launch; l2_read_gbs = best working set <= 64 MB, hbm_read_gbs = 4 GB working set, one pass"}
print(json.dumps(out))
if "--write" in sys.argv:
    with open(os.path.join(ROOT, "profiles", "measured_l2.json"), "w") as f:
        json.dump(out, f, indent=1)
