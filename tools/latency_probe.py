"""Developer probe: per-call latency of the host-buffer entry points by ray count (rtcIntersect1-style use up to
tile-sized streams); the refill of the caller's array is outside the timed region."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench
scene = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
rays = rg.primary_rays(1024, 512, **rg.BOMBERMAN_CAMERA)
for n in (1, 16, 256, 4096, 8192, 8193, 16384, 32768, 65536, 131072, 524288):
    r = rays[:n].copy()
    scene.intersect1M(r)
    ts = []
    for _ in range(60 if n <= 65536 else 15):
        r[:] = rays[:n]
        t0 = time.perf_counter()
        scene.intersect1M(r)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print("M=%6d pageable: %8.1f us per call (%7.2f Mrays/s)" % (n, dt * 1e6, n / dt / 1e6), flush=True)
one = rays[100:101].copy()
ts = []
for _ in range(300):
    one[:] = rays[100:101]
    t0 = time.perf_counter()
    scene.intersect1(one[0])
    ts.append(time.perf_counter() - t0)
print("ecbIntersect1: %.1f us per call" % (float(np.median(ts)) * 1e6))
