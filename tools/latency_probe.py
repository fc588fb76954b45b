"""Developer probe: per-call latency of the host-buffer entry points for tiny ray counts (rtcIntersect1-style use)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench
scene = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
rays = rg.primary_rays(64, 64, **rg.BOMBERMAN_CAMERA)
rays = np.concatenate([rays] * 64)
for n in (1, 16, 256, 4096, 8192, 16384, 65536, 262144):
    r = rays[:n].copy()
    scene.intersect1M(r)
    t0 = time.perf_counter()
    reps = 300
    for _ in range(reps):
        r[:] = rays[:n]
        scene.intersect1M(r)
    dt = (time.perf_counter() - t0) / reps
    print("M=%5d pageable: %.1f us per call (%.3f Mrays/s)" % (n, dt * 1e6, n / dt / 1e6), flush=True)
one = rays[100:101].copy()
t0 = time.perf_counter()
for _ in range(300):
    one[:] = rays[100:101]
    scene.intersect1(one[0])
print("ecbIntersect1: %.1f us per call" % ((time.perf_counter() - t0) / 300 * 1e6))
