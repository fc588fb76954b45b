"""Developer tool: end-to-end Mrays/s of ONE host-buffer call per pass on a scene resident on N devices (ecbSceneUploadMulti),
by transfer strategy.  The bench step (1920x1080 primary + bounce) from pinned or pageable caller memory.

    python tools/e2e_multi.py N [packed|compact|copy|auto] [pinned|pageable] [reps]
(the strategy is read once per process: ECB200_HOST_PATH is set here before the library loads)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
mem = sys.argv[3] if len(sys.argv) > 3 else "pinned"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
if path != "auto":
    os.environ["ECB200_HOST_PATH"] = path
import numpy as np
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench

prim, bounce, _ = bench.config2_rays(rg, "leaf")
scene = ecb.Scene.from_image_file(bench.scene_path("leaf"))
scene.upload_multi(list(range(n_dev))) if n_dev > 1 else scene.upload(0)
if mem == "pinned":
    hp, p1 = ecb.host_alloc(prim.nbytes); hb, p2 = ecb.host_alloc(bounce.nbytes)
    hp = hp.view(np.float32).reshape(prim.shape); hb = hb.view(np.float32).reshape(bounce.shape)
else:
    hp, hb = np.empty_like(prim), np.empty_like(bounce)
n_rays = len(prim) + int((bounce[:, 3] <= bounce[:, 8]).sum())
ts = []
for i in range(2 + reps):
    hp[:] = prim; hb[:] = bounce
    t0 = time.perf_counter()
    scene.intersect1M(hp); scene.intersect1M(hb)
    ts.append(time.perf_counter() - t0)
t = float(np.mean(ts[2:]))
hits = int((hp.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum())
print("devices=%d path=%-7s mem=%-8s threads=%s  %7.1f Mrays/s  (%.3f ms per step, best %.3f; %d primary hits)"
      % (n_dev, path, mem, os.environ.get("ECB200_HOST_THREADS", "default"), n_rays / t / 1e6, t * 1e3, min(ts[2:]) * 1e3, hits), flush=True)
