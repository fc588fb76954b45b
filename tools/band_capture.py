"""Developer tool for ncu: a few launches of the closest-hit kernel on ONE band of image rows of the bench frame (the launch is
latency-bound: 480 warps on 148 SMs; what do the longest warps wait for?).

    python tools/band_capture.py [y0=488] [rows=8] [launches=3]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench
y0 = int(sys.argv[1]) if len(sys.argv) > 1 else 488
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prim = rg.primary_rays(1920, 1080, **rg.BOMBERMAN_CAMERA)
scene = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
src = torch.from_numpy(prim[y0 * 1920:(y0 + rows) * 1920]).to(dev); n = src.shape[0]
geom = torch.empty(n, dtype=torch.int32, device=dev); hit = torch.empty((n, 4), dtype=torch.float32, device=dev)
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); scene.intersect_device_compact(src.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream); b.record()
    torch.cuda.synchronize()
    print("rows %d-%d, %d rays: %.4f ms" % (y0, y0 + rows - 1, n, a.elapsed_time(b)), flush=True)
# the slowest chunks of 32 rays, each launched alone (one warp on the whole device: the pure dependent chain)
ts = []
for c in range(0, n, 32):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); scene.intersect_device_compact(src.data_ptr() + c * 80, 32, geom.data_ptr(), hit.data_ptr(), 80, stream); b.record()
    ts.append((a, b))
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in ts])
order = np.argsort(ms)[::-1]
print("single-chunk launches: median %.4f ms, p90 %.4f, max %.4f (chunk %d); five slowest: %s" % (np.median(ms), np.percentile(ms, 90), ms.max(), order[0], ", ".join("%d: %.4f" % (i, ms[i]) for i in order[:5])))
