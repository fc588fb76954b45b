"""Developer probe: device time of the closest-hit kernel per band of 8 image rows of the bench frame (15 360 primary rays each):
where the rays with the longest dependent chains are (the floor of small launches, profiles/r2n_small_launch.txt)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench
prim = rg.primary_rays(1920, 1080, **rg.BOMBERMAN_CAMERA)
scene = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
out = []
for y0 in range(0, 1080, 8):
    rays = prim[y0 * 1920:(y0 + 8) * 1920]
    src = torch.from_numpy(rays).to(dev); n = len(rays)
    geom = torch.empty(n, dtype=torch.int32, device=dev); hit = torch.empty((n, 4), dtype=torch.float32, device=dev)
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    tmp = src.clone(); scene.intersect_device_stats(tmp.data_ptr(), n, cnt.data_ptr(), 80, stream)
    for _ in range(2): scene.intersect_device_compact(src.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): scene.intersect_device_compact(src.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream)
    b.record(); torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    out.append((y0, a.elapsed_time(b) / 10, c[2] / n, c[1] / n, c[5] / n))
for y0, ms, blk, top, h in out:
    print("rows %4d-%4d: %.4f ms   blocks/ray %.2f  top nodes/ray %.2f  hit fraction %.2f" % (y0, y0 + 7, ms, blk, top, h))
