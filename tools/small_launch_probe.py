"""Developer probe: device time of the closest-hit kernel on 1/8 of the bench frame (the per-rank launch of --gpus 8) by grid
size (ECB200_RAYS_PER_THREAD), with launches enqueued back to back (20 per timing) so that host enqueue gaps do not count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg, sharding as sh
import bench
prim, bounce, _ = bench.config2_rays(rg, "leaf")
scene = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
for world in (1, 2, 4, 8, 16):
    idx, _, _ = bench.interleaved_tiles(1920, 1080, 0, world, sh) if world > 1 else (np.arange(len(prim)), 0, 0)
    for what, rays in (("primary", prim[idx]), ("bounce", bounce[idx])):
        src = torch.from_numpy(rays).to(dev)
        n = len(rays)
        geom = torch.empty(n, dtype=torch.int32, device=dev); hit = torch.empty((n, 4), dtype=torch.float32, device=dev)
        for _ in range(3): scene.intersect_device_compact(src.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): scene.intersect_device_compact(src.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print("rays/thread>=%s  1/%-2d of the frame %-8s %8d rays: %.4f ms  (%.0f Mrays/s; ideal at the full-frame rate %.4f ms)"
              % (os.environ.get("ECB200_RAYS_PER_THREAD", "1"), world, what, n, ms, n / ms / 1e3, n / (5300.0 if what == "primary" else 5800.0) / 1e3), flush=True)
