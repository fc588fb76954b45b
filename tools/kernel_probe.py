"""Developer tool for ncu captures: one launch of the chosen closest-hit kernel per ray set of a scene.

    python tools/kernel_probe.py [lane|quad] [scene] [reps]
Launch order: warm-up primary, warm-up bounce, then `reps` x (primary, bounce).  Prints CUDA-event times.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "lane"
name = sys.argv[2] if len(sys.argv) > 2 else "bomberman_s6c3_leaf"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
scene = ecb.Scene.from_image_file(os.path.join(ROOT, "scenes", "_built", name + ".rimg")).upload(0)
if name.startswith("bomberman"):   # the benchmark frame (bounce rays of the S6/C3 hit file: the same surface at any S)
    prim, bounce, _ = bench.config2_rays(rg, "leaf")
else:
    cam = rg.CUBE_CAMERA
    W, H = (3840, 2160) if "s8" in name else (1024, 1024)
    prim = rg.primary_rays(W, H, **cam)
    ecb.set_traversal_kernel("lane")
    d = torch.from_numpy(prim).to(dev)
    scene.intersect_device(d.data_ptr(), len(prim), 80, stream)
    torch.cuda.synchronize()
    first = d.cpu().numpy()
    b = scene.info()["bounds"]
    xs, ys = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
    bounce = rg.diffuse_rays(first, -first[:, 4:7], float(np.linalg.norm(np.array(b[3:]) - np.array(b[:3]))), sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
ecb.set_traversal_kernel(mode)
sp, sb = torch.from_numpy(prim).to(dev), torch.from_numpy(bounce).to(dev)
for r in range(1 + reps):
    for what, src in (("primary", sp), ("bounce", sb)):
        tmp = src.clone(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); scene.intersect_device(tmp.data_ptr(), tmp.shape[0], 80, stream); b_.record(); torch.cuda.synchronize()
        if r > 0:
            print("%s %s %s: %.4f ms" % (mode, name, what, a.elapsed_time(b_)), flush=True)
