"""Developer tool: time the traversal kernels on device-resident rays and print a checksum of the results.

    ECB200_LIB=embree_compressed_b200/libecb200_<variant>.so python tools/kbench.py [leaf] [--check]

--check compares against the oracle port (bit-exact) -- dev tool only, never part of the product path.
"""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench

leaf = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "leaf"
bench.WIDTH, bench.HEIGHT = 1920, 1080
dev = torch.device("cuda", 0)
scene = ecb.Scene.from_image_file(bench.scene_path(leaf)).upload(0)
prim, bounce, n_hits = bench.make_rays(scene, ecb, rg, torch, dev)
stream = torch.cuda.current_stream().cuda_stream
out = {}
for name, rays in (("primary", prim), ("bounce", bounce)):
    src = torch.from_numpy(rays).to(dev)
    ts = []
    for _ in range(8):
        tmp = src.clone(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); scene.intersect_device(tmp.data_ptr(), len(rays), 80, stream); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res = tmp.cpu().numpy()
    active = int((rays[:, 3] <= rays[:, 8]).sum())
    out[name] = res
    print("%s %-8s %s: %.4f ms (min %.4f)  %.0f Mrays/s  crc=%08x" % (os.path.basename(ecb.LIB_PATH), leaf, name, float(np.median(ts[2:])), min(ts), active / np.median(ts[2:]) / 1e3, zlib.crc32(res.tobytes())))
if "--check" in sys.argv:
    from oracle import pyoracle as po
    port = po.PortScene.from_image(np.fromfile(bench.scene_path(leaf), np.uint8))
    for name, rays in (("primary", prim), ("bounce", bounce)):
        want = rays.copy(); port.intersect(want)
        bad = (want.view(np.uint32) != out[name].view(np.uint32)).any(axis=1).sum()
        print("  check %s vs port: %d mismatching rays" % (name, bad))
