"""Developer tool: time the traversal kernels on device-resident rays and print a checksum of the results.

    ECB200_LIB=embree_compressed_b200/libecb200_<variant>.so python tools/kbench.py [leaf|grid|box] [--scene NAME] [--check]

Default scene: the benchmark's bomberman S6/C3 image of the given leaf kind, 1920x1080.  --scene cube_s6c4_leaf /
cube_s8c4_leaf: the displaced cube of BASELINE configs 1 / 3 (1024x1024 / 3840x2160, tutorial camera).
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

args = sys.argv[1:]
leaf = args[0] if args and not args[0].startswith("-") else "leaf"
scene_name = args[args.index("--scene") + 1] if "--scene" in args else None
dev = torch.device("cuda", 0)
if scene_name:
    path = os.path.join(ROOT, "scenes", "_built", scene_name + ".rimg")
    cam = rg.CUBE_CAMERA if scene_name.startswith("cube") else rg.BOMBERMAN_CAMERA
    W, H = (3840, 2160) if "s8" in scene_name else (1024, 1024) if scene_name.startswith("cube") else (1920, 1080)
else:
    path, cam, (W, H) = bench.scene_path(leaf), rg.BOMBERMAN_CAMERA, (1920, 1080)
scene = ecb.Scene.from_image_file(path).upload(0)
stream = torch.cuda.current_stream().cuda_stream
# primary rays and one cosine-diffuse bounce per hit (block-frame normals), as bench.make_rays does for the benchmark scene
prim = rg.primary_rays(W, H, **cam)
d = torch.from_numpy(prim).to(dev)
blocks = torch.empty(len(prim), dtype=torch.int32, device=dev)
scene.intersect_device(d.data_ptr(), len(prim), 80, stream, blocks.data_ptr())
torch.cuda.synchronize()
first = d.cpu().numpy()
hit = first.view(np.uint32)[:, 18] != 0xFFFFFFFF
normals = np.zeros((len(prim), 3), np.float32)
normals[hit] = scene.block_normals()[blocks.cpu().numpy().view(np.uint32)[hit]]
b = scene.info()["bounds"]
xs, ys = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
bounce = rg.diffuse_rays(first, normals, float(np.linalg.norm(np.array(b[3:]) - np.array(b[:3]))), sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
tag = scene_name or leaf
out = {}
for name, rays in (("primary", prim), ("bounce", bounce)):
    src = torch.from_numpy(rays).to(dev)
    ts = []
    for _ in range(8):
        tmp = src.clone(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); scene.intersect_device(tmp.data_ptr(), len(rays), 80, stream); b_.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b_))
    res = tmp.cpu().numpy()
    active = int((rays[:, 3] <= rays[:, 8]).sum())
    out[name] = res
    print("%s %-14s %s: %.4f ms (min %.4f)  %.0f Mrays/s  crc=%08x" % (os.path.basename(ecb.LIB_PATH), tag, name, float(np.median(ts[2:])), min(ts), active / np.median(ts[2:]) / 1e3, zlib.crc32(res.tobytes())), flush=True)
if "--check" in args:
    from oracle import pyoracle as po
    port = po.PortScene.from_image(np.fromfile(path, np.uint8))
    for name, rays in (("primary", prim), ("bounce", bounce)):
        want = rays.copy(); port.intersect(want)
        bad = (want.view(np.uint32) != out[name].view(np.uint32)).any(axis=1).sum()
        print("  check %s vs port: %d mismatching rays" % (name, bad))
