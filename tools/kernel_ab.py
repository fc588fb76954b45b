"""Developer tool: the two closest-hit kernels (one lane per ray / four lanes per ray) side by side on device-resident rays:
time, bit-equality of the results and of the work counters.

    python tools/kernel_ab.py [scene names ...]      (default: the bench scenes)
"""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg

names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["bomberman_s6c3_leaf", "bomberman_s6c3_grid", "bomberman_s6c3_box",
                                                               "cube_s6c4_leaf", "cube_s8c4_leaf", "cubeplane_s6c4_leaf"]
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def run(scene, rays, mode, reps=6):
    ecb.set_traversal_kernel(mode)
    src = torch.from_numpy(rays).to(dev)
    ts = []
    for _ in range(reps):
        tmp = src.clone(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); scene.intersect_device(tmp.data_ptr(), len(rays), 80, stream); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    t2 = src.clone()
    scene.intersect_device_stats(t2.data_ptr(), len(rays), cnt.data_ptr(), 80, stream)
    torch.cuda.synchronize()
    return float(np.median(ts[1:])), tmp.cpu().numpy(), cnt.cpu().numpy(), t2.cpu().numpy()


for name in names:
    path = os.path.join(ROOT, "scenes", "_built", name + ".rimg")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "tests", "golden", name + ".rimg")
    cam = rg.CUBE_CAMERA if name.startswith("cube") else rg.BOMBERMAN_CAMERA
    W, H = (3840, 2160) if "s8" in name else (1024, 1024) if name.startswith("cube") else (1920, 1080)
    scene = ecb.Scene.from_image_file(path).upload(0)
    prim = rg.primary_rays(W, H, **cam)
    ecb.set_traversal_kernel("lane")
    d = torch.from_numpy(prim).to(dev)
    blocks = torch.empty(len(prim), dtype=torch.int32, device=dev)
    mixed = "plane" in name
    if not mixed:
        scene.intersect_device(d.data_ptr(), len(prim), 80, stream, blocks.data_ptr())
    else:
        scene.intersect_device(d.data_ptr(), len(prim), 80, stream)
    torch.cuda.synchronize()
    first = d.cpu().numpy()
    hit = first.view(np.uint32)[:, 18] != 0xFFFFFFFF
    normals = np.zeros((len(prim), 3), np.float32)
    if not mixed:
        normals[hit] = scene.block_normals()[blocks.cpu().numpy().view(np.uint32)[hit]]
    else:
        normals[hit] = -first[hit, 4:7]
    b = scene.info()["bounds"]
    xs, ys = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
    bounce = rg.diffuse_rays(first, normals, float(np.linalg.norm(np.array(b[3:]) - np.array(b[:3]))), sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
    for what, rays in (("primary", prim), ("bounce", bounce)):
        active = int((rays[:, 3] <= rays[:, 8]).sum())
        res = {}
        for mode in ("lane", "quad"):
            res[mode] = run(scene, rays, mode)
        ml, mq = res["lane"][0], res["quad"][0]
        same = (res["lane"][1].view(np.uint32) == res["quad"][1].view(np.uint32)).all(axis=1)
        same_stats = (res["lane"][2] == res["quad"][2]).all()
        print("%-22s %-8s %9d rays  lane %8.4f ms %7.0f Mrays/s | quad %8.4f ms %7.0f Mrays/s  x%.2f | results %s (%d differ) counters %s %s"
              % (name, what, active, ml, active / ml / 1e3, mq, active / mq / 1e3, ml / mq, "SAME" if same.all() else "DIFFER", int((~same).sum()),
                 "same" if same_stats else "DIFFER", "" if same_stats else str(res["lane"][2]) + " vs " + str(res["quad"][2])), flush=True)
        if not same.all():
            bad = np.nonzero(~same)[0][:5]
            for i in bad:
                print("   ray", i, "lane", res["lane"][1][i, [8, 15, 16]], res["lane"][1].view(np.uint32)[i, 17:20], "quad", res["quad"][1][i, [8, 15, 16]], res["quad"][1].view(np.uint32)[i, 17:20])
    scene.release()
ecb.set_traversal_kernel(None)
