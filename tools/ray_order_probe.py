"""Developer probe: does the ORDER of the same rays in a stream matter?  Times the closest-hit kernels on the BASELINE config 2
frame (primary + bounce) and on config 4 (instancing, every 2nd pixel of the 8K frame) with the pixels in row-major order, in
the order a tile renderer submits them (the reference viewer's 8x8 renderTile, tutorials/viewer/viewer_device.cpp:448-451) and
in Morton order.  The hit records, put back in pixel order, must be the same bits in every order.

    python tools/ray_order_probe.py
"""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg, raygen_device as rgd, sharding as sh
import __graft_entry__ as ge
import bench

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def morton_order(w, h):
    ys, xs = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")

    def spread(v):
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    code = (spread(xs) | (spread(ys) << np.uint64(1))).ravel()
    return np.argsort(code, kind="stable")


def orders(w, h):
    yield "row-major", np.arange(w * h)
    for t in (8, 16, 64):
        yield "%dx%d tiles" % (t, t), sh.tile_pixel_indices(w, sh.tile_grid(w, h, t))
    tiles = [(x0, y0, min(8, w - x0), min(4, h - y0)) for y0 in range(0, h, 4) for x0 in range(0, w, 8)]
    yield "8x4 tiles", sh.tile_pixel_indices(w, tiles)
    yield "morton", morton_order(w, h)


def run(scene, rays_t, label, w, h):
    base = None
    for name, idx in orders(w, h):
        it = torch.from_numpy(idx.astype(np.int64)).to(dev)
        src = rays_t[it].contiguous()
        ts = []
        for _ in range(6):
            tmp = src.clone(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); scene.intersect_device(tmp.data_ptr(), tmp.shape[0], 80, stream); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        back = torch.empty_like(tmp)
        back[it] = tmp
        crc = zlib.crc32(back.cpu().numpy().tobytes())
        base = crc if base is None else base
        ms = float(np.median(ts[1:]))
        print("%-22s %-12s %8d rays %.4f ms %6.0f Mrays/s  %s" % (label, name, src.shape[0], ms, src.shape[0] / ms / 1e3,
                                                                  "same bits" if crc == base else "DIFFERENT"), flush=True)


prim, bounce, _ = bench.config2_rays(rg, "leaf")
sc2 = ecb.Scene.from_image_file(os.path.join(ROOT, "scenes", "_built", "bomberman_s6c3_leaf.rimg")).upload(0)
run(sc2, torch.from_numpy(prim).to(dev), "config 2 primary", 1920, 1080)
run(sc2, torch.from_numpy(bounce).to(dev), "config 2 bounce", 1920, 1080)
sc2.release()
sc4 = ecb.Scene.load(os.path.join(ROOT, "scenes", "_built", ge.CONFIG4_BLOB)).upload(0)
ys, xs = torch.meshgrid(torch.arange(0, 4320, 2, device=dev), torch.arange(0, 7680, 2, device=dev), indexing="ij")
rays = rgd.primary_rays(7680, 4320, device=dev, pixel_index=(ys * 7680 + xs).reshape(-1).to(torch.int64), **ge.CONFIG4_CAMERA)
run(sc4, rays, "config 4 (3840x2160)", 3840, 2160)
