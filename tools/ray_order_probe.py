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


def spread3(v):
    v = v.astype(np.uint64) & np.uint64(0x3FF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x030000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x0300F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x030C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x09249249)
    return v


def sorted_orders(rays, w, h):
    """Orders computed from the rays themselves (what a sorting pre-pass on the device would produce)."""
    org, d = rays[:, 0:3].astype(np.float64), rays[:, 4:7]
    octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2)))
    tile = np.empty(w * h, np.uint64)
    tile[sh.tile_pixel_indices(w, sh.tile_grid(w, h, 8))] = np.arange(w * h, dtype=np.uint64)
    yield "octant,8x8 tiles", np.argsort((octant << np.uint64(40)) | tile, kind="stable")
    act = (rays[:, 3] <= rays[:, 8]) & np.isfinite(org).all(1)
    lo, hi = org[act].min(0), org[act].max(0)
    org = np.where(act[:, None], org, lo)
    for bits in (4, 6):
        q = np.clip(((org - lo) / np.maximum(hi - lo, 1e-30) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        m = spread3(q[:, 0]) | (spread3(q[:, 1]) << np.uint64(1)) | (spread3(q[:, 2]) << np.uint64(2))
        yield "morton%d(org),octant" % bits, np.argsort((m << np.uint64(3)) | octant, kind="stable")
        yield "octant,morton%d(org)" % bits, np.argsort((octant << np.uint64(40)) | m, kind="stable")
    # direction cell (6 faces x 8x8) then origin morton
    a = np.abs(d)
    ax = np.argmax(a, 1)
    face = (ax * 2 + (d[np.arange(len(d)), ax] < 0)).astype(np.uint64)
    u = d[np.arange(len(d)), (ax + 1) % 3] / np.maximum(a[np.arange(len(d)), ax], 1e-30)
    v = d[np.arange(len(d)), (ax + 2) % 3] / np.maximum(a[np.arange(len(d)), ax], 1e-30)
    cu = np.clip(((u + 1) * 2).astype(np.int64), 0, 3).astype(np.uint64)
    cv = np.clip(((v + 1) * 2).astype(np.int64), 0, 3).astype(np.uint64)
    yield "dir cell(6x4x4),tiles", np.argsort((((face << np.uint64(4)) | (cu << np.uint64(2)) | cv) << np.uint64(40)) | tile, kind="stable")


def run(scene, rays_t, label, w, h, by_ray=False):
    base = None
    import itertools
    extra = sorted_orders(rays_t.cpu().numpy(), w, h) if by_ray else ()
    for name, idx in itertools.chain(orders(w, h), extra):
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
        print("%-22s %-24s %8d rays %.4f ms %6.0f Mrays/s  %s" % (label, name, src.shape[0], ms, src.shape[0] / ms / 1e3,
                                                                  "same bits" if crc == base else "DIFFERENT"), flush=True)


prim, bounce, _ = bench.config2_rays(rg, "leaf")
sc2 = ecb.Scene.from_image_file(os.path.join(ROOT, "scenes", "_built", "bomberman_s6c3_leaf.rimg")).upload(0)
run(sc2, torch.from_numpy(prim).to(dev), "config 2 primary", 1920, 1080)
run(sc2, torch.from_numpy(bounce).to(dev), "config 2 bounce", 1920, 1080, by_ray=True)
sc2.release()
sc4 = ecb.Scene.load(os.path.join(ROOT, "scenes", "_built", ge.CONFIG4_BLOB)).upload(0)
ys, xs = torch.meshgrid(torch.arange(0, 4320, 2, device=dev), torch.arange(0, 7680, 2, device=dev), indexing="ij")
rays = rgd.primary_rays(7680, 4320, device=dev, pixel_index=(ys * 7680 + xs).reshape(-1).to(torch.int64), **ge.CONFIG4_CAMERA)
run(sc4, rays, "config 4 (3840x2160)", 3840, 2160)
