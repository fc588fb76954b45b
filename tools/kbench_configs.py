"""Developer tool: device-timed throughput of the kernels behind BASELINE configs 1, 3 and 4 (any-hit, instanced, mixed).

    python tools/kbench_configs.py [c1] [c3] [c4]

Scenes are committed here by the compiled reference's CPU builder (the host builder of this path; dev tool only,
needs oracle/_ref) and imported from live memory; rays resident on the device, CUDA events around each call.
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
from oracle import pyoracle as po

which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c3", "c4"]
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, rays, reps=6):
    src = torch.from_numpy(rays).cuda()
    ts = []
    for _ in range(reps):
        tmp = src.clone(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(tmp); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[1:])), tmp.cpu().numpy()


def report(name, rays, ms):
    active = int((rays[:, 3] <= rays[:, 8]).sum())
    print("%-44s %9d rays  %8.3f ms  %8.0f Mrays/s" % (name, active, ms, active / ms / 1e3), flush=True)


if "c1" in which:  # config 1 as the tutorial builds it: displaced cube S6/C4 + ground plane, 1024^2 primary
    ref = po.RefScene("leaf", 6, 4); ref.add_cube(1, 1.0); ref.add_ground_plane(); ref.commit()
    sc = ecb.Scene.import_bvh4(ref.root()[0], "leaf", 4, 6, po.ref_layout()["proj_col_major"], 0, ref.bounds())
    sc.attach_triangles4v(ref.triangle_root()[0]); sc.upload(0)
    rays = rg.primary_rays(1024, 1024, **rg.CUBE_CAMERA)
    ms, out = timed(lambda t: sc.intersect_device(t.data_ptr(), len(rays), 80, stream), rays)
    report("c1 cube S6/C4 + plane, primary (2 kernels)", rays, ms)
    plain = ecb.Scene.import_bvh4(ref.root()[0], "leaf", 4, 6, po.ref_layout()["proj_col_major"], 0, ref.bounds()).upload(0)
    ms, _ = timed(lambda t: plain.intersect_device(t.data_ptr(), len(rays), 80, stream), rays)
    report("c1 cube S6/C4 without plane, primary", rays, ms)

if "c3" in which:  # config 3: S=8 single object, 4K primary + shadow rays (any-hit)
    ref = po.RefScene("leaf", 8, 4); ref.add_cube(1, 1.0); ref.commit()
    sc = ecb.Scene.import_bvh4(ref.root()[0], "leaf", 4, 8, po.ref_layout()["proj_col_major"], 0, ref.bounds()).upload(0)
    rays = rg.primary_rays(3840, 2160, **rg.CUBE_CAMERA)
    ms, out = timed(lambda t: sc.intersect_device(t.data_ptr(), len(rays), 80, stream), rays)
    report("c3 cube S8/C4 4K primary", rays, ms)
    shadow = rg.shadow_rays(out, light_center=(3.0, 5.0, -2.0), light_radius=1.0)
    ms, _ = timed(lambda t: sc.occluded_device(t.data_ptr(), len(shadow), 48, stream), shadow)
    report("c3 shadow rays, rtcOccluded semantics", shadow, ms)
    import ctypes as C
    def exact(t):
        ecb._check(ecb._lib().ecbOccludedExactDevice(sc._h, 0xFFFFFFFF, t.data_ptr(), len(shadow), 48, stream), "exact")
    ms, _ = timed(exact, shadow)
    report("c3 shadow rays, exact any-hit (opt-in)", shadow, ms)

if "c4" in which:  # config 4: instanced bomberman + cube, cGrid, 8K frame (one rank's share = the whole frame here)
    from instancing_util import build_instanced, import_instanced
    top, objs = build_instanced(po, "grid", [("bomberman", 6, 3), ("cube", 6, 3)], grid=2, spacing=30.0)
    inst, scenes = import_instanced(ecb, po, top, objs)
    inst.upload(0)
    cam = dict(vp=(90.0, 70.0, -100.0), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=60.0)
    for (W, H) in ((1920, 1080), (7680, 4320)):
        rays = rg.primary_rays(W, H, **cam)
        ms, out = timed(lambda t: inst.intersect_device(t.data_ptr(), len(rays), 80, stream), rays, reps=4)
        hits = int((out.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum())
        report("c4 25 instances cGrid %dx%d primary (%d hits)" % (W, H, hits), rays, ms)
