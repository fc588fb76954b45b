"""Developer tool: end-to-end Mrays/s of the SoA stream entry point (ecbIntersectNp == rtcIntersectNp) next to the AoS one
(ecbIntersect1M) on the bench step, pinned caller memory.    python tools/e2e_np.py [N devices] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prim, bounce, _ = bench.config2_rays(rg, "leaf")
scene = ecb.Scene.from_image_file(bench.scene_path("leaf"))
scene.upload_multi(list(range(n_dev))) if n_dev > 1 else scene.upload(0)
n_rays = len(prim) + int((bounce[:, 3] <= bounce[:, 8]).sum())
want_p, want_b = prim.copy(), bounce.copy()
scene.intersect1M(want_p); scene.intersect1M(want_b)
for mem in ("pinned", "pageable"):
    alloc = (lambda nb: ecb.host_alloc(nb)) if mem == "pinned" else None
    sp, sb = ecb.soa_from_rayhits(prim, alloc), ecb.soa_from_rayhits(bounce, alloc)
    names = ecb.Scene.RAY_FIELDS + ecb.Scene.HIT_FIELDS
    ts = []
    for i in range(2 + reps):
        for k, f in enumerate(names):   # fresh rays
            sp[f].view(np.float32)[:] = prim[:, k]; sb[f].view(np.float32)[:] = bounce[:, k]
        t0 = time.perf_counter()
        scene.intersectNp(sp, len(prim)); scene.intersectNp(sb, len(bounce))
        ts.append(time.perf_counter() - t0)
    ok = (ecb.rayhits_from_soa(sp, len(prim)).view(np.uint32) == want_p.view(np.uint32)).all() and \
         (ecb.rayhits_from_soa(sb, len(bounce)).view(np.uint32) == want_b.view(np.uint32)).all()
    t = float(np.mean(ts[2:]))
    print("devices=%d ecbIntersectNp %-8s %7.1f Mrays/s (%.3f ms per step, best %.3f)  bit-identical to ecbIntersect1M: %s"
          % (n_dev, mem, n_rays / t / 1e6, t * 1e3, min(ts[2:]) * 1e3, bool(ok)), flush=True)

# device-timed: SoA arrays vs AoS records as kernel input (compact results both)
import torch
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
single = ecb.Scene.from_image_file(bench.scene_path("leaf")).upload(0)
for what, rays in (("primary", prim), ("bounce", bounce)):
    n = len(rays)
    aos = torch.from_numpy(rays).to(dev)
    soa = [torch.from_numpy(np.ascontiguousarray(rays[:, c])).to(dev) for c in (0, 1, 2, 3, 4, 5, 6, 8)]
    geom = torch.empty(n, dtype=torch.int32, device=dev); hit = torch.empty((n, 4), dtype=torch.float32, device=dev)
    res = {}
    for mode, fn in (("AoS 80-byte records", lambda: single.intersect_device_compact(aos.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), 80, stream)),
                     ("SoA arrays", lambda: single.intersect_device_soa([t.data_ptr() for t in soa], n, geom.data_ptr(), hit.data_ptr(), stream))):
        ts = []
        for _ in range(8):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[mode] = float(np.median(ts[2:]))
    print("device input %-8s: %s" % (what, ", ".join("%s %.4f ms" % kv for kv in res.items())), flush=True)
