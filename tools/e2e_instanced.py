"""Developer probe: e2e throughput of the host-buffer entry point on an instancing scene (BASELINE config 4 in small:
25 instances of bomberman + displaced cube, S6/C3 cGrid, 1920x1080 primary rays), for the transfer strategy named by
ECB200_HOST_PATH.  Needs the compiled reference (oracle/_ref) to build the scene."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
from oracle import pyoracle as po
from instancing_util import build_instanced, import_instanced

top, objs = build_instanced(po, "grid", [("bomberman", 6, 3), ("cube", 6, 3)], spacing=40.0)
inst, scenes = import_instanced(ecb, po, top, objs)
b = top.bounds()
c = [(b[i] + b[i + 3]) / 2 for i in range(3)]
cam = dict(vp=(c[0] + 120.0, c[1] + 90.0, c[2] - 140.0), vi=tuple(c), vu=(0.0, 1.0, 0.0), fov=60.0)
rays = rg.primary_rays(1920, 1080, **cam)
for memory in ("pinned", "pageable"):
    if memory == "pinned":
        buf, ptr = ecb.host_alloc(rays.nbytes)
        h = buf.view(np.float32).reshape(rays.shape)
    else:
        h, ptr = np.empty_like(rays), None
    ts = []
    for _ in range(7):
        h[:] = rays
        t0 = time.perf_counter()
        inst.intersect1M(h)
        ts.append(time.perf_counter() - t0)
    hits = int((h.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum())
    print("instanced e2e %-8s path=%-7s: %.2f ms  %.0f Mrays/s  (%d hits)" % (memory, os.environ.get("ECB200_HOST_PATH", ecb.host_path_for(memory == "pinned")),
                                                                          1e3 * float(np.median(ts[2:])), len(rays) / float(np.median(ts[2:])) / 1e6, hits), flush=True)
    if ptr is not None:
        ecb.host_free(ptr)
