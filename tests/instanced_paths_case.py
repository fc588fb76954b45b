"""Run by tests/test_instancing_gpu.py in a subprocess (the transfer strategy is read from the environment once per
process): an instancing scene through the host-buffer entry points -- closest hit with per-ray instID and any-hit,
pageable and pinned caller memory -- bit-exact against the oracle port.  Needs the compiled reference (oracle/_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import embree_compressed_b200 as ecb  # noqa: E402
from embree_compressed_b200 import raygen as rg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from instancing_util import build_instanced, import_instanced  # noqa: E402

CAM = dict(vp=(9.0, 7.0, -10.0), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=70.0)


def main(leaf):
    top, objs = build_instanced(po, leaf, [("cube", 4, 2), ("cube", 4, 3)])
    port = po.PortInstancedScene(top, objs)
    inst, scenes = import_instanced(ecb, po, top, objs)
    rays = rg.primary_rays(320, 200, **CAM)
    want = rays.copy()
    port.intersect(want)
    occ_want = rg.rays_only(rays)
    port.occluded(occ_want)
    hits = want.view(np.uint32)[:, 18] != 0xFFFFFFFF
    assert hits.sum() > 3000 and len(np.unique(want.view(np.uint32)[hits, 19])) > 10
    for memory in ("pageable", "pinned"):
        if memory == "pinned":
            buf, ptr = ecb.host_alloc(rays.nbytes)
            got = buf.view(np.float32).reshape(rays.shape)
            obuf, optr = ecb.host_alloc(occ_want.nbytes)
            og = obuf.view(np.float32).reshape(occ_want.shape)
        else:
            got, og, ptr, optr = np.empty_like(rays), np.empty_like(occ_want), None, None
        got[:] = rays
        inst.intersect1M(got, inst_id=7)   # the context's instID is overridden by the instance's geomID on a hit
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), (memory, "closest hit")
        og[:] = rg.rays_only(rays)
        inst.occluded1M(og)
        assert (og.view(np.uint32) == occ_want.view(np.uint32)).all(), (memory, "any hit")
        if ptr is not None:
            ecb.host_free(ptr)
            ecb.host_free(optr)
    print("instanced host paths ok: %s, %d rays, %d hits, path=%s" % (leaf, len(rays), int(hits.sum()), os.environ.get("ECB200_HOST_PATH", "auto")))


if __name__ == "__main__":
    main(sys.argv[1])
