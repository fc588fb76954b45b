"""Run by tests/test_host_paths_gpu.py in a subprocess (the transfer strategy and the pipeline chunk are read from the
environment once per process): the host-buffer entry points on a golden scene, pinned and pageable caller memory, 80- and
96-byte strides, intersect and occluded, many more chunks than pipeline stages -- every result bit-exact against the
recorded output of the oracle port / the compiled reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import embree_compressed_b200 as ecb  # noqa: E402
from embree_compressed_b200 import raygen as rg  # noqa: E402
from test_oracle_cpu import COLS, _load  # noqa: E402


def main(name, reps):
    g, image, prim, bounce = _load(name)
    scene = ecb.Scene.from_image(image)
    rays = np.concatenate([prim, bounce] * reps)   # reps x 8192 rays: the pool path, dozens of chunks at ECB200_HOST_CHUNK=1024
    want = np.concatenate([g["port_primary"], g["port_bounce"]] * reps)
    untouched = [c for c in range(20) if c not in COLS and c not in (12, 13, 14)]
    n = len(rays)
    for memory in ("pageable", "pinned"):
        for width in (20, 24):
            if memory == "pinned":
                buf, ptr = ecb.host_alloc(n * width * 4)
                arr = buf.view(np.float32).reshape(n, width)
            else:
                arr, ptr = np.empty((n, width), np.float32), None
            arr[:] = 7.0
            arr[:, :20] = rays
            scene.intersect1M(arr)
            assert (arr[:, COLS].view(np.uint32) == want.view(np.uint32)).all(), (memory, width, "results")
            assert (arr[:, untouched].view(np.uint32) == rays[:, untouched].view(np.uint32)).all(), (memory, width, "inputs modified")
            assert (arr[:, 20:] == 7.0).all(), (memory, width, "padding modified")
            hit = arr.view(np.uint32)[:, 18] != 0xFFFFFFFF
            if "port_primary_ng" in g:   # mixed scene: triangle hits carry their real normal, compressed hits the dummy (1,0,0)
                want_ng = np.concatenate([g["port_primary_ng"], g["port_bounce_ng"]] * reps)
                assert (arr[hit, 12:15].view(np.uint32) == want_ng[hit].view(np.uint32)).all(), (memory, width, "Ng")
                assert len(np.unique(arr.view(np.uint32)[hit, 18])) >= 2, "both accels must be hit"
            else:
                assert (arr[hit, 12:15] == np.float32([1, 0, 0])).all()
            assert (arr[~hit, 12:15].view(np.uint32) == rays[~hit, 12:15].view(np.uint32)).all()
            if ptr is not None:
                ecb.host_free(ptr)
        # any-hit: 48-byte RTCRay records
        occ_rays = rg.rays_only(np.concatenate([bounce] * (2 * reps)))
        occ_want = np.concatenate([g["ref_occluded_tfar"]] * (2 * reps))
        if memory == "pinned":
            buf, ptr = ecb.host_alloc(occ_rays.nbytes)
            occ = buf.view(np.float32).reshape(occ_rays.shape)
            occ[:] = occ_rays
        else:
            occ, ptr = occ_rays.copy(), None
        scene.occluded1M(occ)
        assert (occ[:, 8].view(np.uint32) == occ_want.view(np.uint32)).all(), (memory, "occluded")
        rest = [c for c in range(12) if c != 8]
        assert (occ[:, rest].view(np.uint32) == occ_rays[:, rest].view(np.uint32)).all()
        if ptr is not None:
            ecb.host_free(ptr)
    scene.release()
    print("host paths ok: %s, %d rays, path=%s chunk=%s" % (name, n, os.environ.get("ECB200_HOST_PATH", "packed"), os.environ.get("ECB200_HOST_CHUNK")))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
