"""GPU tests (-m gpu) of the four-lanes-per-ray traversal kernel (csrc/ecb_quad_kernel.cuh): it must return exactly the bits
of the one-lane-per-ray kernel and of the recorded oracle port -- on the golden fixtures (all leaf kinds, mixed scenes,
primary and bounce rays, any-hit modes) and on incoherent rays over the benchmark scenes, with identical work counters."""
import os

import numpy as np
import pytest

from embree_compressed_b200 import raygen as rg
from test_oracle_cpu import CASES, COLS, _load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def quad(ecb):
    ecb.set_traversal_kernel("quad")
    assert ecb.traversal_kernel() == "quad"
    yield ecb
    ecb.set_traversal_kernel(None)


@pytest.mark.parametrize("name", CASES)
def test_quad_kernel_matches_golden_vectors(quad, name):
    ecb = quad
    g, image, prim, bounce = _load(name)
    scene = ecb.Scene.from_image(image)
    for rays, key in ((prim, "primary"), (bounce, "bounce")):
        got = rays.copy()
        scene.intersect1M(got)
        assert (got[:, COLS].view(np.uint32) == g["port_" + key].view(np.uint32)).all(), "%s %s: quad kernel != recorded port output" % (name, key)
        if "port_%s_ng" % key in g:
            assert (got[:, 12:15].view(np.uint32) == g["port_%s_ng" % key].view(np.uint32)).all()
        untouched = [c for c in range(20) if c not in COLS and c not in (12, 13, 14)]
        assert (got[:, untouched].view(np.uint32) == rays[:, untouched].view(np.uint32)).all()
    # exact any-hit runs through the closest-hit kernel: same answer from both kernels
    a, b = rg.rays_only(bounce), rg.rays_only(bounce)
    scene.occluded_exact1M(a)
    ecb.set_traversal_kernel("lane")
    scene.occluded_exact1M(b)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    scene.release()


@pytest.mark.parametrize("name", ["bomberman_s6c3_leaf", "bomberman_s6c3_grid", "bomberman_s6c3_box", "cube_s6c4_leaf"])
def test_quad_and_lane_kernels_agree_on_incoherent_rays(ecb, name):
    import torch
    path = os.path.join(ROOT, "scenes", "_built", name + ".rimg")
    if not os.path.exists(path):
        pytest.skip("scene image not built")
    scene = ecb.Scene.from_image_file(path).upload(0)
    b = scene.info()["bounds"]
    rays = rg.incoherent_rays(400000, b, seed=11)
    if name.startswith("bomberman"):   # the AABB is dominated by the ground quad: aim half of the rays at the figure
        lo, hi = np.array([-6.0, 0.0, -6.0]), np.array([6.0, 14.0, 6.0])
        rng = np.random.default_rng(12)
        tgt = lo + rng.random((200000, 3)) * (hi - lo)
        d = tgt - rays[:200000, 0:3]
        rays[:200000, 4:7] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays[::11, 3], rays[::11, 8] = 2.0, 1.0   # some inactive rays
    rays[5::13, 8] = 30.0                     # some finite tfar
    dev = torch.device("cuda", 0)
    res, cnt = {}, {}
    try:
        for mode in ("lane", "quad"):
            ecb.set_traversal_kernel(mode)
            t = torch.from_numpy(rays).to(dev)
            c = torch.zeros(6, dtype=torch.int64, device=dev)
            scene.intersect_device_stats(t.data_ptr(), len(rays), c.data_ptr())
            torch.cuda.synchronize()
            res[mode], cnt[mode] = t.cpu().numpy(), c.cpu().numpy()
            h = rays.copy()
            scene.intersect1M(h)
            assert (h.view(np.uint32) == res[mode].view(np.uint32)).all(), "%s: host-buffer path differs from the device path" % mode
    finally:
        ecb.set_traversal_kernel(None)
    assert (res["lane"].view(np.uint32) == res["quad"].view(np.uint32)).all()
    assert (cnt["lane"] == cnt["quad"]).all(), (cnt["lane"], cnt["quad"])
    assert cnt["lane"][5] > 1000, "test rays should hit something"
    scene.release()
