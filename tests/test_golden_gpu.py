"""GPU test (-m gpu) that needs nothing but the committed fixtures: the CUDA path on the golden scene images against the
recorded outputs of the oracle port (bit-exact) and of the compiled reference (parity bar)."""
import glob
import os

import numpy as np
import pytest

import parity
from embree_compressed_b200 import raygen as rg
from test_oracle_cpu import CASES, COLS, _load, _with_results

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden_vectors(ecb, name):
    g, image, prim, bounce = _load(name)
    scene = ecb.Scene.from_image(image)
    for rays, key in ((prim, "primary"), (bounce, "bounce")):
        got = rays.copy()
        scene.intersect1M(got)
        assert (got[:, COLS].view(np.uint32) == g["port_" + key].view(np.uint32)).all(), "%s %s: CUDA != recorded port output" % (name, key)
        if "port_%s_ng" % key in g:  # mixed scenes: Ng of triangle hits is the real normal, bit-exact vs the port
            assert (got[:, 12:15].view(np.uint32) == g["port_%s_ng" % key].view(np.uint32)).all()
        untouched = [c for c in range(20) if c not in COLS and c not in (12, 13, 14)]
        assert (got[:, untouched].view(np.uint32) == rays[:, untouched].view(np.uint32)).all()
        cand = np.zeros((len(rays), 4), np.float32)
        cand[:, :2] = g["candidates_" + key]
        c = parity.classify(got, _with_results(rays, g["ref_" + key]), cand)
        assert c["fragile"] <= max(3, 1e-3 * len(rays)), (name, key, c["fragile"])
    occ = rg.rays_only(bounce)
    scene.occluded1M(occ)
    assert (occ[:, 8].view(np.uint32) == g["ref_occluded_tfar"].view(np.uint32)).all()
    scene.release()
