"""CPU tests (-m "not gpu") of the block-construction oracle (oracle/cbvh_build_port.c): the restated
CompressedBVH constructor against the blocks the reference itself committed -- golden images first (no reference
needed), then live scenes built by the compiled reference."""
import os

import numpy as np
import pytest

import build_util as bu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = {"grid": 0, "leaf": 1, "box": 2}


def _golden(name):
    return np.fromfile(os.path.join(GOLDEN, name + ".rimg"), np.uint8)


@pytest.mark.parametrize("leaf", ["grid", "leaf", "box"])
def test_builder_port_reproduces_golden_blocks(oracle, leaf):
    """cube_s4c2_{grid,leaf,box}.rimg are three commits of the same displaced cube: the cGrid image holds every block's
    displaced vertex grid, from which the port must rebuild the nodes / leaves / homography / frustum box of all three."""
    inp, blocks = bu.inputs_from_images(_golden("cube_s4c2_grid"), _golden("cube_s4c2_" + leaf))
    built = bu.port_build(inp["grids"], inp["space"], inp["world"], KINDS[leaf], 2)
    bad = bu.compare_with_reference(blocks, built, KINDS[leaf])
    assert len(blocks) == 96 and not any(bad.values()), bad
    if leaf != "grid":   # the displaced cube has blocks where the homography is not applicable: both paths are covered
        oks = [o["patch_ok"] for o in built]
        assert 0 < sum(oks) < len(oks)
    else:
        assert not any(o["patch_ok"] for o in built), "cGrid never uses the projective mapping (compressed.h:148)"


@pytest.mark.parametrize("kind,S,C", [("cube", 5, 3), ("cube", 5, 1), ("cube", 6, 4), ("bomberman", 4, 2)])
def test_builder_port_matches_compiled_reference(reference, kind, S, C):
    def image(leaf):
        r = reference.RefScene(leaf, S, C)
        r.add_cube(1, 1.0) if kind == "cube" else r.add_obj()
        r.commit()
        img = r.dump_image()
        r.free()
        return img
    grid = image("grid")
    for leaf in ("grid", "leaf", "box"):
        inp, blocks = bu.inputs_from_images(grid, grid if leaf == "grid" else image(leaf))
        built = bu.port_build(inp["grids"], inp["space"], inp["world"], KINDS[leaf], C)
        bad = bu.compare_with_reference(blocks, built, KINDS[leaf])
        assert len(blocks) > 0 and not any(bad.values()), (kind, S, C, leaf, bad)
