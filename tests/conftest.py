"""pytest configuration: the `gpu` marker and shared scene fixtures.

`-m "not gpu"`: oracle vs golden vectors / compiled reference, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`      : parity of the CUDA path (through the C ABI) against the oracle on a real B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger sizes")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ecb():
    import __graft_entry__ as g
    g.build_product()
    import embree_compressed_b200 as ecb
    return ecb


@pytest.fixture(scope="session")
def oracle():
    """The oracle bindings; builds the C port on first use.  The compiled reference is optional here
    (tests that need it use the `reference` fixture)."""
    from oracle import pyoracle
    pyoracle.build_port()
    return pyoracle


@pytest.fixture(scope="session")
def reference(oracle):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libembree3_ref.so not built (needs /root/reference at build time)")
    return oracle


_SCENES = {}


@pytest.fixture(scope="session")
def ref_scene(reference):
    """Factory: ref_scene(kind, leaf, S, C) -> committed RefScene (cached for the session)."""
    def make(kind, leaf, S, C):
        key = (kind, leaf, S, C)
        if key not in _SCENES:
            s = reference.RefScene(leaf, S, C)
            if kind == "cube":
                s.add_cube(1, 1.0)
            elif kind == "cube_flat":
                s.add_cube(0, 0.0)
            elif kind == "bomberman":
                s.add_obj()
            else:
                raise ValueError(kind)
            s.commit()
            _SCENES[key] = s
        return _SCENES[key]
    return make
