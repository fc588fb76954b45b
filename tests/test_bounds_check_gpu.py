"""GPU test (-m gpu): every kernel of the library on a -DECB_BOUNDS_CHECK build.

compute-sanitizer is closed on the B200 pool ("runs under it have left GPUs needing a reset"), so the memcheck evidence of
this repo is a library variant in which the kernels check every index they derive from scene data (node, block, item and
instance references, both traversal stacks, in-block node-word and leaf offsets, ray indices) against the array it indexes and
count violations (csrc/ecb_kernels.cu ECB_CHECK).  tools/sanitize_case.py drives all kernels (both traversal kernels, all leaf
kinds, any-hit, exact any-hit, mixed scenes, instancing, block construction) against that build in a subprocess; the test
demands bit-exact golden results and zero violations."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_kernels_run_clean_on_the_bounds_checking_build():
    sys.path.insert(0, ROOT)
    from embree_compressed_b200 import build as b
    lib = b.build_variant("check", ["ECB_BOUNDS_CHECK=1"], only_if_stale=True)
    env = dict(os.environ, ECB200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")], env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert "bounds check: enabled=1 violations=[0, 0, 0, 0, 0, 0, 0, 0]" in out, out[-3000:]
    assert r.returncode == 0, out[-3000:]
    assert "bit-exact vs golden port: False" not in out
