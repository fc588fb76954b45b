"""-m gpu: the three transfer strategies of the host-buffer entry points (csrc/ecb_api.cpp host_path(): packed /
compact / copy) give the same bits, for pinned and pageable caller memory, contiguous and padded records, with many more
chunks in flight than pipeline stages.  Needs only the committed fixtures."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("path", ["packed", "compact", "copy"])
@pytest.mark.parametrize("scene,chunk,threads", [("cube_s4c2_leaf", "1024", "3"), ("cube_s4c2_leaf", "65536", "8"),
                                                 ("cube_s4c2_leaf", "8192", "6"),
                                                 ("cubeplane_s4c2_leaf", "2048", "4")])   # a mixed scene
def test_host_transfer_strategies(ecb, path, scene, chunk, threads):
    env = dict(os.environ, ECB200_HOST_PATH=path, ECB200_HOST_CHUNK=chunk, ECB200_HOST_THREADS=threads)
    r = subprocess.run([sys.executable, os.path.join(HERE, "host_paths_case.py"), scene, "5"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host paths ok" in r.stdout
