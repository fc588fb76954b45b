"""CPU test (no GPU): the product's host pipeline under ThreadSanitizer.

csrc/ecb_api.cpp and csrc/ecb_import.cpp are compiled UNMODIFIED with g++ -fsanitize=thread against tests/mock_cuda (a
host-only stand-in for the CUDA runtime in which every stream is a worker thread, and fake kernels whose results are a
function of the ray alone).  tests/cpp/host_pipeline_check.cpp then drives the host-buffer entry points -- plain and mixed
scenes, pinned and pageable caller memory, dense and padded strides, 1 ... 150 001 rays in 4 Ki-ray chunks (many more
chunks than pipeline stages), closest hit / any hit / exact any hit, concurrent callers, scenes resident on three (mock)
devices with calls that fan out over them -- and checks every word of every
record against the API contract.  ThreadSanitizer turns a staging buffer reused too early, a scatter job that outlives
its call or an unprotected counter into a report, which fails the test.  This is test infrastructure: the mock is only
ever found through this build line's -I, and the product never links it."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "embree_compressed_b200", "csrc")
EXE = os.path.join(HERE, "cpp", "host_pipeline_check")
SOURCES = [os.path.join(CSRC, "ecb_api.cpp"), os.path.join(CSRC, "ecb_import.cpp"),
           os.path.join(HERE, "mock_cuda", "mock_kernels.cpp"), os.path.join(HERE, "cpp", "host_pipeline_check.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "mock_cuda", f) for f in ("cuda_runtime.h", "fake_trace.h")] + \
    [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]


def _build(exe, flags):
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in DEPS):
        cmd = ["g++", "-std=c++17", "-O1", "-g"] + flags + ["-I" + os.path.join(HERE, "mock_cuda"), "-I" + os.path.join(ROOT, "include"),
                                                          "-o", exe] + SOURCES + ["-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return r.stderr
    return None


def _run(exe, path):
    # three mock devices: the second half of the check uploads the scenes to all of them (ecbSceneUploadMulti) and makes
    # calls long enough to fan out, one ray range per device
    env = dict(os.environ, ECB200_HOST_PATH=path, ECB200_HOST_CHUNK="4096", ECB200_HOST_THREADS="4", MOCK_CUDA_DEVICES="3",
               TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    r = subprocess.run([exe, os.path.join(HERE, "golden", "cube_s4c2_leaf.rimg"), os.path.join(HERE, "golden", "cubeplane_s4c2_leaf.rimg")],
                       env=env, capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("path", ["packed", "compact", "copy"])
def test_host_pipeline_under_thread_sanitizer(path):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    err = _build(EXE, ["-fsanitize=thread"])
    sanitized = err is None
    if sanitized:
        rc, stdout, out = _run(EXE, path)
        # a ThreadSanitizer runtime that cannot start in this container (address-space layout) is not a finding
        sanitized = "FATAL: ThreadSanitizer" not in out
    if not sanitized:  # no usable sanitizer here: the functional half of the check still runs
        err = _build(EXE + "_plain", [])
        assert err is None, err
        rc, stdout, out = _run(EXE + "_plain", path)
    assert "WARNING: ThreadSanitizer" not in out, out[-4000:]
    assert rc == 0, out[-4000:]
    assert "host pipeline check: ok (path=%s)" % path in stdout
