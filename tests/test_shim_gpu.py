"""GPU test (-m gpu): the embree3 drop-in library.  tests/cpp/shim_demo is an application that only speaks the embree3
C API; it is linked against libembree3_b200.so (host builder = the compiled reference, traversal = CUDA) and its
results are compared with the reference tracing the same scene and rays through its own rtcIntersect1/rtcOccluded1."""
import os
import subprocess

import numpy as np
import pytest

import parity
from embree_compressed_b200 import raygen as rg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "cpp", "shim_demo")


@pytest.mark.parametrize("leaf,S,C", [("leaf", 5, 3), ("grid", 4, 2), ("box", 5, 2)])
def test_unmodified_embree3_application_runs_on_the_shim(ecb, reference, ref_scene, tmp_path, leaf, S, C):
    assert os.path.exists(DEMO), "tests/cpp/shim_demo not built (run __graft_entry__.build())"
    cam = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
    rays = rg.primary_rays(160, 128, **cam)
    n = len(rays)
    rays_path, out_path = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rays_path)
    env = dict(os.environ, ECB200_HOST_EMBREE=reference.REF_SO)
    r = subprocess.run([DEMO, leaf, str(S), str(C), rays_path, out_path], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.float32)
    n_pk = min(n // 16, 32) * 16
    n_one = min(n, 64)
    sizes = [n * 20, n_pk * 20, n_one * 20, n * 16]
    assert raw.size == sum(sizes)
    stream, packets, singles, occ = np.split(raw, np.cumsum(sizes)[:-1])
    stream, packets, singles, occ = stream.reshape(n, 20), packets.reshape(n_pk, 20), singles.reshape(n_one, 20), occ.reshape(n, 16)

    ref = ref_scene("cube", leaf, S, C)      # the same scene committed by the reference in this process
    port = reference.PortScene.from_ref(ref)
    truth = rays.copy()
    ref.intersect(truth)
    cand = port.candidates(rays)
    c = parity.assert_reference_parity(stream, truth, cand, "shim rtcIntersect1M %s" % leaf)
    assert c["hits_ref"] > 1000
    # packets: masked lanes (every 5th) come back untouched, active lanes equal the stream result
    lanes = np.arange(n_pk) % 16
    masked = lanes % 5 == 4
    assert (packets[masked].view(np.uint32) == rays[:n_pk][masked].view(np.uint32)).all()
    assert (packets[~masked].view(np.uint32) == stream[:n_pk][~masked].view(np.uint32)).all()
    assert (singles.view(np.uint32) == stream[:n_one].view(np.uint32)).all()
    # any-hit stream with stride 64: padding untouched, tfar == reference's rtcOccluded1
    o_truth = rg.rays_only(rays)
    ref.occluded(o_truth)
    assert (occ[:, 12:] == 7.0).all()
    assert (occ[:, :12].view(np.uint32) == o_truth.view(np.uint32)).all()


def test_instanced_scene_through_the_shim(ecb, reference, tmp_path):
    """rtcSetGeometryInstancedScene / rtcSetGeometryTransform on the shim: two instances of the displaced cube."""
    cam = dict(vp=(1.0, 2.5, -4.5), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=70.0)
    rays = rg.primary_rays(160, 128, **cam)
    n = len(rays)
    rays_path, out_path = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rays_path)
    env = dict(os.environ, ECB200_HOST_EMBREE=reference.REF_SO)
    r = subprocess.run([DEMO, "leaf", "5", "3", rays_path, out_path, "instanced"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    stream = np.fromfile(out_path, np.float32)[: n * 20].reshape(n, 20)
    obj = reference.RefScene("leaf", 5, 3)
    obj.add_cube(1, 1.0)
    obj.commit()
    top = reference.RefScene("leaf", 5, 3)
    top.add_instance(obj, np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, -1.6, 0, 0], np.float32))
    top.add_instance(obj, np.array([0.5, 0, 0, 0, 0.5, 0, 0, 0, 0.5, 1.4, 0.3, 0], np.float32))
    top.commit()
    truth = rays.copy()
    top.intersect(truth)
    port = reference.PortInstancedScene(top, [obj])
    c = parity.assert_reference_parity(stream, truth, port.candidates(rays), "shim instanced")
    hit = truth.view(np.uint32)[:, 18] != 0xFFFFFFFF
    assert c["hits_ref"] > 500 and set(np.unique(stream.view(np.uint32)[hit, 19])) == {0, 1}


def test_mixed_scene_through_the_shim(ecb, reference, tmp_path):
    """The tutorial scene as the tutorial builds it: subdivision cube + triangle-mesh ground plane in one
    RTCScene.  The shim receives two accels from the host and must reproduce AccelN::intersect / occluded."""
    cam = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
    rays = rg.primary_rays(160, 128, **cam)
    n = len(rays)
    rays_path, out_path = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rays_path)
    env = dict(os.environ, ECB200_HOST_EMBREE=reference.REF_SO)
    r = subprocess.run([DEMO, "leaf", "5", "3", rays_path, out_path, "plane"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.float32)
    stream = raw[: n * 20].reshape(n, 20)
    occ = raw[-n * 16:].reshape(n, 16)
    ref = reference.RefScene("leaf", 5, 3)
    ref.add_cube(1, 1.0)
    plane = ref.add_ground_plane()
    ref.commit()
    port = reference.PortScene.from_ref(ref)
    truth, want = rays.copy(), rays.copy()
    ref.intersect(truth)
    port.intersect(want)
    assert (stream.view(np.uint32) == want.view(np.uint32)).all(), "shim result differs from the oracle port"
    c = parity.assert_reference_parity(stream, truth, port.candidates(rays), "shim mixed scene")
    assert (truth.view(np.uint32)[:, 18] == plane).sum() > 1000 and c["hits_ref"] > 5000
    o_truth = rg.rays_only(rays)
    ref.occluded(o_truth)
    assert (occ[:, :12].view(np.uint32) == o_truth.view(np.uint32)).all()
    ref.free()


def test_all_stream_layouts_through_the_shim(ecb, reference, ref_scene, tmp_path):
    """rtcIntersect1Mp / NM / Np and rtcOccluded1Mp / NM / Np (include/embree3/rtcore_scene.h:112-142): every layout of
    the stream API must give what rtcIntersect1M / rtcOccluded1M give on the same rays."""
    cam = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
    rays = rg.primary_rays(160, 128, **cam)
    n = len(rays)
    rays_path, out_path = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rays_path)
    env = dict(os.environ, ECB200_HOST_EMBREE=reference.REF_SO)
    r = subprocess.run([DEMO, "leaf", "5", "3", rays_path, out_path, "streams"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.float32)
    n_pk, n_one = min(n // 16, 32) * 16, min(n, 64)
    sizes = [n * 20, n_pk * 20, n_one * 20, n * 16, n * 20, (n // 8) * 8 * 20, n * 20, n, n, (n // 4) * 4]
    assert raw.size == sum(sizes), (raw.size, sum(sizes))
    parts = np.split(raw, np.cumsum(sizes)[:-1])
    stream = parts[0].reshape(n, 20)
    occ = parts[3].reshape(n, 16)[:, 8]
    assert (stream.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum() > 5000
    for name, part, m in (("1Mp", parts[4], n), ("NM", parts[5], (n // 8) * 8), ("Np", parts[6], n)):
        got = part.reshape(m, 20)
        assert (got.view(np.uint32) == stream[:m].view(np.uint32)).all(), "rtcIntersect%s differs from rtcIntersect1M" % name
    for name, part in (("Np", parts[7]), ("1Mp", parts[8]), ("NM", parts[9])):
        assert (part.view(np.uint32) == occ[:len(part)].view(np.uint32)).all(), "rtcOccluded%s differs from rtcOccluded1M" % name


def test_rest_of_the_embree3_surface_through_the_shim(ecb, reference, tmp_path):
    """The embree3 calls that never touch a ray (buffers, device properties, builder API ...) are forwarded to the host;
    filter callbacks -- host code the reference runs per hit -- are refused with RTC_ERROR_INVALID_OPERATION and leave the
    rays untouched (rtcore_common.h:214-219, rtcore_geometry.h filter functions)."""
    cam = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
    rays = rg.primary_rays(160, 128, **cam)
    rays_path, out_path = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rays_path)
    env = dict(os.environ, ECB200_HOST_EMBREE=reference.REF_SO)
    base = subprocess.run([DEMO, "leaf", "4", "2", rays_path, out_path], env=env, capture_output=True, text=True, timeout=300)
    assert base.returncode == 0, base.stdout + base.stderr
    want = np.fromfile(out_path, np.float32)
    r = subprocess.run([DEMO, "leaf", "4", "2", rays_path, out_path, "plumbing"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "plumbing ok" in r.stdout
    # vertices handed over through an RTCBuffer give the scene the shared pointer gave
    assert (np.fromfile(out_path, np.float32).view(np.uint32) == want.view(np.uint32)).all()
