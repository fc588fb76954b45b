"""GPU tests (-m gpu) of the multi-GPU plumbing of the path: compact device results (what a frame gather
moves), concurrent host-buffer callers on one scene (the reference's tile-parallel usage, rtcore.cpp:287-304 takes no
lock), and -- on a box with at least two GPUs -- one host-buffer call fanned out over the devices of a scene
(ecbSceneUploadMulti) and the peer-copy gather of the C ABI.  Everything is held to the single-device bits.
"""
import os
import threading

import numpy as np
import pytest

from embree_compressed_b200 import raygen as rg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scene(ecb, name="cube_s5c3_leaf"):
    return ecb.Scene.from_image(np.fromfile(os.path.join(GOLDEN, name + ".rimg"), np.uint8))


def _rays(n_side=512):
    prim = rg.primary_rays(n_side, n_side, **rg.CUBE_CAMERA)
    prim[::7, 3], prim[::7, 8] = 2.0, 1.0   # inactive stream rays (tnear > tfar) stay untouched
    return prim


def test_compact_device_results_equal_the_in_place_records(ecb):
    import torch
    scene = _scene(ecb).upload(0)
    rays = _rays()
    n = len(rays)
    dev = torch.device("cuda", 0)
    a = torch.from_numpy(rays).to(dev)
    scene.intersect_device(a.data_ptr(), n)
    b = torch.from_numpy(rays).to(dev)
    geom = torch.full((n,), 123, dtype=torch.int32, device=dev)
    hit = torch.full((n, 4), -7.0, dtype=torch.float32, device=dev)
    scene.intersect_device_compact(b.data_ptr(), n, geom.data_ptr(), hit.data_ptr())
    torch.cuda.synchronize()
    assert (b.cpu().numpy().view(np.uint32) == rays.view(np.uint32)).all(), "compact mode must not write the ray records"
    want = a.cpu().numpy()
    wi = want.view(np.uint32)
    g, h = geom.cpu().numpy().view(np.uint32), hit.cpu().numpy()
    assert (g == wi[:, 18]).all()
    m = wi[:, 18] != 0xFFFFFFFF
    assert m.sum() > 1000
    assert (h[m].view(np.uint32) == want[m][:, [8, 15, 16, 17]].view(np.uint32)).all()
    assert (h[~m] == -7.0).all(), "misses write no hit record"
    # any-hit: geomID 0 = occluded
    r12 = rg.rays_only(rays)
    c = torch.from_numpy(r12).to(dev)
    scene.occluded_device(c.data_ptr(), n)
    d = torch.from_numpy(r12).to(dev)
    occ = torch.full((n,), 5, dtype=torch.int32, device=dev)
    scene.occluded_device_compact(d.data_ptr(), n, occ.data_ptr())
    torch.cuda.synchronize()
    assert ((occ.cpu().numpy() == 0) == np.isneginf(c.cpu().numpy()[:, 8])).all()


def test_concurrent_host_callers_get_their_own_staging_context(ecb):
    """Four threads on ONE scene, long and short calls mixed: every call leaves the single-caller bits."""
    scene = _scene(ecb).upload(0)
    rays = _rays(640)
    want = rays.copy()
    scene.intersect1M(want)
    errors = []

    def worker(k):
        try:
            for rep in range(3):
                m = (len(rays), 50000, 16, 300001)[(k + rep) % 4]
                got = rays[:m].copy()
                scene.intersect1M(got)
                if not (got.view(np.uint32) == want[:m].view(np.uint32)).all():
                    errors.append("thread %d rep %d: %d rays differ" % (k, rep, m))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_frame_pipeline_over_two_streams_leaves_the_one_stream_bits(ecb):
    """bench.py's FramePipeline: bounce(k) on a second stream behind an event, primary(k+1) overlapping it.  Launches of one
    scene that overlap on the device take their own ray-fetch cursors; every frame must hold the serial result."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    scene = _scene(ecb).upload(0)
    prim = _rays(768)
    n = len(prim)
    dev = torch.device("cuda", 0)
    first = torch.from_numpy(prim).to(dev)
    scene.intersect_device(first.data_ptr(), n)
    torch.cuda.synchronize()
    hits = first.cpu().numpy()
    bounce = rg.diffuse_rays(hits, -hits[:, 4:7], 10.0, sample=0)
    want_b = torch.from_numpy(bounce).to(dev)
    scene.intersect_device(want_b.data_ptr(), n)
    torch.cuda.synchronize()
    pipe = bench.FramePipeline(torch, torch.cuda.current_stream())
    frames = [(torch.from_numpy(prim).to(dev), torch.from_numpy(bounce).to(dev)) for _ in range(12)]
    torch.cuda.synchronize()
    for p, b in frames:
        pipe.frame(lambda st, p=p: scene.intersect_device(p.data_ptr(), n, 80, st.cuda_stream),
                   lambda st, b=b: scene.intersect_device(b.data_ptr(), n, 80, st.cuda_stream))
    torch.cuda.synchronize()
    for k, (p, b) in enumerate(frames):
        assert bool((p.view(torch.int32) == first.view(torch.int32)).all()), "primary pass of frame %d" % k
        assert bool((b.view(torch.int32) == want_b.view(torch.int32)).all()), "bounce pass of frame %d" % k


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


def test_one_host_call_fans_out_over_the_devices_of_a_scene(ecb):
    import torch
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    n_dev = min(4, torch.cuda.device_count())
    single = _scene(ecb).upload(0)
    multi = _scene(ecb).upload_multi(list(range(n_dev)))
    assert multi.resident_devices() == list(range(n_dev))
    rays = _rays(1024)   # 1 Mi rays: above the fan-out threshold
    want = rays.copy()
    single.intersect1M(want)
    for pinned in (False, True):
        if pinned:
            buf, ptr = ecb.host_alloc(rays.nbytes)
            got = buf.view(np.float32).reshape(rays.shape)
            got[:] = rays
        else:
            got = rays.copy()
        multi.intersect1M(got)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), "fan-out over %d devices differs (pinned=%s)" % (n_dev, pinned)
        if pinned:
            ecb.host_free(ptr)
    occ_want, occ_got = rg.rays_only(rays), rg.rays_only(rays)
    single.occluded1M(occ_want)
    multi.occluded1M(occ_got)
    assert (occ_got.view(np.uint32) == occ_want.view(np.uint32)).all()
    # device-buffer calls run on the device that owns the buffer
    for d in range(n_dev):
        with torch.cuda.device(d):
            t = torch.from_numpy(rays[:100000]).to(torch.device("cuda", d))
            multi.intersect_device(t.data_ptr(), t.shape[0])
            torch.cuda.synchronize()
            assert (t.cpu().numpy().view(np.uint32) == want[:100000].view(np.uint32)).all(), "device %d" % d


def test_peer_copy_into_another_devices_buffer(ecb):
    """The gather primitive of the C ABI inside one process: device 1 writes its compact records into device 0's arrays."""
    import torch
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    scene = _scene(ecb).upload_multi([0, 1])
    rays = _rays()
    n = len(rays)
    with torch.cuda.device(0):
        frame_geom = ecb.device_alloc(4 * n)
        a = torch.from_numpy(rays).to("cuda:0")
        scene.intersect_device(a.data_ptr(), n)
        torch.cuda.synchronize()
    with torch.cuda.device(1):
        b = torch.from_numpy(rays).to("cuda:1")
        geom = torch.empty(n, dtype=torch.int32, device="cuda:1")
        hit = torch.empty((n, 4), dtype=torch.float32, device="cuda:1")
        s = torch.cuda.current_stream().cuda_stream
        scene.intersect_device_compact(b.data_ptr(), n, geom.data_ptr(), hit.data_ptr(), stream=s)
        ecb.peer_copy_async(frame_geom, geom.data_ptr(), 4 * n, s)
        torch.cuda.synchronize()
    with torch.cuda.device(0):
        back = torch.empty(n, dtype=torch.int32, device="cuda:0")
        ecb.peer_copy_async(back.data_ptr(), frame_geom, 4 * n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (back.cpu().numpy().view(np.uint32) == a.cpu().numpy().view(np.uint32)[:, 18]).all()
        ecb.device_free(frame_geom)


@pytest.mark.parametrize("name", ["cube_s5c3_leaf", "cube_s4c2_grid", "cubeplane_s4c2_leaf"])
def test_soa_stream_equals_the_aos_stream(ecb, name):
    """ecbIntersectNp / ecbOccludedNp (rtcIntersectNp / rtcOccludedNp): the arrays go to the device as they are and the kernels
    read them coalesced; the results must be the bits of the AoS call -- pinned and pageable arrays, long and short streams."""
    scene = _scene(ecb, name).upload(0)
    rays = _rays(640)
    for n in (len(rays), 20000, 7):
        want = rays[:n].copy()
        scene.intersect1M(want)
        for pinned in (False, True):
            ptrs = []

            def alloc(nbytes):
                buf, p = ecb.host_alloc(nbytes)
                ptrs.append(p)
                return buf, p
            soa = ecb.soa_from_rayhits(rays[:n], alloc if pinned else None)
            scene.intersectNp(soa, n)
            got = ecb.rayhits_from_soa(soa, n)
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), "%s n=%d pinned=%s" % (name, n, pinned)
            for p in ptrs:
                ecb.host_free(p)
        occ_want = rg.rays_only(rays[:n])
        scene.occluded1M(occ_want)
        soa = ecb.soa_from_rayhits(rays[:n])
        scene.occludedNp(soa, n)
        assert (soa["tfar"][:n].view(np.uint32) == occ_want[:, 8].view(np.uint32)).all()
    scene.release()


def test_device_soa_input_equals_the_aos_records(ecb):
    """ecbIntersectDeviceSoA: eight device arrays read coalesced -- same hits as the AoS records, with both traversal kernels."""
    import torch
    scene = _scene(ecb).upload(0)
    rays = _rays()
    n = len(rays)
    dev = torch.device("cuda", 0)
    a = torch.from_numpy(rays).to(dev)
    scene.intersect_device(a.data_ptr(), n)
    torch.cuda.synchronize()
    want = a.cpu().numpy()
    cols = [0, 1, 2, 3, 4, 5, 6, 8]
    soa = [torch.from_numpy(np.ascontiguousarray(rays[:, c])).to(dev) for c in cols]
    try:
        for mode in ("lane", "quad"):
            ecb.set_traversal_kernel(mode)
            geom = torch.full((n,), 9, dtype=torch.int32, device=dev)
            hit = torch.zeros((n, 4), dtype=torch.float32, device=dev)
            scene.intersect_device_soa([t.data_ptr() for t in soa], n, geom.data_ptr(), hit.data_ptr())
            torch.cuda.synchronize()
            g, h = geom.cpu().numpy().view(np.uint32), hit.cpu().numpy()
            assert (g == want.view(np.uint32)[:, 18]).all(), mode
            m = g != 0xFFFFFFFF
            assert (h[m].view(np.uint32) == want[m][:, [8, 15, 16, 17]].view(np.uint32)).all(), mode
    finally:
        ecb.set_traversal_kernel(None)
