"""GPU tests (-m gpu) for the remaining BASELINE.json configurations (3 and 5) at reduced and full ray counts."""
import zlib

import numpy as np
import pytest

import parity
from embree_compressed_b200 import raygen as rg

pytestmark = pytest.mark.gpu


def _pair(ecb, oracle, ref):
    root, _ = ref.root()
    scene = ecb.Scene.import_bvh4(root, ref.leaf, ref.C, ref.S, oracle.ref_layout()["proj_col_major"], 0, ref.bounds())
    return scene, oracle.PortScene.from_ref(ref)


def test_config3_high_res_displacement_with_shadow_rays(ecb, reference):
    """Config 3: single object, S=8 (256x256 grid per face), C=4, primary rays + incoherent shadow rays traced with
    rtcOccluded1M semantics.  1280x720 here; the 4K frame runs in the slow test below."""
    ref = reference.RefScene("leaf", 8, 4)
    ref.add_cube(1, 1.0)
    ref.commit()
    scene, port = _pair(ecb, reference, ref)
    assert scene.info()["n_blocks"] == 6 * 4 ** (8 - 4)
    rays = rg.primary_rays(1280, 720, **rg.CUBE_CAMERA)
    got, want, truth = rays.copy(), rays.copy(), rays.copy()
    scene.intersect1M(got)
    port.intersect(want)
    ref.intersect(truth)
    assert len(parity.bit_exact_mismatches(got, want)) == 0
    parity.assert_reference_parity(got, truth, port.candidates(rays), "config 3 primary")
    shadow = rg.shadow_rays(truth, light_center=(3.0, 5.0, -2.0), light_radius=1.0)
    g, t = shadow.copy(), shadow.copy()
    scene.occluded1M(g)
    ref.occluded(t)
    assert (g.view(np.uint32) == t.view(np.uint32)).all(), "rtcOccluded1M differs from the reference's rtcOccluded1"
    occluded = int((t[:, 8] < 0).sum())
    assert 0 < occluded < len(t)
    scene.release()
    ref.free()


@pytest.mark.slow
def test_config3_full_4k_frame(ecb, reference):
    ref = reference.RefScene("leaf", 8, 4)
    ref.add_cube(1, 1.0)
    ref.commit()
    scene, port = _pair(ecb, reference, ref)
    rays = rg.primary_rays(3840, 2160, **rg.CUBE_CAMERA)
    got = rays.copy()
    scene.intersect1M(got)
    # a random 1/16 of the frame against the oracle, the whole frame through size-independent properties
    pick = np.random.default_rng(1).choice(len(rays), len(rays) // 16, replace=False)
    want = rays[pick].copy()
    port.intersect(want)
    assert (want.view(np.uint32) == got[pick].view(np.uint32)).all()
    again = rays[::-1].copy()
    scene.intersect1M(again)
    assert (again[::-1].view(np.uint32) == got.view(np.uint32)).all(), "result depends on ray order"
    shadow = rg.shadow_rays(got, light_center=(3.0, 5.0, -2.0), light_radius=1.0)
    g = shadow.copy()
    scene.occluded1M(g)
    t = shadow[pick].copy()
    ref.occluded(t)
    assert (g[pick].view(np.uint32) == t.view(np.uint32)).all()
    scene.release()
    ref.free()


def test_config5_incoherent_path_tracing_batch(ecb, reference, ref_scene):
    """Config 5 at 1/64 size: cosine-diffuse rays, many hemisphere samples per primary hit, one flat batch."""
    import torch
    ref = ref_scene("bomberman", "leaf", 6, 3)
    scene, port = _pair(ecb, reference, ref)
    prim = rg.primary_rays(320, 180, **rg.BOMBERMAN_CAMERA)
    truth = prim.copy()
    ref.intersect(truth)
    b = ref.bounds()
    diag = float(np.linalg.norm(b[3:] - b[:3]))
    batch = rg.diffuse_rays(truth, ref.hit_normals(truth), diag, samples_per_hit=18)   # 1.04 M rays
    assert len(batch) == 18 * len(prim)
    d = torch.from_numpy(batch).cuda()
    scene.intersect_device(d.data_ptr(), len(batch), 80, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    want = batch.copy()
    port.intersect(want)
    assert len(parity.bit_exact_mismatches(got, want)) == 0
    ref_out = batch.copy()
    ref.intersect(ref_out)
    parity.assert_reference_parity(got, ref_out, port.candidates(batch), "config 5 batch")
    scene.release()


@pytest.mark.slow
def test_config5_full_64M_rays_determinism(ecb, reference, ref_scene):
    """All 64 Mi rays on the device: two runs give the same bits (the dynamic scheduler assigns rays to lanes in a
    different order every time), and a random sample agrees with the oracle."""
    import torch
    ref = ref_scene("bomberman", "leaf", 6, 3)
    scene, port = _pair(ecb, reference, ref)
    prim = rg.primary_rays(1920, 1080, **rg.BOMBERMAN_CAMERA)
    first = prim.copy()
    scene.intersect1M(first)
    normals = np.zeros((len(prim), 3), np.float32)
    hit = first.view(np.uint32)[:, 18] != 0xFFFFFFFF
    normals[hit] = ref.hit_normals(first)[hit]
    b = ref.bounds()
    diag = float(np.linalg.norm(b[3:] - b[:3]))
    n_target = 64 * 1024 * 1024
    per = -(-n_target // len(prim))
    dev = torch.empty((per * len(prim), 20), dtype=torch.float32, device="cuda")
    host_chunks = []
    for k in range(per):
        chunk = rg.diffuse_rays(first, normals, diag, sample=k)
        dev[k * len(prim):(k + 1) * len(prim)] = torch.from_numpy(chunk).cuda()
        if k in (0, per // 2, per - 1):
            host_chunks.append((k, chunk))
    n = n_target
    pristine = dev[:n].clone()
    stream = torch.cuda.current_stream().cuda_stream
    scene.intersect_device(dev.data_ptr(), n, 80, stream)
    torch.cuda.synchronize()
    run2 = pristine.clone()
    scene.intersect_device(run2.data_ptr(), n, 80, stream)
    torch.cuda.synchronize()
    assert torch.equal(dev[:n].view(torch.int32), run2.view(torch.int32))
    for k, chunk in host_chunks:
        lo = k * len(prim)
        hi = min(lo + len(prim), n)
        if hi <= lo:
            continue
        pick = np.random.default_rng(k).choice(hi - lo, 200_000, replace=False)
        want = chunk[pick].copy()
        port.intersect(want)
        got = dev[lo:hi].cpu().numpy()[pick]
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
    scene.release()


@pytest.mark.parametrize("leaf", ["leaf", "grid", "box"])
def test_exact_shadow_query_opt_in(ecb, reference, ref_scene, leaf):
    """the opt-in exact any-hit.  It is NOT parity (the reference answers `occluded` from block boxes);
    it must agree with the closest-hit kernel: occluded <=> a closest hit exists inside [tnear, tfar]."""
    ref = ref_scene("bomberman", leaf, 6, 3)
    scene, port = _pair(ecb, reference, ref)
    prim = rg.primary_rays(480, 270, **rg.BOMBERMAN_CAMERA)
    first = prim.copy()
    scene.intersect1M(first)
    shadow = rg.shadow_rays(first, light_center=(40.0, 60.0, 25.0), light_radius=4.0)
    exact = shadow.copy()
    scene.occluded_exact1M(exact)
    as_hits = rg.new_rayhits(len(shadow))
    as_hits[:, :12] = shadow
    closest = as_hits.copy()
    scene.intersect1M(closest)
    active = shadow[:, 3] <= shadow[:, 8]
    want = (closest.view(np.uint32)[:, 18] != 0xFFFFFFFF) & active
    got = exact[:, 8] < 0
    assert (got == want).all()
    assert (exact[~got].view(np.uint32) == shadow[~got].view(np.uint32)).all()
    conservative = shadow.copy()
    scene.occluded1M(conservative)
    cons = conservative[:, 8] < 0
    assert (got <= cons).all(), "exact occlusion must imply box-level occlusion"
    assert 0 < got.sum() < cons.sum()
    scene.release()
