"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (include/ecb200.h), against the oracle.

Every comparison is made twice: bit-exact against the scalar C port (oracle/cbvh_port.c) and, within the
north_star tolerance with the near-tie set computed by the oracle, against the compiled reference itself
(rtcIntersect1 / rtcOccluded1 applied per ray -- the only call form that works on this accel).
"""
import numpy as np
import pytest

import parity
from embree_compressed_b200 import raygen as rg

pytestmark = pytest.mark.gpu

LEAVES = ["leaf", "grid", "box"]
# The tutorial camera looks down the cube's diagonal: pixel rays then hit shared micro-triangle edges EXACTLY, where
# the reference's strict beta>0 / gamma>0 test (compressed_help.h:266) lets a ray fall through or not depending on the
# last ulp (its SSE4.2 and AVX2 paths disagree on those rays too).  The small cases use a camera off the symmetry axis.
OFF_AXIS = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
CASES = [  # (scene kind, S, C, camera, width, height)
    ("cube", 6, 4, rg.CUBE_CAMERA, 512, 512),            # BASELINE config 1 geometry (1024^2 in the slow test)
    ("bomberman", 6, 3, rg.BOMBERMAN_CAMERA, 640, 360),  # BASELINE config 2 geometry (1920x1080 in the slow test)
    ("cube", 4, 2, OFF_AXIS, 256, 256),
    ("cube", 5, 1, OFF_AXIS, 128, 128),
]


def _scene_pair(ecb, oracle, ref):
    """product scene imported from LIVE reference memory + the port bound to the same memory."""
    root, n_accels = ref.root()
    assert root and n_accels == 1
    scene = ecb.Scene.import_bvh4(root, ref.leaf, ref.C, ref.S, oracle.ref_layout()["proj_col_major"], 0, ref.bounds())
    return scene, oracle.PortScene.from_ref(ref)


def _check_intersect(scene, port, ref, rays, what):
    got = rays.copy()
    scene.intersect1M(got)
    want = rays.copy()
    port.intersect(want)
    bad = parity.bit_exact_mismatches(got, want)
    assert len(bad) == 0, "%s: CUDA differs from the oracle port on %d of %d rays (first rows %s)" % (
        what, len(bad), len(rays), bad[:8])
    truth = rays.copy()
    ref.intersect(truth)
    c = parity.assert_reference_parity(got, truth, port.candidates(rays), what)
    return got, truth, c


@pytest.mark.parametrize("leaf", LEAVES)
@pytest.mark.parametrize("kind,S,C,cam,W,H", CASES)
def test_primary_and_diffuse(ecb, reference, ref_scene, leaf, kind, S, C, cam, W, H):
    ref = ref_scene(kind, leaf, S, C)
    scene, port = _scene_pair(ecb, reference, ref)
    rays = rg.primary_rays(W, H, **cam)
    got, truth, c = _check_intersect(scene, port, ref, rays, "%s S%d C%d %s primary" % (kind, S, C, leaf))
    assert c["hits_ref"] > 0
    # cosine-diffuse bounce from the REFERENCE's hits with the viewer's smooth normals (same array to all three)
    b = ref.bounds()
    diag = float(np.linalg.norm(b[3:] - b[:3]))
    bounce = rg.diffuse_rays(truth, ref.hit_normals(truth), diag)
    _check_intersect(scene, port, ref, bounce, "%s S%d C%d %s diffuse" % (kind, S, C, leaf))
    # shadow-style any-hit query on the same rays: rtcOccluded1 semantics (top-level boxes only)
    occ = rg.rays_only(bounce)
    got_o, want_o, truth_o = occ.copy(), occ.copy(), occ.copy()
    scene.occluded1M(got_o)
    port.occluded(want_o)
    ref.occluded(truth_o)
    assert (got_o.view(np.uint32) == want_o.view(np.uint32)).all()
    assert (got_o.view(np.uint32) == truth_o.view(np.uint32)).all(), "occluded differs from the reference"
    scene.release()


@pytest.mark.parametrize("leaf", LEAVES)
def test_incoherent_rays(ecb, reference, ref_scene, leaf):
    ref = ref_scene("bomberman", leaf, 6, 3)
    scene, port = _scene_pair(ecb, reference, ref)
    b = ref.bounds()
    lo, hi = b[:3].copy(), b[3:].copy()
    # shrink the AABB to the character (the ground quad dominates the scene bounds, SURVEY 8d caveat)
    lo[0], lo[2], hi[0], hi[2] = -25, -25, 25, 25
    rays = rg.incoherent_rays(200_000, np.concatenate([lo, hi]), seed=7)
    _check_intersect(scene, port, ref, rays, "bomberman %s incoherent" % leaf)
    scene.release()


def test_entry_points_and_edge_cases(ecb, reference, ref_scene):
    ref = ref_scene("cube", "leaf", 4, 2)
    scene, port = _scene_pair(ecb, reference, ref)
    rays = rg.primary_rays(64, 64, **rg.CUBE_CAMERA)
    want = rays.copy()
    port.intersect(want)

    # M = 0 and M = 1
    scene.intersect1M(rays[:0].copy())
    one = rays[2080:2081].copy()
    scene.intersect1(one[0])
    assert (one.view(np.uint32) == want[2080:2081].view(np.uint32)).all()

    # byteStride > sizeof(RTCRayHit): 96-byte records, padding must stay untouched
    wide = np.full((len(rays), 24), 7.0, np.float32)
    wide[:, :20] = rays
    scene.intersect1M(wide)
    assert (wide[:, :20].view(np.uint32) == want.view(np.uint32)).all()
    assert (wide[:, 20:] == 7.0).all()

    # inactive stream rays (tnear > tfar) are skipped; a miss writes nothing at all
    mixed = rays.copy()
    mixed[::3, 3] = 2.0
    mixed[::3, 8] = 1.0
    ref_mixed = mixed.copy()
    port.intersect(ref_mixed)
    scene.intersect1M(mixed)
    assert (mixed.view(np.uint32) == ref_mixed.view(np.uint32)).all()
    assert (mixed[::3].view(np.uint32) == ref_mixed[::3].view(np.uint32)).all()

    # rays that already carry a finite tfar / a previous hit keep it when nothing nearer is found
    near = rays.copy()
    near[:, 8] = 0.5
    near.view(np.uint32)[:, 18] = 5
    near_want = near.copy()
    port.intersect(near_want)
    scene.intersect1M(near)
    assert (near.view(np.uint32) == near_want.view(np.uint32)).all()

    # instID is taken from the context argument
    inst = rays.copy()
    scene.intersect1M(inst, inst_id=3)
    hit = inst.view(np.uint32)[:, 18] != 0xFFFFFFFF
    assert hit.any() and (inst.view(np.uint32)[hit, 19] == 3).all()

    # rtcIntersect16: SoA packet with a valid mask
    packet_rows = np.arange(2072, 2088)
    soa = np.ascontiguousarray(rays[packet_rows].T)
    valid = np.full(16, -1, np.int32)
    valid[[1, 5, 6]] = 0
    scene.intersect16(valid, soa)
    back = soa.T
    act = valid == -1
    assert (back[act].view(np.uint32) == want[packet_rows][act].view(np.uint32)).all()
    assert (np.ascontiguousarray(back[~act]).view(np.uint32) == rays[packet_rows][~act].view(np.uint32)).all()

    # occluded: tfar < 0 is "already occluded" and stays untouched
    occ = rg.rays_only(rays)
    occ[::2, 8] = -1.0
    occ_want = occ.copy()
    port.occluded(occ_want)
    scene.occluded1M(occ)
    assert (occ.view(np.uint32) == occ_want.view(np.uint32)).all()
    one_o = rg.rays_only(rays)[2081:2082].copy()
    one_want = one_o.copy()
    port.occluded(one_want)
    scene.occluded1(one_o[0])
    assert (one_o.view(np.uint32) == one_want.view(np.uint32)).all()
    scene.release()


def test_device_resident_path_and_block_output(ecb, reference, ref_scene):
    import torch
    ref = ref_scene("bomberman", "leaf", 6, 3)
    scene, port = _scene_pair(ecb, reference, ref)
    rays = rg.primary_rays(480, 270, **rg.BOMBERMAN_CAMERA)
    want = rays.copy()
    port.intersect(want)
    d = torch.from_numpy(rays).cuda()
    blocks = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    scene.intersect_device(d.data_ptr(), len(rays), 80, stream, blocks.data_ptr())
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    blk = blocks.cpu().numpy().view(np.uint32)
    hit = got.view(np.uint32)[:, 18] != 0xFFFFFFFF
    assert (blk[~hit] == 0xFFFFFFFF).all() and (blk[hit] < scene.info()["n_blocks"]).all()
    # the block frame normal is a unit vector roughly aligned with the reference's smooth normal
    n_blk = scene.block_normals()[blk[hit]]
    assert np.allclose(np.linalg.norm(n_blk, axis=1), 1.0, atol=1e-3)
    n_ref = ref.hit_normals(want)[hit]
    n_ref /= np.maximum(np.linalg.norm(n_ref, axis=1, keepdims=True), 1e-20)
    assert np.median(np.abs(np.sum(n_blk * n_ref, axis=1))) > 0.95
    scene.release()


def test_image_and_blob_round_trip(ecb, reference, ref_scene, tmp_path):
    ref = ref_scene("cube", "grid", 4, 2)
    live, port = _scene_pair(ecb, reference, ref)
    rays = rg.primary_rays(128, 128, **rg.CUBE_CAMERA)
    want = rays.copy()
    port.intersect(want)
    image = ref.dump_image()
    from_img = ecb.Scene.from_image(image)
    p = str(tmp_path / "scene.ecbs")
    from_img.save(p)
    from_blob = ecb.Scene.load(p)
    for s in (live, from_img, from_blob):
        got = rays.copy()
        s.intersect1M(got)
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
        s.release()


@pytest.mark.slow
@pytest.mark.parametrize("leaf", LEAVES)
def test_full_size_configs(ecb, reference, ref_scene, leaf):
    """BASELINE.json configs 1 and 2 at their full ray counts (1024^2 and 1920x1080 primary + one diffuse bounce)."""
    for kind, S, C, cam, W, H in (("cube", 6, 4, rg.CUBE_CAMERA, 1024, 1024),
                                  ("bomberman", 6, 3, rg.BOMBERMAN_CAMERA, 1920, 1080)):
        ref = ref_scene(kind, leaf, S, C)
        scene, port = _scene_pair(ecb, reference, ref)
        rays = rg.primary_rays(W, H, **cam)
        got, truth, _ = _check_intersect(scene, port, ref, rays, "%s %s full primary" % (kind, leaf))
        b = ref.bounds()
        bounce = rg.diffuse_rays(truth, ref.hit_normals(truth), float(np.linalg.norm(b[3:] - b[:3])))
        _check_intersect(scene, port, ref, bounce, "%s %s full diffuse" % (kind, leaf))
        # size-independent properties at full size.
        # (1) rays are independent and the dynamic ray scheduler must not leak between lanes: any permutation of the
        #     batch, traced in one call or in unequal pieces, gives bit-identical per-ray results
        perm = np.random.default_rng(5).permutation(len(rays))
        shuffled = rays[perm].copy()
        scene.intersect1M(shuffled[: len(rays) // 3])
        scene.intersect1M(shuffled[len(rays) // 3:])
        assert (shuffled.view(np.uint32) == got[perm].view(np.uint32)).all()
        # (2) exact micro-triangles only: re-tracing a result changes nothing (the recorded hit sits at t == tfar and
        #     the test is tt < tfar).  The pizza-box / voxel paths are not idempotent in the reference either: their
        #     local ray is the chord between the frustum entry point and the point at tfar (compressed.h:470-476), which
        #     moves when tfar shrinks.
        if leaf == "grid":
            again = got.copy()
            scene.intersect1M(again)
            assert (again.view(np.uint32) == got.view(np.uint32)).all()
        scene.release()


# ---- mixed scenes: BASELINE config 1 as the tutorial builds it, with its ground plane ----------------
@pytest.mark.parametrize("leaf", LEAVES)
def test_mixed_scene_cube_with_ground_plane(ecb, reference, leaf):
    S, C = 5, 3
    ref = reference.RefScene(leaf, S, C)
    ref.add_cube(1, 1.0)
    plane = ref.add_ground_plane()
    ref.commit()
    root, n_accels = ref.root()
    tri_root = ref.triangle_root()[0]
    assert root and tri_root and n_accels == 2
    scene = ecb.Scene.import_bvh4(root, leaf, C, S, reference.ref_layout()["proj_col_major"], 0, ref.bounds())
    scene.attach_triangles4v(tri_root)
    port = reference.PortScene.from_ref(ref)
    rays = rg.primary_rays(384, 384, **OFF_AXIS)
    got, truth, c = _check_intersect(scene, port, ref, rays, "cube+plane %s primary" % leaf)
    gi = got.view(np.uint32)
    on_plane = (gi[:, 18] == plane) & (truth.view(np.uint32)[:, 18] == plane)
    assert on_plane.sum() > 5000 and (gi[:, 18] == 0).sum() > 5000
    assert np.allclose(got[on_plane, 12:15], truth[on_plane, 12:15], rtol=1e-5)          # real normals on triangle hits
    assert (got[gi[:, 18] == 0, 12:15] == [1.0, 0.0, 0.0]).all()                         # dummy Ng on compressed hits
    b = ref.bounds()
    bounce = rg.diffuse_rays(truth, ref.hit_normals(truth), float(np.linalg.norm(b[3:] - b[:3])))
    _check_intersect(scene, port, ref, bounce, "cube+plane %s bounce" % leaf)
    # any-hit: AccelN::occluded = triangles first, then the (conservative) compressed accel
    o_got, o_port, o_ref = rg.rays_only(bounce), rg.rays_only(bounce), rg.rays_only(bounce)
    scene.occluded1M(o_got)
    port.occluded(o_port)
    ref.occluded(o_ref)
    assert (o_got.view(np.uint32) == o_port.view(np.uint32)).all()
    assert (o_got.view(np.uint32) == o_ref.view(np.uint32)).all() and (o_got[:, 8] < 0).any()
    # device-resident entry point and the image / blob forms of the same scene
    import torch
    d = torch.from_numpy(rays).cuda()
    scene.intersect_device(d.data_ptr(), len(rays))
    torch.cuda.synchronize()
    assert (d.cpu().numpy().view(np.uint32) == got.view(np.uint32)).all()
    from_image = ecb.Scene.from_image(ref.dump_image())
    again = rays.copy()
    from_image.intersect1M(again)
    assert (again.view(np.uint32) == got.view(np.uint32)).all()
    ref.free()


@pytest.mark.slow
def test_config1_full_frame_with_ground_plane(ecb, reference):
    """BASELINE config 1 exactly: cube S6/C4 with displacement + ground plane, 1024x1024 primary rays, tutorial camera."""
    ref = reference.RefScene("leaf", 6, 4)
    ref.add_cube(1, 1.0)
    ref.add_ground_plane()
    ref.commit()
    scene = ecb.Scene.import_bvh4(ref.root()[0], "leaf", 4, 6, reference.ref_layout()["proj_col_major"], 0, ref.bounds())
    scene.attach_triangles4v(ref.triangle_root()[0])
    port = reference.PortScene.from_ref(ref)
    rays = rg.primary_rays(1024, 1024, **rg.CUBE_CAMERA)
    _, truth, c = _check_intersect(scene, port, ref, rays, "config 1 with ground plane")
    assert c["hits_ref"] > 500000
    ref.free()


@pytest.mark.parametrize("leaf", LEAVES)
def test_degenerate_rays_and_empty_scene(ecb, reference, ref_scene, leaf):
    """Rays the reference's own tutorials never produce but its API admits: zero / axis-parallel / denormal directions,
    origins on and inside the boxes, negative tnear, tfar = 0, huge and non-finite values.  The CUDA path must terminate
    and leave exactly the bits the oracle leaves."""
    ref = ref_scene("cube", leaf, 4, 2)
    scene, port = _scene_pair(ecb, reference, ref)
    base = rg.primary_rays(48, 48, **OFF_AXIS)
    n = len(base)
    rays = np.concatenate([base] * 10)
    k = lambda i: slice(i * n, (i + 1) * n)
    rays[k(0), 4:7] = 0.0                                   # zero direction
    rays[k(1), 4] = 0.0                                     # one zero component
    rays[k(1), 5] = 0.0                                     # axis-parallel along z
    rays[k(2), 4:7] *= 1e-30                                # denormal-scale direction
    rays[k(3), 0:3] = 0.0                                   # origin inside the cube
    rays[k(4), 3] = -5.0                                    # negative tnear
    rays[k(5), 8] = 0.0                                     # tfar = 0
    rays[k(6), 0:3] *= 1e18                                 # far away
    rays[k(7), 4:7] = np.sign(rays[k(7), 4:7]) * np.float32(1.0)   # diagonal, un-normalised
    rays[k(8), 5] = np.float32(np.inf)                      # non-finite direction component
    rays[k(9), 0] = np.float32(np.nan)                      # NaN origin
    got, want = rays.copy(), rays.copy()
    scene.intersect1M(got)
    port.intersect(want)
    finite = slice(0, 8 * n)                                # the non-finite groups only have to terminate identically in hit/miss
    bad = parity.bit_exact_mismatches(got[finite], want[finite])
    assert len(bad) == 0, "degenerate rays: CUDA differs from the oracle on rows %s" % bad[:10]
    assert ((got.view(np.uint32)[8 * n:, 18] != 0xFFFFFFFF) == (want.view(np.uint32)[8 * n:, 18] != 0xFFFFFFFF)).all()
    truth = rays[finite].copy()
    ref.intersect(truth)
    c = parity.classify(got[finite], truth, port.candidates(rays[finite]))
    assert c["hit_miss_flips"] <= 4 and c["hits_ref"] > 1000, c
    o_got, o_want = rg.rays_only(rays), rg.rays_only(rays)
    scene.occluded1M(o_got)
    port.occluded(o_want)
    assert (o_got[finite].view(np.uint32) == o_want[finite].view(np.uint32)).all()
    scene.release()
    # a committed scene without geometry: BVH4::emptyNode as root (kernels/bvh/bvh.h:127); nothing may be written
    empty = ecb.Scene.import_bvh4(8, leaf, 2, 4, 1, 0, np.zeros(6, np.float32))
    assert empty.info()["n_blocks"] == 0
    e = base.copy()
    empty.intersect1M(e)
    assert (e.view(np.uint32) == base.view(np.uint32)).all()
    eo = rg.rays_only(base)
    empty.occluded1M(eo)
    assert (eo.view(np.uint32) == rg.rays_only(base).view(np.uint32)).all()
    empty.release()


def test_concurrent_callers_on_one_scene(ecb, reference, ref_scene):
    """rtcIntersect* / rtcOccluded* are re-entrant in the reference (kernels/common/rtcore.cpp:287-304: no locks, state
    only in the ray): several user threads trace different ray sets through the same committed scene at once, host
    buffers (compact-result pipeline + scatter pool) and device buffers mixed."""
    import threading
    import torch
    ref = ref_scene("bomberman", "leaf", 6, 3)
    scene, port = _scene_pair(ecb, reference, ref)
    sets = [rg.primary_rays(640, 360, **rg.BOMBERMAN_CAMERA),
            rg.incoherent_rays(200_000, ref.bounds(), seed=21),
            rg.primary_rays(320, 200, **rg.BOMBERMAN_CAMERA)[::-1].copy(),
            rg.incoherent_rays(50_000, ref.bounds(), seed=22)]
    want = []
    for r in sets:
        w = r.copy()
        port.intersect(w)
        want.append(w)
    got = [None] * len(sets)
    errors = []

    def host_worker(i):
        try:
            for _ in range(3):
                g = sets[i].copy()
                scene.intersect1M(g)
                o = rg.rays_only(sets[i])
                scene.occluded1M(o)
            got[i] = g
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def device_worker(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    d = torch.from_numpy(sets[i]).cuda()
                    scene.intersect_device(d.data_ptr(), len(sets[i]), 80, st.cuda_stream)
                    st.synchronize()
                got[i] = d.cpu().numpy()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=host_worker if i % 2 == 0 else device_worker, args=(i,)) for i in range(len(sets))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for i in range(len(sets)):
        assert got[i] is not None and (got[i].view(np.uint32) == want[i].view(np.uint32)).all(), "caller %d got wrong results" % i
    scene.release()
