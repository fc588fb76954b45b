"""CPU tests (-m "not gpu"): ray generators and the multi-GPU sharding / gather plumbing (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from embree_compressed_b200 import raygen as rg
from embree_compressed_b200 import sharding as sh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_primary_rays_follow_the_tutorial_camera():
    r = rg.primary_rays(64, 32, **rg.BOMBERMAN_CAMERA)
    assert r.shape == (64 * 32, 20) and r.dtype == np.float32
    assert np.allclose(np.linalg.norm(r[:, 4:7], axis=1), 1.0, atol=1e-6)
    assert (r[:, 3] == 0).all() and np.isinf(r[:, 8]).all()
    assert (r.view(np.uint32)[:, 18] == 0xFFFFFFFF).all()
    centre = r[16 * 64 + 32, 4:7].astype(np.float64)
    want = -np.array(rg.BOMBERMAN_CAMERA["vp"])
    want /= np.linalg.norm(want)
    assert np.dot(centre, want) > 0.999
    win = rg.primary_rays(64, 32, x0=8, y0=4, w=16, h=8, **rg.BOMBERMAN_CAMERA)   # a tile of the same frame
    full = r.reshape(32, 64, 20)[4:12, 8:24].reshape(-1, 20)
    assert (win.view(np.uint32) == full.view(np.uint32)).all()


def test_sampler_is_the_tutorial_murmur_lcg():
    s = rg.sampler_init(np.array([3, 3, 4]), np.array([7, 7, 7]), 0)
    assert s[0] == s[1] and s[0] != s[2]
    st, f = rg.sampler_float(s)
    assert ((f >= 0) & (f < 1)).all()
    # known answer: MurmurHash3 mix/finalize of id = 3 | 7<<16 then sample 0, followed by one LCG step
    h = np.uint32(0)
    def mix(h, k):
        k = np.uint32((int(k) * 0xcc9e2d51) & 0xffffffff); k = np.uint32(((int(k) << 15) | (int(k) >> 17)) & 0xffffffff)
        k = np.uint32((int(k) * 0x1b873593) & 0xffffffff); h = np.uint32(int(h) ^ int(k))
        h = np.uint32(((int(h) << 13) | (int(h) >> 19)) & 0xffffffff); return np.uint32((int(h) * 5 + 0xe6546b64) & 0xffffffff)
    h = mix(mix(h, 3 | (7 << 16)), 0); x = int(h)
    x ^= x >> 16; x = (x * 0x85ebca6b) & 0xffffffff; x ^= x >> 13; x = (x * 0xc2b2ae35) & 0xffffffff; x ^= x >> 16
    assert int(s[0]) == x
    assert int(st[0]) == (x * 1664525 + 1013904223) & 0xffffffff


def test_diffuse_rays_leave_the_surface_on_the_incoming_side():
    prim = rg.primary_rays(32, 32, **rg.CUBE_CAMERA)
    hits = prim.copy()
    hits[:, 8] = 1.5
    hits.view(np.uint32)[::2, 18] = 0          # every second ray "hit"
    n = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (len(prim), 1))
    b = rg.diffuse_rays(hits, n, scene_diag=5.0)
    hit = hits.view(np.uint32)[:, 18] != 0xFFFFFFFF
    assert (b[~hit, 3] > b[~hit, 8]).all(), "misses must become inactive stream rays"
    d_in = prim[hit, 4:7]
    facing = np.where((d_in @ np.array([0, 1.0, 0]))[:, None] > 0, -n[hit], n[hit])
    assert (np.sum(b[hit, 4:7] * facing, axis=1) >= -1e-6).all()
    assert np.allclose(b[hit, 0:3], prim[hit, 0:3] + 1.5 * prim[hit, 4:7], atol=1e-5)
    assert np.allclose(b[hit, 3], 5e-3)


def test_tile_sharding_covers_the_frame_exactly_once():
    W, H = 200, 120
    for world in (1, 2, 3, 8):
        seen = np.zeros(W * H, np.int32)
        sizes = []
        for rank in range(world):
            tiles, idx = sh.shard_frame(W, H, rank, world, tile=32)
            seen[idx] += 1
            sizes.append(len(tiles))
        assert (seen == 1).all()
        assert max(sizes) - min(sizes) <= 1
    assert sh.index_range(10, 0, 3) == (0, 4) and sh.index_range(10, 2, 3) == (7, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tiles, idx = sh.shard_frame(W, H, rank, world, tile=16)
        # stand-in for the traversal of this rank's tiles: a deterministic function of the pixel index
        rec = np.stack([idx * 0.5, idx % 7, idx % 5, idx, np.full_like(idx, rank), np.zeros_like(idx)], 1).astype(np.float32)
        got = sh.gather_hit_records(torch.from_numpy(rec), dst=0)
        if rank == 0:
            frame = sh.assemble_frame(got, W, H, world, tile=16)
            q.put(frame)
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_of_hit_records_over_gloo(world):
    import torch.multiprocessing as mp
    W, H = 80, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    idx = np.arange(W * H)
    assert np.allclose(frame[:, 0], idx * 0.5) and (frame[:, 3] == idx).all()
    tiles = sh.tile_grid(W, H, 16)
    owner = np.zeros(W * H)
    for r in range(world):
        b, e = sh.tile_range(len(tiles), r, world)
        owner[sh.tile_pixel_indices(W, tiles[b:e])] = r
    assert (frame[:, 4] == owner).all()


def test_interleaved_tiles_cover_the_frame_once_and_match_the_tile_loop():
    """bench.py's sharding of a frame (tile t -> rank t mod N): the transposed fast path for frames without partial tiles and
    the tile loop give the same pixel order; the ranks' index sets partition the frame; offsets are rank-major."""
    for (w, h, t) in ((128, 64, 16), (100, 70, 16), (1920, 1080, 64)):
        for world in (1, 2, 3, 8):
            tiles = sh.tile_grid(w, h, t)
            seen = np.zeros(w * h, np.int64)
            for r in range(world):
                idx, off, counts = sh.interleaved_tile_indices(w, h, r, world, t)
                assert (idx == sh.tile_pixel_indices(w, tiles[r::world])).all()
                assert off == sum(counts[:r]) and counts[r] == len(idx)
                seen[idx] += 1
            assert (seen == 1).all() and sum(counts) == w * h
