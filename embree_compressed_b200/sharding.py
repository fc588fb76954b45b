"""Multi-GPU plumbing of the traversal path (one process per GPU, torch.distributed).

Rays are independent and the scene is read-only after commit, so the path shards with NO data-path collective:
the scene is replicated on every GPU and every rank traces a contiguous range of image tiles (or of ray indices).
The single exchange step is the optional gather of compact hit records to one rank (NCCL on GPUs, gloo on CPU in
the tests).  Nothing here touches the kernels.
"""
import numpy as np

HIT_RECORD_COLUMNS = [8, 15, 16, 17, 18, 19]  # tfar, u, v, primID, geomID, instID of an RTCRayHit = 24 bytes


def tile_grid(width, height, tile=64):
    """Tiles of a frame in row-major tile order: list of (x0, y0, w, h)."""
    tiles = []
    for y0 in range(0, height, tile):
        for x0 in range(0, width, tile):
            tiles.append((x0, y0, min(tile, width - x0), min(tile, height - y0)))
    return tiles


def tile_range(n_tiles, rank, world):
    """Contiguous, balanced tile range [begin, end) of `rank` (differs by at most one tile between ranks)."""
    base, rem = divmod(n_tiles, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def index_range(n, rank, world):
    """Contiguous, balanced index range of a flat ray batch (BASELINE config 5)."""
    return tile_range(n, rank, world)


def tile_pixel_indices(width, tiles):
    """Flat row-major pixel indices (y*width + x) covered by `tiles`, tile after tile, row-major inside a tile."""
    out = []
    for x0, y0, w, h in tiles:
        ys, xs = np.meshgrid(np.arange(y0, y0 + h), np.arange(x0, x0 + w), indexing="ij")
        out.append((ys * width + xs).ravel())
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def interleaved_tile_indices(width, height, rank, world, tile=64):
    """Interleaved tile sharding (tile t of the row-major tile grid belongs to rank t mod world: sky and geometry tiles
    spread evenly over the ranks).  Returns (flat pixel indices of this rank, tile after tile and row-major inside a
    tile; offset of this rank's records in the rank-major frame arrays; records per rank)."""
    if width % tile == 0 and height % tile == 0:   # no partial tiles: one transpose instead of a loop over the tiles
        ty, tx = height // tile, width // tile
        by_tile = np.arange(width * height, dtype=np.int64).reshape(ty, tile, tx, tile).transpose(0, 2, 1, 3).reshape(ty * tx, tile * tile)
        counts = [len(range(r, ty * tx, world)) * tile * tile for r in range(world)]
        idx = by_tile[rank::world].reshape(-1)
    else:
        tiles = tile_grid(width, height, tile)
        counts = [sum(w * h for (_, _, w, h) in tiles[r::world]) for r in range(world)]
        idx = tile_pixel_indices(width, tiles[rank::world])
    return idx, int(sum(counts[:rank])), counts


def shard_frame(width, height, rank, world, tile=64):
    """(tiles of this rank, flat pixel indices of this rank)."""
    tiles = tile_grid(width, height, tile)
    b, e = tile_range(len(tiles), rank, world)
    mine = tiles[b:e]
    return mine, tile_pixel_indices(width, mine)


def gather_hit_records(local_records, dst=0, group=None):
    """Gathers per-rank hit records (tensor [n_local, 6] float32, any device) on rank `dst`.

    Ranks may hold different counts; returns the list of per-rank tensors on dst, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local_records.shape[0]], dtype=torch.int64, device=local_records.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = local_records.shape[1]
    if rank == dst:
        bufs = [torch.empty((c, width), dtype=local_records.dtype, device=local_records.device) for c in counts]
        reqs = []
        for r in range(world):
            if r == dst:
                bufs[r].copy_(local_records)
            elif counts[r]:
                reqs.append(dist.irecv(bufs[r], src=r, group=group))
        for q in reqs:
            q.wait()
        return bufs
    if local_records.shape[0]:
        dist.send(local_records.contiguous(), dst=dst, group=group)
    return None


def assemble_frame(per_rank_records, width, height, world, tile=64):
    """Inverse of shard_frame on the gathering rank: [height*width, 6] array in pixel order."""
    frame = np.zeros((height * width, len(HIT_RECORD_COLUMNS)), np.float32)
    tiles = tile_grid(width, height, tile)
    for rank, rec in enumerate(per_rank_records):
        b, e = tile_range(len(tiles), rank, world)
        idx = tile_pixel_indices(width, tiles[b:e])
        frame[idx] = rec if isinstance(rec, np.ndarray) else rec.cpu().numpy()
    return frame
