"""Helpers for the block-construction tests: read the reference's committed blocks out of an RIMG image
(reference layout, oracle/rimg.h) and drive the builder port."""
import ctypes as C

import numpy as np

from oracle import pyoracle as po

# CompressedBVH header offsets (probe-verified: tests/test_oracle_cpu.py::test_layout_probe_matches_hard_coded_offsets)
OFF = dict(geomID=0, primID=4, uv=8, rcp_edges=24, nodes=32, leaves=40, grid=48, extent=56, elems=60, grid_width=64,
           space=68, proj=104, box=140)


def rimg_blocks(image):
    """All blocks of an image in DFS order: dict of arrays + the world bounds the top-level BVH4 stores for each."""
    h = po.rimg_header(image)
    buf = image.tobytes()
    col_major = int(h["proj_col_major"])
    out = []

    def f32(off, n):
        return np.frombuffer(buf, np.float32, n, off).copy()

    def u32(off):
        return int(np.frombuffer(buf, np.uint32, 1, off)[0])

    def u64(off):
        return int(np.frombuffer(buf, np.uint64, 1, off)[0])

    def visit(ref, bounds):
        if ref == 8:
            return
        if ref & 8:
            b = ref & ~15
            elems, width = u32(b + OFF["elems"]), u32(b + OFF["grid_width"])
            cells = 3 * elems + 1
            proj = f32(b + OFF["proj"], 9)
            if col_major:
                proj = proj.reshape(3, 3).T.reshape(9).copy()
            blk = dict(geomID=u32(b), primID=u32(b + 4), uv=f32(b + OFF["uv"], 4), extent=f32(b + OFF["extent"], 1)[0],
                       elems=elems, width=width, space=f32(b + OFF["space"], 9), proj=proj, box=f32(b + OFF["box"], 10),
                       nodes=np.frombuffer(buf, np.uint8, 4 * elems, u64(b + OFF["nodes"])).copy(), bounds=bounds)
            kind = int(h["leaf_kind"])
            if kind == 1:
                blk["leaves"] = np.frombuffer(buf, np.uint8, 2 * cells, u64(b + OFF["leaves"])).copy()
            if kind == 0:
                blk["grid"] = f32(u64(b + OFF["grid"]), 3 * width * width)
            out.append(blk)
            return
        n = ref & ~15
        bb = f32(n + 32, 24).reshape(6, 4)   # lower_x, upper_x, lower_y, upper_y, lower_z, upper_z
        for i in range(4):
            visit(u64(n + 8 * i), np.array([bb[0, i], bb[2, i], bb[4, i], bb[1, i], bb[3, i], bb[5, i]], np.float32))

    visit(int(h["root"]), None)
    return out


def block_key(b):
    return (b["geomID"], b["primID"], b["uv"][0].tobytes(), b["uv"][1].tobytes())


class _BuildIn(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("space", C.c_float * 9), ("world", C.c_float * 9), ("leaf_kind", C.c_int32),
                ("comp_level", C.c_int32)]


class _BuildOut(C.Structure):
    _fields_ = [("proj", C.c_float * 9), ("box", C.c_float * 10), ("extent", C.c_float), ("bounds", C.c_float * 6),
                ("patch_ok", C.c_int32), ("nodes", C.c_uint8 * (4 * 341)), ("leaves", C.c_uint8 * (2 * 1024))]


def world_of(space):
    """[vx vy vz] = inverse of the stored frame (the reference keeps only the inverse; float64 inverse, rounded)."""
    m = np.asarray(space, np.float64).reshape(3, 3)   # rows = vx, vy, vz of `space`
    return np.linalg.inv(m).astype(np.float32).reshape(9)


def port_build(grids, spaces, worlds, leaf_kind, comp_level, threads=8):
    """grids: (n, (2^C+1)^2*3) float32; spaces, worlds: (n, 9).  Returns list of dicts like rimg_blocks()."""
    L = po.port_lib()
    L.cbvh_build_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    n = len(grids)
    grids = np.ascontiguousarray(grids, np.float32)
    ins = (_BuildIn * n)()
    outs = (_BuildOut * n)()
    for i in range(n):
        ins[i].grid = grids[i].ctypes.data
        for k in range(9):
            ins[i].space[k] = float(spaces[i][k])
            ins[i].world[k] = float(worlds[i][k])
        ins[i].leaf_kind = leaf_kind
        ins[i].comp_level = comp_level
    L.cbvh_build_blocks(ins, outs, n, threads)
    elems = (4 ** comp_level - 1) // 3
    res = []
    for o in outs:
        res.append(dict(proj=np.array(o.proj[:], np.float32), box=np.array(o.box[:], np.float32), extent=np.float32(o.extent),
                        bounds=np.array(o.bounds[:], np.float32), patch_ok=int(o.patch_ok),
                        nodes=np.frombuffer(o.nodes, np.uint8, 4 * elems).copy(),
                        leaves=np.frombuffer(o.leaves, np.uint8, 2 * 4 ** comp_level).copy()))
    return res


NODE_MASK = np.tile(np.array([0xFF, 0xFC, 0xFF, 0xFC], np.uint8), 341)   # low two bits of the x / y bytes are unused (compressed_node.h:264-293)


def inputs_from_images(grid_image, kind_image):
    """Builder inputs for every block of `kind_image` (any leaf kind of the SAME geometry/S/C as `grid_image`, whose
    cGrid payload is the displaced vertex grid): grids, space, world, uv, geomIDs, primIDs + the parsed reference blocks."""
    grids = {block_key(b): b["grid"] for b in rimg_blocks(grid_image)}
    blocks = rimg_blocks(kind_image)
    g = np.stack([grids[block_key(b)] for b in blocks])
    space = np.stack([b["space"] for b in blocks])
    world = np.stack([world_of(b["space"]) for b in blocks])
    uv = np.stack([b["uv"] for b in blocks])
    geom = np.array([b["geomID"] for b in blocks], np.uint32)
    prim = np.array([b["primID"] for b in blocks], np.uint32)
    return dict(grids=g, space=space, world=world, uv=uv, geomIDs=geom, primIDs=prim), blocks


def compare_with_reference(blocks, built, leaf_kind):
    """Counts of blocks whose built pieces differ from the reference's committed bytes."""
    bad = dict(proj=0, box=0, nodes=0, leaves=0, extent=0, bounds=0)
    for b, o in zip(blocks, built):
        bad["proj"] += int(not (b["proj"].view(np.uint32) == o["proj"].view(np.uint32)).all())
        bad["box"] += int(not (b["box"].view(np.uint32) == o["box"].view(np.uint32)).all())
        m = NODE_MASK[:len(b["nodes"])]
        bad["nodes"] += int(not ((b["nodes"] & m) == (o["nodes"][:len(m)] & m)).all())
        if leaf_kind == 1:
            bad["leaves"] += int(not (b["leaves"] == o["leaves"][:len(b["leaves"])]).all())
            bad["extent"] += int(np.float32(b["extent"]).view(np.uint32) != np.float32(o["extent"]).view(np.uint32))
        if b["bounds"] is not None:   # `world` is re-derived from the stored inverse here, hence a tolerance
            bad["bounds"] += int(not np.allclose(b["bounds"], o["bounds"], rtol=1e-5, atol=1e-5))
    return bad


def device_blocks_as_pieces(blocks, comp_level, leaf_kind):
    """Parse device-format blocks (uint8 (n, stride), csrc/ecb_format.h) into the same pieces port_build() returns."""
    elems = (4 ** comp_level - 1) // 3
    cells = 4 ** comp_level
    leaf_offset = (144 + 4 * (elems - 1) + 15) & ~15
    out = []
    for row in blocks:
        f = row[:144].view(np.float32)
        nodes = np.concatenate([row[140:144], row[144:144 + 4 * (elems - 1)]])
        d = dict(space=f[0:9].copy(), proj=f[9:18].copy(), box=f[18:28].copy(), uv=f[28:32].copy(), extent=f[32].copy(),
                 geomID=int(row[132:136].view(np.uint32)[0]), primID=int(row[136:140].view(np.uint32)[0]), nodes=nodes)
        if leaf_kind == 1:
            d["leaves"] = row[leaf_offset:leaf_offset + 2 * cells].copy()
        if leaf_kind == 0:
            w = 2 ** comp_level + 1
            d["grid"] = row[leaf_offset:leaf_offset + 12 * w * w].view(np.float32).copy()
        out.append(d)
    return out
