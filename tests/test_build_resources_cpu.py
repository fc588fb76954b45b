"""CPU test: the occupancy assumptions of DESIGN.md section 3, read from the built library (`cuobjdump --dump-resource-usage`, no GPU
needed): the lane kernel must fit 5 CTAs of 128 threads per SM (<= 96 registers, the `__launch_bounds__` cap, with its stacks in
the per-thread stack frame, not in spilled `LOCAL` arrays), the quad kernel 6 CTAs (<= 80), the instancing kernel for compressed
grids / leaves at least 3 (<= 144).  A build that silently drops below these changes every number in profiles/."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "embree_compressed_b200", "libecb200.so")


def _resources():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not found")
    if not os.path.exists(LIB):
        pytest.skip("libecb200.so not built")
    out = subprocess.run([exe, "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    res = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
            name = None
    return res


def test_kernels_keep_the_register_budgets_their_occupancy_rests_on():
    res = _resources()
    lane = {k: v for k, v in res.items() if "ecb_intersect_kernel" in k}
    quad = {k: v for k, v in res.items() if "ecb_intersect_quad_kernel" in k}
    inst = {k: v for k, v in res.items() if "ecb_instanced_kernel" in k}
    assert len(lane) == 12 and len(quad) == 6 and len(inst) >= 5, sorted(res)
    for k, (reg, stack, shared, local) in lane.items():
        assert reg <= 96 and local == 0 and shared <= 2048, (k, reg, stack, shared, local)   # 5 x 128 threads x 96 <= 65536
        assert stack <= 1600, (k, stack)   # two 96-entry top-level stacks + the quadtree stack; more means spills
    for k, (reg, stack, shared, local) in quad.items():
        assert reg <= 80 and local == 0 and stack <= 512, (k, reg, stack)
    for k, (reg, stack, shared, local) in inst.items():
        assert reg <= 144 and local == 0, (k, reg)
    for k, (reg, stack, shared, local) in res.items():
        if "ecb_occluded_kernel" in k or "ecb_triangle_kernel" in k:
            assert reg <= 96 and local == 0, (k, reg)


def test_lane_kernel_reads_nodes_and_block_headers_with_128_bit_read_only_loads():
    """north_star: "nodes read with vectorised 128-bit __ldg".  SASS of the shipping lane kernel (cLeaf): the seven 16-byte loads
    of a top-level node and the nine of a block header are LDG.E.128.CONSTANT; the ray records, which an in-place call writes
    back, do not go through the non-coherent path."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or libecb200.so missing")
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    body, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            on = "ecb_intersect_kernelILi1ELb0ELb0E" in line
        elif on:
            body.append(line)
    text = "\n".join(body)
    assert text.count("LDG.E.128.CONSTANT") >= 16, text.count("LDG.E.128.CONSTANT")
    ray_loads = [l for l in body if re.search(r"\bLDG\.E(\.\d+)?\.STRONG\.GPU\b", l)]
    assert len(ray_loads) >= 3, "ray records are read with ld.global.cg (coherent at L2)"
    assert "ATOMG" in text or "ATOM" in text   # the dynamic ray-fetch cursor
