"""CPU test: the order in which the hit children of a top-level node are visited (csrc/ecb_order.h, shared by every closest-hit
kernel) against a literal push-and-sort restatement of traverseClosestHit (kernels/bvh/bvh_traverser1.h:436-506) and the sort
networks of kernels/common/stack_item.h:39-60 -- every hit mask x every pattern of equal distances (the networks are not
stable: on ties the visiting order is a property of the exact compare-exchange sequence), compared on the unsigned bits."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_hit_children_order_equals_the_reference_for_every_mask_and_tie_pattern(tmp_path):
    exe = str(tmp_path / "order_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(HERE, "cpp", "order_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "36015 cases, 0 mismatches" in r.stdout
