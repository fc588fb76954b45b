"""Generates tests/golden/parity_report.json: the parity bar of this repo as a committed artifact.

For every BASELINE config x leaf kind x ray set the oracle port (which the CUDA kernels equal bit for bit -- the -m gpu
tests assert that separately) is compared with the compiled reference's AVX2 path on the same rays, every ray is classified
(tests/parity.py: ok / near-tie / fragile) and every FRAGILE ray is given a proof class:

  A  the reference disagrees with ITSELF on that ray: its SSE4.2 path (device config max_isa=sse4.2) or its AVX-512 build
     does not meet the same tolerance against its AVX2 path (rcpss / rsqrtss + Newton and FMA contraction differ per ISA);
  B  the oracle finds a second leaf candidate within EPS of the nearest one (relative t): the two implementations picked
     different members of a numerical tie;
  C  same primitive and the same hit distance within tolerance, only (u, v) deviate, by at most EPS_UV: the cell-local
     coordinate is (entry point - cell box) / cell width, and the quantized cell boxes of a compressed block can be orders of
     magnitude thinner than the coordinates the entry point is computed from, so the last ulp of the entry point is
     amplified by 1 / width (pizza-box and voxel leaves only; the reference's ISAs share their rcpss-based arithmetic and
     agree with each other on these rays, the oracle divides in IEEE arithmetic).

  D  ill-conditioned ray: the discrepancy lies inside the ORACLE'S OWN sensitivity to perturbations of the ray (direction and
     origin moved by at most 3e-6 relative, a third of the tolerance): some perturbed ray reproduces the reference's answer
     within tolerance, or the perturbed answers stay on the reference's primitive and their t / u / v ranges contain the
     reference's values.  These are rays along block boundaries and silhouettes (the frustum edge test of
     compressed_help.h:93-133 decides hit or miss in the last ulp), rays that start on the surface and grazing hits of a
     pizza-box sheet; the reference's ISAs share their rcpss-based reciprocal and agree with each other on them, the oracle
     divides in IEEE arithmetic.

A fragile ray in none of A / B / C / D is UNEXPLAINED; tests/test_parity_report.py fails if there is one.

    python tests/golden/make_parity_report.py [--write] [--quick]
Needs oracle/_ref (the compiled reference); run where /root/reference was mounted at build() time.
"""
import json, os, sys, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from embree_compressed_b200 import raygen as rg
from oracle import pyoracle as po

EPS_B = 1e-3   # class B: second candidate within this relative distance of the nearest
EPS_UV = 1e-3  # class C: same primitive, same t, |du|, |dv| at most this
PERTURB = 3e-6 # class D: relative perturbation of the ray direction under which the oracle's answer becomes the reference's


def perturbation_explains(port, rays, truth, trials=96):
    """For each ray: is the discrepancy inside the oracle's OWN sensitivity to input perturbations below the tolerance?  The
    ray's direction and origin are moved by up to PERTURB (relative) in random directions, at four magnitudes; the ray is
    explained if (a) some perturbed ray gives the reference's answer within the parity tolerance, or (b) the perturbed answers
    hit the same primitive as the reference and their t, u and v ranges (widened by the tolerance) contain the reference's."""
    n = len(rays)
    done = np.zeros(n, bool)
    rng = np.random.default_rng(7)
    ti = truth.view(np.uint32)
    ref_hit = ti[:, 18] != 0xFFFFFFFF
    lo = np.full((n, 3), np.inf)
    hi = np.full((n, 3), -np.inf)
    same_prim_seen = np.zeros(n, bool)
    with np.errstate(all="ignore"):
        scale = np.maximum(np.abs(truth[:, 8]), np.linalg.norm(truth[:, 0:3], axis=1))
    for mag in (PERTURB / 30, PERTURB / 10, PERTURB / 3, PERTURB):
        for _ in range(trials):
            d = rng.standard_normal(3)
            o = rng.standard_normal(3)
            r = rays.copy()
            nd = np.linalg.norm(r[:, 4:7], axis=1, keepdims=True)
            no = np.maximum(np.linalg.norm(r[:, 0:3], axis=1, keepdims=True), 1.0)
            r[:, 4:7] = (r[:, 4:7].astype(np.float64) + mag * nd * (d / np.linalg.norm(d))).astype(np.float32)
            r[:, 0:3] = (r[:, 0:3].astype(np.float64) + mag * no * (o / np.linalg.norm(o))).astype(np.float32)
            port.intersect(r)
            c = parity.classify(r, truth, None)
            ok = np.ones(n, bool)
            ok[c["fragile_rows"]] = False
            done |= ok
            ri = r.view(np.uint32)
            same = ref_hit & (ri[:, 18] == ti[:, 18]) & (ri[:, 17] == ti[:, 17]) & (ri[:, 19] == ti[:, 19])
            vals = np.stack([r[:, 8], r[:, 15], r[:, 16]], 1).astype(np.float64)
            lo = np.where(same[:, None], np.minimum(lo, vals), lo)
            hi = np.where(same[:, None], np.maximum(hi, vals), hi)
            same_prim_seen |= same
        if done.all():
            break
    tv = np.stack([truth[:, 8], truth[:, 15], truth[:, 16]], 1).astype(np.float64)
    tol = np.stack([parity.TOL * scale, np.full(n, parity.TOL), np.full(n, parity.TOL)], 1)
    inside = same_prim_seen & ((tv >= lo - tol) & (tv <= hi + tol)).all(axis=1)
    return done | inside


def make_scene(kind, leaf, S, C, lib=None, isa=None):
    cfg = leaf + (",max_isa=" + isa if isa else "")
    s = po.RefScene(cfg, S, C, lib_path=lib)
    if kind == "bomberman":
        s.add_obj()
    elif kind == "cubeplane":
        s.add_cube(1, 1.0)
        s.add_ground_plane()
    else:
        s.add_cube(1, 1.0)
    s.commit()
    return s


def bad_rows(x, truth, cand):
    c = parity.classify(x, truth, cand)
    n = len(x)
    ok_mask = np.ones(n, bool)
    xi, ri = x.view(np.uint32), truth.view(np.uint32)
    # recompute the masks parity.classify uses (it returns counts and the fragile rows)
    frag = np.zeros(n, bool)
    frag[c["fragile_rows"]] = True
    return c, frag


def not_ok_mask(x, truth):
    """rows where x does not meet the tolerance against truth (ids, t, u, v), ignoring the tie bookkeeping"""
    c = parity.classify(x, truth, None)   # without candidates every bad row is 'fragile'
    m = np.zeros(len(x), bool)
    m[c["fragile_rows"]] = True
    return m


class InstancedCase:
    """BASELINE config 4: the 25-instance scene of the benchmark (__graft_entry__.CONFIG4_SPEC) built by one ISA / library of
    the reference; same interface as a RefScene for the report (intersect, bounds, hit_normals, free)."""

    def __init__(self, isa=None, lib=None):
        import __graft_entry__ as ge
        from instancing_util import affine
        spec = ge.CONFIG4_SPEC
        cfg = spec["leaf"] + (",max_isa=" + isa if isa else "")
        self.objs = []
        for kind, S, C in spec["objects"]:
            o = po.RefScene(cfg, S, C, lib_path=lib)
            o.add_obj() if kind == "bomberman" else o.add_cube(1, 1.0)
            o.commit()
            self.objs.append(o)
        self.top = po.RefScene(cfg, spec["objects"][0][1], spec["objects"][0][2], lib_path=lib)
        k, g = 0, spec["grid"]
        for ix in range(-g, g + 1):
            for iz in range(-g, g + 1):
                o = self.objs[(ix + iz) % len(self.objs)]
                self.top.add_instance(o, affine(0.6 + 0.1 * ((ix * 5 + iz) % 3), 0.3 * k, (spec["spacing"] * ix, 0.2 * ((k % 4) - 1.5), spec["spacing"] * iz)))
                k += 1
        self.top.commit()

    def intersect(self, rays, threads=None):
        return self.top.intersect(rays, threads)

    def bounds(self):
        return self.top.bounds()

    def port(self):
        return po.PortInstancedScene(self.top, self.objs)

    def free(self):
        self.top.free()
        for o in self.objs:
            o.free()


def case_report(name, kind, leaf, S, C, ray_sets, avx512=True, make_tracer=None):
    """ray_sets: callable(ref_avx2) -> list of (label, rays).  make_tracer(ref_avx2, port) -> function(rays) tracing them in
    place with the implementation under test (default: the oracle port; the -m gpu test passes the CUDA path)."""
    t0 = time.time()
    if kind == "instanced":
        ref, sse = InstancedCase(), InstancedCase(isa="sse4.2")
        a512 = InstancedCase(lib=po.REF_SO_AVX512) if (avx512 and po.have_avx512_reference()) else None
        port = ref.port()
    else:
        ref = make_scene(kind, leaf, S, C)
        sse = make_scene(kind, leaf, S, C, isa="sse4.2")
        a512 = make_scene(kind, leaf, S, C, lib=po.REF_SO_AVX512) if (avx512 and po.have_avx512_reference()) else None
        port = po.PortScene.from_ref(ref)
    trace = make_tracer(ref, port) if make_tracer else port.intersect
    out = {}
    for label, rays in ray_sets(ref):
        truth, x, s, = rays.copy(), rays.copy(), rays.copy()
        ref.intersect(truth)
        trace(x)
        sse.intersect(s)
        cand = port.candidates(rays)
        c, frag = bad_rows(x, truth, cand)
        disagree = not_ok_mask(s, truth)
        isas = ["sse4.2"]
        if a512 is not None:
            z = rays.copy()
            a512.intersect(z)
            disagree |= not_ok_mask(z, truth)
            isas.append("avx512skx")
        with np.errstate(all="ignore"):
            gap = (cand[:, 1] - cand[:, 0]) / np.maximum(np.abs(cand[:, 0]), 1e-30)
        xi, ti = x.view(np.uint32), truth.view(np.uint32)
        with np.errstate(all="ignore"):
            scale_t = np.maximum(np.abs(truth[:, 8]), np.linalg.norm(truth[:, 0:3], axis=1))
            same_hit = (xi[:, 18] != 0xFFFFFFFF) & (xi[:, 17] == ti[:, 17]) & (xi[:, 18] == ti[:, 18]) & (xi[:, 19] == ti[:, 19]) & \
                       (np.abs(x[:, 8] - truth[:, 8]) <= parity.TOL * np.maximum(scale_t, 1e-30))
            uv_close = (np.abs(x[:, 15] - truth[:, 15]) <= EPS_UV) & (np.abs(x[:, 16] - truth[:, 16]) <= EPS_UV)
        cls_a = frag & disagree
        cls_b = frag & ~cls_a & (gap <= EPS_B)
        cls_c = frag & ~cls_a & ~cls_b & same_hit & uv_close
        rest = np.nonzero(frag & ~cls_a & ~cls_b & ~cls_c)[0]
        cls_d = np.zeros(len(rays), bool)
        if len(rest):
            cls_d[rest[perturbation_explains(port, rays[rest], truth[rest])]] = True
        unexplained = frag & ~cls_a & ~cls_b & ~cls_c & ~cls_d
        with np.errstate(all="ignore"):
            scale = np.maximum(np.abs(truth[:, 8]), np.linalg.norm(truth[:, 0:3], axis=1))
            rel = np.abs(x[:, 8] - truth[:, 8]) / np.maximum(scale, 1e-30)
            both = (x.view(np.uint32)[:, 18] != 0xFFFFFFFF) & (truth.view(np.uint32)[:, 18] != 0xFFFFFFFF)
        out[label] = {"rays": int(len(rays)), "active": int((rays[:, 3] <= rays[:, 8]).sum()), "hits_reference": c["hits_ref"],
                      "ok": c["ok"], "near_tie": c["near_tie"], "fragile": c["fragile"], "hit_miss_flips": c["hit_miss_flips"],
                      "fragile_class_A_reference_disagrees_with_itself": int(cls_a.sum()),
                      "fragile_class_B_second_candidate_within_eps": int(cls_b.sum()),
                      "fragile_class_C_same_hit_uv_within_1e-3": int(cls_c.sum()),
                      "fragile_class_D_oracle_flips_under_3e-6_perturbation": int(cls_d.sum()),
                      "fragile_unexplained": int(unexplained.sum()), "unexplained_rows": [int(i) for i in np.nonzero(unexplained)[0][:64]],
                      "reference_own_disagreement_rows": int(disagree.sum()), "reference_isas_compared": isas,
                      "rel_dt_p999": float(np.quantile(rel[both], 0.999)) if both.any() else 0.0}
        print("%-34s %-8s n=%8d ok=%8d tie=%6d fragile=%5d (A %5d  B %4d  C %4d  D %4d  unexplained %3d)  ref-vs-itself %6d  [%4.1fs]"
              % (name, label, len(rays), c["ok"], c["near_tie"], c["fragile"], cls_a.sum(), cls_b.sum(), cls_c.sum(), cls_d.sum(), unexplained.sum(), disagree.sum(), time.time() - t0), flush=True)
    for s_ in (ref, sse, a512):
        if s_ is not None:
            s_.free()
    return out


def primary_and_bounce(W, H, cam, hits_file=None):
    def make(ref):
        prim = rg.primary_rays(W, H, **cam)
        first = prim.copy()
        ref.intersect(first)
        b = ref.bounds()
        xs, ys = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
        bounce = rg.diffuse_rays(first, ref.hit_normals(first), float(np.linalg.norm(b[3:] - b[:3])), sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
        return [("primary", prim), ("bounce", bounce)]
    return make


def primary_only(W, H, cam):
    return lambda ref: [("primary", rg.primary_rays(W, H, **cam))]


def diffuse_samples(W, H, cam, samples):
    """BASELINE config 5: cosine-diffuse rays leaving the primary hit points, several sample indices of the tutorial sampler."""
    def make(ref):
        prim = rg.primary_rays(W, H, **cam)
        first = prim.copy()
        ref.intersect(first)
        b = ref.bounds()
        xs, ys = np.meshgrid(np.arange(W, dtype=np.uint32), np.arange(H, dtype=np.uint32))
        n = ref.hit_normals(first)
        diag = float(np.linalg.norm(b[3:] - b[:3]))
        return [("diffuse sample %d" % k, rg.diffuse_rays(first, n, diag, sample=k, pixel_xy=(xs.ravel(), ys.ravel()))) for k in samples]
    return make


def main():
    quick = "--quick" in sys.argv
    cases = []
    sz = (480, 270) if quick else (1920, 1080)
    for leaf in ("leaf", "grid", "box"):
        cases.append(("config2 bomberman S6/C3 %s %dx%d" % (leaf, sz[0], sz[1]), "bomberman", leaf, 6, 3, primary_and_bounce(sz[0], sz[1], rg.BOMBERMAN_CAMERA)))
    c1 = 256 if quick else 1024
    for leaf in ("leaf", "grid", "box"):
        cases.append(("config1 cube S6/C4 %s %dx%d" % (leaf, c1, c1), "cube", leaf, 6, 4, primary_and_bounce(c1, c1, rg.CUBE_CAMERA)))
    cases.append(("config1 cube+plane S6/C4 leaf %dx%d" % (c1, c1), "cubeplane", "leaf", 6, 4, primary_and_bounce(c1, c1, rg.CUBE_CAMERA)))
    if not quick:
        cases.append(("config3 cube S8/C4 leaf 1920x1080", "cube", "leaf", 8, 4, primary_and_bounce(1920, 1080, rg.CUBE_CAMERA)))
    import __graft_entry__ as ge
    c4 = (480, 270) if quick else (1920, 1080)
    cases.append(("config4 25 instances grid %dx%d" % c4, "instanced", "grid", 6, 3, primary_only(c4[0], c4[1], ge.CONFIG4_CAMERA)))
    c5 = (480, 270) if quick else (1920, 1080)
    cases.append(("config5 bomberman S6/C3 leaf %dx%d" % c5, "bomberman", "leaf", 6, 3, diffuse_samples(c5[0], c5[1], rg.BOMBERMAN_CAMERA, (1, 17, 38))))
    report = {"generated_by": "tests/golden/make_parity_report.py", "tolerance": {"t_u_v": parity.TOL, "near_tie_eps": parity.EPS_TIE, "class_B_eps": EPS_B, "class_C_eps_uv": EPS_UV, "class_D_perturbation": PERTURB},
              "cases": {}}
    for name, kind, leaf, S, C, rays in cases:
        report["cases"][name] = case_report(name, kind, leaf, S, C, rays)
    tot = {"fragile": 0, "A": 0, "B": 0, "C": 0, "D": 0, "unexplained": 0, "rays": 0}
    for c in report["cases"].values():
        for r in c.values():
            tot["rays"] += r["rays"]; tot["fragile"] += r["fragile"]; tot["A"] += r["fragile_class_A_reference_disagrees_with_itself"]
            tot["B"] += r["fragile_class_B_second_candidate_within_eps"]; tot["C"] += r["fragile_class_C_same_hit_uv_within_1e-3"]
            tot["D"] += r["fragile_class_D_oracle_flips_under_3e-6_perturbation"]
            tot["unexplained"] += r["fragile_unexplained"]
    report["totals"] = tot
    print(tot)
    if "--write" in sys.argv:
        with open(os.path.join(HERE, "parity_report.json"), "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
