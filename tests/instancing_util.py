"""Shared builder of the instanced test scene (BASELINE config 4: instanced multi-object subdiv scene)."""
import numpy as np


def affine(scale, angle, t):
    """Column-major 3x4 (vx, vy, vz, p): rotation about y, uniform scale, translation."""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([c * scale, 0, -s * scale, 0, scale, 0, s * scale, 0, c * scale, t[0], t[1], t[2]], np.float32)


def build_instanced(po, leaf, objects_spec, grid=2, spacing=3.0):
    """objects_spec: list of (kind, S, C).  Returns (top RefScene, [object RefScenes])."""
    objs = []
    for kind, S, C in objects_spec:
        o = po.RefScene(leaf, S, C)
        if kind == "cube":
            o.add_cube(1, 1.0)
        else:
            o.add_obj()
        o.commit()
        objs.append(o)
    top = po.RefScene(leaf, objects_spec[0][1], objects_spec[0][2])
    k = 0
    for ix in range(-grid, grid + 1):
        for iz in range(-grid, grid + 1):
            o = objs[(ix + iz) % len(objs)]
            top.add_instance(o, affine(0.6 + 0.1 * ((ix * 5 + iz) % 3), 0.3 * k, (spacing * ix, 0.2 * ((k % 4) - 1.5), spacing * iz)))
            k += 1
    top.commit()
    return top, objs


def import_instanced(ecb, po, top, objs):
    """Product side: object scenes and the instancing scene imported from LIVE reference memory."""
    lay = po.ref_layout()
    scenes = []
    for o in objs:
        r, _ = o.root()
        scenes.append(ecb.Scene.import_bvh4(r, o.leaf, o.C, o.S, lay["proj_col_major"], 0, o.bounds()))
    root, ids, xf, handles = top.instance_info()
    by_handle = {o.handle(): i for i, o in enumerate(objs)}
    inst = ecb.Scene.import_instances(root, ids, xf, [scenes[by_handle[h]] for h in handles], 0, top.bounds())
    return inst, scenes
