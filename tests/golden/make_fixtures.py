"""Generates tests/golden/ref_meshes.npz from the reference's shipped scene assets.

Run ONCE in the build container (where /root/reference exists); the .npz is committed so that
nothing on the GPU box reads /root/reference.  Geometry only (binary arrays), parsed with the
OBJ rules of ref: src/obj.cpp:43-112 (see nori_b200.scene.load_obj).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nori_b200.scene import load_obj  # noqa: E402

REF = "/root/reference/scenes"
SOURCES = {
    "bunny": "pa1/bunny.obj",
    "cbox_walls": "pa4/cbox/meshes/walls.obj",
    "cbox_leftwall": "pa4/cbox/meshes/leftwall.obj",
    "cbox_rightwall": "pa4/cbox/meshes/rightwall.obj",
    "cbox_light": "pa4/cbox/meshes/light.obj",
    "cbox_sphere1": "pa4/cbox/meshes/sphere1.obj",
    "cbox_sphere2": "pa4/cbox/meshes/sphere2.obj",
    "ajax_light": "pa5/ajax/light.obj",
    "test_floor": "pa4/tests/meshes/floor.obj",
    "test_furnace": "pa4/tests/meshes/furnace.obj",
    **{f"test_polylum{i}": f"pa4/tests/meshes/polylum{i}.obj" for i in range(1, 6)},
}

out = {}
for name, rel in SOURCES.items():
    m = load_obj(os.path.join(REF, rel))
    out[name + ".V"] = m.V
    out[name + ".F"] = m.F
    if m.N is not None:
        out[name + ".N"] = m.N
    if m.UV is not None:
        out[name + ".UV"] = m.UV
    print(f"{name}: V={m.V.shape[0]} F={m.F.shape[0]} N={m.N is not None} UV={m.UV is not None}")
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_meshes.npz"), **out)
