"""Writes the BASELINE.json workloads as Nori XML scenes (+ OBJ meshes) for the `nori` CLI:

    python scenes/make_scenes.py [out_dir]      ->  out_dir/{bunny-normals,ajax-ao,cbox-mis,ajax-rough}.xml
    nori_b200/lib/nori out_dir/ajax-ao.xml      ->  out_dir/ajax-ao.exr / .png

Geometry comes from tests/golden/ref_meshes.npz (extracted from the reference's shipped assets) and from the
deterministic generators in nori_b200/scene.py (the Ajax stand-in: ajax.obj is not shipped with the reference).
The generated OBJ files are large (ajax stand-in: 512 k triangles) and are not committed.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import host, scene as S  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "generated")
for name, sc in [("bunny-normals", S.config_bunny()), ("ajax-ao", S.config_ajax_ao()),
                 ("cbox-mis", S.config_cbox()), ("ajax-rough", S.config_ajax_microfacet())]:
    print(host.write_xml(sc, out, name))
