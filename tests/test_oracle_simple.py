"""The oracle's `simple` point-light integrator (ORC_INT_SIMPLE; named by ref scenes/pa3/ajax-simple.xml:8-11, no source
in the reference) against its closed form: Li(x) = Phi / (4 pi^2) * max(0, cos theta) / |x - p|^2 * V(x <-> p)."""
import numpy as np

from nori_b200 import scene as S

PHI = (100.0, 50.0, 25.0)
H = 4.0


def floor_scene(extra=None, spp=64):
    V = np.array([[-5, -5, 0], [5, -5, 0], [5, 5, 0], [-5, 5, 0]], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    meshes = [S.Mesh(V, F, name="floor")] + ([extra] if extra is not None else [])
    cam = S.Camera(S.lookat(origin=[0, 0, 10], target=[0, 0, 0], up=[0, 1, 0]).astype(np.float32), 20.0, 32, 32)
    sc = S.Scene(meshes, cam, S.INT_SIMPLE, spp, name="floor-simple", light_pos=(0.0, 0.0, H), light_energy=PHI)
    sc.filter_table, sc.filter_radius = S.box_table()
    return sc


def expected_rgb(o, sc):
    exp = np.zeros((32, 32, 3))
    for y in range(32):
        for x in range(32):
            r = o.sample_ray(x + 0.5, y + 0.5)
            t = -r["o"][2] / r["d"][2]
            p = r["o"] + t * r["d"]
            d2 = p[0] ** 2 + p[1] ** 2 + H ** 2
            exp[y, x] = np.array(PHI) / (4 * np.pi ** 2) * (H / np.sqrt(d2)) / d2
    return exp


def test_simple_matches_closed_form(oracle):
    sc = floor_scene()
    o = oracle.OracleScene(sc)
    film, st = o.render(accel=1)
    film0, st0 = o.render(accel=0)
    assert st.rays == st0.rays == 2 * 32 * 32 * 64          # every camera ray hits the floor and sends one shadow ray
    assert S.rel_l2(film, film0) == 0.0                     # BVH and the reference's brute-force loop agree exactly
    rgb = oracle.film_to_rgb(film, 32, 32, sc.border)
    exp = expected_rgb(o, sc)
    assert np.allclose(rgb, exp, rtol=1e-2)                 # pixel mean over 64 jittered samples vs the centre value
    c = rgb[16, 16] * (4 * np.pi ** 2) * H * H
    assert np.allclose(c, PHI, rtol=5e-3)                   # under the light: cos = 1, distance = H


def test_simple_shadow_and_backface(oracle):
    # an occluder between light and floor: the pixels in its umbra receive nothing (and see the occluder's unlit back)
    Vo = np.array([[-.3, -.3, 2], [.3, -.3, 2], [.3, .3, 2], [-.3, .3, 2]], dtype=np.float32)
    occ = S.Mesh(Vo, np.array([[0, 2, 1], [0, 3, 2]], dtype=np.uint32), name="occluder")   # wound to face -z
    sc = floor_scene(occ, spp=4)
    o = oracle.OracleScene(sc)
    film, st = o.render(accel=1)
    rgb = oracle.film_to_rgb(film, 32, 32, sc.border)
    assert rgb[16, 16].max() == 0.0                         # the camera sees the occluder's top, whose normal (-z) faces away from the light
    assert rgb[16, 20].max() == 0.0                         # floor inside the umbra (|x| < 0.6): lit side up, but occluded
    assert rgb[2, 2].min() > 0.0                            # floor outside the umbra is lit
    assert st.rays < 2 * 32 * 32 * 4                        # back-facing hits send no shadow ray (cos <= 0)
