"""Golden vectors produced by the CPU ORACLE (not by the reference, which cannot be built here): fixed-seed rays ->
hit records, a small fixed-seed film per integrator.  Regenerate with `python tests/golden/make_oracle_vectors.py`.
They pin the oracle against accidental change and give the CUDA path size-independent known answers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nori_b200 import scene as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def golden_rays(n=512, seed=11):
    sc = S.config_bunny()
    V = sc.meshes[0].V
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=po.RAY_DTYPE)
    lo, hi = V.min(0), V.max(0)
    o = 0.5 * (lo + hi) + rng.normal(size=(n, 3)) * float(np.max(hi - lo)) * 1.5
    d = (lo + rng.random((n, 3)) * (hi - lo)) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["o"], rays["d"], rays["mint"], rays["maxt"] = o.astype(np.float32), d.astype(np.float32), 1e-4, np.inf
    return sc, rays


def golden_films():
    out = {}
    for name in ("whitted", "path_mats", "path_ems", "path_mis"):
        sc = S.config_cbox(32, 32, 4, S.INTEGRATORS[name])
        sc.meshes[3] = S.with_(sc.meshes[3], S.microfacet((0.2, 0.2, 0.4), 0.28, 1.7))
        sc.meshes[4] = S.with_(sc.meshes[4], S.dielectric())
        out[name] = sc
    return out


def golden_extras():
    """Vectors of the entry points added late in round 1 (kept in their own file so that oracle_vectors.npz and the
    GPU test that was validated against it stay untouched): a `simple` film and per-path luminances (li_samples)."""
    cam = S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 48, 36)
    simple = S.Scene([S.ajax_standin(1)], cam, S.INT_SIMPLE, 2, name="golden-simple",
                     light_pos=(-20.0, 40.0, 20.0), light_energy=(3.76e4, 3.76e4, 3.76e4))
    li = golden_films()["path_mis"]
    li.seed = 5
    return simple, li


if __name__ == "__main__":
    sc, rays = golden_rays()
    o = po.OracleScene(sc)
    hits, _ = o.intersect(rays, accel=0)
    full = o.intersect_full(rays, accel=0)
    data = {"rays": rays, "hits": hits, "records": full}
    for name, fsc in golden_films().items():
        film, st = po.OracleScene(fsc).render(accel=1, nthreads=1)
        data["film_" + name] = film
        data["rays_" + name] = np.array([st.rays], dtype=np.uint64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz"), **data)
    print({k: v.shape for k, v in data.items()})
    simple, li = golden_extras()
    film, st = po.OracleScene(simple).render(accel=1, nthreads=1)
    extra = {"film_simple": film, "rays_simple": np.array([st.rays], dtype=np.uint64), "li_path_mis": po.OracleScene(li).li_samples(4096)}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors_extra.npz"), **extra)
    print({k: v.shape for k, v in extra.items()})
