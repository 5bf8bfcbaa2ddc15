"""Tiny end-to-end run for compute-sanitizer (memcheck): every kernel, ragged sizes, both builders, TMA path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S
ctx = abi.Context(0)
for integ in (S.INT_NORMALS, S.INT_AO, S.INT_WHITTED, S.INT_PATH_MATS, S.INT_PATH_EMS, S.INT_PATH_MIS):
    sc = S.config_cbox(37, 29, 2, integ)
    sc.meshes[3] = S.with_(sc.meshes[3], S.microfacet((0.2, 0.2, 0.4), 0.28, 1.7)); sc.meshes[4] = S.with_(sc.meshes[4], S.dielectric())
    ctx.load(sc); f, st = ctx.render(); assert np.isfinite(f).all()
sc = S.config_cbox(37, 29, 2, S.INT_PATH_MIS); sc.seed_mode = S.SEED_PER_BLOCK
ctx.load(sc); ctx.render()
sc = S.config_bunny(); sc.camera.width = 70; sc.camera.height = 45
ctx.load(sc); ctx.render()
ctx.set_option("smem_nodes", 64); ctx.render(); ctx.set_option("smem_nodes", 0)
ctx.set_option("count", 1); ctx.render(); ctx.set_option("count", 0)
ctx.set_option("builder", 1); ctx.load(sc); f, st = ctx.render(); ctx.set_option("builder", 0); ctx.load(sc)
rays = np.zeros(100, dtype=abi.RAY_DTYPE); rays["d"] = [0, 0, -1]; rays["o"] = [0, 0.1, 1]; rays["mint"] = 1e-4; rays["maxt"] = np.inf
ctx.intersect(rays); ctx.intersect(rays, shadow=True); ctx.intersect_full(rays); ctx.film_to_rgb(f)
print("sanitize_smoke: ok")
