"""Tail-cut sweep (run on the GPU box): time the render at several values of the "tail" option and check that the film
and the ray count do not depend on it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nori_b200 import abi, scene as S
import bench
class A: width=height=spp=tris=0
tails = [int(x) for x in os.environ.get("TAILS", "0 4 8 12 16").split()]
for wl in sys.argv[1:] or ["ajax-ao", "cbox-mis"]:
    a = A()
    if wl == "ajax-rough": a.spp = 64
    sc = bench.build_scene(wl, a)
    ctx = abi.Context(0); ctx.load(sc)
    ref = None
    for t in tails:
        ctx.set_option("tail", t)
        best = 1e30
        for i in range(6):
            film, st = ctx.render()
            if i >= 2: best = min(best, st.kernel_ms)
        if ref is None: ref = (film.copy(), int(st.rays))
        print(json.dumps({"workload": wl, "tail": t, "ms": round(best, 3), "mrays": round(st.rays / best / 1e3, 1),
                          "rays_equal": int(st.rays) == ref[1], "rel_l2_vs_tail0": float(S.rel_l2(film, ref[0]))}), flush=True)
    ctx.close()
