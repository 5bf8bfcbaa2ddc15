"""Host binned-SAH vs device LBVH: build seconds and render-kernel ms on the same scene (results are identical)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S
import bench
class A: width=height=spp=tris=0
for wl, spp in [("ajax-ao", 64), ("cbox-mis", 64), ("random10m-ao", 4)]:
    a = A(); a.spp = spp
    sc = bench.WORKLOADS[wl](a)
    for b in (0, 1):
        ctx = abi.Context(0); ctx.set_option("builder", b)
        t0 = time.time(); ctx.load(sc); wall = time.time() - t0
        ctx.render(); _, st = ctx.render()
        print(json.dumps({"workload": wl, "spp": spp, **ctx.build_stats(), "load_wall_s": round(wall, 3), **ctx.scene_info(),
                          "kernel_ms": st.kernel_ms, "mrays_s": st.rays / st.kernel_ms / 1e3}), flush=True)
        ctx.close()
