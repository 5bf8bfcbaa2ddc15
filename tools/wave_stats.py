"""Lock-step diagnostics: how much of the warp's walk time is lost to the longest ray of each wave."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S
import bench
class A: width=height=spp=tris=0
for wl in sys.argv[1:] or ["ajax-ao", "cbox-mis"]:
    a = A()
    if wl == "ajax-rough": a.spp = 64
    sc = bench.WORKLOADS[wl](a)
    ctx = abi.Context(0); ctx.load(sc); ctx.set_option("count", 1)
    _, st = ctx.render()
    c = ctx.debug_counters()
    waves, wmax = int(c[6]), int(c[5])
    print(json.dumps({"workload": wl, "rays": int(st.rays), "node_visits": int(st.node_visits), "waves": waves,
                      "mean_longest_walk_per_wave": wmax / max(waves, 1), "mean_walk_per_ray": st.node_visits / max(st.rays, 1),
                      "rays_per_wave": st.rays / max(waves, 1),
                      "lockstep_efficiency_nodes": st.node_visits / max(32 * wmax, 1)}))
    ctx.close()
