"""Tuning sweep of the render kernel's launch/scheduling knobs on one GPU (prints kernel ms per setting)."""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ajax-ao")
    ap.add_argument("--width", type=int, default=0); ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0); ap.add_argument("--tris", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tiles", default="0,1", help="rank,nranks tile shard to emulate one rank of an N-GPU run")
    ap.add_argument("--grid", default="blocks_per_sm=0;smem_nodes=0;chunk=0")
    a = ap.parse_args()
    sc = bench.WORKLOADS[a.workload](a)
    ctx = abi.Context(0)
    ctx.load(sc)
    r, n = (int(v) for v in a.tiles.split(","))
    ctx.set_tiles(r, n)
    print(json.dumps({"scene": sc.name, "tiles": a.tiles, **ctx.scene_info()}))
    axes = []
    for part in a.grid.split(";"):
        k, vs = part.split("=")
        axes.append([(k, int(v)) for v in vs.split(",")])
    defaults = {"blocks_per_sm": 0, "smem_nodes": 0, "chunk": 0}
    for combo in itertools.product(*axes):
        rebuild = False
        for k, v in combo:
            ctx.set_option(k, v)
            rebuild |= k in ("max_leaf", "bfs_nodes")
        if rebuild:
            ctx.load(sc)
            print(json.dumps({"rebuilt": dict(combo), **ctx.scene_info()}))
        ctx.render()
        ms = []
        for _ in range(a.reps):
            _, st = ctx.render()
            ms.append(st.kernel_ms)
        print(json.dumps({"opts": dict(combo), "kernel_ms": min(ms), "mrays_s": st.rays / min(ms) / 1e3}), flush=True)
    for k, v in defaults.items():
        try:
            ctx.set_option(k, v)
        except Exception:
            pass


if __name__ == "__main__":
    main()
