"""GPU-vs-oracle parity on the BASELINE.json scenes: rel-L2 of the un-normalised film and of the normalised RGB image,
ray-count equality, timings of both sides.  Writes one JSON line per config (profiles/r1_parity_report.jsonl)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/parity_report.jsonl")
ap.add_argument("--full", action="store_true", help="full spp for configs 3/4 (minutes of CPU time)")
a = ap.parse_args()

configs = [
    ("configs[0] bunny normals 768x768x1 (per-block seeding)", S.config_bunny()),
    ("configs[1] ajax-ao 800x600x64", S.config_ajax_ao()),
    ("configs[2] cbox path_mis 512x512x%d" % (256 if a.full else 64), S.config_cbox(512, 512, 256 if a.full else 64)),
    ("configs[3] ajax-rough path_mis 768x768x%d" % (1024 if a.full else 64), S.config_ajax_microfacet(768, 768, 1024 if a.full else 64)),
    ("configs[4] random 1M tris ao 1920x1080x2 (10M: see bench)", S.config_random_tris(1_000_000, 1920, 1080, 2, S.INT_AO)),
]
ctx = abi.Context(0)
with open(a.out, "w") as fh:
    for name, sc in configs:
        ctx.load(sc)
        film, st = ctx.render()
        t0 = time.time()
        o = po.OracleScene(sc)
        ofilm, ost = o.render(accel=1)
        o.close()
        W, H, b = sc.camera.width, sc.camera.height, sc.border
        rgb = ctx.film_to_rgb(film)
        orgb = po.film_to_rgb(ofilm, W, H, b)
        row = {"config": name, "triangles": sc.n_tris, "samples": int(st.samples), "gpu_rays": int(st.rays), "oracle_rays": int(ost.rays),
               "rel_l2_film": S.rel_l2(film, ofilm), "rel_l2_rgb": S.rel_l2(rgb, orgb), "max_abs_rgb": float(np.max(np.abs(rgb - orgb))),
               "gpu_kernel_ms": st.kernel_ms, "oracle_seconds": ost.seconds, "oracle_threads": os.cpu_count(), "pass_1e-4": bool(S.rel_l2(film, ofilm) <= 1e-4)}
        print(json.dumps(row), flush=True)
        fh.write(json.dumps(row) + "\n")
