"""One rank's share of an N-GPU frame on ONE GPU (tile shard rank/N through nb_set_tiles): render-kernel time with the L2
flushed before every frame, for a list of option sets.  Lets a 1-GPU box tune what an 8-GPU box will run.
   python tools/shard_probe.py ajax-ao 8 "guided=0" "guided=75" "chunk=1" ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from nori_b200 import abi  # noqa: E402

wl, N = sys.argv[1], int(sys.argv[2])
optsets = sys.argv[3:] or ["guided=0"]
sc = bench.build_scene(wl)
ctx = abi.Context(0)
ctx.load(sc)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
for spec in optsets:
    for kv in spec.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    per_rank = []
    for r in (range(N) if N <= 8 else [0]):
        ctx.set_tiles(r, N)
        ms = []
        for i in range(8):
            flush.fill_(i); torch.cuda.synchronize()
            film, st = ctx.render()
            if i >= 3:
                ms.append(st.kernel_ms)
        per_rank.append(float(np.mean(ms)))
    ctx.set_tiles(0, 1)
    print(f"SHARD {wl} N={N} [{spec}] kernel ms: max over ranks {max(per_rank):.4f}  mean {np.mean(per_rank):.4f}  per rank {[round(x, 3) for x in per_rank]}", flush=True)
    for kv in spec.split(","):          # back to defaults
        k, v = kv.split("=")
        ctx.set_option(k, -1 if k == "guided" else 0)
ctx.close()
