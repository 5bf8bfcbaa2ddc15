"""One frame of a workload under a profiler window (run on the GPU box): two warm-up frames, then cudaProfilerStart, ONE frame,
cudaProfilerStop -- for `ncu --profile-from-start off ...`.   python tools/probe.py <workload> [--spp N] [--opt k=v ...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (cudaProfilerStart/Stop through torch.cuda.profiler)

import bench  # noqa: E402
from nori_b200 import abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()


class A:
    width = height = tris = 0
    spp = a.spp


sc = bench.build_scene(a.workload, A())
ctx = abi.Context(0)
for kv in a.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.load(sc)
for _ in range(2):
    film, st = ctx.render()
torch.cuda.synchronize()
torch.cuda.profiler.start()
film, st = ctx.render()
torch.cuda.profiler.stop()
print(f"probe {a.workload} {a.opt}: kernel {st.kernel_ms:.3f} ms, {st.rays / st.kernel_ms / 1e3:.1f} Mrays/s, launches {st.launches}")
ctx.close()
