"""bench.py -- BASELINE.json's metric on BASELINE.json's config, on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the CPU arm: the Nori-structured host loop, all cores)

A "step" is one pass of the render hot path over one frame: BVH traversal + integrator + film splat for every
(pixel, sample) of the workload, merged into the un-normalised weighted film.  Workload at every N is
BASELINE configs[1]: Ajax (stand-in mesh, ajax.obj is not shipped) ambient occlusion, 800x600, 64 spp;
32x32 image tiles are sharded tile_id % N across ranks (strong scaling) and the finished ImageBlocks are
gathered to rank 0 over NCCL at frame end.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from nori_b200 import scene as S  # noqa: E402

WORKLOADS = {
    "ajax-ao": lambda a: S.config_ajax_ao(a.width or 800, a.height or 600, a.spp or 64),
    "cbox-mis": lambda a: S.config_cbox(a.width or 512, a.height or 512, a.spp or 256, S.INT_PATH_MIS),
    "ajax-rough": lambda a: S.config_ajax_microfacet(a.width or 768, a.height or 768, a.spp or 1024),
    "random10m-ao": lambda a: S.config_random_tris(a.tris or 10_000_000, a.width or 1920, a.height or 1080, a.spp or 4, S.INT_AO),
    "random10m-normals": lambda a: S.config_random_tris(a.tris or 10_000_000, a.width or 1920, a.height or 1080, a.spp or 4, S.INT_NORMALS),
    "bunny": lambda a: S.config_bunny(),
}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.thread = [], None, None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                c, m = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = m
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(c)
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:] if r[1]]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload (None if not captured)."""
    p = os.path.join(REPO, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get("bytes")
    except Exception:
        return None


def algorithmic_bytes(st, scene):
    """64 B per BVH node visit + 48 B per triangle test + 36 B of normals per shaded closest hit
    (+24 B UVs when present) + one film write; rays are generated and consumed in registers (0 B)."""
    b = scene.border
    film = (scene.camera.width + 2 * b) * (scene.camera.height + 2 * b) * 16
    has_uv = any(m.UV is not None for m in scene.meshes)
    return 64 * st.node_visits + 48 * st.tri_tests + (36 + (24 if has_uv else 0)) * st.hits_shaded + film


def run_reference(args, rank, world):
    """CPU arm: the Nori-structured host loop (32x32 tiles, private ImageBlock + sampler per worker, merged film;
    ref: src/main.cpp:85-113) with a CPU BVH in place of the brute-force Accel -- the oracle port, all host cores.
    The reference binary itself cannot be built here (empty ext/ submodules), hence kind = "port"."""
    if rank != 0:
        return
    from oracle import pyoracle
    scene = WORKLOADS[args.workload](args)
    full_spp = scene.spp
    scene.spp = max(1, min(full_spp, args.ref_spp))
    cores = os.cpu_count() or 1
    o = pyoracle.OracleScene(scene)
    accel = 0 if args.accel == "brute" else 1
    for _ in range(max(1, args.warmup) if args.warmup else 0):
        o.render(accel=accel, nthreads=cores)
    secs, rays, samples = 0.0, 0, 0
    for _ in range(args.steps):
        _, st = o.render(accel=accel, nthreads=cores)
        secs += st.seconds; rays += st.rays; samples += st.samples
    mrays = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/sec", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "msamples_per_sec": samples / secs / 1e6,
        "config": {"workload": f"{scene.name} (Ajax stand-in mesh, {scene.n_tris} tris)", "width": scene.camera.width,
                   "height": scene.camera.height, "spp": full_spp, "integrator": "ao" if scene.integrator == 1 else scene.integrator},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "accel": args.accel,
                         "sample": f"full {scene.camera.width}x{scene.camera.height} frame at {scene.spp} of {full_spp} spp per step (Mrays/s is spp-independent)"},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ajax-ao", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--tris", type=int, default=0)
    ap.add_argument("--ref-spp", type=int, default=32, help="spp per step of the CPU arm (bounded sample)")
    ap.add_argument("--accel", default="bvh", choices=["bvh", "brute"],
                    help="CPU arm only: 'brute' = the reference's shipped brute-force Accel (ref: src/accel.cpp:30-43), feasible on --workload bunny")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="key=value tuning option (nb_set_option)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from nori_b200 import abi, multigpu as MG

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    scene = WORKLOADS[args.workload](args)
    ctx = abi.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.load(scene)
    info = ctx.scene_info()
    ctx.set_tiles(rank, world)
    n_mine, edge = ctx.tile_count(rank, world)
    n_max = max(ctx.tile_count(r, world)[0] for r in range(world))
    H2, W2, _ = scene.film_shape
    tstream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: the library enqueues on the handle we pass
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    blocks = torch.zeros((n_max, edge, edge, 4), dtype=torch.float32, device=dev)
    film = torch.zeros((H2, W2, 4), dtype=torch.float32, device=dev)
    film_host = torch.zeros((H2, W2, 4), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    kev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def step_device(want_stats=False):
        """render my tiles -> (gather finished blocks over NCCL) -> merge into the film on rank 0.  Without stats the whole
        step is enqueued without a host synchronisation; the render kernel is bracketed by CUDA events on the launch stream."""
        kev[0].record()
        st = ctx.render_blocks_device(blocks.data_ptr(), stream, want_stats)
        kev[1].record()
        gathered = MG.gather_blocks(blocks, world, rank, dst=0)      # ONE exchange per frame (NCCL send/recv over NVLink)
        if rank == 0:
            film.zero_()
            ctx.merge_all_blocks_device(MG.gathered_base(gathered).data_ptr(), world, n_max, film.data_ptr(), stream)   # one launch
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- instrumented pass (untimed): node visits / triangle tests / hits of this exact workload
    ctx.set_option("count", 1)
    st_count = step_device(want_stats=True)
    ctx.set_option("count", 0)
    counts = torch.tensor([st_count.rays, st_count.node_visits, st_count.tri_tests, st_count.hits_shaded, st_count.samples],
                          dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    tot_rays, tot_nodes, tot_tris, tot_hits, tot_samples = (int(x) for x in counts.tolist())

    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed region: K steps, device events on the launching stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        barrier()
        ev[i][0].record()
        step_device()
        ev[i][1].record()
        ev[i][1].synchronize()
        kernel_ms.append(kev[0].elapsed_time(kev[1]))     # block clear + render_kernel on the launch stream
    barrier()
    t_wall1 = time.time()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    kern = torch.tensor(kernel_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    total_ms = float(step_ms.sum())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- end to end through the reference-facing call: scene arrays H2D from pinned memory + render + film D2H
    e2e_steps = max(3, min(args.steps, 10))
    h2d = info["bytes"]
    d2h = H2 * W2 * 16
    e2e_ms = []
    for i in range(e2e_steps + 1):
        flush.fill_(i & 0xff)
        barrier()
        t0 = time.perf_counter()
        ctx.upload()                                  # host -> device: BVH nodes, triangles, vertex/normal/index arrays (pinned)
        if world == 1:
            ctx.render_host_ptr(film_host.data_ptr()) # nb_render: the call Nori's render() makes; film lands in host memory
        else:
            step_device()
            if rank == 0:
                film_host.copy_(film, non_blocking=True)
        barrier()
        dt = (time.perf_counter() - t0) * 1e3
        if i > 0:
            e2e_ms.append(dt)
    e2e_t = torch.tensor(e2e_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_mean_ms = float(e2e_t.mean())

    if rank == 0:
        peak, peak_src = measured_peak()
        class _S: pass
        agg = _S(); agg.node_visits, agg.tri_tests, agg.hits_shaded = tot_nodes, tot_tris, tot_hits
        alg_bytes = algorithmic_bytes(agg, scene)
        kern_ms_mean = float(kern.mean())
        achieved = alg_bytes / world / (kern_ms_mean * 1e-3) / 1e9      # per GPU: bytes one launch accounts for / its duration
        mrays = tot_rays * args.steps / (total_ms * 1e-3) / 1e6
        line = {
            "metric": "Mrays/sec", "value": mrays, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "msamples_per_sec": tot_samples * args.steps / (total_ms * 1e-3) / 1e6,
            "config": {"workload": f"{scene.name}: BASELINE configs[1] (Ajax stand-in mesh -- ajax.obj is not shipped)" if args.workload == "ajax-ao" else scene.name,
                       "triangles": info["tris"], "bvh_nodes": info["nodes"], "width": scene.camera.width, "height": scene.camera.height,
                       "spp": scene.spp, "integrator": {v: k for k, v in S.INTEGRATORS.items()}[scene.integrator],
                       "seeding": "pcg32 per (pixel, sample)", "parallelism": f"tiles%{world}" if world > 1 else "1 GPU",
                       "l2": "flushed between timed steps (256 MiB write)", "rays_per_step": tot_rays, "samples_per_step": tot_samples,
                       "scene_bytes": info["bytes"]},
            "clocks": clocks,
            "e2e": {"value": tot_rays / (e2e_mean_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_mean_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "nb_upload_scene (scene arrays from pinned host memory) + nb_render into a host film"},
            "gpu_launches": int(args.steps * (world + 1)),   # per step: one render_kernel per rank + one merge kernel on rank 0
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload) if world == 1 else None, "peak_source": peak_src,
                         "kernel": "render_kernel<%s>" % {v: k for k, v in S.INTEGRATORS.items()}[scene.integrator],
                         "kernel_ms": kern_ms_mean, "algorithmic_bytes_per_launch": alg_bytes / world,
                         "node_visits": tot_nodes, "tri_tests": tot_tris, "hits_shaded": tot_hits,
                         "note": "scene (%.0f MB) is L2-resident: this is EFFECTIVE bandwidth of the traversal, see profiles/ for DRAM bytes" % (info["bytes"] / 1e6)},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle
            cores = os.cpu_count() or 1
            cs = WORKLOADS[args.workload](args)
            full_spp = cs.spp
            cs.spp = max(1, min(full_spp, args.ref_spp))
            _, ost = pyoracle.OracleScene(cs).render(accel=1, nthreads=cores)
            line["cpu_baseline"] = {"value": ost.rays / ost.seconds / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "seconds": ost.seconds,
                                    "sample": f"full frame at {cs.spp} of {full_spp} spp (Mrays/s is spp-independent)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
