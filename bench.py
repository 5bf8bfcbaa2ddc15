"""bench.py -- BASELINE.json's metric on BASELINE.json's configs, on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the CPU arm: the Nori-structured host loop, all cores)

A "step" is one pass of the render hot path over one frame: BVH traversal + integrator + film splat for every
(pixel, sample) of the workload, merged into the un-normalised weighted film.  The HEADLINE workload at every N is
BASELINE configs[1]: Ajax (stand-in mesh, ajax.obj is not shipped) ambient occlusion, 800x600, 64 spp.  The other four
BASELINE configs are measured in the same run and reported as sub-records under "configs" in the ONE JSON line rank 0
prints (VERDICT r1 "Next" 1): configs[0] bunny normals 768x768x1 (reference per-block seeding), configs[2] Cornell box
path tracer 512x512x256, configs[3] Ajax microfacet path tracer 768x768x1024, configs[4] 10 M random triangles
1920x1080x4096.

N > 1: 32x32 image tiles are sharded tile_id % N across ranks (strong scaling); the exchange is INSIDE libnori_b200.so
(nb_comm_init_rank / nb_render_gather: one grouped ncclSend/ncclRecv of the finished ImageBlocks to rank 0 per frame, one
merge launch); torch.distributed only ships the 128-byte communicator id, the barriers and the max-over-ranks timings.
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

# name -> (BASELINE.json configs index, builder(args, geometry))
WORKLOADS = {
    "bunny": (0, lambda a, g=True: S.config_bunny()),
    "ajax-ao": (1, lambda a, g=True: S.config_ajax_ao(a.width or 800, a.height or 600, a.spp or 64, geometry=g)),
    "cbox-mis": (2, lambda a, g=True: S.config_cbox(a.width or 512, a.height or 512, a.spp or 256, S.INT_PATH_MIS)),
    "ajax-rough": (3, lambda a, g=True: S.config_ajax_microfacet(a.width or 768, a.height or 768, a.spp or 1024, geometry=g)),
    "random10m-ao": (4, lambda a, g=True: S.config_random_tris(a.tris or 10_000_000, a.width or 1920, a.height or 1080, a.spp or 4096, S.INT_AO, geometry=g)),
    "random10m-normals": (4, lambda a, g=True: S.config_random_tris(a.tris or 10_000_000, a.width or 1920, a.height or 1080, a.spp or 4096, S.INT_NORMALS, geometry=g)),
}
SUB_WORKLOADS = ["bunny", "cbox-mis", "ajax-rough", "random10m-ao"]     # configs[0], [2], [3], [4]
INT_NAMES = {v: k for k, v in S.INTEGRATORS.items()}


class _NoOverride:
    width = height = spp = tris = 0


def build_scene(name, args=None, geometry=True):
    return WORKLOADS[name][1](args or _NoOverride(), geometry)


def host_cores():
    """Cores this process may actually use: the affinity mask, capped by the cgroup CPU quota (not os.cpu_count())."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    n = min(n, max(1, int(q / int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read()) + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


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


def ncu_capture(workload):
    """Summary of the committed `ncu --set full` capture of this workload's render kernel (profiles/ncu_traffic.json):
    DRAM bytes per launch plus the counters that name the kernel's REAL bound (issue slots x active lanes, LSU wavefronts)."""
    try:
        return json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json"))).get(workload) or {}
    except Exception:
        return {}


def node_bytes():
    from nori_b200 import abi
    return int(abi.lib().nb_node_bytes())


def algorithmic_bytes(node_visits, tri_tests, hits_shaded, scene, has_uv):
    """the bytes of one node per BVH node visit (64 B for the binary node the survey assumed, 80 B for the
    8-wide compressed node this build walks -- nb_node_bytes()) + 48 B per triangle test + 36 B of normals per shaded
    closest hit (+24 B UVs when present) + one film write; rays are generated and consumed in registers (0 B)."""
    b = scene.border
    film = (scene.camera.width + 2 * b) * (scene.camera.height + 2 * b) * 16
    return node_bytes() * node_visits + 48 * tri_tests + (36 + (24 if has_uv else 0)) * hits_shaded + film


def config_dict(name, scene, n_tris):
    """The workload description both arms print identically (the driver compares the two lines' `config`)."""
    idx = WORKLOADS[name][0]
    note = " (Ajax stand-in mesh -- ajax.obj is not shipped)" if name.startswith("ajax") else ""
    return {"workload": f"BASELINE configs[{idx}] {name}{note}", "triangles": int(n_tris), "width": scene.camera.width,
            "height": scene.camera.height, "spp": int(scene.spp), "integrator": INT_NAMES[scene.integrator],
            "seeding": "pcg32 per 32x32 block (reference)" if scene.seed_mode == S.SEED_PER_BLOCK else "pcg32 per (pixel, sample)"}


def cpu_sample_spp(name, scene, cap):
    """Samples per pixel of the bounded CPU sample: whole frame, the first k sample streams of every pixel (sample i of a
    pixel owns the stream seed(pixel, i), so Mrays/s does not depend on k)."""
    return max(1, min(int(scene.spp), int(cap)))


# ----------------------------------------------------------------------------------------------- CPU arm
def run_reference(args, rank, world):
    """CPU arm: the Nori-structured host loop (32x32 tiles, private ImageBlock + sampler per worker, merged film;
    ref: src/main.cpp:85-113) with a CPU BVH in place of the brute-force Accel -- the oracle port, all usable host cores,
    on the SAME config (full spp) as the GPU arm.  The reference binary itself cannot be built here (empty ext/
    submodules), hence kind = "port"."""
    if rank != 0:
        return
    from oracle import pyoracle
    scene = build_scene(args.workload, args)
    cfg = config_dict(args.workload, scene, scene.n_tris)
    full_spp = scene.spp
    if args.ref_spp > 0:
        scene.spp = max(1, min(full_spp, args.ref_spp))
    cores = host_cores()
    o = pyoracle.OracleScene(scene)
    accel = 0 if args.accel == "brute" else 1
    for _ in range(args.warmup):
        o.render(accel=accel, nthreads=cores)
    secs, rays, samples = 0.0, 0, 0
    for _ in range(args.steps):
        _, st = o.render(accel=accel, nthreads=cores)
        secs += st.seconds; rays += st.rays; samples += st.samples
    mrays = rays / secs / 1e6
    sample = (f"every step renders the whole {scene.camera.width}x{scene.camera.height} frame at the full {full_spp} spp" if scene.spp == full_spp
              else f"whole frame at {scene.spp} of {full_spp} spp per step (Mrays/s is spp-independent)")
    line = {
        "impl": "reference", "metric": "Mrays/sec", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "msamples_per_sec": samples / secs / 1e6,
        "config": cfg,
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "os_cpu_count": os.cpu_count(), "kind": "port", "accel": args.accel,
                         "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class Harness:
    """One rank's view of the job: torch for device memory / streams / barriers, the C-ABI for everything else."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from nori_b200 import abi
        self.torch, self.dist, self.abi, self.args = torch, dist, abi, args
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the b200 arm has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.tstream = torch.cuda.Stream(device=self.dev)   # a real (non-default) stream: the library enqueues on the handle we pass
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)    # > 126 MB L2
        self.ctx = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="sum"):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()

    def open(self, name, spp=None):
        """Scene on every rank of the group: rank 0 generates + builds, the others receive the arrays over NVLink."""
        torch, abi = self.torch, self.abi
        if self.ctx is not None:
            self.ctx.close()
        scene = build_scene(name, self.args if name == self.args.workload else None, geometry=(self.rank == 0))
        if spp:
            scene.spp = spp
        ctx = abi.Context(self.local)
        if self.world > 1:
            uid = torch.tensor(list(abi.Context.comm_unique_id()) if self.rank == 0 else [0] * 128, dtype=torch.uint8, device=self.dev)
            self.dist.broadcast(uid, 0)
            ctx.comm_init_rank(bytes(uid.cpu().tolist()), self.rank, self.world)
        for kv in self.args.opt:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        ctx.load(scene)
        self.ctx, self.scene, self.name = ctx, scene, name
        self.info = ctx.scene_info()
        H2, W2, _ = scene.film_shape
        self.film = torch.zeros((H2, W2, 4), dtype=torch.float32, device=self.dev) if self.rank == 0 else None
        self.film_host = torch.zeros((H2, W2, 4), dtype=torch.float32).pin_memory() if self.rank == 0 else None
        return scene

    def step(self, want_stats=False):
        """render my tiles -> ONE NCCL gather of finished blocks on rank 0 -> ONE merge launch there (all inside the library)."""
        fp = self.film.data_ptr() if self.rank == 0 else 0
        return self.ctx.render_gather(fp, self.stream, want_stats)

    def count_pass(self):
        """Instrumented pass (untimed): rays / node visits / triangle tests / hits of this exact workload, over all ranks."""
        self.ctx.set_option("count", 1)
        st = self.step(want_stats=True)
        self.ctx.set_option("count", 0)
        return [int(x) for x in self.reduce([st.rays, st.node_visits, st.tri_tests, st.hits_shaded, st.samples])]

    def timed(self, steps, warmup, sampler=None):
        """W warm-up steps, then K timed steps: CUDA events on the launching stream, L2 flushed between steps, barrier on
        both sides, max over ranks.  Returns (ms per step [K], render-kernel ms per step [K], clocks)."""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.barrier()
        if sampler is not None and self.rank == 0:
            sampler.start()
            time.sleep(0.25)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kernel_ms = []
        self.barrier()
        t0 = time.time()
        for i in range(steps):
            self.flush.fill_(i & 0xff)
            self.barrier()
            ev[i][0].record()
            self.step()
            ev[i][1].record()
            ev[i][1].synchronize()
            kernel_ms.append(self.ctx.last_kernel_ms())       # the library's own events around its render kernel, same stream
        self.barrier()
        t1 = time.time()
        step_ms = self.reduce([a.elapsed_time(b) for a, b in ev], "max")
        kern = self.reduce(kernel_ms, "max")
        clocks = sampler.stop(t0, t1) if (sampler is not None and self.rank == 0) else None
        return step_ms, kern, clocks

    def e2e(self, steps):
        """End to end through the reference-facing calls with HOST buffers: scene arrays host->device from pinned memory
        (rank 0 over PCIe, then NVLink broadcast inside the library) + render + film device->host on rank 0."""
        ms = []
        for i in range(steps + 1):
            self.flush.fill_(i & 0xff)
            self.barrier()
            t0 = time.perf_counter()
            self.ctx.upload()
            if self.world == 1:
                self.ctx.render_host_ptr(self.film_host.data_ptr())   # nb_render: the call Nori's render() makes; film lands in host memory
            else:
                self.step()
                if self.rank == 0:
                    self.film_host.copy_(self.film, non_blocking=True)
            self.barrier()
            if i > 0:
                ms.append((time.perf_counter() - t0) * 1e3)
        return float(np.mean(self.reduce(ms, "max")))

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None


def roofline_record(name, scene, counts, kern_ms, world, has_uv, info):
    peak, peak_src = measured_peak()
    rays, nodes, tris, hits, samples = counts
    alg = algorithmic_bytes(nodes, tris, hits, scene, has_uv)
    achieved = alg / world / (kern_ms * 1e-3) / 1e9          # per GPU: bytes one launch accounts for / its duration
    cap = dict(ncu_capture(name))
    if cap.get("capture_spp") and cap.get("bytes"):
        k = scene.spp / cap["capture_spp"]        # the capture rendered fewer samples per pixel than the frame: DRAM bytes and time scale with them
        cap["bytes"] = int(cap["bytes"] * k); cap["captured_ms"] = cap.get("captured_ms", 0) * k
        cap["capture"] = f'{cap.get("capture")} (captured at {cap["capture_spp"]} spp, bytes and time scaled x{k:g} to the {scene.spp}-spp frame)'
    rec = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "frac_is": "EFFECTIVE (cache-served): algorithmic bytes of SURVEY 8d / kernel time; the scene is L2-resident, see `secondary` for the real bound",
           "traffic": cap.get("bytes") if world == 1 else None, "peak_source": peak_src,
           "kernel": "render_kernel<%s>" % INT_NAMES[scene.integrator], "kernel_ms": kern_ms,
           "algorithmic_bytes_per_launch": alg / world, "node_visits": nodes, "node_bytes": node_bytes(), "tri_tests": tris, "hits_shaded": hits,
           "scene_mb": info["bytes"] / 1e6}
    if cap:
        sec = {k: cap[k] for k in ("issue_active_pct", "lanes_per_inst", "lsu_wavefronts_pct_of_peak", "l1_hit_pct", "l2_hit_pct",
                                   "no_instruction_stall_pct", "capture", "captured_ms") if k in cap}
        if "issue_active_pct" in cap and "lanes_per_inst" in cap:
            sec["simd_lane_issue_utilisation"] = cap["issue_active_pct"] / 100.0 * cap["lanes_per_inst"] / 32.0
        if cap.get("bytes") and cap.get("captured_ms"):
            sec["dram_gbs"] = cap["bytes"] / (cap["captured_ms"] * 1e-3) / 1e9
            sec["dram_frac_of_peak"] = sec["dram_gbs"] / peak
        sec["reading"] = "the kernel is bound by instruction issue x active lanes and by L1 (LSU wavefronts of lane-divergent node fetches), not by HBM"
        rec["secondary"] = sec
    return rec


def cpu_check(name, scene, h, cap_spp):
    """The CPU leg of one config (rank 0, N = 1): the oracle port timed on a bounded sample of the SAME workload, and the
    GPU film of that same sample compared with the oracle's (rel-L2) -- the checker, outside every timed region."""
    from oracle import pyoracle
    cores = host_cores()
    k = cpu_sample_spp(name, scene, cap_spp)
    full = scene.spp
    scene.spp = k
    try:
        h.ctx.configure(scene)
        film, st = h.ctx.render()
        o = pyoracle.OracleScene(scene)
        ofilm, ost = o.render(accel=1, nthreads=cores)
        o.close()
    finally:
        scene.spp = full
        h.ctx.configure(scene)
    return {"value": ost.rays / ost.seconds / 1e6, "unit": "Mrays/s", "cores": cores, "os_cpu_count": os.cpu_count(), "kind": "port",
            "seconds": ost.seconds, "sample": f"whole frame at {k} of {full} spp (Mrays/s is spp-independent)"}, \
        {"rel_l2": S.rel_l2(film, ofilm), "rays_equal": int(st.rays) == int(ost.rays), "at_spp": k}


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
    ap.add_argument("--ref-spp", type=int, default=0, help="CPU arm: spp per step (0 = the workload's full spp, i.e. the same config as the GPU arm)")
    ap.add_argument("--accel", default="bvh", choices=["bvh", "brute"],
                    help="CPU arm only: 'brute' = the reference's shipped brute-force Accel (ref: src/accel.cpp:30-43), feasible on --workload bunny")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--configs", default=",".join(SUB_WORKLOADS), help="comma-separated sub-record workloads")
    ap.add_argument("--opt", action="append", default=[], help="key=value tuning option (nb_set_option)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    h = Harness(args)
    rank, world = h.rank, h.world
    args.warmup = max(args.warmup, 3)

    # ------------------------------------------------------------------ headline: BASELINE configs[1]
    scene = h.open(args.workload)
    has_uv = any(m.UV is not None for m in scene.meshes)
    counts = h.count_pass()
    step_ms, kern, clocks = h.timed(args.steps, args.warmup, ClockSampler(h.local))
    total_ms = float(np.sum(step_ms))
    e2e_ms = h.e2e(max(3, min(args.steps, 10)))
    tot_rays, tot_nodes, tot_tris, tot_hits, tot_samples = counts
    line = None
    if rank == 0:
        H2, W2, _ = scene.film_shape
        cfg = config_dict(args.workload, scene, h.info["tris"])
        line = {
            "metric": "Mrays/sec", "value": tot_rays * args.steps / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "msamples_per_sec": tot_samples * args.steps / (total_ms * 1e-3) / 1e6,
            "config": cfg,
            "details": {"bvh_nodes": h.info["nodes"], "scene_bytes": h.info["bytes"], "rays_per_step": tot_rays, "samples_per_step": tot_samples,
                        "parallelism": f"tiles%{world}, blocks gathered with one grouped ncclSend/ncclRecv inside libnori_b200.so" if world > 1 else "1 GPU",
                        "l2": "flushed between timed steps (256 MiB write)"},
            "clocks": clocks,
            "e2e": {"value": tot_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h.info["bytes"]), "d2h_bytes_per_step": int(H2 * W2 * 16),
                    "what": "nb_upload_scene (the scene arena from pinned host memory; N>1: sharded, 1/N per rank over its own PCIe link + one in-place ncclAllGather over NVLink) + nb_render into a host film (N>1: nb_render_gather + film device->host on rank 0)"},
            "gpu_launches": int(args.steps * (world + 1)),   # per step: one render_kernel per rank + one merge kernel on rank 0
            "roofline": roofline_record(args.workload, scene, counts, float(np.mean(kern)), world, has_uv, h.info),
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, chk = cpu_check(args.workload, scene, h, scene.spp)      # the headline's CPU leg runs the SAME config, full spp
            line["cpu_baseline"] = cb
            line["parity"] = chk

    # ------------------------------------------------------------------ the other BASELINE configs, same run
    subs = {}
    names = [] if args.no_configs else [n for n in args.configs.split(",") if n and n != args.workload]
    for name in names:
        if name not in WORKLOADS:
            raise SystemExit(f"unknown workload {name}")
        t_sub = time.time()
        sc = h.open(name)
        full_spp = sc.spp
        # long frames: warm up on a short version of the same frame, time a few (or one) full frames
        frame_ms_guess = {"bunny": 1, "cbox-mis": 70, "ajax-rough": 420, "random10m-ao": 13000, "random10m-normals": 6000}.get(name, 100) / world
        k_steps = int(max(1, min(args.steps, 5, 3000 // max(1, frame_ms_guess))))
        if frame_ms_guess > 1500:
            sc.spp = max(1, full_spp // 256); h.ctx.configure(sc)
            for _ in range(3):
                h.step()
            sc.spp = full_spp; h.ctx.configure(sc)
            cnt_spp = max(1, full_spp // 64)                  # instrumented pass on 1/64 of the samples, scaled: counts are per-sample sums
            sc.spp = cnt_spp; h.ctx.configure(sc)
            c_part = h.count_pass()
            sc.spp = full_spp; h.ctx.configure(sc)
            scale = full_spp / cnt_spp
            cnts = [int(round(c * scale)) for c in c_part]
            cnt_note = f"node/triangle/hit counts from an instrumented pass at {cnt_spp} spp, scaled x{scale:g}; rays and time are of the full frame"
            s_ms, s_kern, _ = h.timed(k_steps, 0)
            w_used = 3
        else:
            cnts = h.count_pass()
            cnt_note = "counts from an instrumented pass of the full frame"
            s_ms, s_kern, _ = h.timed(k_steps, 3)
            w_used = 3
        # exact ray count of the timed configuration (uninstrumented statistics pass)
        st_full = h.step(want_stats=True)
        rays_full, samples_full = [int(x) for x in h.reduce([st_full.rays, st_full.samples])]
        cnts[0], cnts[4] = rays_full, samples_full
        if rank == 0:
            ms = float(np.mean(s_ms))
            rec = {"config": config_dict(name, sc, h.info["tris"]), "ms_per_step": ms, "steps": k_steps, "warmup": w_used,
                   "value": rays_full / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "msamples_per_sec": samples_full / (ms * 1e-3) / 1e6,
                   "rays_per_step": rays_full, "n_gpus": world,
                   "roofline": roofline_record(name, sc, cnts, float(np.mean(s_kern)), world, any(m.UV is not None for m in sc.meshes), h.info),
                   "counts": cnt_note}
            if world == 1 and not args.no_cpu_baseline:
                cap = {"bunny": 1, "cbox-mis": 16, "ajax-rough": 8, "random10m-ao": 1, "random10m-normals": 1}.get(name, 4)
                rec["cpu_baseline"], rec["parity"] = cpu_check(name, sc, h, cap)
            rec["wall_s"] = time.time() - t_sub
            subs[f"configs[{WORKLOADS[name][0]}]"] = rec
        h.barrier()

    if rank == 0:
        if subs:
            line["configs"] = subs
            line["gpu_launches"] += int(sum(r["steps"] * (world + 1) for r in subs.values()))
        print(json.dumps(line), flush=True)
    if world > 1:
        h.dist.barrier()
    h.close()
    if world > 1:
        h.dist.destroy_process_group()


if __name__ == "__main__":
    main()
