"""Builds profiles/r1_report.md -- the per-config table SURVEY.md 8(d) asks for -- from the committed bench lines."""
import json
import os

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def load(name):
    p = os.path.join(P, name)
    return json.load(open(p)) if os.path.exists(p) else None


rows = [("configs[1] Ajax(stand-in) ao 800x600x64", "r1_final_bench.json", "configs[1]"),
        ("configs[2] Cornell box path_mis 512x512x256", "r1_final_bench_cbox-mis.json", "configs[2]"),
        ("configs[3] Ajax(stand-in) microfacet path_mis 768x768x1024", "r1_final_bench_ajax-rough.json", "configs[3]"),
        ("configs[4] 10 M random triangles ao 1920x1080x4", "r1_v4_bench_random10m-ao.json", "configs[4]"),
        ("configs[4] 10 M random triangles normals 1920x1080x4", "r1_v4_bench_random10m-normals.json", None)]
parity = {}
pr = os.path.join(P, "r1_parity_report.jsonl")
if os.path.exists(pr):
    for line in open(pr):
        d = json.loads(line)
        parity[d["config"].split(" ")[0]] = d
traffic = load("ncu_traffic.json") or {}
out = ["# Round 1 report -- BASELINE.json configs on one B200 (per SURVEY.md 8d)", "",
       "GPU numbers: CUDA events on the launch stream, L2 flushed between steps, inputs resident; algorithmic bytes = 64 B x node visits + 48 B x triangle tests",
       "+ 36 B (+24 B UV) x shaded hits + one film write, counted by the instrumented instantiation of the same kernel.", "",
       "| config | ms / frame | Mrays/s | Msamples/s | alg. GB/s | / 6575.8 (measured) | / 8000 (nominal) | ncu DRAM GB/s | e2e Mrays/s | rel-L2 vs oracle |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for title, fn, pk in rows:
    d = load(fn)
    if not d:
        continue
    r = d["roofline"]
    wl = {"r1_final_bench.json": "ajax-ao", "r1_v4_bench_random10m-ao.json": "random10m-ao"}.get(fn)
    dram = "n/a"
    if wl and wl in traffic:
        dram = "%.0f" % (traffic[wl]["bytes"] / (r["kernel_ms"] * 1e-3) / 1e9)
    rel = parity.get(pk, {}).get("rel_l2_film") if pk else None
    out.append("| %s | %.2f | %.0f | %.0f | %.0f | %.3f | %.3f | %s | %.0f | %s |" % (
        title, d["ms_per_step"], d["value"], d["msamples_per_sec"], r["achieved"], r["achieved"] / 6575.8, r["achieved"] / 8000.0, dram,
        d["e2e"]["value"], ("%.1e" % rel) if rel is not None else "n/a"))
b0 = parity.get("configs[0]")
if b0:
    out.append("| configs[0] bunny normals 768x768x1, reference per-block seeding (plumbing) | %.2f | %.1f | %.1f | - | - | - | - | - | %.1e |" % (
        b0["gpu_kernel_ms"], b0["gpu_rays"] / b0["gpu_kernel_ms"] / 1e3, b0["samples"] / b0["gpu_kernel_ms"] / 1e3, b0["rel_l2_film"]))
out += ["", "rel-L2 for configs[2..4] was measured at reduced spp / 1 M triangles so that the CPU oracle finishes in seconds (`r1_parity_report.jsonl`);",
        "configs[1] and configs[0] at full size.  Ray counts of GPU and oracle are identical in every case.", "",
        "## CPU arm on the same box (128 host cores; oracle port of the Nori tile loop, `bench.py --impl reference`)", "",
        "| workload | accel | Mrays/s | Msamples/s | s / step |", "|---|---|---|---|---|"]
for title, fn, acc in [("configs[1] ajax-ao at 32 of 64 spp per step", "r1_final_bench_ref.json", "CPU binned-SAH BVH"),
                       ("configs[0] bunny 768x768x1", "r1_v6_bench_ref_brute_bunny.json", "brute force (the reference's shipped Accel)")]:
    d = load(fn)
    if d:
        out.append("| %s | %s | %.2f | %.2f | %.3f |" % (title, acc, d["value"], d["msamples_per_sec"], d["ms_per_step"] / 1e3))
out += ["", "## Scaling (configs[1], tiles % N, one NCCL gather of finished blocks per frame)", "",
        "| N | Mrays/s | ms / frame | kernel ms (max over ranks) | e2e Mrays/s | efficiency vs N=1 |", "|---|---|---|---|---|---|"]
base = None
for n in (1, 2, 4, 8):
    d = load("r1_v6_scale2_n%d.json" % n) or load("r1_v6_scale_n%d.json" % n)
    if not d:
        continue
    base = base or d["value"]
    out.append("| %d | %.0f | %.3f | %.3f | %.0f | %.2f |" % (n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["value"] / (base * n)))
open(os.path.join(P, "r1_report.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
