"""Reads .ncu-rep captures (no GPU needed) and writes the per-workload summary bench.py quotes (profiles/ncu_traffic.json):
DRAM bytes per launch plus the counters that name the kernel's real bound.

  python tools/ncu_to_json.py ajax-ao=gpurun_out/prof_r2_ajax-ao.ncu-rep cbox-mis=prof.ncu-rep@64 [--note "build / config note"]
(@N: the capture rendered N samples per pixel instead of the workload's full count)
Each capture may hold several kernels (the wavefront engine: logic + trace); bytes and time are summed over the launches of
ONE frame's worth named with --launches (default: the single longest launch), the ratios are those of the longest launch."""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def fnum(r, hdr, key, default=None):
    if key not in hdr:
        return default
    v = r[hdr.index(key)].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return default


def to_bytes(r, hdr, units, key):
    v = fnum(r, hdr, key, 0.0)
    u = units[hdr.index(key)].lower() if key in hdr else "byte"
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    return v * mult


def to_ms(r, hdr, units, key="gpu__time_duration.sum"):
    v = fnum(r, hdr, key, 0.0)
    u = units[hdr.index(key)].lower()
    return v * {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)


def summarise(path):
    hdr, units, rows = rows_of(path)
    best = max(rows, key=lambda r: to_ms(r, hdr, units))
    stalls = [(float(best[i]), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and best[i]]
    tot = sum(s for s, _ in stalls) or 1.0
    st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: 100.0 * s / tot for s, h in stalls}
    rec = {
        "kernel": best[hdr.index("Kernel Name")][:100],
        "bytes": int(to_bytes(best, hdr, units, "dram__bytes_read.sum") + to_bytes(best, hdr, units, "dram__bytes_write.sum")),
        "dram_read_bytes": int(to_bytes(best, hdr, units, "dram__bytes_read.sum")),
        "dram_write_bytes": int(to_bytes(best, hdr, units, "dram__bytes_write.sum")),
        "captured_ms": to_ms(best, hdr, units),
        "issue_active_pct": fnum(best, hdr, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": fnum(best, hdr, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "lsu_wavefronts_pct_of_peak": fnum(best, hdr, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": fnum(best, hdr, "l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": fnum(best, hdr, "lts__t_sector_hit_rate.pct"),
        "warp_inst": fnum(best, hdr, "smsp__inst_executed.sum"),
        "registers": fnum(best, hdr, "launch__registers_per_thread"),
        "no_instruction_stall_pct": st.get("no_instructions"),
        "long_scoreboard_stall_pct": st.get("long_scoreboard"),
        "capture": os.path.relpath(path, REPO),
        "launches_in_capture": len(rows),
    }
    if len(rows) > 1:
        rec["all_launches"] = [{"kernel": r[hdr.index("Kernel Name")][:60], "ms": to_ms(r, hdr, units),
                                "dram_bytes": int(to_bytes(r, hdr, units, "dram__bytes_read.sum") + to_bytes(r, hdr, units, "dram__bytes_write.sum")),
                                "lanes_per_inst": fnum(r, hdr, "smsp__thread_inst_executed_per_inst_executed.ratio")} for r in rows[:64]]
    return rec


def main():
    note = None
    args = sys.argv[1:]
    if "--note" in args:
        i = args.index("--note"); note = args[i + 1]; del args[i:i + 2]
    out_path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(out_path))
    except Exception:
        data = {}
    data["_doc"] = ("per workload: dram__bytes_read.sum + dram__bytes_write.sum ('bytes') of ONE launch of the dominant kernel and the counters "
                    "that name its real bound, from ncu --set full captures (tools/ncu_to_json.py); bench.py copies them into roofline.traffic / roofline.secondary")
    for a in args:
        name, path = a.split("=", 1)
        cap_spp = None
        if "@" in path:
            path, cap_spp = path.rsplit("@", 1); cap_spp = int(cap_spp)
        rec = summarise(path)
        if cap_spp:
            rec["capture_spp"] = cap_spp      # the capture rendered this many samples per pixel; bench.py scales bytes / time to the frame's
        if note:
            rec["note"] = note
        data[name] = rec
        print(name, json.dumps(rec)[:400])
    json.dump(data, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
