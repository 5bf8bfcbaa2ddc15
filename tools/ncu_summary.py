"""Reads an .ncu-rep (no GPU needed) and prints the handful of metrics the profile summaries quote."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')][:80])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:75s} {r[i]:>16s} {units[i]}')
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued') and r[i]]
        tot = sum(s for s, _ in stalls) or 1
        for s, h in sorted(stalls, reverse=True)[:8]:
            print(f'  stall {h[len("smsp__pcsamp_warps_issue_stalled_"):]:40s} {100 * s / tot:5.1f} %')


if __name__ == '__main__':
    main(sys.argv[1])
