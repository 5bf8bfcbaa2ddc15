"""Per source line: warp instructions, lanes, stall samples from an .ncu-rep captured with --import-source on (no GPU needed).
   python tools/ncu_lines.py prof.ncu-rep [top N]"""
import csv, io, subprocess, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; recs = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try:
        recs.append((cur, int(r[0]), r[1].strip()[:90], float(r[hdr['Instructions Executed']]), float(r[hdr['Thread Instructions Executed']]), float(r[hdr['# Samples']] or 0)))
    except Exception:
        pass
ti = sum(x[3] for x in recs); tt = sum(x[4] for x in recs); ts = sum(x[5] for x in recs)
print('total warp-inst %.4g  thread-inst %.4g  lanes %.2f  samples %d' % (ti, tt, tt / ti, ts))
for x in sorted(recs, key=lambda x: -x[3])[:topn]:
    print('%5.2f%% inst  lanes %5.1f  %5.2f%% smp  %s:%d  %s' % (100 * x[3] / ti, x[4] / max(x[3], 1), 100 * x[5] / max(ts, 1), x[0], x[1], x[2]))
# regions of nb_kernels.cuh by line range (function boundaries), plus nb_device.cuh as a whole
import re
src = open('nori_b200/csrc/nb_kernels.cuh').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:template <[^>]*>\s*)?(?:__device__|__global__)[^(]*?\b(\w+)\(', l)
    if m: marks.append((i, m.group(1)))
def region(line):
    name = 'top'
    for i, n in marks:
        if i <= line: name = n
        else: break
    return name
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for f, ln, s, ie, te, sm in recs:
    key = region(ln) if f == 'nb_kernels.cuh' else f
    a = agg[key]; a[0] += ie; a[1] += te; a[2] += sm
print('--- by function (innermost inlined frame)')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] / ti < 0.002: continue
    print('%-28s inst %5.1f%%  lanes %5.1f  thread-inst %5.1f%%  samples %5.1f%%' % (k, 100 * a[0] / ti, a[1] / max(a[0], 1), 100 * a[1] / tt, 100 * a[2] / max(ts, 1)))
