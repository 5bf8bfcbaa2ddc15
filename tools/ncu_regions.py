"""Per-region instruction / lane-utilisation breakdown from the SASS page of an .ncu-rep (blocks of N instructions)."""
import csv, io, subprocess, sys
path = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    try:
        recs.append((r[ix['Source']], float(r[ix['Instructions Executed']]), float(r[ix['Thread Instructions Executed']]), float(r[ix['# Samples']] or 0)))
    except Exception:
        pass
ti = sum(x[1] for x in recs); tt = sum(x[2] for x in recs); ts = sum(x[3] for x in recs)
print('total warp-inst %.4g lane-inst %.4g avg lanes %.2f  n_sass %d' % (ti, tt, tt / ti, len(recs)))
for b in range(0, len(recs), blk):
    seg = recs[b:b + blk]
    ie = sum(x[1] for x in seg); te = sum(x[2] for x in seg); sm = sum(x[3] for x in seg)
    if ie / ti < 0.003: continue
    ops = {}
    for x in seg:
        t = x[0].split()
        op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print('%5d  inst %5.1f%%  lanes %5.1f  lane-inst %5.1f%%  samples %5.1f%%  %s' % (b, 100 * ie / ti, te / max(ie, 1), 100 * te / tt, 100 * sm / ts, top))
