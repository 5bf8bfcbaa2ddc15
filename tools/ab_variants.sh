set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
for v in "" _nophase _p8; do
  for w in cbox-mis ajax-rough; do
    NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RES','$v','$w',round(d['ms_per_step'],2),round(d['value'],1))"
  done
done
NORI_B200_LIB=nori_b200/lib/libnori_b200.so python tools/wave_stats.py cbox-mis 2>&1 | tail -3
