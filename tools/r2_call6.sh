#!/bin/bash
# Round-2 GPU call 6: the path tracer's local-memory footprint.  DRAM spill traffic appears when the
# local memory of all resident threads (stack + spills, ~620 B each) exceeds L2; fewer resident CTAs or a smaller stack
# bring it under.  Time and DRAM bytes per launch for each setting.
set -x
nvidia-smi -L
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in "" _s32 _smst; do
  for o in "blocks_per_sm=0" "blocks_per_sm=10" "blocks_per_sm=9" "blocks_per_sm=8" "blocks_per_sm=6"; do
    for w in cbox-mis; do
      NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('FOOT','[$v]','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
    done
    NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so ncu --clock-control none --profile-from-start off --metrics $M -k regex:render_kernel -c 1 --csv --log-file gpurun_out/foot_tmp.csv python tools/probe.py cbox-mis --spp 64 --opt $o > /dev/null 2>&1
    python - <<PY
import csv,io
rows=[r for r in csv.reader(io.StringIO("\n".join(l for l in open("gpurun_out/foot_tmp.csv").read().splitlines() if l.startswith('"'))))]
h=rows[0]; mi=h.index("Metric Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
print("NCU [$v] [$o]", {r[mi].split("__")[-1][:28]: r[vi]+" "+r[ui] for r in rows[1:]})
PY
  done
done
for v in "" _s32; do
  for o in "blocks_per_sm=0" "blocks_per_sm=9" "blocks_per_sm=8"; do
    NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 300 python bench.py --workload ajax-rough --spp 128 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('FOOT','[$v]','ajax-rough128','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
  done
done
