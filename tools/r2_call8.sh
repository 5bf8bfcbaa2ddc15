#!/bin/bash
# Round-2 GPU call 8: guided scheduling (coarse work units first, fine units last) at N = 1 on every workload, coarse unit 4 / 8.
set -x
nvidia-smi -L
for w in "ajax-ao" "cbox-mis" "ajax-rough --spp 128" "random10m-ao --spp 4" "bunny"; do
  for o in "guided=0" "guided=75" "guided=75 --opt coarse=4" "guided=62" "guided=88 --opt coarse=4" "chunk=4"; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GUIDED','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
  done
done
python tools/shard_probe.py ajax-ao 2 "guided=0" "guided=75,coarse=4" "guided=62" 
python tools/shard_probe.py ajax-rough 8 "guided=0" "guided=75" "guided=75,coarse=4"
