#!/bin/bash
# Round-2 GPU call 9: DYNAMIC guided self-scheduling (a warp claims k = clamp(units left / (warps * G), 1, coarse) consecutive
# sample units of one patch per atomic): every rank's share of the 8-GPU frame on one GPU, the 1-GPU frame, all workloads.
set -x
nvidia-smi -L
export NORI_B200_LIB=nori_b200/lib/libnori_b200_gd.so
(time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -3)
python tools/shard_probe.py ajax-ao 1 "guided=0" "guided=1" "guided=2" "guided=4" "guided=8" "guided=2,coarse=16" "guided=2,coarse=4" "guided=4,coarse=16"
python tools/shard_probe.py ajax-ao 8 "guided=0" "guided=1" "guided=2" "guided=4" "guided=8" "guided=2,coarse=16" "guided=2,coarse=4" "guided=4,coarse=4"
python tools/shard_probe.py ajax-ao 4 "guided=0" "guided=2" "guided=4"
python tools/shard_probe.py ajax-ao 2 "guided=0" "guided=2" "guided=4"
python tools/shard_probe.py cbox-mis 8 "guided=0" "guided=2" "guided=4"
for w in "cbox-mis" "ajax-rough --spp 128" "random10m-ao --spp 4" "bunny"; do
  for o in "guided=0" "guided=2" "guided=4"; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GUIDED','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
  done
done
