#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): N devices behind the C-ABI (nb_create_multi, nori --gpus), the process-per-GPU group
# (nb_comm_init_rank / nb_render_gather) under torchrun, and the bench line at N.
set -x
N=${1:-2}
nvidia-smi -L; nproc
(time timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(time timeout 600 $TR tools/check_multigpu.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -8)
(time timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err); tail -5 gpurun_out/bench_r2_n$N.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_n$N.json").read().strip().splitlines()[-1])
print("N=$N headline", round(d["ms_per_step"],3), "ms", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "ms")
for k,v in d.get("configs",{}).items(): print(" ", k, round(v["ms_per_step"],3), "ms", round(v["value"],1), "Mrays/s")
PY
for o in "prefetch=0" "prefetch=1"; do
  timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-configs --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MULTI N=$N','[$o]',round(d['ms_per_step'],3),round(d['value'],1),'e2e',round(d['e2e']['ms_per_step'],3))"
done
