#!/bin/bash
# Round-2 GPU call 14 (gpurun --gpus 2): the closing tree on two devices -- N devices behind the C-ABI (nb_create_multi,
# nori --gpus), the process-per-GPU group under torchrun with the Latin-pattern tile table, and the headline bench line at N = 2.
set -x
N=2
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(time timeout 300 $TR tools/check_multigpu.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -8)
(time timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-configs > gpurun_out/bench_r2_final_n$N.json 2> gpurun_out/bench_r2_final_n$N.err); tail -3 gpurun_out/bench_r2_final_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_final_n$N.json").read().strip().splitlines()[-1])
print("N=$N headline", round(d["ms_per_step"],3), "ms", round(d["value"],1), "Mrays/s  kernel", d["roofline"].get("kernel_ms"), " e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "ms")
PY
