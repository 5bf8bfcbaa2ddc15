#!/bin/bash
# Round-2 GPU call 15 (gpurun --gpus 8): the headline bench line of the closing tree at N = 8 and N = 4 on one 8-GPU box.
set -x
nvidia-smi -L | head -8
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
(time timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r2_final_n$N.json 2> gpurun_out/bench_r2_final_n$N.err); tail -2 gpurun_out/bench_r2_final_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_final_n$N.json").read().strip().splitlines()[-1])
print("N=$N headline", round(d["ms_per_step"],3), "ms", round(d["value"],1), "Mrays/s  kernel", d["roofline"].get("kernel_ms"), " e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "ms")
PY
done
