#!/bin/bash
# Round-2 scaling call (gpurun --gpus 8): the bench line at N = 1, 2, 4, 8 on ONE box (torchrun uses the first N devices),
# the L2-prefetch A/B where the frame is shortest, nb_create_multi on 4 / 8 devices, `nori --gpus 8`.
set -x
nvidia-smi -L | head -8; nproc
run() {  # N, extra args, tag
  local N=$1; shift; local tag=$1; shift
  if [ $N -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_r2_$tag.json 2> gpurun_out/scale_r2_$tag.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/scale_r2_$tag.json 2> gpurun_out/scale_r2_$tag.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_r2_$tag.json").read().strip().splitlines()[-1])
    print("SCALE $tag N=$N", round(d["ms_per_step"],4), "ms", round(d["value"],1), "Mrays/s kern", round(d["roofline"]["kernel_ms"],4), " e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "ms")
    for k,v in d.get("configs",{}).items(): print("   ", k, round(v["ms_per_step"],3), "ms", round(v["value"],1), "Mrays/s")
except Exception as e:
    print("SCALE $tag N=$N FAILED", e); print(open("gpurun_out/scale_r2_$tag.err").read()[-600:])
PY
}
run 8 n8 --steps 20 --warmup 3 --configs cbox-mis,ajax-rough,random10m-ao
run 1 n1 --steps 20 --warmup 3 --no-configs --no-cpu-baseline
run 2 n2 --steps 20 --warmup 3 --no-configs
run 4 n4 --steps 20 --warmup 3 --no-configs
run 8 n8_prefetch --steps 20 --warmup 3 --no-configs --opt prefetch=1
run 8 n8_b --steps 20 --warmup 3 --no-configs
run 4 n4_rough --steps 5 --warmup 3 --workload ajax-rough --no-configs
(time timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4)
python scenes/make_scenes.py /tmp/scenes > /dev/null 2>&1; ls /tmp/scenes | head -3
for g in 1 8; do (time nori_b200/lib/nori /tmp/scenes/ajax-ao.xml --no-gui --gpus $g 2>&1 | grep -v "^\s\|Configuration\|^\]\|^$" | tail -2); done
