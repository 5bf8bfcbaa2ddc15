#!/bin/bash
# Round-2 GPU call 7: why the sharded frame loses 18 % at N = 8 -- work-unit size.  One GPU renders each rank's share
# (tools/shard_probe.py) with the L2 flushed, for unit sizes and for guided scheduling (coarse units first, fine units last).
set -x
nvidia-smi -L
python tools/shard_probe.py ajax-ao 1 "chunk=0" "chunk=1" "chunk=2" "chunk=4" "chunk=8" "chunk=16" "guided=50" "guided=75" "guided=88"
python tools/shard_probe.py ajax-ao 8 "guided=0" "chunk=1" "chunk=2" "chunk=4" "chunk=8" "guided=50" "guided=75" "guided=88" "guided=100" "prefetch=1" "prefetch=1,guided=75"
python tools/shard_probe.py ajax-ao 4 "guided=0" "guided=50" "guided=75" "guided=88"
python tools/shard_probe.py ajax-ao 2 "guided=0" "guided=75" "guided=88"
python tools/shard_probe.py cbox-mis 8 "guided=0" "guided=75" "guided=88"
(time NB_UNUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3)
