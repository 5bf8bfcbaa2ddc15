#!/bin/bash
# Round-2 GPU call 10: the shipped default (static guided schedule, Latin-pattern tile table) -- full GPU suite, smoke, the bench
# line with every config, and every rank's share of the 2/4/8-GPU frames on one GPU against the plain schedule.
set -x
nvidia-smi -L
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5)
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_call10_bench.json 2> gpurun_out/r2_call10_bench.err; tail -c 3000 gpurun_out/r2_call10_bench.json
python tools/shard_probe.py ajax-ao 1 "guided=0" "guided=75"
python tools/shard_probe.py ajax-ao 2 "guided=0" "guided=75"
python tools/shard_probe.py ajax-ao 4 "guided=0" "guided=75"
python tools/shard_probe.py ajax-ao 8 "guided=0" "guided=75" "guided=75,coarse=4" "guided=50"
python tools/shard_probe.py cbox-mis 8 "guided=0" "guided=75"
python tools/shard_probe.py ajax-rough 8 "guided=75"
python tools/shard_probe.py random10m-ao 8 "guided=75"
