#!/bin/bash
# Round-2 closing call: the committed tree on a fresh box -- GPU suite, smoke, the launch list and the remaining ncu captures
# of the shipped build, the whole bench line and the CPU arm.
set -x
nvidia-smi -L; nproc
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6)
python -c "import __graft_entry__ as g; g.smoke()"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
NCUP="ncu --clock-control none --profile-from-start off"
$NCUP --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_ajax-rough python tools/probe.py ajax-rough --spp 32 > gpurun_out/ncu_r2_ajax-rough.log 2>&1; tail -1 gpurun_out/ncu_r2_ajax-rough.log
$NCUP --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_random10m-ao python tools/probe.py random10m-ao --spp 8 > gpurun_out/ncu_r2_random10m-ao.log 2>&1; tail -1 gpurun_out/ncu_r2_random10m-ao.log
$NCUP --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_final_ajax-ao python tools/probe.py ajax-ao > gpurun_out/ncu_r2_final_ajax-ao.log 2>&1; tail -1 gpurun_out/ncu_r2_final_ajax-ao.log
(time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err); tail -c 300 gpurun_out/bench_r2_final.err; head -c 400 gpurun_out/bench_r2_final.json
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_final_ref.json 2>&1); head -c 600 gpurun_out/bench_r2_final_ref.json
