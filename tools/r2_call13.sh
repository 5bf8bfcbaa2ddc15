#!/bin/bash
# Round-2 GPU call 13: the committed tree (guided schedule, Latin-pattern tile table, 32-bit unit decode) on a fresh box -- GPU
# suite, smoke, launch list, ncu captures of the two headline kernels, the whole bench line and the CPU arm.
set -x
nvidia-smi -L; nproc
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6)
python -c "import __graft_entry__ as g; g.smoke()"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
NCUP="ncu --clock-control none --profile-from-start off"
$NCUP --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_final_ajax-ao python tools/probe.py ajax-ao > gpurun_out/ncu_r2_final_ajax-ao.log 2>&1; tail -1 gpurun_out/ncu_r2_final_ajax-ao.log
$NCUP --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_final_cbox-mis python tools/probe.py cbox-mis --spp 64 > gpurun_out/ncu_r2_final_cbox-mis.log 2>&1; tail -1 gpurun_out/ncu_r2_final_cbox-mis.log
(time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err); tail -c 300 gpurun_out/bench_r2_final.err; head -c 400 gpurun_out/bench_r2_final.json
(time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r2_final_ref.json 2>&1); head -c 600 gpurun_out/bench_r2_final_ref.json
