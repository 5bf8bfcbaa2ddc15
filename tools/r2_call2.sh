#!/bin/bash
# Round-2 GPU call 2: ncu captures of the shipped fused kernel (source-level), launch lists of the wavefront engine, the
# remaining A/Bs (stack 32, warp-tile splat, L2 prefetch, wavefront refill threshold), the whole bench line.
set -x
nvidia-smi -L; nproc
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8)
NCU="ncu --clock-control none --profile-from-start off"
$NCU --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_ajax-ao python tools/probe.py ajax-ao > gpurun_out/ncu_r2_ajax-ao.log 2>&1; tail -2 gpurun_out/ncu_r2_ajax-ao.log
$NCU --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_cbox-mis python tools/probe.py cbox-mis --spp 64 > gpurun_out/ncu_r2_cbox-mis.log 2>&1; tail -2 gpurun_out/ncu_r2_cbox-mis.log
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
$NCU --metrics $M --csv --log-file gpurun_out/launches_r2_wave_ajax-ao.csv python tools/probe.py ajax-ao --opt engine=2 --opt occ_tail=12 > gpurun_out/ncu_wave_ao.log 2>&1; tail -1 gpurun_out/ncu_wave_ao.log
$NCU --metrics $M --csv --log-file gpurun_out/launches_r2_wave_cbox-mis.csv python tools/probe.py cbox-mis --spp 16 --opt engine=2 --opt occ_tail=12 > gpurun_out/ncu_wave_cbox.log 2>&1; tail -1 gpurun_out/ncu_wave_cbox.log
bash tools/ab_variants.sh "default _s32 _stile default _s32 _stile" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "default _stile" "random10m-normals" "--spp 8"
for w in ajax-ao cbox-mis; do
  for o in "prefetch=0" "prefetch=1" "engine=2 --opt occ_tail=0" "engine=2 --opt occ_tail=4" "engine=2 --opt occ_tail=8" "engine=2 --opt occ_tail=8 --opt wf_pool=4194304" "engine=2 --opt occ_tail=8 --opt wf_pool=524288"; do
    timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
  done
done
(time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err); tail -c 400 gpurun_out/bench_r2_b.err; head -c 600 gpurun_out/bench_r2_b.json
