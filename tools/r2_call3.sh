#!/bin/bash
# Round-2 GPU call 3: the 8-wide compressed hierarchy (NB_WIDE) against the binary walk -- parity first, then A/B.
set -x
nvidia-smi -L; nproc
(time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6)
for v in _wide _wide_p16; do
  (time NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_wavefront.py tests/test_gpu_entry_points.py -q -k "not smem_nodes" 2>&1 | tail -12)
done
bash tools/ab_variants.sh "default _wide _wide_p8 _wide_p16 default _wide" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "default _wide _wide_p16" "random10m-ao" "--spp 4"
bash tools/ab_variants.sh "default _wide _wide_p16" "ajax-rough" "--spp 128"
for o in "engine=2 --opt occ_tail=8 --opt wf_pool=4194304"; do
  for v in "" _wide; do
    for w in ajax-ao cbox-mis; do
      NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','[$v]','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
    done
  done
done
NCU="ncu --clock-control none --profile-from-start off"
NORI_B200_LIB=nori_b200/lib/libnori_b200_wide.so $NCU --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_wide_ajax-ao python tools/probe.py ajax-ao > gpurun_out/ncu_r2_wide_ajax-ao.log 2>&1; tail -2 gpurun_out/ncu_r2_wide_ajax-ao.log
NORI_B200_LIB=nori_b200/lib/libnori_b200_wide.so $NCU --set full --import-source on -k regex:render_kernel -c 1 -f -o gpurun_out/prof_r2_wide_cbox-mis python tools/probe.py cbox-mis --spp 64 > gpurun_out/ncu_r2_wide_cbox-mis.log 2>&1; tail -2 gpurun_out/ncu_r2_wide_cbox-mis.log
