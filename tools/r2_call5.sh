#!/bin/bash
# Round-2 GPU call 5: latency / instruction-diet variants of the binary walk (sentinel stack, L1 prefetch of the pushed child
# and of the parked leaf, SAH bin count), each against a default build of the same sources on the same box.
set -x
nvidia-smi -L; nproc
bash tools/ab_variants.sh "_base _sent _pf1 _pf2 _pf3 _base _sent _pf1 _pf2 _pf3" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "_base _sent _pf1 _pf2 _pf3" "random10m-ao" "--spp 4"
bash tools/ab_variants.sh "_base _pf3" "ajax-rough" "--spp 128"
for w in ajax-ao cbox-mis; do
  for o in "sah_bins=16" "sah_bins=32" "sah_bins=8"; do
    NORI_B200_LIB=nori_b200/lib/libnori_b200_base.so timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1), d['roofline']['node_visits'], d['roofline']['tri_tests'])" || tail -3 gpurun_out/ab_err.log
  done
done
for v in _pf3 _sent; do
  (time NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wavefront.py -q -x 2>&1 | tail -4)
done
