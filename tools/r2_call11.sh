#!/bin/bash
# Round-2 GPU call 11: work-unit claim one fetch ahead (NB_UNIT_AHEAD=1, "_ua") and the 32-bit unit decode alone ("_u32") against
# the library of call 10 ("_r10"): parity first, then the frame times at N = 1 and every rank's share of the 8-GPU frame.
set -x
nvidia-smi -L
for v in _ua _u32; do
  export NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so
  (time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_entry_points.py -q -x 2>&1 | tail -25)
done
unset NORI_B200_LIB
bash tools/ab_variants.sh "_r10 _u32 _ua _r10 _u32 _ua" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "_r10 _u32 _ua" "ajax-rough" "--spp 128"
bash tools/ab_variants.sh "_r10 _u32 _ua" "random10m-ao" "--spp 4"
bash tools/ab_variants.sh "_r10 _u32 _ua" "bunny"
for v in _r10 _u32 _ua; do
  export NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so
  echo "== $v"
  python tools/shard_probe.py ajax-ao 1 "guided=75" "guided=0" "chunk=1" "chunk=2" "chunk=8"
  python tools/shard_probe.py ajax-ao 8 "guided=75" "guided=0" "chunk=1" "chunk=2" "guided=75,coarse=4"
  python tools/shard_probe.py cbox-mis 8 "guided=75" "guided=0"
done
export NORI_B200_LIB=nori_b200/lib/libnori_b200_ua.so
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5)
