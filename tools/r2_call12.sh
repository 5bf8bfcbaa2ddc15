#!/bin/bash
# Round-2 GPU call 12: late claim-ahead (NB_UNIT_AHEAD=1: the next unit is claimed when the current one is down to one wave of
# items, "_ua2") against the 32-bit decode alone ("_u32" = the default source): parity, frame times, every rank's share at N = 8.
set -x
nvidia-smi -L
export NORI_B200_LIB=nori_b200/lib/libnori_b200_ua2.so
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25)
unset NORI_B200_LIB
bash tools/ab_variants.sh "_u32 _ua2 _u32 _ua2" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "_u32 _ua2" "ajax-rough" "--spp 128"
bash tools/ab_variants.sh "_u32 _ua2" "random10m-ao" "--spp 4"
bash tools/ab_variants.sh "_u32 _ua2" "bunny"
for v in _u32 _ua2; do
  export NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so
  echo "== $v"
  python tools/shard_probe.py ajax-ao 1 "guided=75" "chunk=1" "chunk=2" "chunk=8"
  python tools/shard_probe.py ajax-ao 2 "guided=75"
  python tools/shard_probe.py ajax-ao 4 "guided=75"
  python tools/shard_probe.py ajax-ao 8 "guided=75" "chunk=2"
  python tools/shard_probe.py cbox-mis 8 "guided=75"
done
