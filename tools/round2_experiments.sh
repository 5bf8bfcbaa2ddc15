#!/bin/bash
# Round-2 opening measurement (one gpurun call, ~4 GPU-minutes).  Build the variants HERE first:
#   bash tools/round2_experiments.sh --build      (cross-compiles; no GPU needed)
# then on the box:   gpurun --timeout 420 -- 'bash tools/round2_experiments.sh > gpurun_out/r2_exp.log 2>&1'
# Variants (all default-off switches in nori_b200/csrc/nb_kernels.cuh):
#   _compact : NB_COMPACT_PATH=1  -- tile rectangle re-derived at splat time (static spill bytes 650 -> 606 path, 222 -> 182 ao)
#   _tail    : NB_TAIL_CUT=1      -- resumable walks; run time option "tail" (tools/tail_sweep.py sweeps it)
#   _both    : both
#   _p10/_p9 : register cap of the path tracers re-swept now that waves are phased (default 11 CTAs/SM = 40 registers,
#              as is 12; 10 -> 48, 9 -> 56, 8 -> 64 which measured slower)
if [ "$1" = "--build" ]; then
python - <<'PY'
from nori_b200 import build
build.build_cuda()
build.build_cuda(force=True, variant="_compact", extra_flags=("-DNB_COMPACT_PATH=1",))
build.build_cuda(force=True, variant="_tail", extra_flags=("-DNB_TAIL_CUT=1",))
build.build_cuda(force=True, variant="_both", extra_flags=("-DNB_TAIL_CUT=1", "-DNB_COMPACT_PATH=1"))
build.build_cuda(force=True, variant="_p10", extra_flags=("-DNB_MIN_BLOCKS_PATH=10",))
build.build_cuda(force=True, variant="_p9", extra_flags=("-DNB_MIN_BLOCKS_PATH=9",))
PY
exit $?
fi
set -x
# first of all: the entry points that have never run on hardware (simple integrator, nb_li_samples, ttest object)
timeout 600 python -m pytest tests/test_zz_gpu_late_entry_points.py -q -rxX 2>&1 | tail -15
NB_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_zzz_gpu_deferred_engine.py -q -x 2>&1 | tail -15
bash tools/ab_variants.sh "default _compact default _compact" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "default _p10 _p9" "cbox-mis ajax-rough"
# a walk that never ends must not take the box with it: the earlier tail-cut build livelocked
for lib in _tail _both; do
  NORI_B200_LIB=nori_b200/lib/libnori_b200$lib.so TAILS="0 2 4 8 12" timeout 90 python tools/tail_sweep.py ajax-ao cbox-mis
done
# parity of the winning variants against the oracle
for lib in _compact _tail; do
  NORI_B200_LIB=nori_b200/lib/libnori_b200$lib.so timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "film_parity" 2>&1 | tail -2
done

# deferred-occlusion engine (nb_wavefront.cu): A/B against the fused kernel, refill threshold sweep
for w in ajax-ao cbox-mis ajax-rough; do
  for o in "engine=0" "engine=1" "engine=1 --opt occ_tail=12" "engine=1 --opt occ_tail=24" "engine=1 --opt occ_tail=28"; do
    python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))"
  done
done
