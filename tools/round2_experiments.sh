#!/bin/bash
# Round-2 opening measurement (one gpurun call).  Build the variants HERE first:  bash tools/round2_experiments.sh --build
# Variants (default-off switches in nori_b200/csrc/nb_kernels.cuh):
#   _l256    : NB_LDG256=1        -- 64 B nodes fetched with two 256-bit loads (LDG.E.ENL2.256) instead of four 128-bit ones
#   _compact : NB_COMPACT_PATH=1  -- tile rectangle re-derived at splat time (fewer spills)
#   _tail    : NB_TAIL_CUT=1      -- resumable walks; run time option "tail"
#   _p10/_p9 : register cap of the path tracers (default 11 CTAs/SM = 40 registers)
#   _s32     : NB_STACK=32        -- per-lane traversal stack of 32 entries (half the local-memory footprint; host SAH trees only)
if [ "$1" = "--build" ]; then
python - <<'PY'
from concurrent.futures import ThreadPoolExecutor
from nori_b200 import build
V = [("", ()), ("_l256", ("-DNB_LDG256=1",)), ("_compact", ("-DNB_COMPACT_PATH=1",)), ("_tail", ("-DNB_TAIL_CUT=1",)),
     ("_p10", ("-DNB_MIN_BLOCKS_PATH=10",)), ("_p9", ("-DNB_MIN_BLOCKS_PATH=9",)), ("_s32", ("-DNB_STACK=32",)),
     ("_l256s32", ("-DNB_LDG256=1", "-DNB_STACK=32"))]
with ThreadPoolExecutor(2) as ex:
    print(list(ex.map(lambda v: build.build_cuda(force=True, variant=v[0], extra_flags=v[1]), V)))
print(build.build_host(force=True))
PY
exit $?
fi
set -x
nvidia-smi -L; nproc
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25)
NB_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_zzz_gpu_deferred_engine.py -q 2>&1 | tail -15
bash tools/ab_variants.sh "default _l256 _compact default _l256 _compact" "ajax-ao cbox-mis"
bash tools/ab_variants.sh "default _l256" "random10m-ao" "--spp 4"
bash tools/ab_variants.sh "default _l256 _p10 _p9" "ajax-rough" "--spp 128"
bash tools/ab_variants.sh "_p10 _p9 _s32 _l256s32" "cbox-mis"
bash tools/ab_variants.sh "_s32 _l256s32" "ajax-ao"
# a walk that never ends must not take the box with it (device watchdog + timeout)
NORI_B200_LIB=nori_b200/lib/libnori_b200_tail.so TAILS="0 2 4 8 12" timeout 120 python tools/tail_sweep.py ajax-ao cbox-mis
# deferred-occlusion engine (nb_wavefront.cu): A/B against the fused kernel, refill threshold sweep
for w in ajax-ao cbox-mis; do
  for o in "engine=0" "engine=1" "engine=1 --opt occ_tail=12" "engine=1 --opt occ_tail=24" "engine=1 --opt occ_tail=28" \
           "engine=2" "engine=2 --opt wf_pool=1048576" "engine=2 --opt wf_pool=4194304" "engine=2 --opt wf_pool=8388608 --opt wf_check=8" "engine=2 --opt occ_tail=12" "engine=2 --opt occ_tail=26"; do
    timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','$w','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
  done
done
for o in "engine=0" "engine=1" "engine=2" "engine=2 --opt wf_pool=8388608"; do
  timeout 300 python bench.py --workload ajax-rough --spp 128 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --opt $o 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ENG','ajax-rough128','[$o]',round(d['ms_per_step'],3),round(d['value'],1))" || tail -3 gpurun_out/ab_err.log
done
# the whole default bench line (all five BASELINE configs) + the CPU arm
(time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err); tail -c 600 gpurun_out/bench_r2_a.err; head -c 3000 gpurun_out/bench_r2_a.json
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_a_ref.json 2>&1); cat gpurun_out/bench_r2_a_ref.json
