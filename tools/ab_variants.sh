# A/B harness for kernel build variants (run on the GPU box): usage: bash tools/ab_variants.sh "<variants>" "<workloads>" [extra bench args]
# a variant "default" is nori_b200/lib/libnori_b200.so, "_x" is libnori_b200_x.so (built with nori_b200.build.build_cuda(variant=..., extra_flags=...))
VARS=${1:-"default"}; WLS=${2:-"ajax-ao cbox-mis"}; EXTRA=${3:-""}
for v in $VARS; do
  [ "$v" = default ] && v=""
  for w in $WLS; do
    NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs $EXTRA 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RES','[$v]','$w','$EXTRA',round(d['ms_per_step'],3),round(d['value'],1),'kern',round(d['roofline']['kernel_ms'],3))" || tail -3 gpurun_out/ab_err.log
  done
done
