# A/B harness for kernel build variants (run on the GPU box): usage: bash tools/ab_variants.sh "<variants>" "<workloads>"
# a variant "" is nori_b200/lib/libnori_b200.so, "_x" is libnori_b200_x.so (built with nori_b200.build.build_cuda(variant=..., extra_flags=...))
VARS=${1:-"default"}; WLS=${2:-"ajax-ao cbox-mis"}
for v in $VARS; do
  [ "$v" = default ] && v=""
  for w in $WLS; do
    NORI_B200_LIB=nori_b200/lib/libnori_b200$v.so python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RES','[$v]','$w',round(d['ms_per_step'],3),round(d['value'],1))"
  done
done
