# The same workloads on builds of older commits and of HEAD, on one box (trees under _variants/trees/t_<commit>, each
# built there with python -m rayrs_b200.build):  bash scripts/gpu_cmd_bisect.sh <configs> <spp> <commit> [<commit> ...]
CFG=$1; SPP=$2; shift 2
for round in 1 2; do
for c in "$@" HEAD; do
if [ $c = HEAD ]; then dir=.; else dir=_variants/trees/t_$c; fi
(cd $dir && timeout 300 python scripts/gpu_dev.py $CFG 0 $SPP 0 2>&1 | grep -v "^ *$" | sed "s/^/[$c] /")
done
done | tee gpurun_out/bisect.log
