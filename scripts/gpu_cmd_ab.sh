# A/B of library builds on the same box: bash scripts/gpu_cmd_ab.sh <tag> <configs> <spp> <variant> [<variant> ...]
# (variants are _variants/lib_<name>.so; two interleaved rounds; the log keeps its own name under gpurun_out/)
TAG=$1; CFG=$2; SPP=$3; shift 3
cp rayrs_b200/librayrs_b200.so /tmp/keep.so
for round in 1 2; do
for v in "$@"; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
timeout 120 python scripts/gpu_dev.py $CFG 0 $SPP 0 2>&1 | grep -v "scene build" | sed "s/^/[$v] /"
done
done | tee gpurun_out/ab_$TAG.log
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
