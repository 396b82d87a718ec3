# A/B of two builds on the same box: bash scripts/gpu_cmd_ab.sh <configs> <spp> <variant> [<variant> ...]
CFG=$1; SPP=$2; shift 2
cp rayrs_b200/librayrs_b200.so /tmp/keep.so
for round in 1 2; do
for v in "$@"; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
python scripts/gpu_dev.py $CFG 0 $SPP 2>&1 | grep -v "scene build" | sed "s/^/[$v] /"
done
done | tee gpurun_out/ab.log
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
