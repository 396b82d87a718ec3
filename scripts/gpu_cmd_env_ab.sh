# A/B of an environment toggle on the same box: bash scripts/gpu_cmd_env_ab.sh <configs> <spp> <VAR>
CFG=$1; SPP=$2; VAR=$3
for round in 1 2; do
python scripts/gpu_dev.py $CFG 0 $SPP 2>&1 | grep -v "scene build" | sed "s/^/[on ] /"
env $VAR=1 python scripts/gpu_dev.py $CFG 0 $SPP 2>&1 | grep -v "scene build" | sed "s/^/[off] /"
done | tee gpurun_out/env_ab.log
