# usage: gpu_cmd_ncu_cfg.sh <tag> <config> <spp> [skip]  -- one ncu --set full capture of k_wavefront on one configuration
TAG=$1; CFG=$2; SPP=$3; SKIP=${4:-1}
python scripts/ncu_one.py $CFG $SPP > gpurun_out/plain_$CFG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG}_$CFG python scripts/ncu_one.py $CFG $SPP > gpurun_out/ncu_$CFG.log 2>&1
echo ${CFG}_exit=$?; cat gpurun_out/plain_$CFG.log
