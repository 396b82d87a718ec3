# usage: gpu_cmd_ncu2.sh <tag> <config> <spp> <flagsA> <flagsB>  -- two ncu --set full captures of k_wavefront (render flags A and B)
TAG=$1; CFG=$2; SPP=$3; FA=$4; FB=$5
python scripts/ncu_one.py $CFG $SPP $FA > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
for F in $FA $FB; do
ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_${CFG}_f$F python scripts/ncu_one.py $CFG $SPP $F > gpurun_out/ncu_${TAG}_f$F.log 2>&1
echo ncu_exit_f$F=$?
done
cat gpurun_out/plain_${TAG}.log
