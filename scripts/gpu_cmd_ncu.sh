TAG=$1
python scripts/ncu_one.py c4 32 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_wavefront' -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_c4 python scripts/ncu_one.py c4 32 > gpurun_out/ncu_c4.log 2>&1
echo c4_exit=$?; cat gpurun_out/plain_c4.log
