TAG=$1
python scripts/ncu_one.py c2 64 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_wavefront' -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_c2 python scripts/ncu_one.py c2 64 > gpurun_out/ncu_c2.log 2>&1
echo c2_exit=$?; cat gpurun_out/plain_c2.log
