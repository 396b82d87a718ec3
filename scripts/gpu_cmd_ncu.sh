TAG=$1
python scripts/ncu_one.py c2 64 4 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade' -s 40 -c 4 -f -o gpurun_out/prof_${TAG}_c2split python scripts/ncu_one.py c2 64 4 > gpurun_out/ncu_c2.log 2>&1
echo c2_exit=$?; cat gpurun_out/plain_c2.log
python scripts/ncu_one.py c4 16 4 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend' -s 20 -c 2 -f -o gpurun_out/prof_${TAG}_c4split python scripts/ncu_one.py c4 16 4 > gpurun_out/ncu_c4.log 2>&1
echo c4_exit=$?; cat gpurun_out/plain_c4.log
