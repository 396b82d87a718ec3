# round 2, first call: GPU suite, then L2-window / queue-hint / queue-size A/B on configs 4 and 5
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_r02a.log
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
cp rayrs_b200/librayrs_b200.so /tmp/keep.so
for round in 1 2; do
for v in cs nocs; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
python scripts/gpu_dev.py c4,c5 0,1048576 32 0,32 2>&1 | grep -v "scene build" | sed "s/^/[$v] /"
done
done | tee gpurun_out/ab_r02a_l2.log
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
python scripts/gpu_dev.py c4,c5 0 16 1 2>&1 | tee gpurun_out/counters_r02a.log
python scripts/gpu_dev.py c4 262144,524288,2097152,8388608 32 0 2>&1 | tee gpurun_out/queue_r02a.log
