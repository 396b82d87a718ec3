python scripts/gpu_dev.py c4 131072,262144,524288,1048576,2097152,4194304,8388608 64 2>&1 | tee gpurun_out/sweep_q.log
python scripts/gpu_dev.py c2 131072,262144,524288,1048576,2097152,4194304 64 2>&1 | tee -a gpurun_out/sweep_q.log
python scripts/gpu_dev.py c5 262144,1048576,4194304 8 2>&1 | tee -a gpurun_out/sweep_q.log
