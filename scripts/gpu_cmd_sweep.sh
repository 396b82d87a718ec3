python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/gpu_dev.py c1,c2,c3 | grep -v "scene build" | tee gpurun_out/sweep_brute.log
RRS_NO_BRUTE=1 python scripts/gpu_dev.py c2 | grep -v "scene build" | sed "s/^/nobrute /" | tee -a gpurun_out/sweep_brute.log
