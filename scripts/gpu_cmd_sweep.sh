python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python scripts/gpu_dev.py c1,c2 | grep -v "scene build" | tee gpurun_out/sweep_twopath.log
python scripts/gpu_dev.py c3 0 0 16 | grep -v "scene build" | sed "s/^/force-pathloop /" | tee -a gpurun_out/sweep_twopath.log
