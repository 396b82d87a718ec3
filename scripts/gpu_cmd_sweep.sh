python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python scripts/gpu_dev.py c1,c2,c3 | grep -v "scene build" | tee gpurun_out/sweep_grouped.log
