python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/gpu_dev.py c1,c2,c3,c4 | grep -v "scene build" | tee gpurun_out/sweep_shade.log
