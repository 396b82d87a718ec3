python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo pytest_exit=$?; tail -5 gpurun_out/pytest_gpu5.log
python scripts/gpu_dev.py c1,c2,c3,c4 > gpurun_out/dev6.log 2>&1; cat gpurun_out/dev6.log
python scripts/gpu_dev.py c5 0 16 >> gpurun_out/dev6.log 2>&1; tail -2 gpurun_out/dev6.log
for r in 1 2 8 16; do RRS_REFILL_LANES=$r python scripts/gpu_dev.py c2,c4 0 64 2>&1 | grep -v "scene build" | sed "s/^/refill=$r /"; done | tee -a gpurun_out/dev6.log
