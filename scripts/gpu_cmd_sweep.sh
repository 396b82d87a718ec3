python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python scripts/gpu_dev.py c4 | grep -v "scene build" | tee gpurun_out/sweep_bvh4.log
python scripts/gpu_dev.py c5 0 16 | grep -v "scene build" | tee -a gpurun_out/sweep_bvh4.log
for ns in 1 2 4 8; do RRS_NODE_STEPS=$ns python scripts/gpu_dev.py c4 0 64 | grep -v "scene build" | sed "s/^/steps=$ns /"; done | tee -a gpurun_out/sweep_bvh4.log
