python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tune in 0 32 48 64; do
RRS_TUNE=$tune python scripts/gpu_dev.py c4 0 64 2>&1 | grep -v "scene build" | sed "s/^/tune=$tune /"
done | tee gpurun_out/sweep_tune2.log
for tune in 0 48; do
RRS_TUNE=$tune python scripts/gpu_dev.py c2 0 64 2>&1 | grep -v "scene build" | sed "s/^/tune=$tune /"
RRS_TUNE=$tune python scripts/gpu_dev.py c5 0 8 2>&1 | grep -v "scene build" | sed "s/^/tune=$tune /"
done | tee -a gpurun_out/sweep_tune2.log
python scripts/gpu_dev.py c1,c3 | grep -v "scene build" | tee -a gpurun_out/sweep_tune2.log
