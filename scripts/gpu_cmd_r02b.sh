python -m pytest tests/test_gpu_intersect.py tests/test_gpu_render.py tests/test_golden.py -m gpu -x -q -k "not config5" > gpurun_out/pytest_r02b.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_r02b.log
bash scripts/gpu_cmd_ab.sh r02b_fma c4,c5 32 cs nofma fma
