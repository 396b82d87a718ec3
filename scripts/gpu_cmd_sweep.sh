python -m pytest tests/test_gpu_intersect.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
python scripts/gpu_dev.py c4 | grep -v "scene build"
python scripts/gpu_dev.py c5 0 16 | grep -v "scene build"
