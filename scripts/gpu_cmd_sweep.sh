python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/gpu_dev.py c3 | grep -v "scene build"
