python -m pytest tests/test_gpu_render.py -m gpu -x -q -s -k "full_size_config" 2>&1 | tail -25
