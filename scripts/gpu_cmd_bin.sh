cp rayrs_b200/librayrs_b200.so /tmp/keep.so
cp _variants/lib_bin.so rayrs_b200/librayrs_b200.so
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "not full_size" 2>&1 | tail -2
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
bash scripts/gpu_cmd_ab.sh r02t_bin c4,c5 32 nobin bin
