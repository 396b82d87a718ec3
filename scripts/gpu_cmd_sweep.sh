python -m pytest tests/test_output_stage.py -m gpu -x -q 2>&1 | tail -3
cp rayrs_b200/librayrs_b200.so /tmp/orig.so
for v in 8_6 6_6 6_5 10_6 8_8; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
python scripts/gpu_dev.py c2,c4 0 64 2>&1 | grep -v "scene build" | sed "s/^/v=$v /"
python scripts/gpu_dev.py c3 0 64 2>&1 | grep -v "scene build" | sed "s/^/v=$v /"
done | tee gpurun_out/sweep_occ.log
cp /tmp/orig.so rayrs_b200/librayrs_b200.so
