cp rayrs_b200/librayrs_b200.so /tmp/orig.so
for v in orig 8_8 8_7; do
if [ $v = orig ]; then cp /tmp/orig.so rayrs_b200/librayrs_b200.so; else cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so; fi
python scripts/gpu_dev.py c5 0 16 2>&1 | grep -v "scene build" | sed "s/^/v=$v /"
python scripts/gpu_dev.py c3 0 64 2>&1 | grep -v "scene build" | sed "s/^/v=$v /"
done | tee gpurun_out/sweep_occ3.log
cp /tmp/orig.so rayrs_b200/librayrs_b200.so
