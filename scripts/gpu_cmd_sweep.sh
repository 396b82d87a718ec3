cp rayrs_b200/librayrs_b200.so /tmp/orig.so
for v in orig 256 64; do
if [ $v = orig ]; then cp /tmp/orig.so rayrs_b200/librayrs_b200.so; else cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so; fi
python scripts/gpu_dev.py c2,c4 0 64 2>&1 | grep -v "scene build" | sed "s/^/block=$v /"
done | tee gpurun_out/sweep_block.log
cp /tmp/orig.so rayrs_b200/librayrs_b200.so
