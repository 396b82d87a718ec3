cp rayrs_b200/librayrs_b200.so /tmp/keep.so
for v in cur ns2 ns3 ns6 lbm4 lbp4 lbp8; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
python scripts/gpu_dev.py c4,c5 0 32 0 2>&1 | grep -v "scene build" | sed "s/^/[$v] /"
done | tee gpurun_out/ab_r02n_sweep.log
cp _variants/lib_cur.so rayrs_b200/librayrs_b200.so
for r in 2 6 8 12; do python scripts/gpu_dev.py c4,c5 0 32 0 0 $r 2>&1 | grep -v "scene build" | sed "s/^/[refill=$r] /"; done | tee -a gpurun_out/ab_r02n_sweep.log
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
