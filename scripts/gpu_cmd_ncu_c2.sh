# usage: gpu_cmd_ncu_c2.sh <tag>   -- one ncu --set full capture of the headline kernel under the bench command
TAG=$1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s 8 -c 2 -f -o gpurun_out/prof_${TAG}_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; echo full_exit=$?
