# usage: gpu_cmd_launches.sh <tag>   -- ncu launch list of the bench command (per-launch durations; the kernel's share of the step)
TAG=$1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo launches_exit=$?
