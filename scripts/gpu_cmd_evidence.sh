# usage: gpu_cmd_evidence.sh <tag>   -- tests + bench lines (no profiler in this call; one profiler per gpurun call:
#        gpu_cmd_launches.sh, gpu_cmd_ncu_c2.sh, gpu_cmd_ncu.sh, each after its command exited 0 without ncu)
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo bench_exit=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo ref_exit=$?
python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err; echo c4_exit=$?
cut -c1-300 gpurun_out/bench_${TAG}_c2.json; cut -c1-300 gpurun_out/bench_${TAG}_ref.json; cut -c1-300 gpurun_out/bench_${TAG}_c4.json
