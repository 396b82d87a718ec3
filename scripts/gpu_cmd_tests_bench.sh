# usage: gpu_cmd_tests_bench.sh <tag>  -- GPU test suite + bench lines of configs 2, 3, 4 (no profiler in this call)
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo bench_exit=$?
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err; echo c3_exit=$?
python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c1.json 2> gpurun_out/bench_${TAG}_c1.err; echo c1_exit=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cut -c1-200 gpurun_out/bench_${TAG}_c2.json; cut -c1-200 gpurun_out/bench_${TAG}_c3.json; cut -c1-200 gpurun_out/bench_${TAG}_c1.json
