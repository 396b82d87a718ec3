# round-end evidence on one B200: GPU suite (-s), the driver's bench command, the reference arm (short), launch list, one full capture
TAG=${1:-r02}
# the per-ray counters first: bench.py's roofline block multiplies them by the rays of the timed launches
bash scripts/gpu_cmd_counters.sh > gpurun_out/counters_${TAG}.log 2>&1; tail -3 gpurun_out/counters_${TAG}.log | cut -c1-200
cp gpurun_out/r02_counters.json profiles/r02_counters.json
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_${TAG}.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err ) 2>&1 | grep real; echo bench_exit=$?; tail -2 gpurun_out/bench_${TAG}_c5.err; cut -c1-300 gpurun_out/bench_${TAG}_c5.json
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | grep real; cut -c1-200 gpurun_out/bench_${TAG}_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wavefront|k_resolve|elementwise|ncclDevKernel" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --spp 64 --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo launches_exit=$?
python scripts/ncu_one.py c5 16 > gpurun_out/plain_${TAG}_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_c5 python scripts/ncu_one.py c5 16 > gpurun_out/ncu_${TAG}_c5.log 2>&1; echo full_exit=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
