# usage: gpu_cmd_evidence.sh <tag>   -- the round's measured evidence: tests, bench lines, ncu launch list + full captures
TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo bench_exit=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo ref_exit=$?
python bench.py --workload c4 --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err; echo c4_exit=$?
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo launches_exit=$?
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_wavefront' -s 8 -c 2 -f -o gpurun_out/prof_${TAG}_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; echo full_exit=$?
python scripts/ncu_one.py c4 32 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_wavefront' -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_c4 python scripts/ncu_one.py c4 32 > gpurun_out/ncu_c4.log 2>&1; echo c4full_exit=$?
cut -c1-300 gpurun_out/bench_${TAG}_c2.json; cut -c1-300 gpurun_out/bench_${TAG}_ref.json; cut -c1-300 gpurun_out/bench_${TAG}_c4.json
