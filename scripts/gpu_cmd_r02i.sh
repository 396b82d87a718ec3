python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_r02i.log 2>&1; echo pytest_exit=$?; tail -14 gpurun_out/pytest_r02i.log
python bench.py --spp 64 --steps 2 --warmup 3 --cpu-seconds 4 > gpurun_out/bench_r02i_c5.json 2> gpurun_out/bench_r02i_c5.err; echo bench_exit=$?; tail -3 gpurun_out/bench_r02i_c5.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
