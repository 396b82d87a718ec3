python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_n2.log 2>&1; echo pytest_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo n2_exit=$?
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4_exit=$?
python bench.py --workload c5 --spp 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo c5_exit=$?
tail -3 gpurun_out/pytest_n2.log; cat gpurun_out/bench_c2_n2.json | cut -c1-600; tail -3 gpurun_out/bench_c2_n2.err; cut -c1-400 gpurun_out/bench_c4.json; cut -c1-400 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
