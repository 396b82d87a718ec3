python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo n2_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo ref_exit=$?
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q 2>&1 | tail -2
cut -c1-700 gpurun_out/bench_c2_n2.json; cut -c1-300 gpurun_out/bench_ref_n2.json
