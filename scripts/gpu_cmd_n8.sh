N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo c2_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --spp 32 --steps 2 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo c5_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/multigpu_check.py > gpurun_out/check_n$N.log 2>&1; echo check_exit=$?
cut -c1-260 gpurun_out/bench_c2_n$N.json; cut -c1-260 gpurun_out/bench_c5_n$N.json; grep multigpu_check gpurun_out/check_n$N.log; tail -3 gpurun_out/bench_c5_n$N.err
