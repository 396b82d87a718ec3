# Several-GPU check of the driver's own launch shape (torchrun, one rank per GPU) through rrs_render_multi.
# usage: gpu_cmd_n8.sh N    (gpurun --gpus N)
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --spp 1024 --steps 3 --warmup 3 > gpurun_out/bench_c5_spp1024_n$N.json 2> gpurun_out/bench_c5_spp1024_n$N.err; echo c5_exit=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/multigpu_check.py > gpurun_out/check_n$N.log 2>&1; echo check_exit=$?
cut -c1-400 gpurun_out/bench_c5_spp1024_n$N.json; grep -E "multigpu_check|PASS|FAIL" gpurun_out/check_n$N.log | head; tail -3 gpurun_out/bench_c5_spp1024_n$N.err
