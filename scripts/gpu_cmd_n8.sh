# Several-GPU check of the driver's own launch shape (torchrun, one rank per GPU) through rrs_render_multi, full size.
# usage: gpu_cmd_n8.sh N [SPP]   (gpurun --gpus N)
N=$1; SPP=${2:-4096}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --spp $SPP --steps 5 --warmup 3 > gpurun_out/bench_c5_spp${SPP}_n$N.json 2> gpurun_out/bench_c5_spp${SPP}_n$N.err; echo c5_exit=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/multigpu_check.py > gpurun_out/check_n$N.log 2>&1; echo check_exit=$?
wc -l gpurun_out/bench_c5_spp${SPP}_n$N.json; cut -c1-300 gpurun_out/bench_c5_spp${SPP}_n$N.json; grep -E "multigpu_check" gpurun_out/check_n$N.log | head; tail -3 gpurun_out/bench_c5_spp${SPP}_n$N.err
