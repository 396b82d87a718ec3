for r in 2 4 6 8 12 16; do timeout 120 python scripts/gpu_dev.py c4,c5 0 32 0 0 $r 2>&1 | grep -v "scene build" | sed "s/^/[refill=$r] /"; done | tee gpurun_out/ab_r02u_refill.log
