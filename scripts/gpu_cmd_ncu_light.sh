# usage: gpu_cmd_ncu_light.sh <tag> <config> <spp> <variant> [...]  -- a few counters of k_wavefront per library variant
TAG=$1; CFG=$2; SPP=$3; shift 3
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
cp rayrs_b200/librayrs_b200.so /tmp/keep.so
for v in "$@"; do
cp _variants/lib_$v.so rayrs_b200/librayrs_b200.so
ncu --metrics $M --clock-control none -k regex:k_wavefront -s 1 -c 1 --csv --log-file gpurun_out/light_${TAG}_$v.csv python scripts/ncu_one.py $CFG $SPP > gpurun_out/light_${TAG}_$v.log 2>&1; echo "$v exit=$?"
python scripts/gpu_dev.py $CFG 0 8 1 2>&1 | grep nodes/ray | sed "s/^/[$v] /"
done
cp /tmp/keep.so rayrs_b200/librayrs_b200.so
