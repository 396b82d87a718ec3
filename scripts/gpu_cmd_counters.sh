# per-ray counters of the render kernel for every workload, from ncu captures of bench.py itself
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
rm -f gpurun_out/r02_counters.json
for spec in c1:64:1 c2:256:2 c3:512:2 c4:256:1 c5:256:1; do
W=${spec%%:*}; R=${spec#*:}; S=${R%%:*}; N=${R#*:}
ncu --metrics $M --clock-control none -k regex:"k_wavefront|k_pathloop" -c $N --csv --log-file gpurun_out/counters_$W.csv \
  python bench.py --workload $W --spp $S --steps 1 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/counters_$W.json 2> gpurun_out/counters_$W.err
echo "$W ncu exit=$?"
python scripts/ncu_counters.py $W gpurun_out/counters_$W.csv gpurun_out/counters_$W.json gpurun_out/r02_counters.json | cut -c1-300
done
