"""Per-ray constants of the render kernel from an ncu capture of bench.py itself (profiles/r02_counters.json, read
back by bench.py's roofline block).

    # on the GPU box, per workload (gpu_cmd_counters.sh does all five):
    ncu --metrics <M> --clock-control none -k regex:k_wavefront -c <scenes> --csv --log-file gpurun_out/counters_cX.csv \
        python bench.py --workload cX --spp S --steps 1 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/counters_cX.json
    python scripts/ncu_counters.py cX gpurun_out/counters_cX.csv gpurun_out/counters_cX.json gpurun_out/r02_counters.json

The first `scenes` launches of the render kernel are one step of the workload (one launch per scene); their counters
are summed and divided by the rays of a step, which the same bench run reports (rays per step do not depend on which
step).  Instruction counts per ray are a property of the code and the workload; DRAM bytes per ray and the hit rates
are measured at the capture's spp (cold caches, serialised launches) and are lower bounds on locality."""
import csv
import json
import sys
from pathlib import Path

key, csv_path, bench_path, out_path = sys.argv[1:5]
line = [ln for ln in Path(bench_path).read_text().splitlines() if ln.startswith("{")][-1]
bench = json.loads(line)
n_scenes = len(bench["config"]["scenes"])
rays = float(bench["config"]["rays_per_step"])
rows = [r for r in csv.reader(open(csv_path)) if len(r) > 8]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    if r[0] == "ID" or not r[ix["ID"]].isdigit():
        continue
    lid = int(r[ix["ID"]])
    if lid >= n_scenes:
        continue
    launches.setdefault(lid, {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
assert len(launches) == n_scenes, (len(launches), n_scenes)
tot = lambda m: sum(v.get(m, 0.0) for v in launches.values())
unit_ns = 1.0
t_ns = tot("gpu__time_duration.sum")
wi, ti = tot("smsp__inst_executed.sum"), tot("smsp__thread_inst_executed.sum")
dram = tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")
avg = lambda m: sum(v.get(m, 0.0) * v.get("gpu__time_duration.sum", 0.0) for v in launches.values()) / max(t_ns, 1e-9)
entry = {
    "kernel": sorted({v["kernel"].split("(")[0] for v in launches.values()}),
    "warp_inst_per_ray": wi / rays, "thread_inst_per_ray": ti / rays, "lanes_per_instruction": ti / wi,
    "dram_bytes_per_ray": dram / rays, "dram_read_bytes_per_ray": tot("dram__bytes_read.sum") / rays,
    "issue_active_pct": avg("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "l1_global_load_hit_pct": avg("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"),
    "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"),
    "local_load_requests_per_ray": tot("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum") / rays,
    "global_load_requests_per_ray": tot("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum") / rays,
    "long_scoreboard_stall_per_issue": avg("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    "l1tex_throughput_pct": avg("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    "l1_lsu_wavefronts_pct": avg("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "l1_global_load_sectors_per_ray": tot("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / rays,
    "l2_throughput_pct": avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "captured": {"command": f"bench.py --workload {key} --spp {bench['config']['spp_total']}", "launches": n_scenes, "rays": rays,
                 "launch_ms_under_ncu": t_ns / 1e6, "spp": bench["config"]["spp_total"]},
    "source": f"profiles/r02/counters_{key}.csv (ncu of bench.py, scripts/ncu_counters.py)",
}
out = Path(out_path)
data = json.loads(out.read_text()) if out.exists() else {}
data[key] = entry
out.write_text(json.dumps(data, indent=1))
print(key, json.dumps(entry))
