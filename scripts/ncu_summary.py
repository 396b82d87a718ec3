"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.

    python scripts/ncu_summary.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]
writes profiles/<tag>_launches.csv (per-kernel totals and shares of the step), profiles/<tag>_metrics.csv
(key metrics per profiled launch) and profiles/<tag>_stalls.txt (most-stalled SASS instructions).
"""
import argparse
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_red.sum",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(path: Path, out: Path):
    lines = [l for l in path.read_text().splitlines() if l and not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with out.open("w") as f:
        f.write("kernel,launches,total_us,avg_us,share_of_profiled_time\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{v[1] / tot:.4f}\n")
    print(out.read_text())


def metrics(rep: Path, out_csv: Path, out_stalls: Path):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with out_csv.open("w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [d[idx["Kernel Name"]].split("(")[0] for d in data])
        for k in KEY:
            if k in idx:
                w.writerow([k, units[idx[k]]] + [d[idx[k]] for d in data])
    print(out_csv.read_text())
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1].split("(")[0], "hdr": None, "inst": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) > 5:
            cur["inst"].append(r)
    seen = set()
    with out_stalls.open("w") as f:
        for b in blocks:
            if b["name"] in seen or not b["hdr"]:
                continue
            seen.add(b["name"])
            h = b["hdr"]
            i_s, i_src, i_ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
            tot = sum(int(r[i_s]) for r in b["inst"]) or 1
            f.write(f"== {b['name']}: {tot} stall samples over {len(b['inst'])} SASS instructions ==\n")
            for k, r in sorted(sorted(enumerate(b["inst"]), key=lambda kv: -int(kv[1][i_s]))[:14]):
                f.write(f"  #{k:5d} {100 * int(r[i_s]) / tot:5.1f}%  executed {r[i_ex]:>9s}  {r[i_src].strip()[:100]}\n")
    print(out_stalls.read_text())


def update_model(metrics_csv: Path, key: str, captured_rays: float = 0.0):
    """Copy the measured per-launch DRAM traffic and the pipe / issue evidence of the profiled kernel into
    profiles/bytes_per_ray.json[key]["ncu"], where bench.py picks up roofline.traffic."""
    import json
    rows = {r[0]: r for r in csv.reader(metrics_csv.open())}
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def mean(name, scale=False):
        r = rows[name]
        vals = [float(x) for x in r[2:]]
        return (sum(vals) / len(vals)) * (unit[r[1]] if scale else 1.0)
    model_path = ROOT / "profiles" / "bytes_per_ray.json"
    model = json.loads(model_path.read_text())
    model[key]["ncu"] = {
        "dram_bytes_per_launch": mean("dram__bytes_read.sum", True) + mean("dram__bytes_write.sum", True),
        "kernel": rows["metric"][2].strip('"'),
        "launch_ms_under_ncu": mean("gpu__time_duration.sum"),
        "warp_instructions_per_launch": mean("smsp__inst_executed.sum"),
        "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": mean("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": mean("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "xu_pipe_pct": mean("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "lanes_per_instruction": mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "dram_pct_of_peak": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "source": f"profiles/{metrics_csv.name}",
    }
    if captured_rays:
        # the capture rendered fewer samples per pixel than the bench command: bench.py scales the per-launch
        # counts (DRAM bytes, warp instructions) by rays of its launch / rays of the captured launch
        model[key]["ncu"]["rays_per_launch_captured"] = captured_rays
    model_path.write_text(json.dumps(model, indent=1) + "\n")
    print("updated", model_path, key, model[key]["ncu"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default="", help="ncu launch-list csv of the same command (omit for captures without one)")
    ap.add_argument("--rep", default=str(ROOT / "gpurun_out" / "prof.ncu-rep"))
    ap.add_argument("--model-key", default="", help="c1..c5: also record the capture in profiles/bytes_per_ray.json")
    ap.add_argument("--captured-rays", type=float, default=0.0,
                    help="rays traced by the captured launch when it was not the bench command itself (reduced spp)")
    a = ap.parse_args()
    prof = ROOT / "profiles"
    prof.mkdir(exist_ok=True)
    if a.launches and Path(a.launches).exists():
        launches(Path(a.launches), prof / f"{a.tag}_launches.csv")
    if Path(a.rep).exists():
        metrics(Path(a.rep), prof / f"{a.tag}_metrics.csv", prof / f"{a.tag}_stalls.txt")
        if a.model_key:
            update_model(prof / f"{a.tag}_metrics.csv", a.model_key, a.captured_rays)
