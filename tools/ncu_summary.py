"""
Summarise .ncu-rep files (read here with `ncu -i ... --page raw --csv`) into one JSON: python tools/ncu_summary.py out.json tag=file.ncu-rep ...
Keeps the metrics DESIGN.md and bench.py cite (duration, DRAM bytes, pipe utilisation, issue activity, L2 hit rate, occupancy).
"""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
out = {}
for arg in sys.argv[2:]:
    tag, path = arg.split("=", 1)
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    entry = {"report": path.split("/")[-1], "launches": []}
    for r in rows[2:]:
        d = dict(zip(header, r))
        u = dict(zip(header, units))
        launch = {"Kernel Name": d.get("Kernel Name"), "Grid Size": d.get("Grid Size"), "Block Size": d.get("Block Size")}
        for k in KEEP:
            if k in d and d[k] not in ("", "n/a"):
                launch[k] = {"value": d[k], "unit": u.get(k, "")}
        entry["launches"].append(launch)
    out[tag] = entry
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:6000])
