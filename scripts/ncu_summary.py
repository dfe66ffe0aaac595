"""Summarise an .ncu-rep (read here, no GPU): python scripts/ncu_summary.py rep.ncu-rep [regex]
Prints one JSON line per profiled launch with the metrics the roofline discussion uses."""
import csv, io, json, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    tensor_cols = [n for n in head if n in ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                                            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                            "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")
                   or n.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        d = {"kernel": name.split("(")[0], "grid": r[col.get("Grid Size", 0)]}
        for k in KEYS + [c for c in tensor_cols if c not in KEYS]:
            if k in col and r[col[k]] != "":
                d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[col["dram__bytes_write.sum"]].replace(",", ""))
            mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            d["dram_bytes"] = rd * mult.get(units[col["dram__bytes_read.sum"]], 1.0) + wr * mult.get(units[col["dram__bytes_write.sum"]], 1.0)
        except Exception:
            pass
        print(json.dumps(d))

main()
