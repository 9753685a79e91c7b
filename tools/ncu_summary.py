"""Turn `ncu -i X.ncu-rep --page raw --csv` into the short per-launch summaries kept under profiles/.

    ncu -i gpurun_out/conv.ncu-rep --page raw --csv > /tmp/conv_raw.csv
    python tools/ncu_summary.py /tmp/conv_raw.csv [kernel-substring] > profiles/ncu_<what>_rNN_summary.txt

One block per launch, `  metric<spaces>value unit` lines (the format bench.py's `read_profile_metrics` parses: the first
launch block whose kernel name contains the requested substring).  Byte / time values are normalised to plain bytes / us
so that a reader does not have to undo ncu's auto-scaling of units."""
import csv
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def norm(val, unit):
    try:
        v = float(val.replace(",", ""))
    except ValueError:
        return val, unit
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte", "Tbyte"):
        return f"{v * SCALE[unit]:.0f}", "byte"
    if unit in ("ns", "us", "ms", "s"):
        return f"{v * SCALE[unit]:.3f}", "us"
    return f"{v:.6f}".rstrip("0").rstrip("."), unit


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    # ncu prints log lines before the header when stdout is shared: find the header row
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[h], rows[h + 1]
    short = [n.split(".", 2)[-1] if n.split(".")[0].isupper() or "Triage" in n else n for n in names]
    col = {n: i for i, n in enumerate(names)}
    n_out = 0
    for r in rows[h + 2:]:
        if len(r) < len(names) or want not in r[col["Kernel Name"]]:
            continue
        print(f"-- launch {n_out}")
        n_out += 1
        for key in ("Kernel Name", "Block Size", "Grid Size"):
            print(f"  {key:<72} {r[col[key]]}")
        seen = set()
        for k in KEEP:
            for i, n in enumerate(names):
                if (n == k or short[i] == k) and k not in seen and r[i] != "":
                    v, u = norm(r[i], units[i])
                    print(f"  {k:<72} {v} {u}")
                    seen.add(k)
    if n_out == 0:
        sys.exit(f"no launch of a kernel matching {want!r}")


if __name__ == "__main__":
    main()
