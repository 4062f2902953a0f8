"""Condense `ncu -i REPORT.ncu-rep --page raw --csv` into the markdown table kept under profiles/.
Usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv > X_raw.csv ; python tools/summarize_ncu_full.py X_raw.csv out.md "title" """
import csv
import sys

path, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "ncu --set full")
rows = list(csv.reader([l for l in open(path) if not l.startswith("==")]))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "dur"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_sectors_op_read.sum", "L2 rd sectors"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__grid_size", "grid"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %")]
cols = [(hdr.index(m), lab) for m, lab in want if m in hdr]
ki = hdr.index("Kernel Name")
seen, lines = set(), [f"# {title}\n", "Isolated launches at the BASELINE shapes (cold cache, serialised; clocks not locked).\n",
                      "| kernel | " + " | ".join(f"{lab} [{units[i]}]" for i, lab in cols) + " |", "|---|" + "---|" * len(cols)]
for r in data:
    key = (r[ki], r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
    if key in seen:
        continue
    seen.add(key)
    lines.append("| `" + r[ki][:60] + "` | " + " | ".join(r[i] for i, _ in cols) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
