"""Markdown table from an `ncu --set full` report: python tools/ncu_summary.py file.ncu-rep
(the tables under profiles/*_ncu_full_summary.md).  Reads `ncu -i <rep> --page raw --csv`, or that page already exported
to a .csv file (the reports of one session exceed what the GPU box copies back, so tools/_scan_job.sh exports there)."""
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}

    def val(r, name, scale_time=False):
        v = num(r[col[name]])
        if scale_time:
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col[name]], 1.0)
        return v

    def to_mb(r, name):
        return val(r, name) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[col[name]], 1e-6)

    stall_cols = [h for h in head if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in head if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
    print("| kernel | time ms | grid x block | regs | static smem KB | FP64 pipe % | issue active % | warp instructions | "
          "DRAM read MB | DRAM write MB | L1 hit % | L2 hit % | top stalls (cycles per issue) |")
    print("|" + "---|" * 13)
    for r in data:
        if len(r) < len(head):
            continue
        stalls = sorted(((val(r, h), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")) for h in stall_cols),
                        reverse=True)[:5]
        smem = val(r, "launch__shared_mem_per_block_static") * {"byte": 1e-3, "Kbyte": 1.0}.get(
            units[col["launch__shared_mem_per_block_static"]].split("/")[0], 1e-3)
        print("| %s | %.4f | %d x %d | %d | %.1f | %.1f | %.1f | %.3e | %.0f | %.0f | %.1f | %.1f | %s |" % (
            r[col["Kernel Name"]][:34], val(r, "gpu__time_duration.sum", True), val(r, "launch__grid_size"),
            val(r, "launch__block_size"), val(r, "launch__registers_per_thread"), smem,
            val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), val(r, "smsp__inst_executed.sum"),
            to_mb(r, "dram__bytes_read.sum"), to_mb(r, "dram__bytes_write.sum"), val(r, "l1tex__t_sector_hit_rate.pct"),
            val(r, "lts__t_sector_hit_rate.pct"), ", ".join("%s %.1f" % (n, v) for v, n in stalls)))


if __name__ == "__main__":
    main(sys.argv[1])
