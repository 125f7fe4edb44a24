"""Markdown summary of `ncu -i <rep> --page raw --csv` exports (made on the GPU box by profiles/run_ncu_r2.sh, because the
reports themselves are too large to bring back): one row per captured launch with the metrics the roofline discussion uses.
    python profiles/ncu_raw_summary.py gpurun_out/prof_r2_gemm512_raw.csv ... > profiles/ncu_full_r2_summary.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem KB"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__sass_inst_executed_op_local_ld.sum", "local ld")]


def fmt(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return "%.2f MB" % (x / 1e6)
    return "%.4g" % x


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"## {path.split('/')[-1]}\n")
    print("| kernel | grid | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for vals in rows[2:]:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        print(f"| `{name}` | {d.get('Grid Size', '')} | " + " | ".join(fmt(d.get(k, ""), u.get(k, "")) for k, _ in COLS) + " |")
    print()
