"""Markdown table of the per-launch metrics of an .ncu-rep (read here with `ncu -i`).  usage: ncu_table.py report.ncu-rep [...]"""
import csv, subprocess, sys, io
METRICS = [("gpu__time_duration.sum", "ms", 1e-6), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
           ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1), ("dram__bytes_read.sum", "DRAM rd MB", 1e-6), ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
           ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1), ("launch__registers_per_thread", "regs", 1),
           ("smsp__warps_active.avg.per_cycle_active", "warps/SMSP", 1)]
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(m for m, _, _ in METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    print(f"### {rep.split('/')[-1]}\n")
    print("| # | kernel | " + " | ".join(n for _, n, _ in METRICS) + " |")
    print("|---|---|" + "---|" * len(METRICS))
    for k, row in enumerate(rows[2:]):
        name = row[head.index("Kernel Name")]
        name = name.replace("void ", "").replace("tc::", "").replace("nrt::", "")
        cells = []
        for m, n, sc in METRICS:
            if m not in head: cells.append("-"); continue
            i = head.index(m); v = row[i].replace(",", "")
            try: f = float(v)
            except ValueError: cells.append(v); continue
            u = units[i]
            if n == "ms": f *= UNIT.get(u, 1e-6)
            elif "MB" in n: f *= UNIT.get(u, 1e-6)
            cells.append(f"{f:.3f}" if n == "ms" else (f"{f:.0f}" if n in ("grid", "block", "regs") else f"{f:.1f}"))
        print(f"| {k} | `{name[:90]}` | " + " | ".join(cells) + " |")
    print()
