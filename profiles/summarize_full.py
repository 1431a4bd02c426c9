"""Markdown table of the headline counters of an `ncu --set full` report (one row per captured launch).
usage: python profiles/summarize_full.py REP.ncu-rep"""
import csv, subprocess, sys
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
     "lts__t_sector_hit_rate.pct", "launch__grid_size"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
conv = {"Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "Gbyte": 1e3}
def num(r, m):
    return float(r[h.index(m)].replace(",", ""))
print("| # | kernel | grid | time us | DRAM rd MB | DRAM wr MB | DRAM rd % of peak | SM % | issue % | warps active % | regs | CTA/SM (smem) | CTA/SM (regs) | warp inst M | L2 hit % |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for n, r in enumerate(rows[2:]):
    name = r[h.index("Kernel Name")].replace("void ", "").split("(")[0]
    t = num(r, M[0]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u[h.index(M[0])], 1.0)
    rd = num(r, M[1]) * conv[u[h.index(M[1])]]; wr = num(r, M[2]) * conv[u[h.index(M[2])]]
    print(f"| {n} | {name} | {int(num(r, M[12]))} | {t:.1f} | {rd:.1f} | {wr:.1f} | {num(r, M[3]):.1f} | {num(r, M[4]):.1f} | {num(r, M[5]):.1f} | {num(r, M[6]):.1f} | "
          f"{int(num(r, M[7]))} | {int(num(r, M[8]))} | {int(num(r, M[9]))} | {num(r, M[10]) / 1e6:.1f} | {num(r, M[11]):.1f} |")
