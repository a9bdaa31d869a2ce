"""Markdown table of the key ncu metrics per kernel launch: python scripts/ncu_md.py <raw.csv | file.ncu-rep>"""
import csv, io, subprocess, sys
src = sys.argv[1]
txt = open(src).read() if src.endswith(".csv") else subprocess.run(
    ["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("us", "gpu__time_duration.sum"), ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1/smem %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
print("| kernel | grid x block | " + " | ".join(c for c, _ in cols) + " |")
print("|---|---|" + "---|" * len(cols))
def num(r, m):
    if m not in idx: return "-"
    v, u = r[idx[m]].replace(",", ""), units[idx[m]]
    try: f = float(v)
    except ValueError: return v
    if u == "ns": f /= 1e3
    if u == "byte": f /= 1e6
    if u == "Kbyte": f /= 1e3
    if u == "Gbyte": f *= 1e3
    return f"{f:.1f}" if f < 1e5 else f"{f:.3g}"
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    print(f"| {name} | {r[idx['Grid Size']]} x {r[idx['Block Size']]} | " + " | ".join(num(r, m) for _, m in cols) + " |")
