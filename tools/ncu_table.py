"""One line per kernel launch of an ncu report (duration, instructions, DRAM traffic, issue / warps / tensor / DRAM utilisation):
    python tools/ncu_table.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
col = {n: h.index(n) for n in h}
def g(r, n, scale=1.0):
    try: return float(r[col[n]].replace(",", "")) * scale
    except Exception: return float("nan")
units = rows[1]
def to_us(r):
    u = units[col["gpu__time_duration.sum"]]
    v = g(r, "gpu__time_duration.sum")
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
def to_mb(r, n):
    u = units[col[n]]
    return g(r, n) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
print(f"{'kernel':50s} {'us':>7s} {'inst(M)':>8s} {'rdMB':>7s} {'wrMB':>7s} {'issue%':>6s} {'warps%':>6s} {'tensor%':>7s} {'dram%':>6s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("clipgp::", "").replace("void ", "")[:50]
    print(f"{name:50s} {to_us(r):7.1f} {g(r, 'smsp__inst_executed.sum', 1e-6):8.2f} {to_mb(r, 'dram__bytes_read.sum'):7.2f} "
          f"{to_mb(r, 'dram__bytes_write.sum'):7.2f} {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f}")
