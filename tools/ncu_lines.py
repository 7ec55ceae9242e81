"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line:
    ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:NAME | python tools/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
cur_file = None; hdr = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        agg.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]),
                    {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}))
    except Exception: pass
tot_i = sum(a[3] for a in agg); tot_s = sum(a[4] for a in agg)
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    st = sorted(a[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{a[0]}:{a[1]:4d} inst {100*a[3]/max(tot_i,1):5.1f}% samp {100*a[4]/max(tot_s,1):5.1f}%  {st}  | {a[2]}")
