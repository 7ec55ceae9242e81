"""Rank CUDA source lines by excessive shared-memory wavefronts (bank conflicts) from an `ncu --page source --csv` dump:
    ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:NAME | python tools/ncu_shared_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 14
cur = None; hdr = None; agg = {}
for r in csv.reader(sys.stdin):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        ex, w, ins = int(d["L1 Wavefronts Shared Excessive"] or 0), int(d["L1 Wavefronts Shared"] or 0), int(d["Instructions Executed"] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1].strip()[:100]])
    a[0] += ex; a[1] += w; a[2] += ins
tot_ex = sum(a[0] for a in agg.values()); tot_w = sum(a[1] for a in agg.values())
print(f"shared-memory wavefronts {tot_w:,}  of which excessive (bank conflicts) {tot_ex:,} = {100 * tot_ex / max(tot_w, 1):.1f}%")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} excess {a[0]:9,d} ({100 * a[0] / max(tot_ex, 1):4.1f}%)  wavefronts {a[1]:9,d}  inst {a[2]:9,d} | {a[3]}")
