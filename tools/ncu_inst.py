"""Per-source-line instruction / sample shares of one kernel in an ncu report.
    python tools/ncu_inst.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = {}; launches = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; launches += 1; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        k = (cur, int(r[0]), r[1].strip()[:100]); v = agg.setdefault(k, [0, 0]); v[0] += int(d["Instructions Executed"]); v[1] += int(d["# Samples"])
    except Exception: pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot:,} samples {ts:,}")
cum = 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    cum += v[0]
    print(f"{k[0]}:{k[1]:4d} inst {100*v[0]/tot:5.1f}% (cum {100*cum/tot:5.1f}%) samp {100*v[1]/ts:5.1f}% | {k[2]}")
