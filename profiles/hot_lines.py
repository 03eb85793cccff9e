"""Aggregate the ncu source page (needs -lineinfo + --import-source on) per CUDA source line:
python profiles/hot_lines.py rep kernel_regex [top]  -> instructions executed, stall samples and the dominant stall reasons"""
import csv, io, subprocess, sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = ""; hdr = None
agg = defaultdict(lambda: defaultdict(float)); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue          # SASS rows have an empty line number
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    text[key] = r[1].strip()[:110]
    def num(x):
        try: return float(x)
        except (TypeError, ValueError): return 0.0
    agg[key]["inst"] += num(d.get("Instructions Executed"))
    agg[key]["samples"] += num(d.get("# Samples"))
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[key][k] += float(v or 0)
            except ValueError: pass
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values())
print(f"total inst {tot_i:.3e} samples {tot_s:.0f}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print(f"{key[0]}:{key[1]:4d} inst {100*a['inst']/tot_i:5.1f}% samp {100*a['samples']/tot_s:5.1f}%  {' '.join(f'{k}={100*v/max(a['samples'],1):.0f}%' for k,v in st)} | {text[key]}")
