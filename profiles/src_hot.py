"""Aggregate warp-stall samples per CUDA source line from `ncu -i REP --page source --csv --print-source cuda,sass`.
usage: python profiles/src_hot.py REP.ncu-rep LAUNCH_INDEX [TOP]"""
import csv, subprocess, sys, collections
rep, launch = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.Counter(); text = {}; func = ""; fpath = ""; cur = None; ni = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": ni = r.index("# Samples"); continue
    if r[0] != "":
        cur = (fpath, r[0]); text[cur] = r[1].strip()[:120]; continue
    if ni is not None and len(r) > ni and cur is not None:
        try: agg[cur] += int(r[ni])
        except ValueError: pass
tot = sum(agg.values())
print(func, "samples", tot)
for k, n in agg.most_common(top):
    print(f"{100*n/tot:5.1f}%  {k[0]}:{k[1]:>4}  {text[k]}")
