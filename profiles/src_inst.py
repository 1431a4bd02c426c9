"""Per CUDA source line: warp instructions executed, stall samples and the dominant stall reasons, from
`ncu -i REP --page source --csv --print-source cuda,sass`.   usage: python profiles/src_inst.py REP.ncu-rep LAUNCH_INDEX [TOP]"""
import csv, subprocess, sys, collections
rep, launch = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
inst = collections.Counter(); samp = collections.Counter(); stalls = collections.defaultdict(collections.Counter); text = {}
func = fpath = ""; cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "":
        cur = (fpath, r[0]); text[cur] = r[1].strip()[:100]; continue
    if hdr is None or cur is None or len(r) < len(hdr): continue
    try:
        inst[cur] += int(r[hdr.index("Instructions Executed")]); samp[cur] += int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "(" not in h:
            try: stalls[cur][h[6:]] += int(r[i])
            except ValueError: pass
ti, ts = sum(inst.values()), sum(samp.values())
print(func, "warp instructions", ti, "samples", ts)
tot_st = collections.Counter()
for k in stalls: tot_st.update(stalls[k])
print("stall mix:", ", ".join(f"{k} {100*v/max(sum(tot_st.values()),1):.0f}%" for k, v in tot_st.most_common(8)))
for k, n in inst.most_common(top):
    st = ", ".join(f"{a} {b}" for a, b in stalls[k].most_common(3))
    print(f"{100*n/ti:5.1f}% inst {100*samp[k]/max(ts,1):5.1f}% samp  {k[0]}:{k[1]:>4}  {text[k]}   [{st}]")
