"""Instructions executed / stall samples per source line of an .ncu-rep (needs -lineinfo + --import-source on).
Usage: ncu_lines.py file.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; cur_fn = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
tot = collections.defaultdict(lambda: [0, 0, 0])
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": cur_fn = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or len(row) < len(hdr) or row[0] == "": continue
    if want and want not in cur_fn: continue
    d = dict(zip(hdr, row))
    try:
        ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); ss = int(d["# Samples"])
    except ValueError:
        continue
    k = (cur_fn[:40], cur_file, int(row[0]))
    agg[k][0] += ie; agg[k][1] += te; agg[k][2] += ss; agg[k][3] = row[1].strip()[:110]
    tot[cur_fn[:40]][0] += ie; tot[cur_fn[:40]][1] += te; tot[cur_fn[:40]][2] += ss
for fn, (ie, te, ss) in tot.items():
    print(f"== {fn}: warp-inst {ie}  thread-inst {te}  samples {ss}")
    rows = sorted(((k, v) for k, v in agg.items() if k[0] == fn), key=lambda kv: -kv[1][2])[:top]
    for (f, file, line), (i, t, s, src) in rows:
        print(f"{100*s/max(ss,1):5.1f}%smp {100*i/max(ie,1):5.1f}%inst  {file}:{line:<5d} {src}")
