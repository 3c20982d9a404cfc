"""Top CUDA source lines by warp-stall samples (needs -lineinfo and --import-source on).
usage: python tools/ncu_cuda_lines.py report.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
out, si, f, fn = [], None, "", ""
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": f = r[1]; continue
    if r[0] == "Function Name": fn = r[1][:60]; continue
    if r[0] == "Line No": si = r.index("# Samples"); continue
    if si is None or not r[0].isdigit(): continue
    try: n = int(r[si])
    except (ValueError, IndexError): continue
    if n > 0: out.append((n, fn, f.split("/")[-1], r[0], r[1].strip()[:120]))
for key in sorted(set(o[1] for o in out)):
    sub = sorted((o for o in out if o[1] == key), reverse=True)
    tot = sum(o[0] for o in sub)
    print("==", key, tot)
    for o in sub[:top]: print(f"{o[0]:6d} {100 * o[0] / tot:5.1f}% {o[2]}:{o[3]}  {o[4]}")
