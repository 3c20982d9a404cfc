"""Summarise `ncu --page source --csv` output: top SASS lines by stall samples + stall-reason totals.
usage: python tools/ncu_src_summary.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = blk.splitlines()
    print("== kernel", lines[0][:120])
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rd[0]; rows = rd[1:]
    idx = {h: i for i, h in enumerate(hdr)}
    samp = idx["# Samples"]; src = idx["Source"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter(); total = 0
    for r in rows:
        try: n = int(r[samp])
        except ValueError: continue
        total += n
        for h in stall_cols:
            try: tot[h] += int(r[idx[h]])
            except ValueError: pass
    print("total samples", total)
    print("stall reasons:", ", ".join(f"{k[6:]}={v} ({100*v/max(total,1):.1f}%)" for k, v in tot.most_common(8)))
    rows2 = []
    for i, r in enumerate(rows):
        try: rows2.append((int(r[samp]), i, r))
        except ValueError: pass
    rows2.sort(reverse=True)
    for n, i, r in rows2[:top]:
        reasons = sorted(((int(r[idx[h]]) if r[idx[h]].isdigit() else 0, h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{n:7d} {100*n/max(total,1):5.1f}%  #{i:4d} {r[src][:90]:90s} {reasons}")
