"""Copy a record run (tools/record_run.sh, merged into gpurun_out/) into profiles/ in the judged formats.
usage: python tools/refresh_profiles.py [tag]      (tag defaults to r1)"""
import csv, io, json, os, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, f"{tag}_bench_cfg2.json"))
shutil.copy(os.path.join(G, "bench_final_reference.json"), os.path.join(P, f"{tag}_bench_cfg2_reference_arm.json"))
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, f"{tag}_launches_bench_cfg2.csv"))
# per-kernel shares of the launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "launches_final.csv")) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    name = r[ki].split("(")[0]
    tot[name] += ms
    cnt[name] += 1
allms = sum(tot.values())
with open(os.path.join(P, f"{tag}_launch_shares_cfg2.csv"), "w") as f:
    f.write(f"# per-kernel totals of profiles/{tag}_launches_bench_cfg2.csv (ncu --metrics gpu__time_duration.sum "
            "--clock-control none; python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline)\n")
    f.write("kernel,launches,total_ms,share\n")
    for n, ms in tot.most_common():
        f.write(f"{n},{cnt[n]},{ms:.3f},{ms / allms:.4f}\n")
# key metrics of the full capture + DRAM traffic per launch
rep = os.path.join(G, "prof_final2.ncu-rep")
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_key_metrics.py"), rep],
                     capture_output=True, text=True).stdout
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_cuda_lines.py"), rep, "12"],
                       capture_output=True, text=True).stdout
with open(os.path.join(P, f"{tag}_final_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:'bwd_data_kernel|umma_gemm_kernel|"
            "gout_tiles|nchw_to_nhwc|nhwc_to_nchw' -c 6 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline\n"
            "# cfg2 (64->64, 3x3, 128x128, B=256), Torch layout, fp32 (gpurun_out/prof_final2.ncu-rep)\n")
    f.write(txt)
    f.write("\n# ---- warp-stall samples by CUDA source line (tools/ncu_cuda_lines.py)\n")
    f.write(lines)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
old = json.load(open(os.path.join(P, "traffic.json"))) if os.path.exists(os.path.join(P, "traffic.json")) else {}
traffic = {"_limiter": old.get("_limiter", {})}
for r in rr[2:]:
    d = dict(zip(h, r))
    def gb(key):
        u = units[h.index(key)]
        v = float(d[key].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    name = d["Kernel Name"]
    key = "umma_fwd_kernel" if "umma_gemm_kernel" in name else ("umma_bwd_data_kernel" if "bwd_data_kernel" in name else None)
    if key:
        traffic[f"cfg2:torch:{key}"] = int(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
traffic["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/{tag}_final_ncu.txt (ncu --set full)"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_shares_cfg2.csv")).read())
print(traffic)
