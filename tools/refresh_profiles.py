"""Copy a record run (tools/record_run.sh, merged into gpurun_out/) into profiles/ in the judged formats.
usage: python tools/refresh_profiles.py [tag [bench_dir [ncu_rep_dir [out_dir]]]]
       (defaults: r2, gpurun_out/, gpurun_out/, profiles/; tools/record_run.sh runs it ON the GPU box with the reports in
       /tmp and out_dir = gpurun_out/profiles_r2, which is then copied into profiles/)"""
import csv, io, json, os, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
REP = sys.argv[3] if len(sys.argv) > 3 else G
P = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def first_json_line(src, dst):
    for line in open(src):
        if line.startswith("{"):
            open(dst, "w").write(line)
            return json.loads(line)
    return None


def copy_bench(src_name, dst_name):
    src = os.path.join(G, src_name)
    if os.path.exists(src) and os.path.getsize(src):
        return first_json_line(src, os.path.join(P, dst_name))
    return None


copy_bench("bench_final.json", f"{tag}_bench_cfg2.json")
copy_bench("bench_final_reference.json", f"{tag}_bench_cfg2_reference_arm.json")
copy_bench("bench_layer.json", f"{tag}_bench_cfg2_layer_scope.json")
for v in ("jittor", "torch"):
    copy_bench(f"bench_cfg3_{v}.json", f"{tag}_bench_cfg3_{v}.json")
    for op in ("bf16", "fp32"):
        copy_bench(f"bench_stack_{v}_{op}.json", f"{tag}_bench_stack_{v}_{op}.json")
copy_bench("det_1gpu_final.json", f"{tag}_bench_detector_1gpu.json")
copy_bench("det_b16_graph.json", f"{tag}_bench_detector_cfg1_batch16.json")
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, f"{tag}_launches_bench_cfg2.csv"))
# single-layer table
rows = []
for w in ("c3", "c4", "c5", "det2", "det5"):
    for v in ("jittor", "torch"):
        f = os.path.join(G, f"layer_{w}_{v}.json")
        if os.path.exists(f) and os.path.getsize(f):
            d = None
            for line in open(f):
                if line.startswith("{"):
                    d = json.loads(line)
            if d:
                k = {n: round(x["avg_ms"], 4) for n, x in d["kernels"].items() if x["avg_ms"] >= 0.05}
                rows.append(f"{w:5s} {v:7s} {d['value']:10.0f} img/s {d['ms_per_step']:8.3f} ms  paths {d['config']['path_fwd']}/{d['config']['path_bwd']}  {k}")
open(os.path.join(P, f"{tag}_sweep_layers.txt"), "w").write(
    "# python bench.py --workload W --variant V --no-cpu-baseline --no-e2e --steps 10 (fp32 operands, fwd+bwd, span scope)\n" + "\n".join(rows) + "\n")
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
            "--clock-control none; python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detector-dp)\n")
    f.write("kernel,launches,total_ms,share\n")
    for n, ms in tot.most_common():
        f.write(f"{n},{cnt[n]},{ms:.3f},{ms / allms:.4f}\n")


def summarise(rep_name, out_name, header, top=12):
    rep = os.path.join(REP, rep_name)
    if not os.path.exists(rep):
        return None
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_key_metrics.py"), rep],
                         capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_cuda_lines.py"), rep, str(top)],
                           capture_output=True, text=True).stdout
    with open(os.path.join(P, out_name), "w") as f:
        f.write(header)
        f.write(txt)
        f.write("\n# ---- warp-stall samples by CUDA source line (tools/ncu_cuda_lines.py)\n")
        f.write(lines)
    return rep


def dram_per_kernel(rep):
    """-> [(kernel name, dram bytes read + written)] in launch order"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, units = rr[0], rr[1]
    out = []
    for r in rr[2:]:
        d = dict(zip(h, r))

        def gb(key):
            u = units[h.index(key)]
            return float(d[key].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        out.append((d["Kernel Name"], gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")))
    return out


def key_of(name):
    if "umma_gemm_kernel" in name:
        return "umma_bwd_weight_kernel" if "<1, 1," in name or "<0, 1," in name else "umma_fwd_kernel"
    if "bwd_data_kernel" in name:
        return "umma_bwd_data_kernel"
    return None


_told = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(_told)) if os.path.exists(_told) else {}
traffic = {"_limiter": old.get("_limiter", {})}
rep = summarise("prof_final2.ncu-rep", f"{tag}_final_ncu.txt",
                "# ncu --set full --clock-control none --import-source on -k regex:'bwd_data_kernel|umma_gemm_kernel|"
                "gout_tiles|nchw_to_nhwc|nhwc_to_nchw' -c 6 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detector-dp\n"
                "# cfg2 (64->64, 3x3, 128x128, B=256), Torch layout, fp32\n")
if rep:
    for name, b in dram_per_kernel(rep):
        k = key_of(name)
        if k:
            traffic[f"cfg2:torch:{k}"] = int(b)
for v in ("jittor", "torch"):
    rep = summarise(f"prof_cfg3_{v}.ncu-rep", f"{tag}_ncu_cfg3_{v}.txt",
                    f"# ncu --set full ... -k regex:'bwd_data_kernel|umma_gemm_kernel' -c 3 python bench.py --workload cfg3 --variant {v} "
                    "--steps 1 --warmup 0 --no-e2e --no-cpu-baseline\n# cfg3 (256->256 @28x28, B=64), fp32\n")
    if rep:
        for name, b in dram_per_kernel(rep):
            k = key_of(name)
            if k:
                traffic[f"cfg3:{v}:{k}"] = int(b)
rep = summarise("prof_stack_jittor_bf16.ncu-rep", f"{tag}_ncu_stack_jittor_bf16.txt",
                "# ncu --set full ... -k regex:'bwd_data_kernel|umma_gemm_kernel' -c 26 python bench.py --workload stack --variant jittor "
                "--operand bf16 --steps 1 --warmup 0\n# configs[3]: the 13 layers of one step (forward + fused backward each)\n", top=8)
if rep:
    agg = collections.Counter()
    for name, b in dram_per_kernel(rep):
        k = key_of(name)
        if k:
            agg[k] += b
    for k, b in agg.items():
        traffic[f"stack:jittor:bf16:{k}"] = int(b)     # all 13 launches of one step
summarise("prof_layer_conv.ncu-rep", f"{tag}_ncu_layer_conv.txt",
          "# ncu --set full ... -k regex:'conv_kernel' -c 3 python bench.py --scope layer --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-detector-dp\n"
          "# cfg2 whole layer: the companion offset conv on the shifted-view kernels (forward, data gradient, weight gradient)\n")
traffic["_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/{tag}_*ncu*.txt (ncu --set full); "
                    "stack:* entries are sums over the 13 launches of one step")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_shares_cfg2.csv")).read())
print(json.dumps(traffic, indent=1))
