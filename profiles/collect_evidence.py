"""Copy the outputs of profiles/gpu_round.sh from gpurun_out/ into profiles/r01_* and rebuild
traffic.json + the ncu summary from the .ncu-rep captures:  python profiles/collect_evidence.py"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_active.avg', 'sm__cycles_active.max', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__waves_per_multiprocessor']
tob = lambda v, u: float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
traffic, out = {}, []
for f, title in (('prof_persist', 'python profiles/prof_rnn_layer.py bf16 lstm  (cfg1 layer: T=64 B=50 H=128, both directions)'),
                 ('prof_gemm', 'python profiles/prof_gemm_one.py 0 1 3200 1024 128 tf32  (cfg1 hoisted input projection, layer 0)')):
    rep = os.path.join(G, f + '.ncu-rep')
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.append(f"# ncu --set full --clock-control none: {title}")
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        out.append(name)
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"  {k} = {r[i]} {units[i]}")
        short = name.split('(')[0].split('<')[0].replace('void ', '').replace('slnlp::', '').strip()
        traffic[short] = tob(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')]) + \
            tob(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
if out:
    open(os.path.join(P, 'r01_ncu_kernels_raw.txt'), 'w').write("\n".join(out) + "\n")
    old = {}
    try:
        old = json.load(open(os.path.join(P, 'traffic.json')))
    except (OSError, ValueError):
        pass
    old.update(traffic)
    json.dump(old, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
for f in ("bench_cfg1_bf16", "bench_cfg1_fp32", "bench_cfg2_bf16", "bench_cfg3_bf16", "bench_reference"):
    src = os.path.join(G, f + ".json")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "r01_" + f + ".json"))
        d = json.loads(open(src).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"]),
              " cpu", (d.get("cpu_baseline") or {}).get("value"), " x", d.get("speedup_e2e_vs_cpu"))
src = os.path.join(G, "launches_cfg1_bf16.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, "r01_launches_cfg1_bf16.csv"))
    agg = subprocess.run([sys.executable, os.path.join(P, "agg_launches.py"), src], capture_output=True, text=True).stdout
    open(os.path.join(P, "r01_launches_cfg1_bf16_by_kernel.txt"), "w").write(agg)
for f in os.listdir(G):
    if f.startswith("multi_") and f.endswith(".json"):
        shutil.copy(os.path.join(G, f), os.path.join(P, "r01_" + f))
