"""Copy the outputs of profiles/gpu_round3.sh (gpurun_out/r03/) into profiles/r02_* and rebuild traffic.json + the ncu
summary from the .ncu-rep captures:  python profiles/collect_evidence_r02.py"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out", "r03"), os.path.join(ROOT, "profiles")
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_active.avg', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor']
tob = lambda v, u: float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
traffic, out = {}, []
for f, title in (('prof_persist', 'python profiles/prof_rnn_layer.py bf16 lstm  (cfg1 layer: T=64 B=50 H=128, both directions)'),
                 ('prof_gemm_pair', 'python profiles/bench_gemm_pair.py inproj  (cfg4 hoisted projection [262144 x 1024] x [1024 x 4096], bf16, CTA pairs)'),
                 ('prof_step_fwd', 'python profiles/prof_step_pair.py 16 4096 512 1  (cfg4 recurrent step, forward: B 4096, H 512, both directions)'),
                 ('prof_step_bwd', 'python profiles/prof_step_pair.py 16 4096 512 1  (cfg4 recurrent step, backward)')):
    rep = os.path.join(G, f + '.ncu-rep')
    if not os.path.exists(rep):
        print("missing", rep)
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
open(os.path.join(P, 'r02_ncu_kernels_raw.txt'), 'w').write("\n".join(out) + "\n")
old = json.load(open(os.path.join(P, 'traffic.json')))
old.update(traffic)
json.dump(old, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
for f, dst in (("bench_cfg1", "r02_bench_cfg1.json"), ("bench_cfg2", "r02_bench_cfg2.json"), ("bench_cfg3", "r02_bench_cfg3.json"),
               ("bench_cfg4", "r02_bench_cfg4.json"), ("bench_reference", "r02_bench_reference.json")):
    shutil.copy(os.path.join(G, f + ".json"), os.path.join(P, dst))
    d = json.loads(open(os.path.join(G, f + ".json")).read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round((d.get("e2e") or {}).get("value") or 0))
for f, dst in (("warm_cfg1_bf16.txt", "r02_warm_cfg1_bf16.txt"), ("warm_cfg1_fp32.txt", "r02_warm_cfg1_fp32.txt"), ("warm_cfg4_bf16.txt", "r02_warm_cfg4_bf16.txt"),
               ("timeline_cfg1.txt", "r02_timeline_cfg1.txt"), ("gemm_pair.txt", "r02_gemm_pair.txt"),
               ("hbm_kernels.txt", "r02_hbm_kernels.txt"), ("launches_cfg1_bf16.csv", "r02_launches_cfg1_bf16.csv")):
    shutil.copy(os.path.join(G, f), os.path.join(P, dst))
agg = subprocess.run([sys.executable, os.path.join(P, "agg_launches.py"), os.path.join(G, "launches_cfg1_bf16.csv")], capture_output=True, text=True).stdout
open(os.path.join(P, "r02_launches_cfg1_bf16_by_kernel.txt"), "w").write(agg)
