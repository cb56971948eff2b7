#!/bin/bash
O=gpurun_out/r04g
mkdir -p $O
python profiles/bench_gemm_dw.py > $O/gemm_dw.txt 2>&1; tail -8 $O/gemm_dw.txt
python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=600 -x -k "gemm" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
B="python bench.py --steps 30 --warmup 5 --legs none --no-cpu-baseline"
for i in 1 2; do
SLNLP_GEMM_DEEP=0 $B > $O/bench_ring4_$i.json 2> $O/err.txt
$B > $O/bench_ring8_$i.json 2> $O/err.txt
done
python - <<Q
import json
for c in ("ring4_1", "ring8_1", "ring4_2", "ring8_2"):
    d = json.loads(open("$O/bench_%s.json" % c).read().strip().splitlines()[-1])
    print(c, round(d["value"]), d["unit"], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]))
Q
