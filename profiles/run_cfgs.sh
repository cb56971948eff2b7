python -m pytest tests/test_gpu_estimator.py -m gpu -q 2>&1 | tail -1
for wl in cfg2 cfg3; do for pr in fp32 bf16; do
python bench.py --workload $wl --precision $pr --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b_${wl}_${pr}.err > gpurun_out/b_${wl}_${pr}.json; echo "$wl $pr rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/b_${wl}_${pr}.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['launches_per_step'],d['e2e']['value'])"
done; done
timeout 600 python bench.py --workload cfg4 --batch 512 --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b_cfg4.err > gpurun_out/b_cfg4.json; echo "cfg4 rc=$?"; tail -c 600 gpurun_out/b_cfg4.json; tail -3 gpurun_out/b_cfg4.err
