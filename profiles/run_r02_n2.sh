#!/bin/bash
# round 2, two GPUs: the driver's N = 2 command (replicas + dp + grid sub-records) and the 2-rank DP parity test
mkdir -p gpurun_out/r02n2
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02n2/bench_n2.json 2> gpurun_out/r02n2/bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r02n2/bench_n2.json; grep -v "Warn\|warn" gpurun_out/r02n2/bench_n2.err | tail -8
( time timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x --timeout=900 ) > gpurun_out/r02n2/pytest_dp.log 2>&1; echo "pytest dp rc=$?"; tail -5 gpurun_out/r02n2/pytest_dp.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 5 --warmup 2 ) > gpurun_out/r02n2/ref_n2.json 2> gpurun_out/r02n2/ref_n2.err; echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/r02n2/ref_n2.json
