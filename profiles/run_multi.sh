#!/bin/bash
# 2-GPU checks: (1) bench default at N=2 (independent fits, no collective); (2) data-parallel
# NCCL all-reduce (cfg1 model, global batch 100 -> 50 per rank; and cfg4 model at global batch 256);
# (3) main.py grid search farmed over 2 GPUs (spawn backend); (4) reference arm under torchrun.
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/multi_fits_n$N.json 2> gpurun_out/multi_fits_n$N.err; echo "fits rc=$?"; tail -c 400 gpurun_out/multi_fits_n$N.json | head -c 400; echo
$TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --dp --batch 100 > gpurun_out/multi_dp_cfg1_n$N.json 2> gpurun_out/multi_dp_cfg1_n$N.err; echo "dp cfg1 rc=$?"; head -c 300 gpurun_out/multi_dp_cfg1_n$N.json; echo; tail -3 gpurun_out/multi_dp_cfg1_n$N.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --dp --workload cfg4 --batch 256 > gpurun_out/multi_dp_cfg4_n$N.json 2> gpurun_out/multi_dp_cfg4_n$N.err; echo "dp cfg4 rc=$?"; head -c 300 gpurun_out/multi_dp_cfg4_n$N.json; echo; tail -3 gpurun_out/multi_dp_cfg4_n$N.err
cd sign-language-nlp_b200 && python main.py --config config/b200-lstm-attn.yaml --workdir /tmp/grid_run --gpus $N --max_epochs 3 --verbose 1 --dataset_args "{synthetic: {n_seq: 600, T: 32, v_src: 500, v_tgt: 12}}" --grid_args "{lr: [0.1, 0.01], model_args: {embedding_size: [128], hidden_size: [128], num_layers: [1, 2], dropout: [0.1]}}" > ../gpurun_out/grid_run.log 2>&1; echo "main rc=$?"; cd ..; grep -E "grid\]|Worker farm|fits_per_hour|test_accuracy" gpurun_out/grid_run.log | tail -8
$TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/multi_ref_n$N.json 2> gpurun_out/multi_ref_n$N.err; echo "ref rc=$?"; head -c 300 gpurun_out/multi_ref_n$N.json; echo
$TR bench.py --gpus $N --workload cfg5 --grid-fraction 0.2 > gpurun_out/multi_grid_n$N.json 2> gpurun_out/multi_grid_n$N.err; echo "grid rc=$?"; head -c 700 gpurun_out/multi_grid_n$N.json; echo
