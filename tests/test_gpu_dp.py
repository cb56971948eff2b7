"""GPU, 2 devices: the data-parallel CUDA path (FusedTrainStep + BucketedGradSync over NCCL, captured in
a CUDA graph) reproduces the single-rank step of the same global batch - through `bench.py`'s own `dp`
leg, so the driver's N > 1 runs carry the same check (`dp.dp_parity`).  Skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_two_rank_data_parallel_step_equals_the_single_rank_step(precision, tol):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3",
           "--legs", "dp", "--dp-batch", "256", "--precision", precision]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    dp = line["dp"]
    assert dp["n_gpus"] == 2 and dp["batch_per_gpu"] == 128 and dp["cuda_graph"] is True
    assert dp["allreduce"]["collectives_per_step"] >= 6 + 3          # one per encoder layer + tail ranges + loss/count
    par = dp["dp_parity"]
    assert par["ranks_identical"] is True
    assert par["loss_rel"] < tol and par["update_rel"] < max(tol, 1e-4), par
