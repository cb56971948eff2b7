"""Data-parallel training step for one large fit (BASELINE.json configs[3]; a capability the
reference does not have - SURVEY.md section 8e): every rank holds the full weights and runs B/N
sequences of each global batch; ONE NCCL all-reduce of the flat fp32 gradient buffer per step,
then the global-norm clip and SGD run identically on every rank.

The criterion is a MEAN over the valid labels (``ignore_index``), so per-rank gradients are means
over different counts n_r.  ``sync_gradients`` first turns them back into sums (x n_r), all-reduces
gradients and counts together, and divides by the global count N: the result is exactly the
single-process gradient of the global batch, whatever the split of ignored labels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def sync_gradients(gflat: torch.Tensor, loss_and_count: torch.Tensor, group=None) -> torch.Tensor:
    """In place: gflat <- gradient of the GLOBAL mean loss; loss_and_count [2] = (local mean loss,
    local valid count) <- (global mean loss, global count).  Returns gflat.  Works on any backend
    (NCCL on the GPUs; gloo in the CPU test)."""
    n_local = loss_and_count[1].clone()
    gflat.mul_(n_local)
    loss_and_count[0].mul_(n_local)
    # one collective for gradients, loss numerator and count: they travel as one flat buffer
    tail = loss_and_count.to(gflat.dtype)
    if gflat.is_cuda and not torch.cuda.is_current_stream_capturing():
        work = dist.all_reduce(gflat, group=group, async_op=True)
        dist.all_reduce(tail, group=group)
        work.wait()
    else:
        dist.all_reduce(gflat, group=group)
        dist.all_reduce(tail, group=group)
    n_global = tail[1].clamp_min(1.0)
    gflat.div_(n_global)
    loss_and_count[0] = tail[0] / n_global
    loss_and_count[1] = tail[1]
    return gflat


class DataParallelStep:
    """FusedTrainStep over this rank's slice of each global batch + the gradient all-reduce."""

    def __init__(self, module, local_batch: int, seq_len: int, lr: float, momentum: float = 0.9,
                 max_norm: float = 0.5, group=None):
        from .rnn import FusedTrainStep
        self.group = group
        self.ts = FusedTrainStep(module, local_batch, seq_len, lr=lr, momentum=momentum, max_norm=max_norm,
                                 grad_sync=lambda g, loss: sync_gradients(g, loss, group))
        self.ts.grad_scale = 1.0      # sync_gradients already normalises by the global count

    def step(self, X, y, lengths):
        return self.ts.step(X, y, lengths)

    @property
    def grad_norm(self):
        return self.ts.grad_norm
