"""Data-parallel training step for one large fit (BASELINE.json configs[3]; a capability the
reference does not have - SURVEY.md section 8e): every rank holds the full weights and runs B/N
sequences of each global batch; the flat fp32 gradient buffer is all-reduced over NCCL, then the
global-norm clip and SGD run identically on every rank.

The criterion is a MEAN over the valid labels (``ignore_index``), so per-rank gradients are means
over different counts n_r.  The exchange turns them back into sums (x n_r), all-reduces gradients
and counts, and divides by the global count N: the result is exactly the single-process gradient
of the global batch, whatever the split of ignored labels.

Two forms:
  * ``sync_gradients``: ONE all-reduce of the whole buffer after backward (any backend; the CPU
    gloo test and modules without gradient hooks);
  * ``BucketedGradSync``: the exchange SURVEY.md section 8e asks for - every encoder layer's gradient
    range is all-reduced as soon as that layer's BPTT and weight-gradient GEMMs are issued
    (``ready(lo, hi)`` is called by the module's backward), asynchronously on NCCL's stream, so
    the transfer of layer l overlaps the BPTT of layer l-1; ``finish`` joins them, exchanges
    whatever range was never announced and normalises by the global count.  Graph-capturable: the
    forks and joins become edges of the captured step.
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


class BucketedGradSync:
    """Per-range overlapped gradient exchange (see the module docstring).

    Protocol, driven by ``FusedTrainStep``:  ``begin(loss_and_count)`` right after the criterion
    kernel (the local valid count is known from then on; the 2-element loss/count exchange is
    started here), ``ready(gflat, lo, hi)`` whenever the range [lo, hi) of the flat gradient buffer
    is final, ``finish(gflat, loss_and_count)`` after backward."""

    bucketed = True

    def __init__(self, group=None):
        self.group = group
        self._works, self._covered = [], []
        self._n_local = self._tail = None
        self.n_collectives = 0          # gradient all-reduces issued by the last step (tests / bench)

    def _all_reduce(self, t):
        # async on CUDA (NCCL's own stream; the caller's stream keeps going), blocking on CPU backends
        if t.is_cuda:
            self._works.append(dist.all_reduce(t, group=self.group, async_op=True))
        else:
            dist.all_reduce(t, group=self.group)

    def begin(self, loss_and_count: torch.Tensor):
        self._works, self._covered = [], []
        self.n_collectives = 0
        self._n_local = loss_and_count[1].clone()
        self._tail = torch.stack([loss_and_count[0] * self._n_local, self._n_local])
        self._all_reduce(self._tail)

    def ready(self, gflat: torch.Tensor, lo: int, hi: int):
        if hi <= lo:
            return
        seg = gflat[lo:hi]
        seg.mul_(self._n_local)
        self._all_reduce(seg)
        self._covered.append((lo, hi))
        self.n_collectives += 1

    def finish(self, gflat: torch.Tensor, loss_and_count: torch.Tensor) -> torch.Tensor:
        pos = 0
        for lo, hi in sorted(self._covered) + [(gflat.numel(), gflat.numel())]:
            if lo > pos:                 # a range nobody announced: exchange it now
                self.ready(gflat, pos, lo)
            pos = max(pos, hi)
        for w in self._works:
            w.wait()
        self._works = []
        n_global = self._tail[1].clamp_min(1.0)
        gflat.div_(n_global)
        loss_and_count[0] = self._tail[0] / n_global
        loss_and_count[1] = self._tail[1]
        return gflat

    # a plain callable too (after a backward that announced nothing): one exchange of everything
    def __call__(self, gflat, loss_and_count):
        self.begin(loss_and_count)
        return self.finish(gflat, loss_and_count)


class DataParallelStep:
    """FusedTrainStep over this rank's slice of each global batch + the overlapped gradient exchange."""

    def __init__(self, module, local_batch: int, seq_len: int, lr: float, momentum: float = 0.9,
                 max_norm: float = 0.5, group=None, bucketed: bool = True, use_graph: bool = True):
        from .rnn import FusedTrainStep
        self.group = group
        self.sync = BucketedGradSync(group) if bucketed else (lambda g, loss: sync_gradients(g, loss, group))
        self.ts = FusedTrainStep(module, local_batch, seq_len, lr=lr, momentum=momentum, max_norm=max_norm,
                                 grad_sync=self.sync, use_graph=use_graph)
        self.ts.grad_scale = 1.0      # the exchange already normalises by the global count

    def step(self, X, y, lengths):
        return self.ts.step(X, y, lengths)

    def release(self):
        """Drop the captured step (it holds NCCL kernels) - call before destroy_process_group()."""
        self.ts.graph = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    @property
    def grad_norm(self):
        return self.ts.grad_norm
