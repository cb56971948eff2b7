"""One flat fp32 buffer behind a module's named parameters (shared by the RNN and the
Transformer host classes): the C-ABI kernels take raw device pointers into it, the fused
clip+SGD kernels walk it in one pass, and ``state_dict`` keeps the reference's names/shapes.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import contextlib
import torch
import torch.nn as nn

from ._lib import check, lib

import os as _os
_TC_GEMM = lib.slnlp_gemm_tf32
_F32_GEMM = lib.slnlp_gemm_tf32x3


def _stream():
    return torch.cuda.current_stream().cuda_stream


# torch.cuda.CUDAGraph registers every capture with the device's default RNG generator and unregisters
# it in its destructor; neither side is synchronised, so several fits that capture / drop step graphs
# from different threads of one process (grid.py fits_per_gpu) corrupt that registry ("The graph
# should be registered to the state", abort).  Captures and graph releases therefore take this lock;
# replays and ordinary launches of the other threads run freely meanwhile (thread-local capture mode).
import threading as _threading
CAPTURE_LOCK = _threading.RLock()


_tls = _threading.local()


def new_stream(priority=None):
    """A CUDA stream of its own (created through the C ABI, wrapped as an external stream).
    ``torch.cuda.Stream()`` draws from a pool of 32 that is handed out round-robin, so two fits of one
    process can end up sharing a stream - fatal when one of them is capturing on it.
    ``priority``: "high" / "low" = the device's extreme stream priorities (kept by captured graph nodes)."""
    ptr = lib.slnlp_stream_create() if priority is None else lib.slnlp_stream_create_priority(1 if priority == "high" else 0)
    if not ptr:
        check(1, "stream_create")
    return torch.cuda.ExternalStream(ptr, device=torch.cuda.current_device())


def thread_stream(role):
    """Per-thread, per-device auxiliary streams ("capture", "warmup", "side", "main"), created once."""
    key = (role, torch.cuda.current_device())
    cache = _tls.__dict__.setdefault("streams", {})
    if key not in cache:
        # "capturehi": the step's dependency chain captured on a high-priority stream (graph nodes keep the priority),
        # so that where its CTAs and those of the weight-gradient side lanes (default = lowest priority) are both
        # pending, the chain's are placed first ($SLNLP_STREAM_PRIO=0: off)
        prio = None
        if _os.environ.get("SLNLP_STREAM_PRIO", "1") != "0":
            prio = "high" if role == "capturehi" else None
        cache[key] = new_stream(prio)
    return cache[key]


@contextlib.contextmanager
def capture_graph(graph, high_priority=False):
    """Capture the CUDA work issued inside into ``graph`` (a torch.cuda.CUDAGraph), under CAPTURE_LOCK.

    ``torch.cuda.graph`` synchronises the whole device and empties the caching allocator before every
    capture; a fit captures three or four graphs (full batch, tail batch, scoring) and several fits may
    share the GPU, so this goes to capture_begin / capture_end directly on a private stream: no device-wide
    sync (invalid anyway while another thread's stream is capturing), no cudaFree storm.
    ``$SLNLP_CAPTURE_MODE`` = "thread_local" when several threads of the process capture (grid.py)."""
    mode = _os.environ.get("SLNLP_CAPTURE_MODE", "global")
    cur = torch.cuda.current_stream()
    # high_priority: H = 128 models, whose recurrent kernels leave SMs free for the weight-gradient GEMMs beside them
    # (cfg1 +2.3 %; where the recurrence fills the GPU it measured -1.9 %, cfg2)
    stream = thread_stream("capturehi" if high_priority else "capture")
    stream.wait_stream(cur)
    with CAPTURE_LOCK:
        with torch.cuda.stream(stream):
            graph.capture_begin(capture_error_mode=mode)
            try:
                yield
            finally:
                graph.capture_end()
    cur.wait_stream(stream)


def _align4(n):
    return (n + 3) & ~3


class _Box(nn.Module):
    """Plain container so parameters get the reference's dotted names."""


class FlatParamModule(nn.Module):
    """nn.Module whose parameters are views of ONE flat buffer (and whose grads are views of
    one flat gradient buffer).  Subclasses call ``_register_flat`` from ``__init__``."""

    precision = "fp32"
    _dead: Tuple[str, ...] = ()      # parameters that never receive a gradient (grad stays None)

    def _register_flat(self, segs: Sequence[Tuple[str, torch.Tensor]], order: Sequence[str],
                       glue_suffix: str = "\0"):
        """segs: (name, init tensor) in MEMORY order; order: names in named_parameters() order.
        A name ending in ``glue_suffix`` is laid out directly after its predecessor (no
        alignment gap) so that the pair forms one contiguous kernel operand."""
        self._names = [n for n, _ in segs]
        self._shapes = {n: tuple(t.shape) for n, t in segs}
        self._off: Dict[str, int] = {}
        off = 0
        for n, t in segs:
            if not n.endswith(glue_suffix):
                off = _align4(off)
            self._off[n] = off
            off += t.numel()
        self._numel = _align4(off)
        flat = torch.zeros(self._numel)
        for n, t in segs:
            flat[self._off[n]:self._off[n] + t.numel()] = t.detach().reshape(-1)
        self._flat = flat
        self._gflat = None
        self._params: Dict[str, nn.Parameter] = {}
        self._ws_cache: Dict = {}
        self._rng = None
        for n in order:
            parts = n.split(".")
            box = self
            for p in parts[:-1]:
                if not hasattr(box, p):
                    setattr(box, p, _Box())
                box = getattr(box, p)
            prm = nn.Parameter(self._view(self._flat, n))
            box.register_parameter(parts[-1], prm)
            self._params[n] = prm

    def _view(self, flat, name):
        shape = self._shapes[name]
        n = math.prod(shape)
        return flat[self._off[name]:self._off[name] + n].view(shape)

    def _apply(self, fn, recurse=True):
        # keep every parameter a view of ONE flat buffer across .to()/.cuda()/.float()
        self._flat = fn(self._flat)
        for n, prm in self._params.items():
            prm.data = self._view(self._flat, n)
            prm.grad = None
        for name, buf in list(self.named_buffers()):
            mod, _, leaf = name.rpartition(".")
            owner = self.get_submodule(mod) if mod else self
            owner._buffers[leaf] = fn(buf)
        self._gflat = None
        self._ws_cache = {}
        self._rng = None
        return self

    def to(self, device=None, *args, **kwargs):                  # bkp:383-386, transformer.py:50-58
        out = super().to(device, *args, **kwargs)
        if device is not None:
            self.device = torch.device(device) if not isinstance(device, torch.device) else device
        return out

    def _ensure_flat(self):
        """Re-flatten if some outside code replaced parameter storage."""
        base = self._flat.data_ptr()
        ok = all(p.data_ptr() == base + 4 * self._off[n] and p.device == self._flat.device
                 for n, p in self._params.items())
        if not ok:
            dev = next(iter(self._params.values())).device
            flat = torch.zeros(self._numel, device=dev)
            for n, p in self._params.items():
                flat[self._off[n]:self._off[n] + p.numel()] = p.data.reshape(-1).to(dev)
            self._flat = flat
            for n, p in self._params.items():
                p.data = self._view(flat, n)
            self._gflat = None
        if not self._flat.is_cuda:
            raise RuntimeError("slnlp_b200 modules compute on CUDA only (no CPU fallback): "
                               "call .to('cuda') first")

    def flat_parameters(self):
        self._ensure_flat()
        return self._flat

    def flat_grads(self):
        """Flat gradient buffer; ``p.grad`` of every live parameter is a view of it."""
        self._ensure_flat()
        if self._gflat is None or self._gflat.device != self._flat.device:
            self._gflat = torch.zeros_like(self._flat)
        for n, p in self._params.items():
            if n in self._dead:
                continue  # dead branch: grad stays None as in the reference
            want = self._view(self._gflat, n)
            if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                p.grad = want
        return self._gflat

    def _ptr(self, name, flat=None):
        flat = self._flat if flat is None else flat
        return flat.data_ptr() + 4 * self._off[name]

    def _rng_state(self):
        if self._rng is None or self._rng.device != self._flat.device:
            self._rng = torch.tensor([self.seed, 0], dtype=torch.int64, device=self._flat.device)
        return self._rng

    def _workspace(self, B, T, train, fresh=False, bwd=None):
        bwd = train if bwd is None else bwd
        key = (B, T, train, bwd)
        if not fresh and key in self._ws_cache:
            return self._ws_cache[key]
        ws = self._make_workspace(B, T, train, bwd)
        if not fresh:
            self._ws_cache[key] = ws
        return ws

    # ------------------------------------------------------------------ kernels
    def _gemm(self, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, bias=None, beta=0.0, big=False, act=None, drop=None):
        """C = op(A) op(B) (+ bias) (+ beta C) on the TMA-fed TF32 tensor-core kernel (fp32 path: the fp32-FMA
        GEMM), then optionally ``act`` ("tanh") and ``drop`` = (p, rng_ptr, site): the element-wise dropout
        slnlp_dropout(site) applies to C.  (A latency-built fp32-FMA "skinny" kernel with both epilogues and an
        in-kernel split-K reduction was measured for the 50-row products of the decoder side: 4 us SLOWER per
        call than the tensor-core kernel's 5.8 - two dependent L2 round trips for the slices cost what the TMA
        pipeline's prologue does - and was dropped; only fusing the chain into fewer kernels pays.)
        (`big` marks the [B*T]-row GEMMs; the library falls back to the fp32 kernel by itself for shapes a
        128-row tile cannot cover)"""
        # fp32 path: the same TMA / tcgen05 kernel with split operands (three tf32 MMAs per k-step, fp32-accurate:
        # slnlp_gemm_tf32x3); $SLNLP_F32_TC=0 keeps the fp32-FMA GEMM
        fn = _TC_GEMM if self.precision == "bf16" else (_F32_GEMM if self._f32_tc_ok() else lib.slnlp_gemm_f32)
        ws = self._gemm_ws()
        check(fn(tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, bias, beta, ws.data_ptr(), ws.numel(), _stream()), "gemm")
        if act == "tanh":
            assert ldc == N
            check(lib.slnlp_tanh_fwd(C, M * N, _stream()), "tanh")
        if drop is not None:
            assert ldc == N
            check(lib.slnlp_dropout(C, C, M * N, drop[0], drop[1], drop[2], _stream()), "dropout")

    # ---- CTA-pair bf16 GEMM (gemm_pair.cu): the [T*B]-row contractions at data-parallel batch sizes
    @staticmethod
    def _pair_ok(M, N, K):
        """Worth a 256 x 256 pair tile: at least ~two waves of tiles (or a long-K accumulation) and a K that
        amortises the accumulator drain."""
        if _os.environ.get("SLNLP_PAIR", "1") == "0" or M % 8 or N % 8 or K % 8:
            return False
        tiles = ((M + 255) // 256) * ((N + 255) // 256)
        return bool(lib.slnlp_gemm_bf16_supported(0, 0, M, N, K)) and K >= 256 and (tiles >= 148 or K >= 16384)

    def _cast_bf16(self, src_ptr, lds, dst, rows, cols, ldd=None):
        check(lib.slnlp_cast_bf16(src_ptr, lds, dst.data_ptr(), cols if ldd is None else ldd, rows, cols, 0, _stream()),
              "cast_bf16")

    def _gemm_bf16(self, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, bias=None, beta=0.0):
        check(lib.slnlp_gemm_bf16(tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, bias, beta, _stream()), "gemm_bf16")

    f32_tensor_cores = True      # class default: the fp32 path's GEMMs run as split-operand tf32 MMAs

    def _f32_tc_ok(self):
        env = _os.environ.get("SLNLP_F32_TC")
        return self.f32_tensor_cores if env is None else env != "0"

    def _gemm_ws(self):
        """Split-K scratch of the GEMMs: one buffer per (thread, device, stream) - kernels on one stream are ordered, so
        the modules a thread steps one after another share it (the main stream and the weight-gradient side lanes never
        share partials).  Per MODULE it cost a grid-search fit ~9 fresh 10 MB allocations: 19 ms of a 71 ms fit at
        emb 128 / hidden 128 / 2 layers, 130 of 254 ms at 512 / 256 / 4 (profiles/prof_grid_fit.py)."""
        key = (self._flat.device, torch.cuda.current_stream().cuda_stream)
        pool = _tls.__dict__.setdefault("gemm_scratch", {})
        ws = pool.get(key)
        if ws is None:
            ws = pool[key] = torch.empty(lib.slnlp_gemm_workspace_floats(), device=self._flat.device)
        return ws

    # ---- weight-gradient side lanes: dW / db kernels are not on the dependency chain of BPTT, so
    # they run on other streams (parallel branches of the captured graph) next to the next layer's
    # recurrent kernel.  Several lanes: the weight-gradient products of one layer are independent of each
    # other too, and each is a few CTAs of mostly fixed latency - side by side they cost one GEMM, in a
    # row (the last layer's, with nothing left to hide behind) they cost four
    N_LANES = 8

    def _side_stream(self, lane=0):
        with torch.cuda.device(self._flat.device):
            side = thread_stream("side" if lane == 0 else f"side{lane}")
        if lane == 0:
            self._side = side
        return side

    def _fork_side(self, lane=0):
        side = self._side_stream(lane)
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        self.__dict__.setdefault("_lanes_pending", set()).add(lane)
        return side

    @contextlib.contextmanager
    def _side_branch(self, lane=0):
        """Kernels launched inside run on side lane ``lane``, ordered after everything issued so far on
        the current stream (and after that lane's earlier work).  Only while a CUDA graph is being
        captured (the fork / join are free graph edges there); an eager step is bound by the host's
        launch rate and the event traffic would slow it down."""
        if not torch.cuda.is_current_stream_capturing():
            yield
            return
        with torch.cuda.stream(self._fork_side(lane % self.N_LANES)):
            yield

    def _join_lane(self, lane):
        """The current stream waits for ONE side lane (an activation branch the chain needs back)."""
        pending = self.__dict__.get("_lanes_pending")
        if not pending or lane not in pending:
            return
        ev = torch.cuda.Event()
        ev.record(self._side_stream(lane))
        torch.cuda.current_stream().wait_event(ev)
        pending.discard(lane)

    def _join_side(self):
        pending = self.__dict__.get("_lanes_pending")
        if not pending:
            return   # nothing forked since the last join (and, under graph capture, no branch to merge)
        for lane in sorted(pending):
            ev = torch.cuda.Event()
            ev.record(self._side_stream(lane))
            torch.cuda.current_stream().wait_event(ev)
        pending.clear()
