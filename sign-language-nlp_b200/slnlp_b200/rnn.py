"""B200 host side of the RNN encoder-decoder-attention classifiers.

Mirrors ``EncoderDecoderAttnBaseBkp`` of the reference
(model/base/encoder_decoder_attn_bkp.py:330-413, "bkp" below): same constructor
keywords, ``.to(device)``, ``forward(X, y, lengths) -> [B, V_tgt]`` log-probs, same
``state_dict`` names and shapes, same default initialisation stream (so the same
``torch.manual_seed`` gives the same initial weights).  All model math runs in the
hand-written sm_100a kernels behind the C ABI (include/slnlp_b200.h); PyTorch only
owns memory, streams and autograd bookkeeping.  CPU tensors raise - no fallback.

Two ways in:
  * ``module(X=..., y=..., lengths=...)``: an autograd.Function, so stock skorch /
    ``loss.backward()`` / ``clip_grad_norm_`` / ``torch.optim.SGD`` work unchanged;
  * ``FusedTrainStep``: the whole skorch train step (forward, CE on the log-probs,
    backward, global-norm clip 0.5, SGD-momentum) as one CUDA graph replay.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import List

import torch
import torch.nn as nn

from ._lib import RnnExtras, check, lib

PAD_WORD, BOS_WORD = "<pad>", "<bos>"  # dataset/constant/tokens.py:1,4
MODE = {"lstm": 0, "gru": 1}
GATES = {"lstm": 4, "gru": 3}


from .flat import CAPTURE_LOCK, FlatParamModule, _Box, _align4, _stream, capture_graph, thread_stream  # noqa: F401


class RnnEncDecB200(FlatParamModule):
    MAX_OUTPUT_LEN = 1  # bkp:332
    accepts_init_generator = True     # __init__(init_generator=torch.Generator): see _build_parameters
    _joins_rng_lane = True   # _run_forward joins side lane 3 (rng advance + dropout factors) before the first RNN layer
    _dead = ("model.decoder.pre_output_layer.weight",)

    def __init__(self, src_vocab, tgt_vocab, batch_first, rnn_type, embedding_size=256,
                 hidden_size=512, num_layers=1, dropout=0.1, precision="fp32", **kwargs):
        super().__init__()
        assert rnn_type in MODE, "Invalid `rnn_type`."          # bkp:347
        assert precision in ("fp32", "bf16")
        self.batch_first = batch_first
        self.src_vocab, self.tgt_vocab = src_vocab, tgt_vocab
        self.rnn_type, self.precision = rnn_type, precision
        self.E, self.H, self.L = int(embedding_size), int(hidden_size), int(num_layers)
        self.G = GATES[rnn_type]
        self.p_drop = float(dropout)
        self.p_rnn = float(dropout) if self.L > 1 else 0.0       # bkp:100,190
        self.src_pad = src_vocab.stoi[PAD_WORD]                  # model/util/util.py:5-6
        self.tgt_pad = tgt_vocab.stoi[PAD_WORD]
        self.bos_idx = tgt_vocab.stoi[BOS_WORD]                  # util.py:8-9 (-> unk = 0)
        self.V_src, self.V_tgt = len(src_vocab), len(tgt_vocab)
        # factored phonological embedding (a variant the reference does not have - it
        # joins the six fields into one token before embedding): ``src_field_vocab_sizes`` = rows of one table
        # per field, ``src_field_widths`` = their widths (sum = embedding_size; default: an even split).  X is
        # then [B, T, F]; the F gathers and the concat are ONE kernel (K1); a frame is padding when its FIRST
        # field is <pad>.  state_dict: model.src_embed.fields.{i}.weight.
        fv = kwargs.get("src_field_vocab_sizes", None)
        self.field_rows = [int(v) for v in fv] if fv else None
        if self.field_rows:
            F = len(self.field_rows)
            fw = kwargs.get("src_field_widths", None)
            self.field_widths = [int(w) for w in fw] if fw else [self.E // F + (1 if i < self.E % F else 0) for i in range(F)]
            assert len(self.field_widths) == F and sum(self.field_widths) == self.E, "field widths must sum to embedding_size"
            assert all(w % 4 == 0 for w in self.field_widths), "field widths must be multiples of 4 (16-byte rows)"
        self.device = kwargs.get("device", None)
        self.validate_inputs = True
        # weight-gradient kernels on a side stream, next to the next layer's BPTT.  It pays when the
        # recurrent kernel leaves most SMs idle (H = 128: 26 CTAs; measured +3 % on cfg1) and not when the
        # cluster kernels fill the GPU (H = 256: -1 % on cfg2).  SLNLP_OVERLAP_DW=0/1 overrides.
        env = os.environ.get("SLNLP_OVERLAP_DW")
        self.overlap_dw = (self.H == 128) if env is None else env != "0"
        # the small decoder / generator / attention / bridge weight gradients always leave the chain
        self.overlap_small = os.environ.get("SLNLP_OVERLAP_SMALL", "1") != "0"
        self._build_parameters(kwargs.get("init_generator", None))
        self.seed = int(kwargs.get("seed", torch.initial_seed() & 0x7FFFFFFF))

    # ------------------------------------------------------------------ parameters
    def _build_parameters(self, gen=None):
        E, H, L, G = self.E, self.H, self.L, self.G
        # default initialisers drawn in the reference's construction order (bkp:362-381): Encoder.rnn, attention (key,
        # query, energy), Decoder.rnn, bridge, pre_output, src Embedding, trg Embedding, Generator - the draws torch.nn's
        # reset_parameters() make (RNNBase: uniform(+-1/sqrt(H)) over its parameters in registration order; Linear:
        # kaiming_uniform(a = sqrt 5) then the bias; Embedding: normal, padding row zeroed), on plain tensors.  `gen` =
        # None draws from the process-global CPU generator like torch.nn would (same torch.manual_seed -> the
        # reference's initial weights, tests/test_host_logic.py); the grid farm passes a PRIVATE generator per fit, so
        # that concurrent fits neither share nor serialise on the global one.
        import math
        from types import SimpleNamespace as NS
        init = nn.init

        def rnn(D_in, bidirectional):
            ns, stdv = NS(), (1.0 / math.sqrt(H) if H > 0 else 0)
            for l in range(L):
                d_in = D_in if l == 0 else H * (2 if bidirectional else 1)
                for suf in (("", "_reverse") if bidirectional else ("",)):
                    for kind, shape in (("weight_ih", (G * H, d_in)), ("weight_hh", (G * H, H)), ("bias_ih", (G * H,)),
                                        ("bias_hh", (G * H,))):
                        setattr(ns, f"{kind}_l{l}{suf}", init.uniform_(torch.empty(*shape), -stdv, stdv, generator=gen))
            return ns

        def linear(d_in, d_out, bias):
            ns = NS(weight=init.kaiming_uniform_(torch.empty(d_out, d_in), a=math.sqrt(5), generator=gen), bias=None)
            if bias:
                bound = 1 / math.sqrt(d_in) if d_in > 0 else 0
                ns.bias = init.uniform_(torch.empty(d_out), -bound, bound, generator=gen)
            return ns

        def embedding(rows, width, padding_idx):
            w = init.normal_(torch.empty(rows, width), generator=gen)
            if padding_idx is not None:
                w[padding_idx].fill_(0)
            return NS(weight=w)

        t_enc = rnn(E, True)
        t_key = linear(2 * H, H, False)
        t_query = linear(H, H, False)
        t_energy = linear(H, 1, False)
        t_dec = rnn(E + 2 * H, False)
        t_bridge = linear(2 * H, H, True)
        t_pre = linear(3 * H + E, H, False)
        if self.field_rows:
            t_src_fields = [embedding(v, w, self.src_pad) for v, w in zip(self.field_rows, self.field_widths)]
        else:
            t_src = embedding(self.V_src, E, self.src_pad)
        t_trg = embedding(self.V_tgt, E, self.tgt_pad)
        t_gen = linear(H, self.V_tgt, False)

        # flat layout: per layer the two directions of each tensor are adjacent
        segs: List = []  # (name, init tensor)
        for l in range(L):
            for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                for suf in ("", "_reverse"):
                    segs.append((f"model.encoder.rnn.{kind}_l{l}{suf}", getattr(t_enc, f"{kind}_l{l}{suf}")))
        segs += [("model.decoder.attention.key_layer.weight", t_key.weight),
                 ("model.decoder.attention.query_layer.weight", t_query.weight),
                 ("model.decoder.attention.energy_layer.weight", t_energy.weight)]
        for l in range(L):
            for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                segs.append((f"model.decoder.rnn.{kind}_l{l}", getattr(t_dec, f"{kind}_l{l}")))
        segs += [("model.decoder.bridge.weight", t_bridge.weight),
                 ("model.decoder.bridge.bias", t_bridge.bias),
                 ("model.decoder.pre_output_layer.weight", t_pre.weight)]
        src_names = ([f"model.src_embed.fields.{i}.weight" for i in range(len(self.field_rows))] if self.field_rows
                     else ["model.src_embed.weight"])
        segs += ([(n, t.weight) for n, t in zip(src_names, t_src_fields)] if self.field_rows
                 else [("model.src_embed.weight", t_src.weight)])
        segs += [("model.trg_embed.weight", t_trg.weight),
                 ("model.generator.proj.weight", t_gen.weight)]
        # registration order = the reference's named_parameters() order
        order = []
        for l in range(L):
            for suf in ("", "_reverse"):
                for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    order.append(f"model.encoder.rnn.{kind}_l{l}{suf}")
        order += [n for n, _ in segs if n.startswith("model.decoder.attention")]
        order += [n for n, _ in segs if n.startswith("model.decoder.rnn")]
        order += ["model.decoder.bridge.weight", "model.decoder.bridge.bias",
                  "model.decoder.pre_output_layer.weight", *src_names,
                  "model.trg_embed.weight", "model.generator.proj.weight"]
        self._src_names = src_names
        self._register_flat(segs, order, glue_suffix="_reverse")

    def _make_workspace(self, B, T, train, bwd):
        return _Workspace(self, B, T, train, bwd)

    @property
    def uses_rng(self):
        return self.p_rnn > 0.0

    def _run_forward(self, ws, X, lengths, y=None):
        """X [B,T] int64 cuda, lengths [B] int64 cuda.  Fills ws; returns ws.logp."""
        E, H, L, G = self.E, self.H, self.L, self.G
        B, T = ws.B, ws.T
        s = _stream()
        mode, prec = MODE[self.rnn_type], (1 if self.precision == "bf16" else 0)
        lp = lengths.data_ptr()
        check(lib.slnlp_embed_gather_fwd(self._ptr(self._src_names[0]), X.data_ptr(), ws.emb.data_ptr(), B, T, ws.F,
                                         ws.f_off, ws.f_w_src, ws.f_rows_src, 1, 1.0, None, s), "embed")
        if ws.F > 1:      # the padding mask of the attention: the first field of every frame
            ws.xmask.copy_(X[..., 0])
            Xp = ws.xmask.data_ptr()
        else:
            Xp = X.data_ptr()
        drop = ws.train and self.p_rnn > 0.0
        rng = self._rng_state().data_ptr() if drop else None
        if drop and ws.rnn_fused_dropout:
            # the inter-layer dropout factors of this step as tensors, generated on a side lane (under capture)
            # while the embedding and the first projection run: the persistent kernels then apply them with
            # one load per step instead of ten dependent Philox rounds
            with self._side_branch(3):
                for l in range(L - 1):
                    check(lib.slnlp_dropout_mask(ws.drop_mask[l].data_ptr(), ws.drop_mask[l].numel(), self.p_rnn, rng, l,
                                                 _stream()), "dropout_mask")
        for l in range(L):
            D = E if l == 0 else 2 * H
            xin = ws.emb if l == 0 else ws.enc_xin[l]
            pre = f"model.encoder.rnn."
            if ws.pair[l]:
                # data-parallel batch sizes: bf16 copies of the layer input (kept for dW_ih) and of W_ih (serves
                # the forward as a K-major and dX as an MN-major operand), CTA-pair tcgen05 GEMM
                if not ws.xin_bf_ready[l]:     # else: the layer below (or its dropout) already wrote it as bf16
                    self._cast_bf16(xin.data_ptr(), D, ws.xin_bf[l], T * B, D)
                self._cast_bf16(self._ptr(f"{pre}weight_ih_l{l}"), D, ws.w_ih_bf[l], 2 * G * H, D)
                self._gemm_bf16(0, 1, T * B, 2 * G * H, D, ws.xin_bf[l].data_ptr(), D, ws.w_ih_bf[l].data_ptr(), D,
                                ws.enc_gates[l].data_ptr(), 2 * G * H, self._ptr(f"{pre}bias_ih_l{l}"))
            else:
                self._gemm(0, 1, T * B, 2 * G * H, D, xin.data_ptr(), D, self._ptr(f"{pre}weight_ih_l{l}"), D,
                           ws.enc_gates[l].data_ptr(), 2 * G * H, self._ptr(f"{pre}bias_ih_l{l}"), 0.0, big=True)
            if l == 0:
                self._join_lane(3)        # rng advanced (FusedTrainStep) and dropout factors ready
            if ws.rnn_extras:
                # persistent kernels: final states straight into the concatenated layout, the next layer's
                # dropped input from the same deferred store (no concat / dropout launches)
                to_drop = l < L - 1 and drop and ws.rnn_fused_dropout
                ex = RnnExtras(1, ws.enc_xin[l + 1].data_ptr() if to_drop else None, self.p_rnn, rng if to_drop else None, l, 0,
                               ws.drop_mask[l].data_ptr() if to_drop else None)
                check(lib.slnlp_rnn_layer_fwd_ex(mode, prec, T, B, H, 2, ws.enc_gates[l].data_ptr(),
                                                 self._ptr(f"{pre}weight_hh_l{l}"), self._ptr(f"{pre}bias_hh_l{l}"),
                                                 lp, None, None, ws.enc_out[l].data_ptr(), ws.enc_stash[l].data_ptr(),
                                                 ws.enc_final[l].data_ptr(), ctypes.byref(ex), s), "rnn_layer_fwd")
                if l < L - 1 and drop and not to_drop:
                    check(lib.slnlp_dropout(ws.enc_out[l].data_ptr(), ws.enc_xin[l + 1].data_ptr(),
                                            ws.enc_out[l].numel(), self.p_rnn, rng, l, s), "dropout")
                continue
            if ws.bf_step:
                # large batch: the per-step kernels read bf16 operands (h_{t-1} from the bf16 copy of `out` they
                # write themselves, W_hh from a bf16 copy) - the recurrence there is bound by operand bytes
                self._cast_bf16(self._ptr(f"{pre}weight_hh_l{l}"), H, ws.w_hh_bf[l], 2 * G * H, H)
                # (gates_bf: the activated gates are stashed as bf16 in the buffer BPTT later overwrites with dG)
                check(lib.slnlp_rnn_layer_fwd_bf16_ex(mode, T, B, H, 2, ws.enc_gates[l].data_ptr(), ws.w_hh_bf[l].data_ptr(),
                                                      self._ptr(f"{pre}bias_hh_l{l}"), lp,
                                                      None if ws.skip_out32[l] else ws.enc_out[l].data_ptr(),
                                                      ws.out_bf[l].data_ptr(), ws.enc_stash[l].data_ptr(),
                                                      ws.enc_hfin[l].data_ptr(),
                                                      ws.dg_bf[l].data_ptr() if ws.gates_bf[l] else None, s), "rnn_layer_fwd_bf16")
            else:
                check(lib.slnlp_rnn_layer_fwd(mode, prec, T, B, H, 2, ws.enc_gates[l].data_ptr(),
                                              self._ptr(f"{pre}weight_hh_l{l}"), self._ptr(f"{pre}bias_hh_l{l}"),
                                              lp, None, None, ws.enc_out[l].data_ptr(), ws.enc_stash[l].data_ptr(),
                                              ws.enc_hfin[l].data_ptr(), s), "rnn_layer_fwd")
            if l < L - 1 and drop:
                if ws.masked[l]:       # bf16 in, bf16 out, and the keep mask as bits for this layer's BPTT
                    check(lib.slnlp_dropout_bf16_masked(ws.out_bf[l].data_ptr(), ws.xin_bf[l + 1].data_ptr(),
                                                        ws.keep_bits[l].data_ptr(), ws.out_bf[l].numel(), self.p_rnn, rng, l, s),
                          "dropout_bf16_masked")
                elif ws.bf_in[l + 1]:    # every consumer of the next layer's input reads bf16: one pass, no fp32 copy
                    check(lib.slnlp_dropout_bf16(ws.enc_out[l].data_ptr(), ws.xin_bf[l + 1].data_ptr(),
                                                 ws.enc_out[l].numel(), self.p_rnn, rng, l, s), "dropout_bf16")
                else:
                    check(lib.slnlp_dropout(ws.enc_out[l].data_ptr(), ws.enc_xin[l + 1].data_ptr(),
                                            ws.enc_out[l].numel(), self.p_rnn, rng, l, s), "dropout")
            check(lib.slnlp_concat_dirs(ws.enc_hfin[l].data_ptr(), ws.enc_final[l].data_ptr(), B, H, 2, 0, s),
                  "concat_dirs")
        # pad_packed_sequence(padding_value=<pad>) (bkp:121-123): a pad-FILLED copy for the key projection and
        # the attention; BPTT and dW_hh keep reading the zero-padded original (no un-fill on the way back)
        enc_out = ws.enc_filled
        # the pad-filled copy and the key projection (bkp:246) need only the encoder output: under capture they
        # run on a side lane next to bridge -> tanh -> query, which need only the final states
        with self._side_branch(2):
            check(lib.slnlp_pad_fill_copy(ws.enc_out[L - 1].data_ptr(), enc_out.data_ptr(), lp, T, B, 2 * H,
                                          float(self.src_pad), _stream()), "pad_fill")
            self._gemm(0, 1, T * B, H, 2 * H, enc_out.data_ptr(), 2 * H,
                       self._ptr("model.decoder.attention.key_layer.weight"), 2 * H, ws.pk.data_ptr(), H, big=True)
        # bridge (bkp:268-280)
        if not ws.fuse_bridge:
            self._gemm(0, 1, L * B, H, 2 * H, ws.enc_final.data_ptr(), 2 * H, self._ptr("model.decoder.bridge.weight"),
                       2 * H, ws.hidden0.data_ptr(), H, self._ptr("model.decoder.bridge.bias"), act="tanh")
        # attention (bkp:304-327)
        if not ws.fuse_query:
            self._gemm(0, 1, B, H, H, ws.hidden0[L - 1].data_ptr(), H,
                       self._ptr("model.decoder.attention.query_layer.weight"), H, ws.q.data_ptr(), H)
        self._join_lane(2)
        bos_row = self._ptr("model.trg_embed.weight") + 4 * self.bos_idx * E
        if ws.dec_head:
            # bridge -> tanh -> query -> attention -> [bos embedding || context]: one launch, one CTA per sequence
            check(lib.slnlp_dec_head_fwd(T, B, H, L, E, ws.enc_final.data_ptr(),
                                         self._ptr("model.decoder.bridge.weight") if ws.fuse_bridge else None,
                                         self._ptr("model.decoder.bridge.bias"),
                                         self._ptr("model.decoder.attention.query_layer.weight") if ws.fuse_query else None,
                                         ws.pk.data_ptr(), self._ptr("model.decoder.attention.energy_layer.weight"),
                                         enc_out.data_ptr(), Xp, self.src_pad, bos_row, ws.hidden0.data_ptr(),
                                         ws.q.data_ptr(), ws.alpha.data_ptr(), ws.ctx.data_ptr(), ws.dec_xin[0].data_ptr(),
                                         s), "dec_head_fwd")
        else:
            check(lib.slnlp_attn_step_fwd(ws.q.data_ptr(), ws.pk.data_ptr(),
                                          self._ptr("model.decoder.attention.energy_layer.weight"),
                                          enc_out.data_ptr(), Xp, self.src_pad, T, B, H, 2 * H,
                                          ws.alpha.data_ptr(), ws.ctx.data_ptr(), s), "attn_fwd")
        # one decoder step (bkp:215-216): at the reference's batch its cell stays on the fused fp32 kernel (a single
        # step of 50 sequences gains nothing from the tensor-core kernels, measured); batches > 256 (or
        # SLNLP_DEC_TC=1) take the tensor-core GEMM + step kernels (the skinny fp32 cell was 1.6 ms per layer at 4096)
        dec_prec = prec if (os.environ.get("SLNLP_DEC_TC", "0") == "1" or B > 256) else 0
        if not ws.dec_head:
            check(lib.slnlp_dec_input_fwd(bos_row, ws.ctx.data_ptr(), ws.dec_xin[0].data_ptr(), B, E, 2 * H, s), "dec_input")
        fused_cell = os.environ.get("SLNLP_DEC_FUSED", "1") != "0" and dec_prec == 0
        for l in range(L):
            D = E + 2 * H if l == 0 else H
            pre = "model.decoder.rnn."
            h0 = ws.hidden0[l].data_ptr()
            if fused_cell:
                # projection + recurrent product + cell + inter-layer dropout: one launch (dec_cell.cu)
                to_drop = l < L - 1 and drop
                check(lib.slnlp_dec_cell_fwd(mode, B, H, D, ws.dec_xin[l].data_ptr(), h0, h0 if mode == 0 else None,
                                             self._ptr(f"{pre}weight_ih_l{l}"), self._ptr(f"{pre}weight_hh_l{l}"),
                                             self._ptr(f"{pre}bias_ih_l{l}"), self._ptr(f"{pre}bias_hh_l{l}"),
                                             ws.dec_gates[l].data_ptr(), ws.dec_stash[l].data_ptr(), ws.dec_h[l].data_ptr(),
                                             ws.dec_xin[l + 1].data_ptr() if to_drop else None, self.p_rnn,
                                             rng if to_drop else None, 100 + l, s), "dec_cell_fwd")
                continue
            self._gemm(0, 1, B, G * H, D, ws.dec_xin[l].data_ptr(), D, self._ptr(f"{pre}weight_ih_l{l}"), D,
                       ws.dec_gates[l].data_ptr(), G * H, self._ptr(f"{pre}bias_ih_l{l}"))
            check(lib.slnlp_rnn_layer_fwd(mode, dec_prec, 1, B, H, 1, ws.dec_gates[l].data_ptr(),
                                          self._ptr(f"{pre}weight_hh_l{l}"), self._ptr(f"{pre}bias_hh_l{l}"),
                                          None, h0, h0 if mode == 0 else None, ws.dec_h[l].data_ptr(),
                                          ws.dec_stash[l].data_ptr(), None, s), "rnn_layer_fwd(dec)")
            if l < L - 1 and drop:
                check(lib.slnlp_dropout(ws.dec_h[l].data_ptr(), ws.dec_xin[l + 1].data_ptr(), B * H,
                                        self.p_rnn, rng, 100 + l, s), "dropout")
        # generator (bkp:73-76) on decoder_states, not pre_output (bkp:40-46)
        self._gemm(0, 1, B, self.V_tgt, H, ws.dec_h[L - 1].data_ptr(), H,
                   self._ptr("model.generator.proj.weight"), H, ws.logits.data_ptr(), self.V_tgt)
        if not getattr(ws, "fused_ce", False):   # the fused train step folds it into the criterion kernel
            check(lib.slnlp_log_softmax_fwd(ws.logits.data_ptr(), ws.logp.data_ptr(), B, self.V_tgt, s), "log_softmax")
        return ws.logp

    def _run_backward(self, ws, X, lengths, gflat, y=None):
        """Consumes ws.dlogits; accumulates parameter gradients into ``gflat``."""
        E, H, L, G = self.E, self.H, self.L, self.G
        B, T, V = ws.B, ws.T, self.V_tgt
        s = _stream()
        mode, prec = MODE[self.rnn_type], (1 if self.precision == "bf16" else 0)
        lp = lengths.data_ptr()
        Xp = ws.xmask.data_ptr() if ws.F > 1 else X.data_ptr()      # ws.xmask: filled by this step's forward
        gp = lambda n: self._ptr(n, gflat)
        drop = ws.train and self.p_rnn > 0.0
        rng = self._rng_state().data_ptr() if drop else None
        GH = G * H
        # Weight gradients are leaves of the dependency graph: each is issued on the side stream as
        # soon as its operands exist, so that the main stream carries only the d(activation) chain.
        # data parallel (dp.py): every finished range of the flat gradient buffer is announced so that its
        # all-reduce overlaps the rest of backward; the announcing stream must own the writes, so the
        # side-stream branches are off then (at data-parallel sizes the GPU is full anyway)
        hook = getattr(self, "_grad_ready", None)
        off, numel = self._off, self._numel
        small = self._side_branch if (self.overlap_small and hook is None) else contextlib.nullcontext
        # generator
        with small():
            self._gemm(1, 0, V, H, B, ws.dlogits.data_ptr(), ws.Vp, ws.dec_h[L - 1].data_ptr(), H,
                       gp("model.generator.proj.weight"), H, None, 1.0)
        cell_bwd = ws.dec_cell_bwd
        self._gemm(0, 0, B, H, V, ws.dlogits.data_ptr(), ws.Vp, self._ptr("model.generator.proj.weight"), H,
                   (ws.d_hl[L - 1] if cell_bwd else ws.d_h).data_ptr(), H)
        # decoder cells, top down
        dec_prec = prec if (os.environ.get("SLNLP_DEC_TC", "0") == "1" or B > 256) else 0
        pre = "model.decoder.rnn."
        for l in range(L - 1, -1, -1):
            D = E + 2 * H if l == 0 else H
            h0 = ws.hidden0[l].data_ptr()
            xin = ws.dec_xin[l].data_ptr()
            # d(input) of the cell; for l > 0 the input was dropout(h_{l-1}): the mask rides in the epilogue
            ddrop = (self.p_rnn, rng, 100 + l - 1) if (l > 0 and drop) else None
            if cell_bwd:
                # gate gradients, d(h0) (+ d(c0): c0 = h0 = hidden0, bkp:278-279) and d(input) of the cell: one launch
                dg = ws.dec_dg[l].data_ptr()
                dst = ws.dec_dnh[l].data_ptr() if mode != 0 else None
                dx = ws.d_decx if l == 0 else ws.d_hl[l - 1]
                check(lib.slnlp_dec_cell_bwd(mode, B, H, D, ws.dec_gates[l].data_ptr(), ws.dec_stash[l].data_ptr(), h0,
                                             h0 if mode == 0 else None, ws.d_hl[l].data_ptr(),
                                             self._ptr(f"{pre}weight_ih_l{l}"), self._ptr(f"{pre}weight_hh_l{l}"), dg, dst,
                                             dx.data_ptr(), ws.d_hidden0[l].data_ptr(), self.p_rnn if ddrop else 0.0,
                                             rng if ddrop else None, 100 + l - 1 if ddrop else 0, s), "dec_cell_bwd")
            else:
                dg, dst = ws.dec_gates[l].data_ptr(), ws.dec_stash[l].data_ptr()
                check(lib.slnlp_rnn_layer_bwd(mode, dec_prec, 1, B, H, 1, dg, dst, ws.dec_h[l].data_ptr(),
                                              self._ptr(f"{pre}weight_hh_l{l}"), None, h0, h0 if mode == 0 else None,
                                              ws.d_h.data_ptr(), None, None, ws.d_hidden0[l].data_ptr(),
                                              ws.d_c0.data_ptr() if mode == 0 else None, ws.carry.data_ptr(), s),
                      "rnn_layer_bwd(dec)")
                if mode == 0:  # LSTM: c0 = h0 = hidden0 (bkp:278-279)
                    check(lib.slnlp_axpy(ws.d_hidden0[l].data_ptr(), ws.d_c0.data_ptr(), 1.0, B * H, s), "axpy")
                dx = ws.d_decx if l == 0 else ws.d_h
                self._gemm(0, 0, B, D, GH, dg, GH, self._ptr(f"{pre}weight_ih_l{l}"), D, dx.data_ptr(), D, drop=ddrop)
            with small():
                ss = _stream()
                self._gemm(1, 0, GH, D, B, dg, GH, xin, D, gp(f"{pre}weight_ih_l{l}"), D, None, 1.0)
                check(lib.slnlp_colsum_f32(dg, B, GH, GH, gp(f"{pre}bias_ih_l{l}"), 1.0, ss), "colsum")
                if mode == 0:
                    self._gemm(1, 0, GH, H, B, dg, GH, h0, H, gp(f"{pre}weight_hh_l{l}"), H, None, 1.0)
                    # LSTM: d b_hh == d b_ih (both add to the same pre-activation)
                    check(lib.slnlp_axpy(gp(f"{pre}bias_hh_l{l}"), gp(f"{pre}bias_ih_l{l}"), 1.0, GH, ss), "axpy")
                else:
                    self._gemm(1, 0, 2 * H, H, B, dg, GH, h0, H, gp(f"{pre}weight_hh_l{l}"), H, None, 1.0)
                    self._gemm(1, 0, H, H, B, dst, H, h0, H, gp(f"{pre}weight_hh_l{l}") + 4 * 2 * H * H, H, None, 1.0)
                    check(lib.slnlp_colsum_f32(dg, B, 2 * H, GH, gp(f"{pre}bias_hh_l{l}"), 1.0, ss), "colsum")
                    check(lib.slnlp_colsum_f32(dst, B, H, H, gp(f"{pre}bias_hh_l{l}") + 4 * 2 * H, 1.0, ss), "colsum")
        # decoder input = [trg_embed[bos] || ctx]
        if self.bos_idx != self.tgt_pad:
            drow = gp("model.trg_embed.weight") + 4 * self.bos_idx * E
        else:
            drow = ws.scratch_row.data_ptr()  # padding_idx row gets no gradient
        enc_out = ws.enc_filled
        att = "model.decoder.attention."
        if ws.dec_head_bwd:
            # the bos row's gradient is a leaf (side lane); attention -> query -> tanh' -> bridge backward read d(ctx)
            # in place from d(decoder input): one launch, one CTA per sequence
            with small():
                check(lib.slnlp_dec_input_bwd(ws.d_decx.data_ptr(), drow, ws.d_ctx.data_ptr(), B, E, 2 * H, _stream()),
                      "dec_input_bwd")
            check(lib.slnlp_dec_head_bwd(T, B, H, L, E, ws.d_decx.data_ptr(), ws.q.data_ptr(), ws.pk.data_ptr(),
                                         self._ptr(att + "energy_layer.weight"), enc_out.data_ptr(), ws.alpha.data_ptr(),
                                         ws.hidden0.data_ptr(),
                                         self._ptr(att + "query_layer.weight") if ws.fuse_query else None,
                                         self._ptr("model.decoder.bridge.weight") if ws.fuse_bridge else None,
                                         ws.d_seq.data_ptr(), ws.d_pk.data_ptr(), ws.d_q.data_ptr(), ws.dv_part.data_ptr(),
                                         ws.d_hidden0.data_ptr(), ws.d_enc_final.data_ptr(), s), "dec_head_bwd")
        else:
            check(lib.slnlp_dec_input_bwd(ws.d_decx.data_ptr(), drow, ws.d_ctx.data_ptr(), B, E, 2 * H, s), "dec_input_bwd")
            # attention
            check(lib.slnlp_attn_step_bwd(ws.d_ctx.data_ptr(), ws.q.data_ptr(), ws.pk.data_ptr(),
                                          self._ptr(att + "energy_layer.weight"), enc_out.data_ptr(),
                                          ws.alpha.data_ptr(), T, B, H, 2 * H, ws.d_seq.data_ptr(), ws.d_pk.data_ptr(),
                                          ws.d_q.data_ptr(), ws.dv_part.data_ptr(), s), "attn_bwd")
        # d(encoder output) through the key projection: only the encoder BPTT waits for it - a side lane under
        # capture, next to query -> tanh' -> bridge
        with self._side_branch(2):
            self._gemm(0, 0, T * B, 2 * H, H, ws.d_pk.data_ptr(), H, self._ptr(att + "key_layer.weight"), 2 * H,
                       ws.d_seq.data_ptr(), 2 * H, None, 1.0, big=True)
        if not ws.fuse_query:
            self._gemm(0, 0, B, H, H, ws.d_q.data_ptr(), H, self._ptr(att + "query_layer.weight"), H,
                       ws.d_hidden0[L - 1].data_ptr(), H, None, 1.0)
        # bridge
        if not ws.fuse_bridge:
            check(lib.slnlp_tanh_bwd(ws.d_hidden0.data_ptr(), ws.hidden0.data_ptr(), L * B * H, s), "tanh_bwd")
            self._gemm(0, 0, L * B, 2 * H, H, ws.d_hidden0.data_ptr(), H, self._ptr("model.decoder.bridge.weight"),
                       2 * H, ws.d_enc_final.data_ptr(), 2 * H)
        with small():
            ss = _stream()
            check(lib.slnlp_colsum_f32(ws.dv_part.data_ptr(), B, H, H, gp(att + "energy_layer.weight"), 1.0, ss), "colsum")
            self._gemm(1, 0, H, H, B, ws.d_q.data_ptr(), H, ws.hidden0[L - 1].data_ptr(), H,
                       gp(att + "query_layer.weight"), H, None, 1.0)
            self._gemm(1, 0, H, 2 * H, L * B, ws.d_hidden0.data_ptr(), H, ws.enc_final.data_ptr(), 2 * H,
                       gp("model.decoder.bridge.weight"), 2 * H, None, 1.0)
            check(lib.slnlp_colsum_f32(ws.d_hidden0.data_ptr(), L * B, H, H, gp("model.decoder.bridge.bias"), 1.0, ss),
                  "colsum")
        # the key-layer gradient reads the pad-filled copy: a leaf like the other weight gradients
        with (self._side_branch(1) if (self.overlap_small and hook is None) else contextlib.nullcontext()):
            self._gemm(1, 0, H, 2 * H, T * B, ws.d_pk.data_ptr(), H, enc_out.data_ptr(), 2 * H,
                       gp(att + "key_layer.weight"), 2 * H, None, 1.0, big=True)
        # encoder BPTT, top down, on the zero-padded encoder output (the 1.0 fill of pad_packed_sequence is a
        # constant and must not enter dW_hh)
        if hook is not None:      # attention + decoder + bridge, and target embedding + generator, are final
            hook(gflat, off["model.decoder.attention.key_layer.weight"], off[self._src_names[0]])
            hook(gflat, off["model.trg_embed.weight"], numel)
        pre = "model.encoder.rnn."
        self._join_lane(2)
        for l in range(L - 1, -1, -1):
            D = E if l == 0 else 2 * H
            dg, st, out = ws.enc_gates[l].data_ptr(), ws.enc_stash[l].data_ptr(), ws.enc_out[l].data_ptr()
            if ws.rnn_extras:
                # d(final states) read in the concatenated layout; for l < L-1 d_seq is the gradient of the DROPPED
                # output of this layer: the forward's mask is applied while the kernel reads it
                undrop = l < L - 1 and drop and ws.rnn_fused_dropout
                ex = RnnExtras(1, None, self.p_rnn if undrop else 0.0, rng if undrop else None, l, 1 if undrop else 0,
                               ws.drop_mask[l].data_ptr() if undrop else None)
                check(lib.slnlp_rnn_layer_bwd_ex(mode, prec, T, B, H, 2, dg, st, out, self._ptr(f"{pre}weight_hh_l{l}"),
                                                 lp, None, None, ws.d_seq.data_ptr(), ws.d_enc_final[l].data_ptr(), None,
                                                 None, None, ws.carry.data_ptr(), ctypes.byref(ex), s), "rnn_layer_bwd")
            elif ws.bf_step:
                check(lib.slnlp_concat_dirs(ws.d_enc_final[l].data_ptr(), ws.d_hfin.data_ptr(), B, H, 2, 1, s),
                      "concat_dirs_inv")
                for d in range(2):       # W_hh^T of each direction as a K-major bf16 operand
                    check(lib.slnlp_cast_bf16(self._ptr(f"{pre}weight_hh_l{l}") + 4 * d * GH * H, H,
                                              ws.w_hhT_bf[l].data_ptr() + 2 * d * GH * H, GH, GH, H, 1, s), "cast_bf16")
                check(lib.slnlp_rnn_layer_bwd_bf16(mode, T, B, H, 2, dg, ws.dg_bf[l].data_ptr(), st, out,
                                                   ws.w_hhT_bf[l].data_ptr(), lp, ws.d_seq.data_ptr(), ws.d_hfin.data_ptr(),
                                                   None, ws.carry.data_ptr(), 0 if ws.dg_bf_only[l] else 1,
                                                   ws.keep_bits[l].data_ptr() if ws.masked[l] else None,
                                                   1.0 / (1.0 - self.p_rnn) if ws.masked[l] else 1.0,
                                                   1 if ws.gates_bf[l] else 0, s),
                      "rnn_layer_bwd_bf16")
            else:
                check(lib.slnlp_concat_dirs(ws.d_enc_final[l].data_ptr(), ws.d_hfin.data_ptr(), B, H, 2, 1, s),
                      "concat_dirs_inv")
                check(lib.slnlp_rnn_layer_bwd(mode, prec, T, B, H, 2, dg, st, out, self._ptr(f"{pre}weight_hh_l{l}"),
                                              lp, None, None, ws.d_seq.data_ptr(), ws.d_hfin.data_ptr(), None,
                                              None, None, ws.carry.data_ptr(), s), "rnn_layer_bwd")
            # weight / bias gradients of this layer: off the chain, on the side lanes.  Under capture they fork
            # HERE, right behind the BPTT kernel that produced their operand (forking after the dx GEMM would make
            # them wait for it: the last layer's weight gradients then trail the whole step)
            par = self.overlap_dw and hook is None and torch.cuda.is_current_stream_capturing()
            if ws.pair[l] and not ws.bf_step:
                # bf16 copies of d(pre-activations) (A of dX, dW_ih, dW_hh) and of the layer output (B of dW_hh);
                # the bf16 step kernels have written both already
                self._cast_bf16(dg, 2 * GH, ws.dg_bf[l], T * B, 2 * GH)
                if mode == 0:
                    self._cast_bf16(out, 2 * H, ws.out_bf[l], T * B, 2 * H)
            if par:
                self._encoder_weight_grads(ws, l, gp, parallel=True)
            # the gradient the next (lower) layer's BPTT is waiting for
            dx_pair = ws.pair[l] and self._pair_ok(T * B, D, 2 * GH)
            if dx_pair:
                self._gemm_bf16(0, 0, T * B, D, 2 * GH, ws.dg_bf[l].data_ptr(), 2 * GH, ws.w_ih_bf[l].data_ptr(), D,
                                (ws.d_seq if l > 0 else ws.d_emb).data_ptr(), D)
            if l > 0:
                if not dx_pair:
                    self._gemm(0, 0, T * B, D, 2 * GH, dg, 2 * GH, self._ptr(f"{pre}weight_ih_l{l}"), D,
                               ws.d_seq.data_ptr(), D, big=True)
                if drop and not ws.rnn_fused_dropout and not ws.masked[l - 1]:    # masked: applied by layer l-1's BPTT
                    check(lib.slnlp_dropout(ws.d_seq.data_ptr(), ws.d_seq.data_ptr(), T * B * D, self.p_rnn,
                                            rng, l - 1, s), "dropout")
            else:
                if not dx_pair:
                    self._gemm(0, 0, T * B, E, 2 * GH, dg, 2 * GH, self._ptr(f"{pre}weight_ih_l{l}"), E,
                               ws.d_emb.data_ptr(), E, big=True)
                check(lib.slnlp_embed_gather_bwd(gp(self._src_names[0]), X.data_ptr(), ws.d_emb.data_ptr(), B, T, ws.F,
                                                 ws.f_off, ws.f_w_src, ws.f_rows_src, 1, 1.0, self.src_pad, s),
                      "embed_bwd")
            if not par:       # eager / data-parallel: after the critical-path kernels, same stream
                self._encoder_weight_grads(ws, l, gp, parallel=False)
            if hook is not None:  # this layer's range goes out while the layer below runs its BPTT
                nxt = f"{pre}weight_ih_l{l + 1}" if l < L - 1 else "model.decoder.attention.key_layer.weight"
                hook(gflat, off[f"{pre}weight_ih_l{l}"], off[nxt])
                if l == 0:
                    hook(gflat, off[self._src_names[0]], off["model.trg_embed.weight"])
        if self.overlap_dw or self.overlap_small:
            self._join_side()

    def _encoder_weight_grads(self, ws, l, gp, parallel=False):
        """dW_ih, db_ih, dW_hh, db_hh of encoder layer l from its d(pre-activation) buffer: four
        independent pieces, each on its own side lane when ``parallel`` (graph capture only)."""
        E, H, G = self.E, self.H, self.G
        B, T = ws.B, ws.T
        GH = G * H
        mode = MODE[self.rnn_type]
        D = E if l == 0 else 2 * H
        pre = "model.encoder.rnn."
        dg, st, out = ws.enc_gates[l].data_ptr(), ws.enc_stash[l].data_ptr(), ws.enc_out[l].data_ptr()
        xin = (ws.emb if l == 0 else ws.enc_xin[l]).data_ptr()
        # consecutive layers alternate between two sets of four lanes: a lane is a FIFO, and layer l+1's weight-gradient
        # GEMMs - still running, slowly, beside layer l's BPTT - would otherwise hold back layer l's (the LAST layer's
        # then trail the step: measured 42 us after the final BPTT for three 10 us GEMMs)
        base = 4 * (l % 2)
        lane = (lambda i: self._side_branch(base + i)) if parallel else (lambda i: contextlib.nullcontext())
        pair = ws.pair[l]
        with lane(0):
            if pair and self._pair_ok(2 * GH, D, T * B):
                self._gemm_bf16(1, 0, 2 * GH, D, T * B, ws.dg_bf[l].data_ptr(), 2 * GH, ws.xin_bf[l].data_ptr(), D,
                                gp(f"{pre}weight_ih_l{l}"), D, None, 1.0)
            else:
                self._gemm(1, 0, 2 * GH, D, T * B, dg, 2 * GH, xin, D, gp(f"{pre}weight_ih_l{l}"), D, None, 1.0, big=True)
        with lane(3):
            s = _stream()
            if pair:
                check(lib.slnlp_colsum_bf16(ws.dg_bf[l].data_ptr(), T * B, 2 * GH, 2 * GH, gp(f"{pre}bias_ih_l{l}"), 1.0, s),
                      "colsum_bf16")
            else:
                check(lib.slnlp_colsum_f32(dg, T * B, 2 * GH, 2 * GH, gp(f"{pre}bias_ih_l{l}"), 1.0, s), "colsum")
            if mode == 0:  # LSTM: d b_hh == d b_ih for both directions at once
                check(lib.slnlp_axpy(gp(f"{pre}bias_hh_l{l}"), gp(f"{pre}bias_ih_l{l}"), 1.0, 2 * GH, s), "axpy")
            else:
                for d in range(2):
                    gb = gp(f"{pre}bias_hh_l{l}") + 4 * d * GH
                    check(lib.slnlp_colsum_f32(dg + 4 * d * GH, T * B, 2 * H, 2 * GH, gb, 1.0, s), "colsum")
                    check(lib.slnlp_colsum_f32(st + 4 * d * H, T * B, H, 2 * H, gb + 4 * 2 * H, 1.0, s), "colsum")
        K = (T - 1) * B
        for d in range(2):
            # dW_hh[d] = sum_t dG_t^T h_{prev(t)}: a GEMM over rows shifted by one timestep
            a_row = B if d == 0 else 0
            b_row = 0 if d == 0 else B
            gw = gp(f"{pre}weight_hh_l{l}") + 4 * d * GH * H
            hb = out + 4 * (b_row * 2 * H + d * H)
            if K <= 0:
                continue
            with lane(1 + d):
                if mode == 0 and pair and self._pair_ok(GH, H, K):
                    self._gemm_bf16(1, 0, GH, H, K, ws.dg_bf[l].data_ptr() + 2 * (a_row * 2 * GH + d * GH), 2 * GH,
                                    ws.out_bf[l].data_ptr() + 2 * (b_row * 2 * H + d * H), 2 * H, gw, H, None, 1.0)
                elif mode == 0:
                    self._gemm(1, 0, GH, H, K, dg + 4 * (a_row * 2 * GH + d * GH), 2 * GH, hb, 2 * H,
                               gw, H, None, 1.0, big=True)
                else:
                    self._gemm(1, 0, 2 * H, H, K, dg + 4 * (a_row * 2 * GH + d * GH), 2 * GH, hb, 2 * H,
                               gw, H, None, 1.0, big=True)
                    self._gemm(1, 0, H, H, K, st + 4 * (a_row * 2 * H + d * H), 2 * H, hb, 2 * H,
                               gw + 4 * 2 * H * H, H, None, 1.0, big=True)

    # ------------------------------------------------------------------ public forward
    def _check_inputs(self, X, lengths):
        if not (X.is_cuda and lengths.is_cuda):
            raise RuntimeError("slnlp_b200: inputs must be CUDA tensors (no CPU fallback)")
        nd = 3 if self.field_rows else 2
        if X.dim() != nd or lengths.dim() != 1 or lengths.numel() != X.shape[0]:
            raise ValueError("expected X [B,T,F] (factored fields) and lengths [B]" if self.field_rows
                             else "expected X [B,T] and lengths [B]")
        if self.field_rows and X.shape[2] != len(self.field_rows):
            raise ValueError(f"expected {len(self.field_rows)} fields per frame, got {X.shape[2]}")
        if self.validate_inputs:
            T = X.shape[1]
            hi = (torch.tensor(self.field_rows, device=X.device) if self.field_rows else self.V_src)
            bad = ((lengths < 1) | (lengths > T)).any() | (X < 0).any() | (X >= hi).any()
            if bool(bad):
                raise ValueError("lengths must be in [1, T] (pack_padded_sequence) and tokens in [0, V_src)")

    def forward(self, X, y=None, lengths=None, **kwargs):        # bkp:388-402
        """Returns log-probabilities [B, V_tgt].  ``y`` is accepted for interface parity;
        its values never reach the output of the RNN models."""
        if not self.batch_first:
            X = X.transpose(0, 1)
        self._ensure_flat()
        dev = self._flat.device
        X = X.to(dev, torch.int64).contiguous()
        if lengths is None:                                       # util.resolve_lengths
            lengths = ((X[..., 0] if self.field_rows else X) != self.src_pad).sum(1)
        lengths = lengths.to(dev, torch.int64).contiguous()
        self._check_inputs(X, lengths)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params.values()):
            names = [n for n in self._params]
            return _ModuleFn.apply(self, X, lengths, None, *[self._params[n] for n in names])
        ws = self._workspace(X.shape[0], X.shape[1], self.training)
        if ws.train and self.p_rnn > 0:
            check(lib.slnlp_rng_advance(self._rng_state().data_ptr(), _stream()), "rng")
        return self._run_forward(ws, X, lengths).clone()

    @torch.no_grad()
    def predict_logp(self, X, lengths, y=None):
        """Inference forward without autograd or input validation."""
        self._ensure_flat()
        was = self.training
        self.training = False
        try:
            ws = self._workspace(X.shape[0], X.shape[1], False)
            return self._run_forward(ws, X, lengths)
        finally:
            self.training = was


class _Workspace:
    """Activation + gradient scratch for one (B, T, train) shape; torch owns the memory."""

    def __init__(self, m: RnnEncDecB200, B, T, train, bwd=None):
        import ctypes
        bwd = train if bwd is None else bwd
        E, H, L, G, V = m.E, m.H, m.L, m.G, m.V_tgt
        dev = m._flat.device
        f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        self.B, self.T, self.train = B, T, train
        self.rnn_extras = (os.environ.get("SLNLP_RNN_EXTRAS", "1") != "0" and
                           bool(lib.slnlp_rnn_extras_supported(1 if m.precision == "bf16" else 0, T, B, H, 2)))
        # the inter-layer dropout inside the recurrent kernels, from factors generated off the critical path
        # (Philox in the step loop - ten dependent rounds - was measured a net loss on the tcgen05 kernel, whose
        # 4 epilogue warps sit one per scheduler)
        env = os.environ.get("SLNLP_RNN_FUSED_DROPOUT")
        self.rnn_fused_dropout = self.rnn_extras and (env is None or env != "0")
        if m.field_rows:      # factored embedding: F tables back to back in the flat buffer
            F = self.F = len(m.field_rows)
            base = m._off[m._src_names[0]]
            self.f_off = (ctypes.c_int64 * F)(*[m._off[n] - base for n in m._src_names])
            self.f_w_src = (ctypes.c_int * F)(*m.field_widths)
            self.f_rows_src = (ctypes.c_int64 * F)(*m.field_rows)
            self.xmask = torch.empty(B, T, dtype=torch.int64, device=dev)
        else:
            self.F = 1
            self.f_off = (ctypes.c_int64 * 1)(0)
            self.f_w_src = (ctypes.c_int * 1)(E)
            self.f_rows_src = (ctypes.c_int64 * 1)(m.V_src)
        drop = train and m.p_rnn > 0
        # CTA-pair bf16 GEMM for the hoisted projections (and their dX / dW twins) where T*B rows are worth it
        bf = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)
        self.pair = [m.precision == "bf16" and m._pair_ok(T * B, 2 * G * H, E if l == 0 else 2 * H) for l in range(L)]
        self.xin_bf = [bf(T * B, E if l == 0 else 2 * H) if self.pair[l] else None for l in range(L)]
        self.w_ih_bf = [bf(2 * G * H, E if l == 0 else 2 * H) if self.pair[l] else None for l in range(L)]
        # ... and bf16-operand per-step recurrent kernels where that family applies (LSTM, batch > 256)
        self.bf_step = (m.precision == "bf16" and all(self.pair) and os.environ.get("SLNLP_BF_STEP", "1") != "0" and
                        bool(lib.slnlp_rnn_bf16_step_supported(MODE[m.rnn_type], T, B, H, 2)))
        self.dg_bf = [bf(T * B, 2 * G * H) if (self.pair[l] and bwd) else None for l in range(L)]
        self.out_bf = [bf(T * B, 2 * H) if (self.pair[l] and ((bwd and G == 4) or self.bf_step)) else None for l in range(L)]
        if self.bf_step:
            self.w_hh_bf = [bf(2, G * H, H) for _ in range(L)]
            self.w_hhT_bf = [bf(2, H, G * H) for _ in range(L)] if bwd else []
            if not drop:       # layer l+1 reads layer l's bf16 output directly
                for l in range(1, L):
                    self.xin_bf[l] = self.out_bf[l - 1]
        dims = [(E if l == 0 else 2 * H) for l in range(L)]
        # bf_in[l]: every consumer of layer l's input (projection, dW_ih) reads the bf16 copy; dg_bf_only[l]: every
        # consumer of its d(pre-activations) (dX, dW_ih, dW_hh, bias sums, the next BPTT step) reads the bf16 copy
        self.bf_in = [self.pair[l] and m._pair_ok(2 * G * H, dims[l], T * B) for l in range(L)]
        self.dg_bf_only = [self.bf_step and self.bf_in[l] and m._pair_ok(T * B, dims[l], 2 * G * H) and
                           m._pair_ok(G * H, H, (T - 1) * B) for l in range(L)]
        self.xin_bf_ready = [l > 0 and ((drop and self.bf_in[l]) or (not drop and self.bf_step)) for l in range(L)]
        # the persistent CTA-pair step kernels also take the inter-layer dropout's keep mask as bits (no dropout pass
        # over the gradient) and can skip the fp32 copy of a lower layer's output when nothing reads it
        pairk = self.bf_step and bool(lib.slnlp_rnn_bf16_pair_supported(MODE[m.rnn_type], T, B, H, 2))
        self.masked = [pairk and drop and bwd and l < L - 1 and self.bf_in[l + 1] and (T * B * 2 * H) % 128 == 0 for l in range(L)]
        self.keep_bits = [torch.empty(T * B * 2 * H // 32, dtype=torch.int32, device=dev) if self.masked[l] else None
                          for l in range(L)]
        self.gates_bf = [pairk and bwd and self.dg_bf[l] is not None and os.environ.get("SLNLP_GATES_BF16", "1") != "0"
                         for l in range(L)]
        self.skip_out32 = [pairk and l < L - 1 and (self.dg_bf_only[l] or not bwd) and
                           (self.masked[l] if (drop and bwd) else (self.bf_in[l + 1] and not drop)) for l in range(L)]
        self.emb = f(T, B, E)
        self.enc_gates = [f(T, B, 2, G, H) for _ in range(L)]
        self.enc_stash = [f(T, B, 2, H) for _ in range(L)]
        self.enc_out = [f(T, B, 2 * H) for _ in range(L)]
        self.enc_filled = f(T, B, 2 * H)
        self.drop_mask = [f(T, B, 2 * H) for _ in range(L - 1)] if (drop and self.rnn_fused_dropout) else []
        self.enc_xin = [None] + [f(T, B, 2 * H) if drop else self.enc_out[l - 1] for l in range(1, L)]
        self.enc_hfin = [f(2, B, H) for _ in range(L)]
        self.enc_final = f(L, B, 2 * H)
        self.hidden0 = f(L, B, H)
        self.pk, self.q = f(T, B, H), f(B, H)
        self.alpha, self.ctx = f(B, T), f(B, 2 * H)
        self.dec_gates = [f(1, B, 1, G, H) for _ in range(L)]
        self.dec_stash = [f(1, B, 1, H) for _ in range(L)]
        self.dec_h = [f(1, B, H) for _ in range(L)]
        self.dec_xin = [f(B, E + 2 * H)] + [f(B, H) if drop else self.dec_h[l - 1] for l in range(1, L)]
        self.logits, self.logp = f(B, V), f(B, V)
        # the decoder side at the reference's batch sizes as per-sequence fused kernels (dec_head.cu, dec_cell.cu):
        # head = attention -> decoder input in one launch (backward: the attention reads d(ctx) in place from d(decoder
        # input), the bos row's gradient is a leaf on a side lane); cell_bwd = one launch per decoder layer instead of
        # step kernels + axpy + d(input) GEMM + split-K reduce + dropout.  $SLNLP_DEC_HEAD_FUSE=1 also folds the query
        # (H <= 256) and bridge (H <= 128) products into the head kernels, one CTA per sequence: measured SLOWER at the
        # reference's batch (cfg1 0.381 -> 0.398 ms/step: the head then waits for the key projection, which ran NEXT to
        # bridge -> tanh -> query before, and 32 warps re-reading 192 KB of weights per sequence cost what the two
        # 6 us GEMM launches did) - kept for the parity test and the record, off by default
        mode = MODE[m.rnn_type]
        small = B <= 256
        self.fuse_query = (small and H <= 256 and os.environ.get("SLNLP_DEC_HEAD", "1") != "0" and
                           os.environ.get("SLNLP_DEC_HEAD_FUSE", "0") == "1")
        self.fuse_bridge = self.fuse_query and H <= 128
        self.dec_head = (small and os.environ.get("SLNLP_DEC_HEAD", "1") != "0" and
                         bool(lib.slnlp_dec_head_supported(T, B, H, L, int(self.fuse_query), int(self.fuse_bridge))))
        if not self.dec_head:
            self.fuse_query = self.fuse_bridge = False
        # the backward twin without the two products is attn_step_bwd reading d(ctx) in place
        self.dec_head_bwd = self.dec_head and (self.fuse_query or os.environ.get("SLNLP_DEC_HEAD_BWD", "0") == "1")
        # (the one-launch cell backward pays while a layer's contraction G*H x (D + H) stays within 2^18 weights: cfg1
        # 0.379 -> 0.368 ms/step; at GRU 512 / 256 - 768 x 1280 and 768 x 512 - it measured 29.7 us per launch and the
        # step 1.782 -> 1.827 ms: those shapes keep the step kernels + tensor-core GEMM.  $SLNLP_DEC_CELL_BWD=2 forces it)
        env_cb = os.environ.get("SLNLP_DEC_CELL_BWD", "1")
        self.dec_cell_bwd = (small and env_cb != "0" and
                             os.environ.get("SLNLP_DEC_TC", "0") != "1" and os.environ.get("SLNLP_DEC_FUSED", "1") != "0" and
                             all(bool(lib.slnlp_dec_cell_bwd_supported(mode, B, H, E + 2 * H if l == 0 else H)) and
                                 (env_cb == "2" or G * H * ((E + 2 * H if l == 0 else H) + H) <= (1 << 18))
                                 for l in range(L)))
        if bwd:
            self.Vp = (V + 3) & ~3                    # row stride of dlogits: 16-byte rows keep its GEMMs on TMA
            self.dlogits = torch.zeros(B, self.Vp, device=dev)
            self.d_h, self.d_c0 = f(B, H), f(B, H)
            if self.dec_cell_bwd:     # d(h_1) per layer (a layer's d(input) must not overwrite what its own CTAs still read)
                self.d_hl = [f(B, H) for _ in range(L)]
                self.dec_dg = [f(B, G, H) for _ in range(L)]
                self.dec_dnh = [f(B, H) if G == 3 else None for _ in range(L)]
            self.d_decx, self.d_ctx = f(B, E + 2 * H), f(B, 2 * H)
            self.d_seq = f(T, B, 2 * H)       # d enc_out / d layer outputs (reused down the stack)
            self.d_pk, self.d_q, self.dv_part = f(T, B, H), f(B, H), f(B, H)
            self.d_hidden0, self.d_enc_final = f(L, B, H), f(L, B, 2 * H)
            self.d_hfin = f(2, B, H)
            self.d_emb = f(T, B, E)
            self.carry = f(4, B, H)
            self.scratch_row = torch.zeros(E, device=dev)
            self.loss = torch.zeros(2, device=dev)
            self.row_ws = f(3 * B)


class _ModuleFn(torch.autograd.Function):
    """Autograd bridge for the drop-in path: forward/backward launch the C-ABI kernels."""

    @staticmethod
    def forward(ctx, module, X, lengths, y, *params):
        ws = module._workspace(X.shape[0], X.shape[1], module.training, fresh=True, bwd=True)
        if ws.train and module.uses_rng:
            check(lib.slnlp_rng_advance(module._rng_state().data_ptr(), _stream()), "rng")
            ctx.rng_step = module._rng_state().clone()
        ctx.module, ctx.ws, ctx.X, ctx.lengths, ctx.y = module, ws, X, lengths, y
        logp = module._run_forward(ws, X, lengths, y)
        return logp.clone()

    @staticmethod
    def backward(ctx, dlogp):
        m, ws = ctx.module, ctx.ws
        if not dlogp.is_cuda:
            raise RuntimeError("slnlp_b200: gradient must be a CUDA tensor")
        dlogp = dlogp.contiguous().float()
        check(lib.slnlp_log_softmax_bwd(dlogp.data_ptr(), ws.logp.data_ptr(), ws.dlogits.data_ptr(),
                                        ws.B, m.V_tgt, ws.Vp, _stream()), "log_softmax_bwd")
        g = torch.zeros_like(m._flat)
        if ws.train and m.uses_rng:  # replay the dropout masks of this forward
            saved = m._rng_state().clone()
            m._rng_state().copy_(ctx.rng_step)
        m._run_backward(ws, ctx.X, ctx.lengths, g, ctx.y)
        if ws.train and m.uses_rng:
            m._rng_state().copy_(saved)
        grads = []
        for n in m._params:
            grads.append(None if n in m._dead else m._view(g, n))
        return (None, None, None, None, *grads)


class OptimState:
    """SGD-momentum + clip state of one fit: hyper = {lr, momentum, max_norm, first-step flag}
    (device, read by the kernels so that a captured graph sees lr changes), momentum buffer,
    gradient-norm scratch."""

    def __init__(self, module: FlatParamModule, lr: float, momentum: float = 0.9, max_norm: float = 0.5):
        module._ensure_flat()
        dev = module._flat.device
        self.hyper = torch.tensor([lr, momentum, max_norm if max_norm else 0.0, 0.0], device=dev)
        self.buf = torch.zeros_like(module._flat)     # zero buffer == "first step" of torch SGD
        self.partials = torch.zeros(lib.slnlp_sumsq_partials(), device=dev)
        self.norm = torch.zeros(1, device=dev)
        self._lr = lr

    def set_lr(self, lr):
        if lr != self._lr:
            self.hyper[0] = lr
            self._lr = lr

    def state_dict(self, module: FlatParamModule = None):
        """With ``module``: the layout of ``torch.optim.SGD(module.parameters()).state_dict()`` - what
        the reference's skorch Checkpoint writes to optimizer.pt and ``load_params(f_optimizer=...)``
        reads - with every momentum buffer a copy of that parameter's slice of the flat buffer."""
        if module is None:
            return {"hyper": self.hyper.clone(), "momentum_buffer": self.buf.clone()}
        lr, momentum = float(self.hyper[0]), float(self.hyper[1])
        names = list(module._params)
        stepped = bool(self.buf.abs().max() > 0)
        state = {i: {"momentum_buffer": module._view(self.buf, n).clone()} for i, n in enumerate(names)
                 if stepped and n not in module._dead}
        group = {"lr": lr, "momentum": momentum, "dampening": 0, "weight_decay": 0, "nesterov": False,
                 "maximize": False, "foreach": None, "differentiable": False, "fused": None,
                 "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd, module: FlatParamModule):
        """Inverse of ``state_dict(module)`` (also accepts the flat form)."""
        if "momentum_buffer" in sd:
            self.buf.copy_(sd["momentum_buffer"]); self.hyper.copy_(sd["hyper"])
            return
        names = list(module._params)
        self.buf.zero_()
        for i, st in sd["state"].items():
            if st.get("momentum_buffer") is not None:
                module._view(self.buf, names[int(i)]).copy_(st["momentum_buffer"])
        grp = sd["param_groups"][0]
        self.hyper[0] = grp["lr"]; self.hyper[1] = grp["momentum"]
        self._lr = grp["lr"]


class FusedTrainStep:
    """The skorch train step of SURVEY.md 3.2 as ONE CUDA-graph replay:

    zero grads -> forward -> CrossEntropyLoss(ignore_index=pad) on the log-probs ->
    backward -> global L2 norm clip (GradientNormClipping, helper.py:227-229) ->
    SGD(momentum, nesterov=False) (config/*.yaml:39-42).
    """

    def __init__(self, module: FlatParamModule, batch_size: int, seq_len: int, lr: float = 0.01,
                 momentum: float = 0.9, max_norm: float = 0.5, use_graph: bool = True,
                 grad_sync=None, state: "OptimState" = None):
        module._ensure_flat()
        self.m, self.B, self.T = module, batch_size, seq_len
        dev = module._flat.device
        self.ws = module._make_workspace(batch_size, seq_len, True, True)
        self.ws.fused_ce = True      # log_softmax + CE + d logits run as one kernel (K10)
        # static inputs of the captured graph, views of one int64 buffer [X | lengths | y].  (Packing a host
        # batch into one pinned staging buffer + one H2D copy was measured: the extra host-side copies
        # and the buffer-reuse event cost as much as the two H2D launches they save.)
        nf = len(getattr(module, "field_rows", None) or []) or 1
        BT = batch_size * seq_len * nf
        self._inputs = torch.empty(BT + 2 * batch_size, dtype=torch.int64, device=dev)
        self.X = self._inputs[:BT].view(batch_size, seq_len, nf) if nf > 1 else self._inputs[:BT].view(batch_size, seq_len)
        self.lengths = self._inputs[BT:BT + batch_size]
        self.y = self._inputs[BT + batch_size:]
        self.X.fill_(module.src_pad); self.lengths.fill_(1); self.y.zero_()
        # optimizer state may be shared by several steps of different batch shape (tail batches)
        self.state = state if state is not None else OptimState(module, lr, momentum, max_norm)
        self.gflat = module.flat_grads()
        self._grads_clean = False                    # True once a step's SGD kernel has zeroed the gradient buffer
        self.grad_sync = grad_sync                   # callable(gflat, loss) for data parallel
        self.grad_scale = 1.0
        self.graph = None
        # The data-parallel step is captured too (NCCL collectives are graph-capturable): measured
        # 165 k vs 95 k seq/s at 2 GPUs (cfg1 model, global batch 100).  The captured graph holds NCCL
        # kernels: release it (ts.graph = None; DataParallelStep.release()) before
        # destroy_process_group().  SLNLP_DP_GRAPH=0 keeps the data-parallel step eager.
        self.use_graph = use_graph and (grad_sync is None or os.environ.get("SLNLP_DP_GRAPH", "1") == "1")

    # optimizer state lives in self.state; these aliases keep call sites short
    hyper = property(lambda self: self.state.hyper)
    buf = property(lambda self: self.state.buf)
    partials = property(lambda self: self.state.partials)
    norm = property(lambda self: self.state.norm)

    def set_lr(self, lr):
        self.state.set_lr(lr)

    def _step(self):
        m, ws = self.m, self.ws
        s = _stream()
        if not self._grads_clean:        # afterwards the SGD kernel leaves the consumed gradient buffer zeroed
            self.gflat.zero_()
        if m.uses_rng:
            with m._side_branch(3):      # joined before the first kernel that draws from the stream
                check(lib.slnlp_rng_advance(m._rng_state().data_ptr(), _stream()), "rng")
            if not getattr(m, "_joins_rng_lane", False):
                m._join_lane(3)
        m._run_forward(ws, self.X, self.lengths, self.y)
        # nothing in backward reads the mean loss (the gradient carries its own 1 / n_valid): under capture its reduction
        # leaves the chain for a side lane; the data-parallel exchange weights by n_valid and keeps it in stream
        loss_aside = self.grad_sync is None and torch.cuda.is_current_stream_capturing()
        check(lib.slnlp_logsoftmax_ce_fused(ws.logits.data_ptr(), self.y.data_ptr(), m.tgt_pad, self.B, m.V_tgt,
                                            ws.logp.data_ptr(), None if loss_aside else ws.loss.data_ptr(),
                                            ws.dlogits.data_ptr(), ws.Vp, ws.row_ws.data_ptr(), s), "logsoftmax_ce")
        if loss_aside:
            with m._side_branch(7):
                check(lib.slnlp_ce_reduce(ws.row_ws.data_ptr(), self.B, ws.loss.data_ptr(), _stream()), "ce_reduce")
        bucketed = getattr(self.grad_sync, "bucketed", False)
        if bucketed:     # data parallel: gradient ranges are exchanged while the rest of backward runs (dp.py)
            self.grad_sync.begin(ws.loss)
            m._grad_ready = self.grad_sync.ready
        try:
            m._run_backward(ws, self.X, self.lengths, self.gflat, self.y)
        finally:
            m._grad_ready = None
        if bucketed:
            self.grad_sync.finish(self.gflat, ws.loss)
        elif self.grad_sync is not None:
            self.grad_sync(self.gflat, ws.loss)
        m._join_side()      # every side lane (weight gradients, the loss reduction) is back before the optimizer
        n = m._numel
        check(lib.slnlp_gradnorm(self.gflat.data_ptr(), n, self.partials.data_ptr(), self.norm.data_ptr(), s), "gradnorm")
        check(lib.slnlp_sgd_momentum_clip_zero(m._flat.data_ptr(), self.gflat.data_ptr(), self.buf.data_ptr(), n,
                                               self.hyper.data_ptr(), self.norm.data_ptr(), self.grad_scale, s), "sgd")
        self._grads_clean = True

    def load_batch(self, X, y, lengths):
        """Stage one batch into the graph's static input buffers (device or pinned host tensors)."""
        self.X.copy_(X, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.lengths.copy_(lengths, non_blocking=True)

    def run(self):
        """One training step on the staged batch.  Returns the device tensor [loss, n_valid]."""
        if not self.use_graph:
            self._step()
            return self.ws.loss
        if self.graph is None:
            # warm-up outside capture (lazy module/func attribute init), on a side stream
            # (stream-level synchronisation only: a device-wide one is invalid while ANY stream of the
            # device is capturing - another fit's thread under grid.py fits_per_gpu - and would also
            # invalidate that capture)
            side = thread_stream("warmup")
            side.wait_stream(torch.cuda.current_stream())
            saved = (self.m._flat.clone(), self.buf.clone(), self.m._rng_state().clone())
            with torch.cuda.stream(side):
                self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.current_stream().synchronize()
            self.m._flat.copy_(saved[0]); self.buf.copy_(saved[1]); self.m._rng_state().copy_(saved[2])
            # "thread_local" when several fits share the process (grid.py fits_per_gpu): another thread's
            # allocations must not invalidate this capture
            graph = torch.cuda.CUDAGraph()
            with capture_graph(graph, high_priority=bool(getattr(self.m, "overlap_dw", False))):
                self._step()
            self.graph = graph
            self.m._flat.copy_(saved[0]); self.buf.copy_(saved[1]); self.m._rng_state().copy_(saved[2])
        self.graph.replay()
        return self.ws.loss

    def release(self):
        """Drop the captured graph (under the capture lock: see flat.CAPTURE_LOCK)."""
        with CAPTURE_LOCK:
            self.graph = None

    def step(self, X, y, lengths):
        self.load_batch(X, y, lengths)
        return self.run()

    @property
    def grad_norm(self):
        return self.norm


class InferStep:
    """The scoring forward (eval mode, no autograd: main.py:116-117 -> skorch predict) on a fixed
    (B, T) as ONE CUDA-graph replay.  Inputs are staged into static buffers exactly like
    ``FusedTrainStep``; ``run`` returns the module's log-prob buffer [B, V_tgt] (valid until the
    next run)."""

    def __init__(self, module: FlatParamModule, batch_size: int, seq_len: int, use_graph: bool = True):
        module._ensure_flat()
        self.m, self.B, self.T = module, batch_size, seq_len
        dev = module._flat.device
        self.ws = module._workspace(batch_size, seq_len, False)
        nf = len(getattr(module, "field_rows", None) or []) or 1
        BT = batch_size * seq_len * nf
        self._inputs = torch.empty(BT + 2 * batch_size, dtype=torch.int64, device=dev)
        self.X = self._inputs[:BT].view(batch_size, seq_len, nf) if nf > 1 else self._inputs[:BT].view(batch_size, seq_len)
        self.lengths = self._inputs[BT:BT + batch_size]
        self.y = self._inputs[BT + batch_size:]
        self.X.fill_(module.src_pad); self.lengths.fill_(1); self.y.zero_()
        self.use_graph, self.graph = use_graph, None

    def load_batch(self, X, y, lengths):
        self.X.copy_(X, non_blocking=True)
        self.lengths.copy_(lengths, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)

    def run(self):
        m = self.m
        if not self.use_graph:
            return m._run_forward(self.ws, self.X, self.lengths, self.y)
        if self.graph is None:
            side = thread_stream("warmup")
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m._run_forward(self.ws, self.X, self.lengths, self.y)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            with capture_graph(graph):
                m._run_forward(self.ws, self.X, self.lengths, self.y)
            self.graph = graph
        self.graph.replay()
        return self.ws.logp

    def release(self):
        with CAPTURE_LOCK:
            self.graph = None

    def step(self, X, y, lengths):
        self.load_batch(X, y, lengths)
        return self.run()
