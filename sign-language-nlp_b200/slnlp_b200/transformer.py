"""B200 host side of the Transformer classifier (``model.Transformer`` of the reference,
model/transformer.py:9-109; north_star's ``EncoderDecoderTransformerAttn``).

Same constructor keywords, ``.to(device)``, ``forward(X, y, lengths) -> [B, V_tgt]``
log-probs, ``state_dict`` names/shapes/order (``transformer.*`` are stock nn.Transformer names)
and default initialisation stream as the reference.  The math runs in the sm_100a kernels
behind the C ABI: fused embedding*sqrt(E)+PE gather, tcgen05 / fp32 GEMMs, the flash-style
attention kernel, fused residual+LayerNorm, fused log-softmax / CE; no torch.nn compute.

Quirks of the reference kept on purpose:
  * the causal mask is applied to the ENCODER self-attention (transformer.py:68,84);
  * the true label ``y`` is the one-token decoder input (:65,78-79);
  * cross-attention sees padded memory (no memory_key_padding_mask, :82-87).
"""
from __future__ import annotations

import contextlib
import ctypes
import math
import os
from typing import List

import torch
import torch.nn as nn

from ._lib import check, lib
from .flat import FlatParamModule, _stream

PAD_WORD = "<pad>"  # dataset/constant/tokens.py
LN_EPS = 1e-5


def _pe_table(max_len, d_model):
    """model/component/positional_encoding.py:22-31 -> buffer [max_len, 1, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


class _PEBox(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.register_buffer("pe", _pe_table(max_len, d_model))


class TransformerB200(FlatParamModule):
    # The split-operand tf32 GEMM (slnlp_gemm_tf32x3) keeps every product fp32-accurate, but the tensor core adds
    # into its accumulator with truncation: ~1e-6 (K = 128) to 1e-5 (K = 1024) of the output scale.  The RNN models'
    # fp32 fixtures hold their 1e-5 / identical-argmax / 1e-4 gradient-norm bars with it; this model's deeper chain of
    # K = 512 ... 2048 products missed the gradient-norm bar (2.3e-4), so its fp32 path stays on the fp32-FMA GEMM.
    f32_tensor_cores = False

    def __init__(self, embedding_size, num_heads, num_layers, hidden_size, dropout, src_vocab, tgt_vocab,
                 device=None, batch_first=False, precision="fp32", **kwargs):
        super().__init__()
        assert precision in ("fp32", "bf16")
        self.model_type = "Transformer"
        self.embedding_size = int(embedding_size)
        self.src_vocab, self.tgt_vocab = src_vocab, tgt_vocab
        self.device, self.batch_first, self.precision = device, batch_first, precision
        self.E, self.F, self.L, self.nhead = int(embedding_size), int(hidden_size), int(num_layers), int(num_heads)
        assert self.E % self.nhead == 0, "embed_dim must be divisible by num_heads"   # nn.MultiheadAttention
        self.dh = self.E // self.nhead
        self.p_drop = float(dropout)
        self.V_src, self.V_tgt = len(src_vocab), len(tgt_vocab)
        self.src_pad = src_vocab.stoi[PAD_WORD]            # generate_padding_mask, model/util/util.py:45-61
        self.tgt_pad = tgt_vocab.stoi[PAD_WORD]
        self.validate_inputs = True
        self.overlap_small = os.environ.get("SLNLP_OVERLAP_SMALL", "1") != "0"
        # the encoder's [B*S]-row weight gradients + bias sums on the side lane too: a dW GEMM at 3,200 rows is a
        # partial wave of mostly fixed latency, and beside the d(activation) chain it costs nothing (cfg3 2.312 -> 2.200
        # ms/step; $SLNLP_TR_SIDE_BIG=0 puts them back in stream)
        self.overlap_big = os.environ.get("SLNLP_TR_SIDE_BIG", "1") != "0"
        self.seed = int(kwargs.get("seed", torch.initial_seed() & 0x7FFFFFFF))
        self._build_parameters()

    # ------------------------------------------------------------------ parameters
    def _build_parameters(self):
        E = self.E
        # torch.nn only draws the default initialisers, in the reference's construction order
        # (transformer.py:32-48): src Embedding, tgt Embedding, nn.Transformer (xavier), Linear.
        t_src = nn.Embedding(self.V_src, E)
        t_tgt = nn.Embedding(self.V_tgt, E)
        t_tr = nn.Transformer(d_model=E, nhead=self.nhead, num_encoder_layers=self.L, num_decoder_layers=self.L,
                              dim_feedforward=self.F, dropout=self.p_drop)
        t_lin = nn.Linear(E, self.V_tgt)
        segs: List = [("src_embedding.weight", t_src.weight), ("tgt_embedding.weight", t_tgt.weight)]
        segs += [("transformer." + n, p) for n, p in t_tr.named_parameters()]
        segs += [("linear.weight", t_lin.weight), ("linear.bias", t_lin.bias)]
        order = [n for n, _ in segs]
        self._register_flat(segs, order)
        # buffers, in the reference's registration order between the embeddings
        self.src_pos_encoding = _PEBox(E)
        self.tgt_pos_encoding = _PEBox(E)
        # state_dict order of the reference: src_embedding, src_pos_encoding, tgt_embedding, ...
        want = ["src_embedding", "src_pos_encoding", "tgt_embedding", "tgt_pos_encoding", "transformer", "linear"]
        mods = self._modules
        for k in want:
            mods[k] = mods.pop(k)
        for n in order:   # norm weight/bias pairs must be adjacent (one colsum reduces both grads)
            if ".norm" in n and n.endswith(".weight"):
                assert self._off[n[:-6] + "bias"] == self._off[n] + E

    def _make_workspace(self, B, T, train, bwd):
        return _TWorkspace(self, B, T, train, bwd)

    @property
    def uses_rng(self):
        return self.p_drop > 0.0

    # ------------------------------------------------------------------ helpers
    def _lin_fwd(self, x, rows, w, b, y, n_out, n_in, ldx=None, ldy=None, big=False, w_off=0, b_off=0):
        """y[rows, n_out] = x[rows, n_in] W[w_off:w_off+n_out]^T + b"""
        self._gemm(0, 1, rows, n_out, n_in, x, ldx or n_in, self._ptr(w) + 4 * w_off * n_in, n_in, y, ldy or n_out,
                   (self._ptr(b) + 4 * b_off) if b else None, 0.0, big=big)

    def _lin_bwd(self, g, dy, x, rows, w, b, dx, n_out, n_in, beta_dx, lddy=None, ldx=None, big=False, w_off=0,
                 b_off=0):
        """gW += dy^T x; gb += colsum(dy); dx = beta_dx*dx + dy W   (dx may be None)"""
        lddy = lddy or n_out
        # The one-token decoder's weight gradients (rows = batch) are leaves of the dependency graph
        # and a fraction of a wave: they run on the side stream next to the d(activation) chain.
        # The previous side branch is joined first, so dy / x buffers recycled by later layers are
        # never overwritten under a pending read.
        self._join_side()
        if dx is not None:
            self._gemm(0, 0, rows, n_in, n_out, dy, lddy, self._ptr(w) + 4 * w_off * n_in, n_in, dx, n_in, None,
                       beta_dx, big=big)
        with (self._side_branch() if (self.overlap_small and (not big or self.overlap_big)) else contextlib.nullcontext()):
            self._gemm(1, 0, n_out, n_in, rows, dy, lddy, x, ldx or n_in, self._ptr(w, g) + 4 * w_off * n_in, n_in, None,
                       1.0, big=big)
            if b:
                check(lib.slnlp_colsum_f32(dy, rows, n_out, lddy, self._ptr(b, g) + 4 * b_off, 1.0, _stream()), "colsum")

    def _ln_fwd(self, x, res, name, y, stats, rows):
        check(lib.slnlp_add_layernorm_fwd(x, res, self._ptr(name + ".weight"), self._ptr(name + ".bias"), y,
                                          stats[0].data_ptr(), stats[1].data_ptr(), rows, self.E, LN_EPS, _stream()),
              "add_layernorm")

    def _ln_bwd(self, g, dy, x, res, name, stats, dx, rows, ws):
        nb = lib.slnlp_ln_bwd_blocks(rows)
        check(lib.slnlp_layernorm_bwd(dy, x, res, self._ptr(name + ".weight"), stats[0].data_ptr(), stats[1].data_ptr(),
                                      dx, ws.ln_part.data_ptr(), rows, self.E, 0, _stream()), "layernorm_bwd")
        # d gamma and d beta are adjacent in the flat gradient buffer
        check(lib.slnlp_colsum_f32(ws.ln_part.data_ptr(), nb, 2 * self.E, 2 * self.E, self._ptr(name + ".weight", g),
                                   1.0, _stream()), "colsum")

    def _drop(self, ws, buf, n, site):
        if ws.train and self.p_drop > 0.0:
            check(lib.slnlp_dropout(buf, buf, n, self.p_drop, self._rng_state().data_ptr(), site, _stream()), "dropout")

    def _mha_fwd(self, ws, q, ldq, k, ldk, v, ldv, o, lse, Sq, Sk, causal, tokens, pad, site):
        p = self.p_drop if ws.train else 0.0
        fn = lib.slnlp_mha_tf32_fwd if self.precision == "bf16" else lib.slnlp_mha_fwd
        check(fn(q, ldq, k, ldk, v, ldv, o, self.E, lse, ws.B, Sq, Sk, self.nhead, self.dh, causal,
                 tokens, pad, p, self._rng_state().data_ptr() if p > 0 else None, site, _stream()), "mha_fwd")

    def _mha_bwd(self, ws, q, ldq, k, ldk, v, ldv, o, do, lse, dq, dk, dv, Sq, Sk, causal, tokens, pad, site):
        p = self.p_drop if ws.train else 0.0
        fn = lib.slnlp_mha_tf32_bwd if self.precision == "bf16" else lib.slnlp_mha_bwd
        check(fn(q, ldq, k, ldk, v, ldv, o, do, self.E, lse, ws.dvec.data_ptr(), dq, dk, dv, ws.B, Sq, Sk,
                 self.nhead, self.dh, causal, tokens, pad, p,
                 self._rng_state().data_ptr() if p > 0 else None, site, _stream()), "mha_bwd")

    def _ffn_fwd(self, ws, pre, x, st, rows, site, big):
        """st.hdn = dropout(relu(x W1^T + b1)); st.f = dropout(hdn W2^T + b2)"""
        E, F = self.E, self.F
        self._lin_fwd(x, rows, pre + "linear1.weight", pre + "linear1.bias", st.hdn.data_ptr(), F, E, big=big)
        if ws.train and self.p_drop > 0.0:     # activation + dropout: one pass
            check(lib.slnlp_relu_dropout_fwd(st.hdn.data_ptr(), rows * F, self.p_drop, self._rng_state().data_ptr(), site,
                                             _stream()), "relu_dropout")
        else:
            check(lib.slnlp_relu_fwd(st.hdn.data_ptr(), rows * F, _stream()), "relu")
        self._lin_fwd(st.hdn.data_ptr(), rows, pre + "linear2.weight", pre + "linear2.bias", st.f.data_ptr(), E, F, big=big)
        self._drop(ws, st.f.data_ptr(), rows * E, site + 1)

    def _ffn_bwd(self, ws, g, pre, x, st, dsum, dhdn, rows, site, big):
        """dsum = d(x + f) on entry; on exit dsum += d x through the FFN.  dhdn: scratch [rows, F]."""
        E, F = self.E, self.F
        df = ws.tmpE_big if rows > ws.B else ws.tmpE_small
        if ws.train and self.p_drop > 0.0:   # d f = dropout-mask(d sum); keep dsum intact for the residual
            check(lib.slnlp_dropout(dsum, df.data_ptr(), rows * E, self.p_drop, self._rng_state().data_ptr(), site + 1,
                                    _stream()), "dropout")
            dfp = df.data_ptr()
        else:
            dfp = dsum
        self._lin_bwd(g, dfp, st.hdn.data_ptr(), rows, pre + "linear2.weight", pre + "linear2.bias", dhdn, E, F, 0.0, big=big)
        if ws.train and self.p_drop > 0.0:     # st.hdn > 0 exactly where the unit was active and kept
            check(lib.slnlp_relu_dropout_bwd(dhdn, st.hdn.data_ptr(), rows * F, self.p_drop, _stream()), "relu_dropout_bwd")
        else:
            check(lib.slnlp_relu_bwd(dhdn, st.hdn.data_ptr(), rows * F, _stream()), "relu_bwd")
        self._lin_bwd(g, dhdn, x, rows, pre + "linear1.weight", pre + "linear1.bias", dsum, F, E, 1.0, big=big)

    # ------------------------------------------------------------------ forward
    def _run_forward(self, ws, X, lengths, y):
        E, L, B, S = self.E, self.L, ws.B, ws.T
        s = _stream()
        R = B * S
        scale = math.sqrt(E)
        pe_src = self.src_pos_encoding.pe.data_ptr()
        pe_tgt = self.tgt_pos_encoding.pe.data_ptr()
        check(lib.slnlp_embed_gather_fwd(self._ptr("src_embedding.weight"), X.data_ptr(), ws.src.data_ptr(), B, S, 1,
                                         ws.f_off, ws.f_w, ws.f_rows_src, 0, scale, pe_src, s), "embed")
        self._drop(ws, ws.src.data_ptr(), R * E, 0)
        x = ws.src
        for l in range(L):
            st = ws.enc[l]
            pre = f"transformer.encoder.layers.{l}."
            self._lin_fwd(x.data_ptr(), R, pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias",
                          st.qkv.data_ptr(), 3 * E, E, big=True)
            q = st.qkv.data_ptr()
            self._mha_fwd(ws, q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, st.o.data_ptr(), st.lse.data_ptr(), S, S, 1,
                          X.data_ptr(), self.src_pad, 10 + 10 * l)
            self._lin_fwd(st.o.data_ptr(), R, pre + "self_attn.out_proj.weight", pre + "self_attn.out_proj.bias",
                          st.a.data_ptr(), E, E, big=True)
            self._drop(ws, st.a.data_ptr(), R * E, 11 + 10 * l)
            self._ln_fwd(x.data_ptr(), st.a.data_ptr(), pre + "norm1", st.x1.data_ptr(), st.ln1, R)
            self._ffn_fwd(ws, pre, st.x1.data_ptr(), st, R, 12 + 10 * l, True)
            self._ln_fwd(st.x1.data_ptr(), st.f.data_ptr(), pre + "norm2", st.x2.data_ptr(), st.ln2, R)
            x = st.x2
        self._ln_fwd(x.data_ptr(), None, "transformer.encoder.norm", ws.memory.data_ptr(), ws.ln_enc, R)
        # decoder: one target token per sequence = the label y (transformer.py:65,78-79)
        check(lib.slnlp_embed_gather_fwd(self._ptr("tgt_embedding.weight"), y.data_ptr(), ws.tgt.data_ptr(), B, 1, 1,
                                         ws.f_off, ws.f_w, ws.f_rows_tgt, 0, scale, pe_tgt, s), "embed")
        self._drop(ws, ws.tgt.data_ptr(), B * E, 1)
        z = ws.tgt
        mem = ws.memory.data_ptr()
        for l in range(L):
            st = ws.dec[l]
            pre = f"transformer.decoder.layers.{l}."
            self._lin_fwd(z.data_ptr(), B, pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias",
                          st.qkv.data_ptr(), 3 * E, E)
            q = st.qkv.data_ptr()
            self._mha_fwd(ws, q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, st.o.data_ptr(), st.lse.data_ptr(), 1, 1, 0,
                          y.data_ptr(), self.tgt_pad, 1000 + 10 * l)
            self._lin_fwd(st.o.data_ptr(), B, pre + "self_attn.out_proj.weight", pre + "self_attn.out_proj.bias",
                          st.a.data_ptr(), E, E)
            self._drop(ws, st.a.data_ptr(), B * E, 1001 + 10 * l)
            self._ln_fwd(z.data_ptr(), st.a.data_ptr(), pre + "norm1", st.x1.data_ptr(), st.ln1, B)
            # cross attention over the (unmasked) memory
            self._lin_fwd(st.x1.data_ptr(), B, pre + "multihead_attn.in_proj_weight", pre + "multihead_attn.in_proj_bias",
                          st.qc.data_ptr(), E, E)
            self._lin_fwd(mem, R, pre + "multihead_attn.in_proj_weight", pre + "multihead_attn.in_proj_bias",
                          st.kvc.data_ptr(), 2 * E, E, big=True, w_off=E, b_off=E)
            kv = st.kvc.data_ptr()
            self._mha_fwd(ws, st.qc.data_ptr(), E, kv, 2 * E, kv + 4 * E, 2 * E, st.oc.data_ptr(), st.lsec.data_ptr(),
                          1, S, 0, None, 0, 1002 + 10 * l)
            self._lin_fwd(st.oc.data_ptr(), B, pre + "multihead_attn.out_proj.weight", pre + "multihead_attn.out_proj.bias",
                          st.a2.data_ptr(), E, E)
            self._drop(ws, st.a2.data_ptr(), B * E, 1003 + 10 * l)
            self._ln_fwd(st.x1.data_ptr(), st.a2.data_ptr(), pre + "norm2", st.x2.data_ptr(), st.ln2, B)
            self._ffn_fwd(ws, pre, st.x2.data_ptr(), st, B, 1004 + 10 * l, False)
            self._ln_fwd(st.x2.data_ptr(), st.f.data_ptr(), pre + "norm3", st.x3.data_ptr(), st.ln3, B)
            z = st.x3
        self._ln_fwd(z.data_ptr(), None, "transformer.decoder.norm", ws.zf.data_ptr(), ws.ln_dec, B)
        self._lin_fwd(ws.zf.data_ptr(), B, "linear.weight", "linear.bias", ws.logits.data_ptr(), self.V_tgt, E)
        if not getattr(ws, "fused_ce", False):   # the fused train step folds it into the criterion kernel
            check(lib.slnlp_log_softmax_fwd(ws.logits.data_ptr(), ws.logp.data_ptr(), B, self.V_tgt, s), "log_softmax")
        return ws.logp

    # ------------------------------------------------------------------ backward
    def _run_backward(self, ws, X, lengths, g, y):
        E, L, B, S, F = self.E, self.L, ws.B, ws.T, self.F
        s = _stream()
        R = B * S
        scale = math.sqrt(E)
        dzA, dzB = ws.dz[0].data_ptr(), ws.dz[1].data_ptr()
        self._lin_bwd(g, ws.dlogits.data_ptr(), ws.zf.data_ptr(), B, "linear.weight", "linear.bias", dzA, self.V_tgt, E, 0.0,
                      lddy=ws.Vp)
        z_top = ws.dec[L - 1].x3
        self._ln_bwd(g, dzA, z_top.data_ptr(), None, "transformer.decoder.norm", ws.ln_dec, dzB, B, ws)
        cur, oth = dzB, dzA      # cur holds d(layer output)
        mem = ws.memory.data_ptr()
        dmem = ws.dmem.data_ptr()
        for l in range(L - 1, -1, -1):
            st = ws.dec[l]
            pre = f"transformer.decoder.layers.{l}."
            zin = (ws.dec[l - 1].x3 if l > 0 else ws.tgt).data_ptr()
            # norm3 / FFN
            self._ln_bwd(g, cur, st.x2.data_ptr(), st.f.data_ptr(), pre + "norm3", st.ln3, oth, B, ws)
            cur, oth = oth, cur
            self._ffn_bwd(ws, g, pre, st.x2.data_ptr(), st, cur, ws.dhdn_small.data_ptr(), B, 1004 + 10 * l, False)
            # norm2 / cross attention
            self._ln_bwd(g, cur, st.x1.data_ptr(), st.a2.data_ptr(), pre + "norm2", st.ln2, oth, B, ws)
            cur, oth = oth, cur
            da = self._masked_grad(ws, cur, ws.tmpE_small.data_ptr(), B * E, 1003 + 10 * l)
            self._lin_bwd(g, da, st.oc.data_ptr(), B, pre + "multihead_attn.out_proj.weight",
                          pre + "multihead_attn.out_proj.bias", ws.do_small.data_ptr(), E, E, 0.0)
            kv = st.kvc.data_ptr()
            dkv = ws.dkvc.data_ptr()
            self._mha_bwd(ws, st.qc.data_ptr(), E, kv, 2 * E, kv + 4 * E, 2 * E, st.oc.data_ptr(), ws.do_small.data_ptr(),
                          st.lsec.data_ptr(), ws.dqc.data_ptr(), dkv, dkv + 4 * E, 1, S, 0, None, 0, 1002 + 10 * l)
            self._lin_bwd(g, ws.dqc.data_ptr(), st.x1.data_ptr(), B, pre + "multihead_attn.in_proj_weight",
                          pre + "multihead_attn.in_proj_bias", cur, E, E, 1.0)
            self._lin_bwd(g, dkv, mem, R, pre + "multihead_attn.in_proj_weight", pre + "multihead_attn.in_proj_bias",
                          dmem, 2 * E, E, 0.0 if l == L - 1 else 1.0, big=True, w_off=E, b_off=E)
            # norm1 / self attention on the single token
            self._ln_bwd(g, cur, zin, st.a.data_ptr(), pre + "norm1", st.ln1, oth, B, ws)
            cur, oth = oth, cur
            da = self._masked_grad(ws, cur, ws.tmpE_small.data_ptr(), B * E, 1001 + 10 * l)
            self._lin_bwd(g, da, st.o.data_ptr(), B, pre + "self_attn.out_proj.weight", pre + "self_attn.out_proj.bias",
                          ws.do_small.data_ptr(), E, E, 0.0)
            q = st.qkv.data_ptr()
            dq = ws.dqkv_small.data_ptr()
            self._mha_bwd(ws, q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, st.o.data_ptr(), ws.do_small.data_ptr(),
                          st.lse.data_ptr(), dq, dq + 4 * E, dq + 8 * E, 1, 1, 0, y.data_ptr(), self.tgt_pad, 1000 + 10 * l)
            self._lin_bwd(g, dq, zin, B, pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias", cur, 3 * E, E, 1.0)
        # target embedding (no padding_idx in the reference's nn.Embedding, transformer.py:32-37)
        self._drop(ws, cur, B * E, 1)
        check(lib.slnlp_embed_gather_bwd(self._ptr("tgt_embedding.weight", g), y.data_ptr(), cur, B, 1, 1, ws.f_off, ws.f_w,
                                         ws.f_rows_tgt, 0, scale, -1, s), "embed_bwd")
        # encoder
        dxA, dxB = ws.dx[0].data_ptr(), ws.dx[1].data_ptr()
        self._ln_bwd(g, dmem, ws.enc[L - 1].x2.data_ptr(), None, "transformer.encoder.norm", ws.ln_enc, dxA, R, ws)
        cur, oth = dxA, dxB
        for l in range(L - 1, -1, -1):
            st = ws.enc[l]
            pre = f"transformer.encoder.layers.{l}."
            xin = (ws.enc[l - 1].x2 if l > 0 else ws.src).data_ptr()
            self._ln_bwd(g, cur, st.x1.data_ptr(), st.f.data_ptr(), pre + "norm2", st.ln2, oth, R, ws)
            cur, oth = oth, cur
            self._ffn_bwd(ws, g, pre, st.x1.data_ptr(), st, cur, ws.dhdn_big.data_ptr(), R, 12 + 10 * l, True)
            self._ln_bwd(g, cur, xin, st.a.data_ptr(), pre + "norm1", st.ln1, oth, R, ws)
            cur, oth = oth, cur
            da = self._masked_grad(ws, cur, ws.tmpE_big.data_ptr(), R * E, 11 + 10 * l)
            self._lin_bwd(g, da, st.o.data_ptr(), R, pre + "self_attn.out_proj.weight", pre + "self_attn.out_proj.bias",
                          ws.do_big.data_ptr(), E, E, 0.0, big=True)
            q = st.qkv.data_ptr()
            dq = ws.dqkv_big.data_ptr()
            self._mha_bwd(ws, q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, st.o.data_ptr(), ws.do_big.data_ptr(),
                          st.lse.data_ptr(), dq, dq + 4 * E, dq + 8 * E, S, S, 1, X.data_ptr(), self.src_pad, 10 + 10 * l)
            self._lin_bwd(g, dq, xin, R, pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias", cur, 3 * E, E,
                          1.0, big=True)
        self._drop(ws, cur, R * E, 0)
        check(lib.slnlp_embed_gather_bwd(self._ptr("src_embedding.weight", g), X.data_ptr(), cur, B, S, 1, ws.f_off, ws.f_w,
                                         ws.f_rows_src, 0, scale, -1, s), "embed_bwd")
        self._join_side()

    def _masked_grad(self, ws, src, tmp, n, site):
        """Gradient through a dropout site that must leave ``src`` intact (it also feeds the
        residual branch): returns ``tmp`` = mask(src) when dropout is live, else ``src``."""
        if ws.train and self.p_drop > 0.0:
            check(lib.slnlp_dropout(src, tmp, n, self.p_drop, self._rng_state().data_ptr(), site, _stream()), "dropout")
            return tmp
        return src

    # ------------------------------------------------------------------ public forward
    def forward(self, X, y, lengths=None, **kwargs):             # transformer.py:60-90
        assert X is not None, "`X` is a required paramenter"
        assert y is not None, "`y` is a required paramenter"
        if not self.batch_first and X.dim() == 2:
            X = X.t()
        self._ensure_flat()
        dev = self._flat.device
        if not (X.is_cuda and y.is_cuda):
            raise RuntimeError("slnlp_b200: inputs must be CUDA tensors (no CPU fallback)")
        X = X.to(dev, torch.int64).contiguous()
        y = y.to(dev, torch.int64).contiguous().view(-1)
        if X.dim() != 2 or y.numel() != X.shape[0]:
            raise ValueError("expected X [B,S] and y [B]")
        if self.validate_inputs:
            bad = (X < 0).any() | (X >= self.V_src).any() | (y < 0).any() | (y >= self.V_tgt).any()
            if bool(bad):
                raise ValueError("tokens must be in [0, V_src) and labels in [0, V_tgt)")
        if X.shape[1] > self.src_pos_encoding.pe.shape[0]:
            raise ValueError("sequence longer than the positional-encoding table (5000)")
        lengths = torch.empty(0, dtype=torch.int64, device=dev)   # unused by the Transformer (transformer.py:60)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params.values()):
            from .rnn import _ModuleFn
            return _ModuleFn.apply(self, X, lengths, y, *[self._params[n] for n in self._params])
        ws = self._workspace(X.shape[0], X.shape[1], self.training)
        if ws.train and self.uses_rng:
            check(lib.slnlp_rng_advance(self._rng_state().data_ptr(), _stream()), "rng")
        return self._run_forward(ws, X, lengths, y).clone()

    @torch.no_grad()
    def predict_logp(self, X, lengths, y):
        self._ensure_flat()
        ws = self._workspace(X.shape[0], X.shape[1], False)
        return self._run_forward(ws, X, lengths, y)


class _LayerState:
    pass


class _TWorkspace:
    """Activations (+ gradient scratch) of one (B, S, train) shape; torch owns the memory."""

    def __init__(self, m: TransformerB200, B, S, train, bwd):
        E, F, L, V, nh = m.E, m.F, m.L, m.V_tgt, m.nhead
        dev = m._flat.device
        f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        self.B, self.T, self.train = B, S, train
        R = B * S
        self.f_off = (ctypes.c_int64 * 1)(0)
        self.f_w = (ctypes.c_int * 1)(E)
        self.f_rows_src = (ctypes.c_int64 * 1)(m.V_src)
        self.f_rows_tgt = (ctypes.c_int64 * 1)(m.V_tgt)
        stats = lambda rows: (f(rows), f(rows))
        self.src, self.memory, self.ln_enc = f(R, E), f(R, E), stats(R)
        self.enc, self.dec = [], []
        for _ in range(L):
            st = _LayerState()
            st.qkv, st.o, st.lse, st.a = f(R, 3 * E), f(R, E), f(B, nh, S), f(R, E)
            st.x1, st.ln1, st.hdn, st.f, st.x2, st.ln2 = f(R, E), stats(R), f(R, F), f(R, E), f(R, E), stats(R)
            self.enc.append(st)
        self.tgt = f(B, E)
        for _ in range(L):
            st = _LayerState()
            st.qkv, st.o, st.lse, st.a = f(B, 3 * E), f(B, E), f(B, nh, 1), f(B, E)
            st.x1, st.ln1 = f(B, E), stats(B)
            st.qc, st.kvc, st.oc, st.lsec, st.a2 = f(B, E), f(R, 2 * E), f(B, E), f(B, nh, 1), f(B, E)
            st.x2, st.ln2, st.hdn, st.f, st.x3, st.ln3 = f(B, E), stats(B), f(B, F), f(B, E), f(B, E), stats(B)
            self.dec.append(st)
        self.zf, self.ln_dec = f(B, E), stats(B)
        self.logits, self.logp = f(B, V), f(B, V)
        if bwd:
            self.Vp = (V + 3) & ~3
            self.dlogits = torch.zeros(B, self.Vp, device=dev)
            self.dz = (f(B, E), f(B, E))
            self.dx = (f(R, E), f(R, E))
            self.dmem = f(R, E)
            self.tmpE_small, self.tmpE_big = f(B, E), f(R, E)
            self.do_small, self.do_big = f(B, E), f(R, E)
            self.dhdn_small, self.dhdn_big = f(B, F), f(R, F)
            self.dqkv_small, self.dqkv_big = f(B, 3 * E), f(R, 3 * E)
            self.dqc, self.dkvc = f(B, E), f(R, 2 * E)
            self.dvec = f(B, nh, max(S, 1))
            self.ln_part = f(lib.slnlp_ln_bwd_blocks(R), 2 * E)
            self.loss = torch.zeros(2, device=dev)
            self.row_ws = f(3 * B)
