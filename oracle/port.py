"""CPU oracle, library form: the reference models rebuilt on stock torch.nn.

TEST INFRASTRUCTURE ONLY (see oracle/restatement.py for the import rule).

This is the form the reference itself runs: torch.nn.LSTM / GRU on packed
sequences, torch.nn.Transformer, nn.Embedding, nn.Linear.  It is what
``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time on the GPU
box's host cores (kind = "port": the Python reference cannot travel to the
box), and what the GPU parity tests compare against at full size.  Parameter
names and shapes equal the reference's ``state_dict``
so weights move between the reference, this port and the CUDA modules with
``load_state_dict``.

Parity status: PINNED by tests/test_oracle_golden.py against golden vectors
generated from the real reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from .restatement import BOS_IDX, PAD_IDX, positional_table


class _Holder(nn.Module):
    """Gives the parameters the reference's ``model.<sub>.<name>`` prefixes."""


class RnnEncDecPort(nn.Module):
    """EncoderDecoderAttnBaseBkp (bkp:330-413) with MAX_OUTPUT_LEN = 1."""

    def __init__(self, rnn_type, src_vocab_size, tgt_vocab_size, embedding_size, hidden_size,
                 num_layers, dropout, pad_idx=PAD_IDX, bos_idx=BOS_IDX):
        super().__init__()
        rnn = {"lstm": nn.LSTM, "gru": nn.GRU}[rnn_type]
        E, H, L = embedding_size, hidden_size, num_layers
        p = dropout if L > 1 else 0.0
        self.pad_idx, self.bos_idx, self.L = pad_idx, bos_idx, L
        m = self.model = _Holder()
        m.encoder = _Holder()
        m.encoder.rnn = rnn(E, H, L, batch_first=True, bidirectional=True, dropout=p)
        m.decoder = _Holder()
        m.decoder.attention = _Holder()
        m.decoder.attention.key_layer = nn.Linear(2 * H, H, bias=False)
        m.decoder.attention.query_layer = nn.Linear(H, H, bias=False)
        m.decoder.attention.energy_layer = nn.Linear(H, 1, bias=False)
        m.decoder.rnn = rnn(E + 2 * H, H, L, batch_first=True, dropout=p)
        m.decoder.bridge = nn.Linear(2 * H, H, bias=True)
        m.decoder.dropout_layer = nn.Dropout(dropout)
        m.decoder.pre_output_layer = nn.Linear(3 * H + E, H, bias=False)
        m.src_embed = nn.Embedding(src_vocab_size, E, padding_idx=pad_idx)
        m.trg_embed = nn.Embedding(tgt_vocab_size, E, padding_idx=pad_idx)
        m.generator = _Holder()
        m.generator.proj = nn.Linear(H, tgt_vocab_size, bias=False)

    def forward(self, X, y=None, lengths=None):
        m = self.model
        B, T = X.shape
        packed = pack_padded_sequence(m.src_embed(X), lengths.cpu(), batch_first=True,
                                      enforce_sorted=False)
        out, hid = m.encoder.rnn(packed)
        if isinstance(hid, tuple):
            hid = hid[0]
        enc_out, _ = pad_packed_sequence(out, batch_first=True, total_length=T,
                                         padding_value=self.pad_idx)
        enc_final = torch.cat([hid[0::2], hid[1::2]], dim=2)
        h0 = torch.tanh(m.decoder.bridge(enc_final))
        state = (h0, h0) if isinstance(m.decoder.rnn, nn.LSTM) else h0
        key = m.decoder.attention.key_layer(enc_out)
        q = m.decoder.attention.query_layer(h0[-1].unsqueeze(1))
        e = m.decoder.attention.energy_layer(torch.tanh(q + key)).squeeze(2)
        e = e.masked_fill(X == self.pad_idx, float("-inf"))
        alpha = F.softmax(e, dim=-1).unsqueeze(1)
        ctx = torch.bmm(alpha, enc_out)
        bos = torch.full((B, 1), self.bos_idx, dtype=torch.long, device=X.device)
        prev = m.trg_embed(bos)
        dec_out, _ = m.decoder.rnn(torch.cat([prev, ctx], dim=2), state)
        # the reference also evaluates (and discards) the pre-output branch, bkp:218-220
        m.decoder.pre_output_layer(m.decoder.dropout_layer(torch.cat([prev, dec_out, ctx], dim=2)))
        return F.log_softmax(m.generator.proj(dec_out), dim=-1)[:, -1]


class _PE(nn.Module):
    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("pe", positional_table(max_len, d_model).unsqueeze(1))

    def forward(self, x):
        return self.dropout(x + self.pe[:x.size(0)])


class TransformerPort(nn.Module):
    """model/transformer.py:9-109."""

    def __init__(self, src_vocab_size, tgt_vocab_size, embedding_size, num_heads, num_layers,
                 hidden_size, dropout, pad_idx=PAD_IDX):
        super().__init__()
        E = embedding_size
        self.E, self.pad_idx = E, pad_idx
        self.src_embedding = nn.Embedding(src_vocab_size, E)
        self.src_pos_encoding = _PE(E, dropout)
        self.tgt_embedding = nn.Embedding(tgt_vocab_size, E)
        self.tgt_pos_encoding = _PE(E, dropout)
        self.transformer = nn.Transformer(d_model=E, nhead=num_heads, num_encoder_layers=num_layers,
                                          num_decoder_layers=num_layers,
                                          dim_feedforward=hidden_size, dropout=dropout)
        self.linear = nn.Linear(E, tgt_vocab_size)

    def forward(self, X, y, lengths=None):
        src, tgt = X.t(), y.unsqueeze(0)
        S = src.size(0)
        src_mask = torch.triu(torch.ones(S, S, dtype=torch.bool, device=X.device), diagonal=1)
        tgt_mask = torch.zeros(1, 1, dtype=torch.bool, device=X.device)
        s = self.src_pos_encoding(self.src_embedding(src) * math.sqrt(self.E))
        t = self.tgt_pos_encoding(self.tgt_embedding(tgt) * math.sqrt(self.E))
        out = self.transformer(src=s, tgt=t, src_mask=src_mask, tgt_mask=tgt_mask,
                               src_key_padding_mask=(X == self.pad_idx),
                               tgt_key_padding_mask=(y == self.pad_idx).unsqueeze(1))
        return F.log_softmax(self.linear(out), dim=-1).squeeze(0)


def build_port(kind, src_vocab_size, tgt_vocab_size, embedding_size, hidden_size, num_layers,
               dropout, num_heads=None):
    if kind in ("lstm", "gru"):
        return RnnEncDecPort(kind, src_vocab_size, tgt_vocab_size, embedding_size, hidden_size,
                             num_layers, dropout)
    return TransformerPort(src_vocab_size, tgt_vocab_size, embedding_size, num_heads, num_layers,
                           hidden_size, dropout)


def reference_train_step(module, optimizer, X, y, lengths, max_norm=0.5, ignore_index=PAD_IDX):
    """skorch train_step_single + GradientNormClipping + optimizer step."""
    module.train()
    optimizer.zero_grad()
    loss = F.cross_entropy(module(X=X, y=y, lengths=lengths), y, ignore_index=ignore_index)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(module.parameters(), max_norm=max_norm, norm_type=2)
    optimizer.step()
    return loss
