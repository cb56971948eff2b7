"""CPU oracle: plain-tensor restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / baseline, never as the thing measured or shipped.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py)
against golden vectors produced by importing the real reference modules from
/root/reference (tests/golden/make_golden.py, committed with its outputs).
The skorch train-step order (zero_grad -> forward -> CrossEntropyLoss on the
returned log-probs -> backward -> clip_grad_norm_(0.5, 2) -> SGD step) is
restated from skorch's public source because skorch is not vendored in the
reference: that ordering alone is "parity unpinned" (SURVEY.md section 8c).

The arithmetic of nn.LSTM / nn.GRU / nn.Transformer / nn.Embedding lives in
torch (pinned by the reference to 1.12.1, pyproject.toml:16); this file spells
it out with explicit per-step cells and per-head attention on plain tensors so
that the CUDA kernels have an op-by-op specification.  All functions take a
``state_dict``-shaped mapping (same keys/shapes as the reference modules).

Reference citations use ``bkp`` = model/base/encoder_decoder_attn_bkp.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

PAD_IDX = 1  # dataset/constant/tokens.py:4 with torchtext specials order (unk=0, pad=1)
BOS_IDX = 0  # '<bos>' is not in the vocab -> defaultdict -> unk (model/util/util.py:8-9)

GATES = {"lstm": 4, "gru": 3}


# --------------------------------------------------------------------------
# recurrent cells (torch.nn.LSTM / torch.nn.GRU math; gate order i,f,g,o / r,z,n)
# --------------------------------------------------------------------------
def lstm_cell(xp, h, c, w_hh, b_hh):
    """xp = x W_ih^T + b_ih (already projected).  Returns (h', c')."""
    gates = xp + h @ w_hh.t() + b_hh
    i, f, g, o = gates.chunk(4, dim=-1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c2 = f * c + i * g
    h2 = o * torch.tanh(c2)
    return h2, c2


def gru_cell(xp, h, w_hh, b_hh):
    hp = h @ w_hh.t() + b_hh
    xr, xz, xn = xp.chunk(3, dim=-1)
    hr, hz, hn = hp.chunk(3, dim=-1)
    r = torch.sigmoid(xr + hr)
    z = torch.sigmoid(xz + hz)
    n = torch.tanh(xn + r * hn)
    return (1.0 - z) * n + z * h


def _run_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, rnn_type, reverse):
    """One direction of one packed RNN layer (bkp:110-114).

    x: [B, T, D].  State is frozen and the output is 0 for t >= len_b
    (pack_padded_sequence semantics); the reverse direction walks
    t = len_b-1 .. 0.  Returns (out [B,T,H], h_final [B,H]).
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        active = (lengths > t).unsqueeze(1)
        if rnn_type == "lstm":
            h2, c2 = lstm_cell(xp[:, t], h, c, w_hh, b_hh)
            c = torch.where(active, c2, c)
        else:
            h2 = gru_cell(xp[:, t], h, w_hh, b_hh)
        h = torch.where(active, h2, h)
        outs[t] = torch.where(active, h2, torch.zeros_like(h2))
    return torch.stack(outs, dim=1), h


def encoder_forward(sd, emb, lengths, rnn_type, num_layers, pad_fill=float(PAD_IDX),
                    dropout_masks: Optional[Sequence[torch.Tensor]] = None,
                    prefix="model.encoder.rnn."):
    """Encoder.forward + concatenate_directions (bkp:102-159).

    Returns (enc_out [B,T,2H] with pad positions filled by ``pad_fill``,
    enc_final [L,B,2H]).  ``dropout_masks[l]`` (already scaled by 1/(1-p)) is
    multiplied into the input of layer l+1 (nn.LSTM inter-layer dropout).
    """
    x = emb
    finals = []
    for l in range(num_layers):
        outs, hs = [], []
        for d, suffix in enumerate(("", "_reverse")):
            out, hfin = _run_direction(
                x, lengths,
                sd[f"{prefix}weight_ih_l{l}{suffix}"], sd[f"{prefix}weight_hh_l{l}{suffix}"],
                sd[f"{prefix}bias_ih_l{l}{suffix}"], sd[f"{prefix}bias_hh_l{l}{suffix}"],
                rnn_type, reverse=(d == 1))
            outs.append(out)
            hs.append(hfin)
        x = torch.cat(outs, dim=2)
        finals.append(torch.cat(hs, dim=1))
        if dropout_masks is not None and l < num_layers - 1:
            x = x * dropout_masks[l]
    T = emb.shape[1]
    valid = (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(2)
    enc_out = torch.where(valid, x, torch.full_like(x, pad_fill))  # bkp:120-123
    return enc_out, torch.stack(finals, dim=0)


def bahdanau_attention(sd, query, proj_key, value, src_mask,
                       prefix="model.decoder.attention."):
    """BahdanauAttention.forward (bkp:304-327).  query [B,H], proj_key [B,T,H],
    value [B,T,2H], src_mask [B,T] bool (True = valid).  Returns (ctx, alphas)."""
    q = query @ sd[prefix + "query_layer.weight"].t()
    e = torch.tanh(q.unsqueeze(1) + proj_key) @ sd[prefix + "energy_layer.weight"].t()
    e = e.squeeze(2).masked_fill(~src_mask, float("-inf"))
    alphas = torch.softmax(e, dim=-1)
    ctx = torch.bmm(alphas.unsqueeze(1), value).squeeze(1)
    return ctx, alphas


def rnn_encdec_forward(sd: Dict[str, torch.Tensor], X, lengths, rnn_type: str,
                       num_layers: int, pad_idx: int = PAD_IDX, bos_idx: int = BOS_IDX,
                       enc_dropout_masks=None, dec_dropout_masks=None,
                       return_intermediates: bool = False):
    """EncoderDecoderAttnBaseBkp.forward (bkp:388-402) for MAX_OUTPUT_LEN = 1.

    X [B,T] int64, lengths [B] int64 -> log-probabilities [B, V_tgt].
    ``y`` is not an argument: its values never reach the RNN models' output
    (SURVEY.md section 0 quirk 2).
    """
    if X.dim() == 3:
        # factored phonological embedding (no reference counterpart): one table per
        # field - orientation / movement / handshape of both hands - gathered and concatenated, torch.cat of
        # per-field nn.Embedding being its oracle; a frame is padding when its FIRST field is <pad>
        fields = X
        emb = torch.cat([sd[f"model.src_embed.fields.{i}.weight"][fields[..., i]] for i in range(fields.shape[-1])], dim=-1)
        X = fields[..., 0]
    else:
        emb = sd["model.src_embed.weight"][X]                               # bkp:49
    B, T = X.shape
    enc_out, enc_final = encoder_forward(sd, emb, lengths, rnn_type, num_layers,
                                         pad_fill=float(pad_idx),
                                         dropout_masks=enc_dropout_masks)
    # Decoder.init_hidden (bkp:268-280)
    hidden0 = torch.tanh(enc_final @ sd["model.decoder.bridge.weight"].t()
                         + sd["model.decoder.bridge.bias"])                 # [L,B,H]
    proj_key = enc_out @ sd["model.decoder.attention.key_layer.weight"].t()  # bkp:246
    src_mask = X != pad_idx                                                  # bkp:404-406
    ctx, alphas = bahdanau_attention(sd, hidden0[-1], proj_key, enc_out, src_mask)
    prev_embed = sd["model.trg_embed.weight"][bos_idx].unsqueeze(0).expand(B, -1)
    x = torch.cat([prev_embed, ctx], dim=1)                                  # bkp:215
    p = "model.decoder.rnn."
    for l in range(num_layers):
        xp = x @ sd[f"{p}weight_ih_l{l}"].t() + sd[f"{p}bias_ih_l{l}"]
        if rnn_type == "lstm":
            h, _ = lstm_cell(xp, hidden0[l], hidden0[l], sd[f"{p}weight_hh_l{l}"],
                             sd[f"{p}bias_hh_l{l}"])
        else:
            h = gru_cell(xp, hidden0[l], sd[f"{p}weight_hh_l{l}"], sd[f"{p}bias_hh_l{l}"])
        x = h
        if dec_dropout_masks is not None and l < num_layers - 1:
            x = x * dec_dropout_masks[l]
    # generator consumes decoder_states, not pre_output (bkp:40-46, 75-76)
    logits = h @ sd["model.generator.proj.weight"].t()
    logp = F.log_softmax(logits, dim=-1)
    if return_intermediates:
        return logp, dict(emb=emb, enc_out=enc_out, enc_final=enc_final, hidden0=hidden0,
                          proj_key=proj_key, alphas=alphas, ctx=ctx, h_top=h, logits=logits)
    return logp


# --------------------------------------------------------------------------
# Transformer (model/transformer.py:60-90 on top of torch.nn.Transformer,
# post-norm, ReLU, LayerNorm eps 1e-5)
# --------------------------------------------------------------------------
def positional_table(max_len: int, d_model: int) -> torch.Tensor:
    """model/component/positional_encoding.py:22-31 -> [max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _mha(x_q, x_kv, w_in, b_in, w_out, b_out, nhead, attn_bias):
    """nn.MultiheadAttention, batch-major here: x_q [B,Sq,E], x_kv [B,Sk,E];
    attn_bias [B,Sq,Sk] additive (0 / -inf) or None."""
    B, Sq, E = x_q.shape
    Sk = x_kv.shape[1]
    dh = E // nhead
    q = x_q @ w_in[:E].t() + b_in[:E]
    k = x_kv @ w_in[E:2 * E].t() + b_in[E:2 * E]
    v = x_kv @ w_in[2 * E:].t() + b_in[2 * E:]
    q = q.view(B, Sq, nhead, dh).transpose(1, 2)
    k = k.view(B, Sk, nhead, dh).transpose(1, 2)
    v = v.view(B, Sk, nhead, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if attn_bias is not None:
        s = s + attn_bias.unsqueeze(1)
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Sq, E)
    return o @ w_out.t() + b_out


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def transformer_forward(sd, X, y, nhead: int, num_layers: int, pad_idx: int = PAD_IDX,
                        return_intermediates: bool = False):
    """Transformer.forward (model/transformer.py:60-90), dropout = 0.

    X [B,S] int64, y [B] int64 -> log-probs [B, V_tgt].  Quirks kept
    (SURVEY.md section 0 quirk 7): causal mask on the ENCODER self-attention,
    the true label y is the one-token decoder input, memory is not
    key-padding-masked in cross-attention.
    """
    B, S = X.shape
    E = sd["src_embedding.weight"].shape[1]
    pe = positional_table(S, E)
    src = sd["src_embedding.weight"][X] * math.sqrt(E) + pe.unsqueeze(0)      # :106-109
    tgt = sd["tgt_embedding.weight"][y].unsqueeze(1) * math.sqrt(E) + pe[:1].unsqueeze(0)
    causal = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)        # util.py:11-42
    keypad = (X == pad_idx)                                                    # util.py:45-61
    bias = torch.zeros(B, S, S)
    bias = bias.masked_fill(causal.unsqueeze(0) | keypad.unsqueeze(1), float("-inf"))
    x = src
    for l in range(num_layers):
        p = f"transformer.encoder.layers.{l}."
        a = _mha(x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                 sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"],
                 nhead, bias)
        x = _ln(x + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        f = torch.relu(x @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
        f = f @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"]
        x = _ln(x + f, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    memory = _ln(x, sd["transformer.encoder.norm.weight"], sd["transformer.encoder.norm.bias"])
    tbias = torch.zeros(B, 1, 1).masked_fill((y == pad_idx).view(B, 1, 1), float("-inf"))
    z = tgt
    for l in range(num_layers):
        p = f"transformer.decoder.layers.{l}."
        a = _mha(z, z, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                 sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"],
                 nhead, tbias)
        z = _ln(z + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        a = _mha(z, memory, sd[p + "multihead_attn.in_proj_weight"],
                 sd[p + "multihead_attn.in_proj_bias"], sd[p + "multihead_attn.out_proj.weight"],
                 sd[p + "multihead_attn.out_proj.bias"], nhead, None)
        z = _ln(z + a, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        f = torch.relu(z @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
        f = f @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"]
        z = _ln(z + f, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    z = _ln(z, sd["transformer.decoder.norm.weight"], sd["transformer.decoder.norm.bias"])
    logits = z.squeeze(1) @ sd["linear.weight"].t() + sd["linear.bias"]
    logp = F.log_softmax(logits, dim=-1)
    if return_intermediates:
        return logp, dict(src=src, memory=memory, logits=logits)
    return logp


# --------------------------------------------------------------------------
# criterion / clip / optimizer (config/*.yaml:19-20,36,39-42; helper.py:62-70,227-229)
# --------------------------------------------------------------------------
def criterion(logp, y, ignore_index: int = PAD_IDX):
    """CrossEntropyLoss(ignore_index=pad) applied on the module's log-probs:
    a second log_softmax, then the mean of -logp[y] over y != ignore_index."""
    lp2 = logp - torch.logsumexp(logp, dim=-1, keepdim=True)
    valid = y != ignore_index
    picked = lp2.gather(1, y.clamp(min=0).unsqueeze(1)).squeeze(1)
    return -(picked * valid).sum() / valid.sum()


def clip_grad_norm(grads: Dict[str, torch.Tensor], max_norm: float = 0.5):
    """torch.nn.utils.clip_grad_norm_(max_norm, norm_type=2) over non-None grads.
    Returns (total_norm, clipped grads)."""
    total = torch.linalg.vector_norm(
        torch.stack([torch.linalg.vector_norm(g) for g in grads.values()]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


def sgd_momentum_step(params, grads, bufs, lr: float, momentum: float = 0.9):
    """torch.optim.SGD(momentum, dampening 0, nesterov False, weight_decay 0).
    ``bufs`` is mutated; a missing buffer means first step (buf = grad)."""
    out = {}
    for k, p in params.items():
        if k not in grads:
            out[k] = p
            continue
        g = grads[k]
        bufs[k] = g.clone() if k not in bufs else bufs[k] * momentum + g
        out[k] = p - lr * bufs[k]
    return out


def train_step(sd, bufs, forward_fn, y, lr, max_norm=0.5, momentum=0.9,
               ignore_index: int = PAD_IDX) -> Tuple[Dict[str, torch.Tensor], float, float]:
    """One skorch-equivalent training step.

    forward_fn(params) -> logp.  Returns (new params, loss, total grad norm).
    Parameters whose gradient is None (the dead pre_output_layer, SURVEY.md
    quirk 1) are excluded from the norm and left untouched, as in torch.
    """
    names = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(".pe")]
    params = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(params)
    loss = criterion(forward_fn(full), y, ignore_index)
    gl = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    grads = {k: g for k, g in zip(names, gl) if g is not None}
    total, grads = clip_grad_norm(grads, max_norm)
    new = sgd_momentum_step({k: params[k].detach() for k in names}, grads, bufs, lr, momentum)
    out = dict(sd)
    out.update(new)
    return out, float(loss.detach()), float(total)
