"""CPU oracle for the piano-a2s hot path (ConvStack -> BiGRU encoder -> hierarchical decoder).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`piano_a2s_b200/`, root
`models.py`) may import this file; it is used by `tests/`, by
`__graft_entry__.smoke()` as the checker, and by `bench.py`'s `cpu_baseline` /
`--impl reference` legs as the timed CPU restatement.

It is a functional restatement, in plain fp32 PyTorch-CPU ops on an explicit
``state_dict``, of the algorithm in the reference's ``models.py`` (file:line
citations are into /root/reference).  It follows the reference *literally* --
including the per-step ``cat`` + ``Linear(1024->256)`` attention recompute, the
absence of an attention length mask, and "finished sequences keep decoding" --
so that it is a faithful CPU baseline as well as the parity checker.

Pinning: `tests/test_oracle_golden.py` imports the unmodified reference
(`/root/reference/models.py`, with a stub for its top-level `music21` import)
in the build container and checks this file against it for eval, greedy and
teacher-forced training forwards and for all parameter gradients; the golden
vectors under `tests/golden/` were produced from the reference by
`tests/golden/make_golden.py` and are what travels to the GPU box.

Randomness.  The reference draws one python ``random.random()`` coin per executed
note step and one per bar (models.py:404, :289) and calls ``F.dropout`` on the token
embeddings (models.py:239, :391) and on the conv features (models.py:541).  All of
that goes through a ``RandomSource`` so that a test can replay the *same* coins and
masks through the CUDA path and through this oracle.
"""
from __future__ import annotations

import random as _pyrandom
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Vocabulary constants (data_processing/humdrum.py:70-97; models.py:9-12).
# 148 base symbols + 25 extended = 173; <sos>,<eos>,<pad> are the last three base symbols.
# --------------------------------------------------------------------------------------
VOCAB_SIZE = 173
SOS = 145
EOS = 146
PAD = 147


class RandomSource:
    """Reference-order randomness: python `random` coins + torch `F.dropout` masks."""

    def coin(self) -> float:                       # models.py:289, :404
        return _pyrandom.random()

    def dropout_mask(self, shape, p: float) -> torch.Tensor:
        # F.dropout(x, p, training=True) == x * mask with mask in {0, 1/(1-p)}; drawing the mask on a
        # tensor of ones consumes the torch CPU generator exactly like the reference call does.
        return F.dropout(torch.ones(shape), p=p, training=True)


class RecordingSource(RandomSource):
    """Draws like RandomSource and remembers everything (to replay through the CUDA path)."""

    def __init__(self):
        self.coins: List[float] = []
        self.masks: List[torch.Tensor] = []

    def coin(self):
        c = super().coin()
        self.coins.append(c)
        return c

    def dropout_mask(self, shape, p):
        m = super().dropout_mask(shape, p)
        self.masks.append(m)
        return m


class ReplaySource(RandomSource):
    """Replays recorded coins / masks in order."""

    def __init__(self, coins: Sequence[float], masks: Sequence[torch.Tensor]):
        self.coins = list(coins)
        self.masks = [m.detach().cpu().float() for m in masks]
        self.ci = 0
        self.mi = 0

    def coin(self):
        c = self.coins[self.ci]
        self.ci += 1
        return c

    def dropout_mask(self, shape, p):
        m = self.masks[self.mi]
        self.mi += 1
        assert tuple(m.shape) == tuple(shape), (m.shape, shape)
        return m


# --------------------------------------------------------------------------------------
# Building blocks
# --------------------------------------------------------------------------------------
def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5, new_stats=None):
    """BatchNorm{1,2}d over dim 1 (models.py:499-505, used at :525-539)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, eps)
        if new_stats is not None:
            new_stats[prefix + ".running_mean"] = rm2
            new_stats[prefix + ".running_var"] = rv2
            new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)


def conv_stack(x, sd, training, rnd: RandomSource, new_stats=None, taps=None, prefix="convstack."):
    """ConvStack.forward (models.py:523-543).  x: (B,1,T,F) -> (B,T,256)."""
    for i in (1, 2, 3, 4):
        x = F.conv2d(x, sd[f"{prefix}conv{i}.weight"], None, 1, 1)
        x = F.relu(_bn(x, sd, f"{prefix}bn{i}", training, new_stats=new_stats))
        if taps is not None:
            taps[f"conv{i}"] = x
    x = x.transpose(1, 2).flatten(2)                                  # (B,T,40*F), feature = c*F+f
    x = F.linear(x, sd[prefix + "out.weight"])                        # (B,T,256)
    x = _bn(x.transpose(1, 2), sd, prefix + "out_bn", training, new_stats=new_stats).transpose(1, 2)
    x = F.relu(x)
    if training:
        # models.py:541.  In the reference x is a transposed view of a (B,256,T) buffer and F.dropout fills
        # its bernoulli mask in that memory order, so the mask is drawn as (B,256,T) and viewed transposed.
        B_, T_, C_ = x.shape
        x = x * rnd.dropout_mask((B_, C_, T_), 0.2).transpose(1, 2)
    return x


def _gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """One torch-convention GRU step, gate order (r, z, n)."""
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    H = h.shape[-1]
    r = torch.sigmoid(gi[..., :H] + gh[..., :H])
    z = torch.sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H])
    n = torch.tanh(gi[..., 2 * H:] + r * gh[..., 2 * H:])
    return (1 - z) * n + z * h


def _gru_dir(x, sd, prefix, suffix, reverse, lengths=None):
    """One direction of one GRU layer over (B,T,I).  With `lengths`, packed-sequence semantics:
    sample b is stepped only for t < lengths[b] (reverse direction starts at lengths[b]-1)."""
    w_ih, w_hh = sd[f"{prefix}weight_ih_{suffix}"], sd[f"{prefix}weight_hh_{suffix}"]
    b_ih, b_hh = sd[f"{prefix}bias_ih_{suffix}"], sd[f"{prefix}bias_hh_{suffix}"]
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        hn = _gru_cell(x[:, t], h, w_ih, w_hh, b_ih, b_hh)
        if lengths is not None:
            m = (t < lengths).to(x.dtype).unsqueeze(1)
            hn = m * hn + (1 - m) * h
        h = hn
        outs[t] = h
    return torch.stack(outs, 1), h


def encoder(x, sd, prefix="encoder."):
    """Encoder.forward (models.py:75-82): 2-layer BiGRU + shared-fc tanh bridge."""
    hs = []
    for layer in (0, 1):
        of, hf = _gru_dir(x, sd, prefix + "gru.", f"l{layer}", False)
        ob, hb = _gru_dir(x, sd, prefix + "gru.", f"l{layer}_reverse", True)
        x = torch.cat([of, ob], -1)
        hs += [hf, hb]
    fcw, fcb = sd[prefix + "fc.weight"], sd[prefix + "fc.bias"]
    h1 = torch.tanh(F.linear(torch.cat((hs[0], hs[1]), 1), fcw, fcb))
    h2 = torch.tanh(F.linear(torch.cat((hs[2], hs[3]), 1), fcw, fcb))
    return x, torch.cat((h1, h2), 1).unsqueeze(0)                      # (B,T,2H), (1,B,2H)


def attention(hidden, enc, sd, prefix):
    """AttentionLayer.forward (models.py:452-461); literal form, no length mask."""
    T = enc.shape[1]
    h = hidden.transpose(0, 1).repeat(1, T, 1)
    energy = torch.tanh(F.linear(torch.cat((h, enc), 2), sd[prefix + "attn.weight"], sd[prefix + "attn.bias"]))
    att = F.linear(energy, sd[prefix + "v.weight"]).squeeze(2)
    return F.softmax(att, dim=1)


def note_decoder(enc, hidden, sd, prefix, max_steps, inference, ground_truth, tf_ratio, training,
                 rnd: RandomSource, trace=None):
    """NoteDecoder.decode_notes (models.py:366-420)."""
    B = enc.shape[0]
    emb = sd[prefix + "embedding.weight"]
    token = emb[torch.full((B, 1), SOS, dtype=torch.long)]                       # (B,1,E)
    score_probs = torch.zeros(B, max_steps, VOCAB_SIZE)
    eos = torch.zeros(B)
    lengths = torch.full((B,), max_steps, dtype=torch.long)
    steps = 0
    for t in range(max_steps):
        if eos.sum() == B:                                                       # models.py:389
            break
        steps += 1
        if training:
            token = token * rnd.dropout_mask(token.shape, 0.1)                   # models.py:391
        a = attention(hidden, enc, sd, prefix + "attn.").unsqueeze(1)
        context = torch.bmm(a, enc)                                              # (B,1,2H)
        x = torch.cat([token, context], 2)
        h = _gru_cell(x[:, 0], hidden[0], sd[prefix + "gru.weight_ih_l0"], sd[prefix + "gru.weight_hh_l0"],
                      sd[prefix + "gru.bias_ih_l0"], sd[prefix + "gru.bias_hh_l0"])
        hidden = h.unsqueeze(0)
        out = F.linear(torch.cat([h.unsqueeze(1), context], -1), sd[prefix + "out.weight"], sd[prefix + "out.bias"])
        prob = F.log_softmax(out, dim=-1)
        score_probs[:, t, :] = prob.squeeze(1)
        tf = rnd.coin() < tf_ratio                                               # models.py:404
        am = torch.argmax(prob, dim=-1)                                          # (B,1)
        if (not inference) and tf:
            token = emb[ground_truth[:, t].unsqueeze(1)]
        else:
            token = emb[am]
        if ground_truth is not None:                                             # models.py:411-419
            hit = ground_truth[:, t] == EOS
        else:
            hit = am[:, 0] == EOS
        eos[hit] = 1
        lengths[hit] = t + 1
    if trace is not None:
        trace.append(steps)
    return score_probs, lengths


def _staff_token(tokens, lengths, sd, prefix="decoder."):
    """get_staff_token_from_{probs,gt} (models.py:164-189): note_emb -> packed BiGRU(16->32) -> h_n."""
    x = sd[prefix + "note_emb.weight"][tokens]
    _, hf = _gru_dir(x, sd, prefix + "staff_emb.", "l0", False, lengths)
    _, hb = _gru_dir_packed_reverse(x, sd, prefix + "staff_emb.", lengths)
    return torch.cat([hf, hb], 1).unsqueeze(1)                                   # (B,1,2*S)


def _gru_dir_packed_reverse(x, sd, prefix, lengths):
    """Reverse direction of a packed sequence: sample b runs t = lengths[b]-1 .. 0 from h=0."""
    w_ih, w_hh = sd[prefix + "weight_ih_l0_reverse"], sd[prefix + "weight_hh_l0_reverse"]
    b_ih, b_hh = sd[prefix + "bias_ih_l0_reverse"], sd[prefix + "bias_hh_l0_reverse"]
    B, T, _ = x.shape
    h = x.new_zeros(B, w_hh.shape[1])
    for t in range(int(lengths.max()) - 1, -1, -1):
        hn = _gru_cell(x[:, t], h, w_ih, w_hh, b_ih, b_hh)
        m = (t < lengths).to(x.dtype).unsqueeze(1)
        h = m * hn + (1 - m) * h
    return None, h


def _mlp_head(x, sd, prefix):
    """time_sig_out / key_out (models.py:123-132)."""
    x = F.relu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    x = F.relu(F.linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"]))
    return F.linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def hierarchical_decoder(enc, hidden, sd, cfg, inference, ground_truth, tf_ratio, training,
                         rnd: RandomSource, prefix="decoder.", trace=None):
    """HierarchicalDecoder.decode_bars (models.py:191-316)."""
    B = enc.shape[0]
    max_bars, n_ts, n_key = cfg["max_bars"], cfg["num_time_sig"], cfg["num_keys"]
    L_up, L_lo = cfg["max_length"]
    if inference:
        assert tf_ratio == 0 and ground_truth is None
    if ground_truth is not None:
        ts_gt, key_gt, up_gt, up_len_gt, lo_gt, lo_len_gt = ground_truth
    # SOS token (models.py:141-162): staff summary of [<sos>,<eos>] for both staves + SOS embeddings.
    se = torch.tensor([[SOS, EOS]]).repeat(B, 1)
    st = _staff_token(se, torch.full((B,), 2, dtype=torch.long), sd, prefix)
    ts_tok = sd[prefix + "time_sig_emb.weight"][torch.full((B, 1), n_ts, dtype=torch.long)]
    key_tok = sd[prefix + "key_emb.weight"][torch.full((B, 1), n_key, dtype=torch.long)]
    token = torch.cat([st, st, ts_tok, key_tok], -1)
    ts_outs = torch.zeros(B, max_bars, n_ts)
    key_outs = torch.zeros(B, max_bars, n_key)
    up_outs = torch.zeros(B, max_bars, L_up, VOCAB_SIZE)
    lo_outs = torch.zeros(B, max_bars, L_lo, VOCAB_SIZE)
    for bar in range(max_bars):
        if training:
            token = token * rnd.dropout_mask(token.shape, 0.1)                   # models.py:239
        a = attention(hidden, enc, sd, prefix + "attn.").unsqueeze(1)
        context = torch.bmm(a, enc)
        x = torch.cat([token, context], 2)
        h = _gru_cell(x[:, 0], hidden[0], sd[prefix + "gru.weight_ih_l0"], sd[prefix + "gru.weight_hh_l0"],
                      sd[prefix + "gru.bias_ih_l0"], sd[prefix + "gru.bias_hh_l0"])
        hidden = h.unsqueeze(0)
        bar_summary = h.unsqueeze(1)
        g_up = up_gt[:, bar, :] if ground_truth is not None else None
        g_lo = lo_gt[:, bar, :] if ground_truth is not None else None
        tf_in = tf_ratio if ground_truth is not None else 0.0
        up_p, up_len = note_decoder(enc, bar_summary.transpose(0, 1), sd, prefix + "upper_decoder.", L_up,
                                    inference, g_up, tf_in, training, rnd, trace)
        lo_p, lo_len = note_decoder(enc, bar_summary.transpose(0, 1), sd, prefix + "lower_decoder.", L_lo,
                                    inference, g_lo, tf_in, training, rnd, trace)
        up_outs[:, bar] = up_p
        lo_outs[:, bar] = lo_p
        head_in = torch.cat([bar_summary.squeeze(1), context.squeeze(1)], 1)
        ts_outs[:, bar] = F.log_softmax(_mlp_head(head_in, sd, prefix + "time_sig_out."), -1)
        key_outs[:, bar] = F.log_softmax(_mlp_head(head_in, sd, prefix + "key_out."), -1)
        tf = rnd.coin() < tf_ratio                                               # models.py:289
        if tf and not inference:
            us = _staff_token(up_gt[:, bar, :], up_len_gt[:, bar], sd, prefix)
            ls = _staff_token(lo_gt[:, bar, :], lo_len_gt[:, bar], sd, prefix)
            ts_tok = sd[prefix + "time_sig_emb.weight"][ts_gt[:, bar]].unsqueeze(1)
            key_tok = sd[prefix + "key_emb.weight"][key_gt[:, bar]].unsqueeze(1)
        else:
            us = _staff_token(torch.argmax(up_p, -1), up_len, sd, prefix)
            ls = _staff_token(torch.argmax(lo_p, -1), lo_len, sd, prefix)
            ts_tok = sd[prefix + "time_sig_emb.weight"][torch.argmax(ts_outs[:, bar], -1)].unsqueeze(1)
            key_tok = sd[prefix + "key_emb.weight"][torch.argmax(key_outs[:, bar], -1)].unsqueeze(1)
        token = torch.cat([us, ls, ts_tok, key_tok], -1)
    return ts_outs, key_outs, up_outs, lo_outs


DEFAULT_CFG = dict(in_channels=1, freq_bins=480, conv_feature_size=256, hidden_size=256, max_bars=5,
                   num_time_sig=7, num_keys=14, max_length=(398, 189), note_emb_size=16, staff_emb_size=32,
                   time_sig_emb_size=5, key_emb_size=8)               # hparams/pretrain.yaml:84-96


def score_transcription(sd, spectrogram, cfg=None, inference=True, ground_truth=None, tf_ratio=0.0,
                        training=False, rnd: Optional[RandomSource] = None, new_stats=None, taps=None, trace=None):
    """ScoreTranscription.forward (models.py:26-51) on an explicit state_dict."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    rnd = rnd or RandomSource()
    conv = conv_stack(spectrogram, sd, training, rnd, new_stats, taps)
    enc, hidden = encoder(conv, sd)
    if taps is not None:
        taps["convstack"] = conv
        taps["encoder"] = enc
        taps["hidden"] = hidden
    return hierarchical_decoder(enc, hidden, sd, cfg, inference, ground_truth, tf_ratio, training, rnd, trace=trace)


def training_loss(outs, ground_truth):
    """ASR.compute_objectives (pretrain.py:56-93): NLL(time)+NLL(key)+NLL_ignore147(upper)+NLL_ignore147(lower)."""
    ts, key, up, lo = outs
    ts_gt, key_gt, up_gt, _, lo_gt, _ = ground_truth
    loss = F.nll_loss(ts.permute(0, 2, 1), ts_gt) + F.nll_loss(key.permute(0, 2, 1), key_gt)
    up2 = up.reshape(up.shape[0] * up.shape[1], -1, up.shape[3])
    lo2 = lo.reshape(lo.shape[0] * lo.shape[1], -1, lo.shape[3])
    loss = loss + F.nll_loss(up2.permute(0, 2, 1), up_gt.reshape(up2.shape[0], -1), ignore_index=PAD)
    loss = loss + F.nll_loss(lo2.permute(0, 2, 1), lo_gt.reshape(lo2.shape[0], -1), ignore_index=PAD)
    return loss


def unpad(seq):
    """pretrain.py:245-249: cut a token row at its first <eos>."""
    seq = [int(v) for v in seq]
    return seq[:seq.index(EOS)] if EOS in seq else seq


def greedy_tokens(outs):
    """argmax -> unpad, as recorded by compute_objectives (pretrain.py:97-117)."""
    ts, key, up, lo = outs
    return dict(upper=[[unpad(r) for r in b] for b in up.argmax(-1).tolist()],
                lower=[[unpad(r) for r in b] for b in lo.argmax(-1).tolist()],
                key=key.argmax(-1).tolist(), time_sig=ts.argmax(-1).tolist())


def adadelta_step(params, grads, state, lr=1.0, rho=0.95, eps=1e-8, max_grad_norm=5.0):
    """fit_batch tail (pretrain.py:125-128 + pretrain.yaml:44-47): clip_grad_norm_(5.0) [speechbrain default,
    recalled] then torch.optim.Adadelta(lr=1, rho=.95, eps=1e-8).  `state[k] = (square_avg, acc_delta)`."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_grad_norm / (total + 1e-6), max=1.0)
    for k, p in params.items():
        g = grads[k] * coef
        sq, acc = state[k]
        sq.mul_(rho).addcmul_(g, g, value=1 - rho)
        delta = (acc + eps).sqrt() / (sq + eps).sqrt() * g
        acc.mul_(rho).addcmul_(delta, delta, value=1 - rho)
        p.sub_(lr * delta)
    return total
