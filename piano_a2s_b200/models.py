"""Drop-in replacement for the reference's `models.py` (wei-zeng98/piano-a2s @ ca8bc59), B200-native.

Same public names, constructor arguments (hparams/pretrain.yaml:84-96 == finetune.yaml), forward signatures and
state_dict keys as /root/reference/models.py, so `!new:models.ScoreTranscription` and checkpoints keep working.
The nn.Module children are *parameter holders* created in the reference's order with the reference's initialisers
(identical weights under the same torch seed); every forward dispatches to libpa2s kernels (piano_a2s_b200.ops).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import hashlib
import math

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, rng

# ---------------------------------------------------------------------------------------------------------------
# Vocabulary (data_processing/humdrum.py:70-131 `LabelsMultiple(extended=True)`): 148 + 25 = 173 symbols.
# ---------------------------------------------------------------------------------------------------------------
def _pitches(letters, n):
    return [l * n + acc for l in letters for acc in ("-", "", "#")]


class LabelsMultiple:
    def __init__(self, extended=False):
        durations = ["1", "1.", "2", "2.", "4", "4.", "8", "8.", "16", "16.", "32", "32.", "64", "64.", "3", "6", "12", "24", "48", "96"]
        low = ["BBB#"] + _pitches("CDEFGAB", 2)[1:]                       # "CC-" only exists in the extended set
        mid = _pitches("CDEFGAB", 1) + _pitches("cdefgab", 1) + _pitches("cdefgab", 2) + _pitches("cdefgab", 3)
        top = _pitches("cdef", 4)[:-1]                                     # ... "ffff-", "ffff"
        control = ["r", ".", "[", "_", "]", ";", "\t", "\n", "<b>", "<sos>", "<eos>", "<pad>"]
        self.labels = durations + low + mid + top + control
        if extended:
            self.labels += ["128", "20", "40", "176", "112"] + _pitches("CDEFGAB", 3)[1:-1] + ["CC-"]
        self.labels_map = {c: i for i, c in enumerate(self.labels)}
        self.labels_map_inv = {i: c for i, c in enumerate(self.labels)}

    def decode(self, tokens):
        decoded = list(filter(None, [self.labels_map_inv.get(t) for t in tokens]))
        return [i if i != "<b>" else " " for i in decoded]

    def checksum(self):
        return hashlib.sha256("\x00".join(self.labels).encode()).hexdigest()


labels = LabelsMultiple(extended=True)
SOS = labels.labels_map["<sos>"]
EOS = labels.labels_map["<eos>"]
PAD = labels.labels_map["<pad>"]
vocab_size = len(labels.labels_map)


def _dev(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("piano_a2s_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    return d


class ScoreTranscription(nn.Module):
    def __init__(self, in_channels=1, freq_bins=480, conv_feature_size=256, hidden_size=256, max_bars=5, num_time_sig=7,
                 num_keys=14, max_length=(437, 129), note_emb_size=16, staff_emb_size=32, time_sig_emb_size=5, key_emb_size=8):
        super().__init__()
        self.convstack = ConvStack(in_channels, freq_bins, conv_feature_size)
        self.encoder = Encoder(conv_feature_size, hidden_size)
        self.decoder = HierarchicalDecoder(max_bars, num_time_sig, num_keys, hidden_size, max_length, note_emb_size,
                                           staff_emb_size, time_sig_emb_size, key_emb_size)

    def forward(self, spectrogram, inference=True, ground_truth=None, teacher_forcing_ratio=0.,
                device=torch.device("cuda" if torch.cuda.is_available() else "cpu")):
        self.device = device
        with ops.module_precision(self):
            if ground_truth is not None and not inference:
                # executed decoder steps per (staff, bar) are a function of the targets alone: fetch them on a side stream now,
                # so that the one host read the decoder needs does not wait for the ConvStack / encoder kernels queued below
                self.decoder.prefetch_steps(ground_truth)
            # (always reassigned: a forward that raised before the decoder must not leave its aliases to the next graph)
            grad_cuda = torch.is_grad_enabled() and spectrogram.is_cuda
            if ops.DECODER_IMPL == "multi":
                self.decoder._bar_lins = self.decoder.bar_linears() if grad_cuda else None
                self.decoder._presunk = None
            else:
                self.decoder._presunk = self.decoder.weight_sinks() if grad_cuda else None
            conv_outputs = self.convstack(spectrogram)                   # (B, T, conv_feature_size)
            encoder_outputs, hidden = self.encoder(conv_outputs)         # (B, T, 2H), (1, B, 2H)
            return self.decoder(encoder_outputs, hidden, inference, ground_truth, teacher_forcing_ratio, device)


class Encoder(nn.Module):
    def __init__(self, input_size=200, hidden_size=200):
        super().__init__()
        self.gru = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=2, batch_first=True, bidirectional=True)
        self.fc = nn.Linear(hidden_size * 2, hidden_size)
        self.init_weight()

    def init_weight(self):
        init_gru(self.gru)
        init_layer(self.fc)

    def forward(self, x):
        with ops.module_precision(self):
            return self._forward(x)

    def _forward(self, x):
        g = self.gru
        B = x.shape[0]
        bg = 4 if B >= 4 else (2 if B >= 2 else 1)
        bg = min(bg, int(os.environ.get("PA2S_GRU_BG", bg)))               # clips per cluster of the recurrence kernels (measurement switch)
        hs = []
        for layer in range(g.num_layers):
            p = [getattr(g, f"{n}_l{layer}{sfx}") for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
            x, hN = ops.BiGRULayerFn.apply(x, *p, bg)
            hs.append(hN)
        hidden1 = torch.tanh(ops.linear(torch.cat((hs[0][0], hs[0][1]), dim=1), self.fc.weight, self.fc.bias))
        hidden2 = torch.tanh(ops.linear(torch.cat((hs[1][0], hs[1][1]), dim=1), self.fc.weight, self.fc.bias))
        hidden = torch.cat((hidden1, hidden2), dim=1).unsqueeze(0)       # (1, B, 2H)
        return x, hidden


class HierarchicalDecoder(nn.Module):
    def __init__(self, max_bars=8, num_time_sig=9, num_keys=14, hidden_size=200, max_length=(437, 129), note_emb_size=16,
                 staff_emb_size=16, time_sig_emb_size=8, key_emb_size=8):
        super().__init__()
        self.max_bars = max_bars
        self.num_time_sig = num_time_sig
        self.time_sig_SOS = num_time_sig
        self.num_keys = num_keys
        self.key_SOS = num_keys
        self.hidden_size = hidden_size
        self.max_length = max_length
        self.note_emb_size = note_emb_size
        self.staff_emb_size = staff_emb_size
        self.time_sig_emb_size = time_sig_emb_size
        self.key_emb_size = key_emb_size

        self.note_emb = nn.Embedding(vocab_size, note_emb_size)
        self.time_sig_emb = nn.Embedding(num_time_sig + 1, time_sig_emb_size)
        self.key_emb = nn.Embedding(num_keys + 1, key_emb_size)
        self.staff_emb = nn.GRU(note_emb_size, staff_emb_size, num_layers=1, batch_first=True, bidirectional=True)

        self.upper_decoder = NoteDecoder(max_length[0], note_emb_size, hidden_size)
        self.lower_decoder = NoteDecoder(max_length[1], note_emb_size, hidden_size)
        self.attn = AttentionLayer(hidden_size)
        self.gru = nn.GRU(staff_emb_size * 4 + time_sig_emb_size + key_emb_size + hidden_size * 2, hidden_size * 2,
                          num_layers=1, batch_first=True)
        self.time_sig_out = nn.Sequential(nn.Linear(hidden_size * 4, hidden_size * 4), nn.ReLU(),
                                          nn.Linear(hidden_size * 4, hidden_size * 2), nn.ReLU(),
                                          nn.Linear(hidden_size * 2, num_time_sig))
        self.key_out = nn.Sequential(nn.Linear(hidden_size * 4, hidden_size * 4), nn.ReLU(),
                                     nn.Linear(hidden_size * 4, hidden_size * 2), nn.ReLU(),
                                     nn.Linear(hidden_size * 2, num_keys))
        self.consume_python_rng = True      # keep the reference's python-`random` consumption count in inference too
        self.parallel_staves = True
        self._side_streams = None
        self._aux_stream = None
        self._presunk = None
        self._bar_lins = None
        self._defer_stream = None
        self._steps_host = None
        self._steps_pending = None
        self.init_weight()

    def init_weight(self):
        init_gru(self.gru)

    # -- staff summariser ---------------------------------------------------------------------------------------
    def _staff_summary(self, tokens, lengths):
        g = self.staff_emb
        p = [getattr(g, f"{n}_l0{sfx}") for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return ops.StaffGRUFn.apply(tokens, lengths, self.note_emb.weight, *p)          # (B, 2*staff_emb)

    def _staff_summary_pair(self, tok_a, len_a, tok_b, len_b):
        """The summaries of two token matrices (upper / lower staff) in ONE staff-GRU call: the shorter matrix is padded (the lengths
        bound what the recurrence reads), the rows are stacked, the result is split again -- one launch forward and one backward
        instead of two latency-bound launches each."""
        La, Lb = tok_a.shape[1], tok_b.shape[1]
        L = max(La, Lb)
        if La < L:
            tok_a = F.pad(tok_a, (0, L - La), value=PAD)
        if Lb < L:
            tok_b = F.pad(tok_b, (0, L - Lb), value=PAD)
        n = tok_a.shape[0]
        both = self._staff_summary(torch.cat([tok_a, tok_b]), torch.cat([len_a.reshape(-1), len_b.reshape(-1)]))
        return both[:n], both[n:]

    def get_SOS_token(self, batch_size):
        dev = self.note_emb.weight.device
        # constant index tensors, built once per (device, batch size): `torch.tensor(list, device=cuda)` is a BLOCKING copy that waits
        # for everything queued on the stream (the whole ConvStack + encoder of the step: 28 ms of a 61 ms step were spent here)
        key = (str(dev), int(batch_size))
        cache = getattr(self, "_sos_cache", None)
        if cache is None or cache[0] != key:
            se = torch.tensor([[SOS, EOS]], dtype=torch.long).repeat(batch_size, 1).to(dev)
            cache = (key, se, torch.full((batch_size,), 2, device=dev, dtype=torch.long),
                     torch.full((batch_size, 1), self.time_sig_SOS, device=dev, dtype=torch.long),
                     torch.full((batch_size, 1), self.key_SOS, device=dev, dtype=torch.long))
            self._sos_cache = cache
        _, se, two, ts_idx, key_idx = cache
        staff_token = self._staff_summary(se, two).unsqueeze(1)
        bar_token = torch.cat([staff_token, staff_token], dim=-1)
        ts = self.time_sig_emb(ts_idx)
        ky = self.key_emb(key_idx)
        return torch.cat([bar_token, ts, ky], dim=-1), staff_token

    def get_staff_token_from_probs(self, score_probs, lengths):
        return self._staff_summary(torch.argmax(score_probs, dim=-1), lengths).unsqueeze(1)

    def get_staff_token_from_gt(self, score_gt, lengths):
        return self._staff_summary(score_gt, lengths).unsqueeze(1)

    @staticmethod
    def _steps_from_gt(gt_staff):
        """(B, bars, L) -> (bars,) executed steps: the reference leaves the loop once every clip's GT row has shown
        <eos> (models.py:389,412-415), i.e. after max_b(first eos index)+1 steps, or never (L steps)."""
        L = gt_staff.shape[-1]
        is_eos = gt_staff == EOS
        first = torch.where(is_eos.any(-1), is_eos.int().argmax(-1), torch.full_like(gt_staff[..., 0], L - 1))
        return first.max(0).values + 1

    def prefetch_steps(self, ground_truth):
        """Starts the device->host read of `_steps_from_gt` for both staves on an auxiliary stream (which only waits for the work
        queued so far, i.e. the arrival of the targets).  decode_bars() picks the result up; without a prefetch it computes the
        same numbers with a blocking read."""
        upper_gt, lower_gt = ground_truth[2], ground_truth[4]
        self._steps_pending = None
        if not upper_gt.is_cuda or (hasattr(upper_gt, "_pa2s_steps") and hasattr(lower_gt, "_pa2s_steps")):
            return
        main = torch.cuda.current_stream()
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream()
        side = self._aux_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            st = torch.stack([self._steps_from_gt(upper_gt), self._steps_from_gt(lower_gt)])
            if self._steps_host is None or self._steps_host.shape != st.shape:
                self._steps_host = torch.empty(st.shape, dtype=st.dtype, pin_memory=True)
            self._steps_host.copy_(st, non_blocking=True)
            ev = side.record_event()
        upper_gt.record_stream(side)
        lower_gt.record_stream(side)
        self._steps_pending = (ev, upper_gt, lower_gt)

    def _make_streams(self):
        # the persistent note-decoder kernels (64 co-resident CTAs that own their SMs) go ahead of the wide parallel kernels that fill
        # the remaining SMs: two high-priority streams for the staves, one normal-priority stream for the deferred dEp / dv kernels
        prio = -1 if os.environ.get("PA2S_SIDE_PRIO", "0") == "1" else 0
        # the two staves run on two CUDA streams; their gradients meet in AccumulateGrad nodes of the default stream by design, so
        # autograd's stream-mismatch warning is switched off -- here, when the two-stream decoder is first used, not at import
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self._side_streams = (torch.cuda.Stream(priority=prio), torch.cuda.Stream(priority=prio))
        self._defer_stream = torch.cuda.Stream() if os.environ.get("PA2S_DEFER_STREAM", "0") == "1" else None

    def streams(self):
        """The CUDA streams the note decoders of the two staves run on (created on first use): for train.reserve_memory."""
        if self._side_streams is None:
            self._make_streams()
        return tuple(self._side_streams)

    def weight_sinks(self):
        """[(DecoderGradSink, weight aliases) or None] for the upper / lower note decoder (ops.DecoderWeightSinkFn).  Autograd runs
        nodes in reverse creation order: ScoreTranscription.forward creates the sinks BEFORE the ConvStack, so their node -- which
        only hands over gradients that were formed on a side stream long before -- is the last of a backward pass instead
        of sitting between the decoder and the encoder on the main stream."""
        sunk = [None, None]
        if torch.is_grad_enabled():
            for si, dec in enumerate((self.upper_decoder, self.lower_decoder)):
                if all(w.requires_grad for w in dec._weights()):
                    sink = ops.DecoderGradSink()
                    sunk[si] = (sink, ops.DecoderWeightSinkFn.apply(sink, *dec._weights()))
        return sunk

    def _heads(self, seq, x):
        x = F.relu(ops.linear(x, seq[0].weight, seq[0].bias))
        x = F.relu(ops.linear(x, seq[2].weight, seq[2].bias))
        return F.log_softmax(ops.linear(x, seq[4].weight, seq[4].bias), dim=-1)

    def decode_bars(self, encoder_outputs, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0):
        if encoder_outputs.is_cuda and ops.DECODER_IMPL == "multi":
            return self._decode_bars_multi(encoder_outputs, hidden, inference, ground_truth, teacher_forcing_ratio)
        return self._decode_bars_per_bar(encoder_outputs, hidden, inference, ground_truth, teacher_forcing_ratio)

    def _steps(self, ground_truth):
        """[upper steps per bar, lower steps per bar] a forward executes for these targets (one host read at most, see prefetch_steps)"""
        upper_gt, lower_gt = ground_truth[2], ground_truth[4]
        pend, self._steps_pending = self._steps_pending, None
        hint = (getattr(upper_gt, "_pa2s_steps", None), getattr(lower_gt, "_pa2s_steps", None))
        if hint[0] is not None and hint[1] is not None:
            return [list(hint[0]), list(hint[1])]                        # counted on the host by the loader (train.targets_to_device)
        if pend is not None and pend[1] is upper_gt and pend[2] is lower_gt:
            pend[0].synchronize()                                        # waits for the tiny side-stream read only
            return self._steps_host.tolist()
        return torch.stack([self._steps_from_gt(upper_gt), self._steps_from_gt(lower_gt)]).cpu().tolist()   # one sync

    def bar_linears(self, D=None):
        """The three Linear maps applied once per bar (attention query, bar GRU input / hidden) with deferred weight gradients
        (ops.DeferredLinear).  ScoreTranscription.forward creates them BEFORE the ConvStack, so that their gradient nodes run at the very
        end of a backward pass instead of between the note decoders' reverse pass and the encoder's."""
        D = D or 2 * self.hidden_size
        g = self.gru
        return (ops.DeferredLinear(self.attn.attn.weight[:, :D], None), ops.DeferredLinear(g.weight_ih_l0, g.bias_ih_l0),
                ops.DeferredLinear(g.weight_hh_l0, g.bias_hh_l0))

    def _bar_step(self, token, h, enc, Ep_bar, lins):
        """bar-level attention + GRU cell of one bar (models.py:239-247) -> (bar_summary, context)"""
        lin_q, lin_ih, lin_hh = lins
        context = ops.AttnStepFn.apply(lin_q(h), Ep_bar, enc, self.attn.v.weight)
        h = ops.GRUGatesFn.apply(lin_ih(torch.cat([token, context], dim=1)), lin_hh(h), h)
        return h, context

    def _decode_bars_multi(self, enc, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0):
        """decode_bars (models.py:191-316) on the multi-sequence decoder kernels (ops.StaffRun / ops.DecodersFn).

        A bar whose input token is built from the ground truth (bar-level teacher forcing, models.py:289-299) does not depend on the
        previous bar's note decoders; all coins are pre-drawn (in the reference's order), so the bars split into SEGMENTS: runs of
        bars of which only the first needs the previous bar's predictions.  Per segment the bar-level chain runs first, then ONE
        launch per staff decodes all its bars (both staves concurrently on two streams); the reverse pass is one launch per staff
        over all bars."""
        B, T, D = enc.shape
        dev = enc.device
        training = self.training
        src = rng.source()
        nb = self.max_bars
        if inference:
            assert teacher_forcing_ratio == 0
            assert ground_truth is None
        have_gt = ground_truth is not None
        if have_gt:
            time_sig_gt, key_gt, upper_gt, upper_len_gt, lower_gt, lower_len_gt = ground_truth
            steps = self._steps(ground_truth)
        else:
            steps = [[self.max_length[0]] * nb, [self.max_length[1]] * nb]
        decs = (self.upper_decoder, self.lower_decoder)
        E = self.note_emb_size
        tokdim = self.staff_emb_size * 4 + self.time_sig_emb_size + self.key_emb_size
        tf_bars = have_gt and not inference                              # the next bar token may come from the ground truth
        # ---- all randomness of the forward pass, drawn in the reference's order (models.py:239, 391, 404, 289)
        iid = getattr(src, "iid", False)
        Smax = [max(steps[0]), max(steps[1])]
        bar_masks = None
        note_masks = [None, None]
        tf_bits = [[0] * nb, [0] * nb] if (have_gt and not inference) else [None, None]
        if training and iid:
            bar_masks = src.dropout_mask((nb, B, tokdim), 0.1, dev, "bar_token")
            note_masks = [src.dropout_mask((Smax[si], nb * B, E), 0.1, dev, "note_steps") for si in (0, 1)]
        elif training:
            bar_masks = []
            note_masks = [torch.ones(Smax[si], nb * B, E, device=dev) for si in (0, 1)]
        bar_tf = []
        for bar in range(nb):
            if training and not iid:
                bar_masks.append(src.dropout_mask((B, 1, tokdim), 0.1, dev, "bar_token").view(B, tokdim))
            for si in (0, 1):
                S = steps[si][bar]
                if have_gt:
                    coins = src.coins(S)
                    if not inference:
                        bits = 0
                        for st_, c in enumerate(coins):
                            if c < teacher_forcing_ratio:
                                bits |= 1 << st_
                        tf_bits[si][bar] = bits
                if training and not iid:
                    note_masks[si][:S, bar * B:(bar + 1) * B] = src.dropout_mask((S, B, E), 0.1, dev, "note_steps")
            bar_tf.append(src.coin() < teacher_forcing_ratio)
        # bars k >= 1 whose token is built from bar k-1's predictions start a new segment
        nqmax = ops.lib.pa2s_decm_max_queries()
        segs, k0 = [], 0
        for k in range(1, nb + 1):
            if k == nb or not (tf_bars and bar_tf[k - 1]) or k - k0 == nqmax:
                segs.append((k0, k))
                k0 = k

        # step-invariant encoder half of the three attention layers (models.py:458): Ep = enc W_e^T + b
        enc2 = enc.reshape(B * T, D)
        # (one N = 3*A contraction for the three layers: enc is split into bf16 pieces once, and the backward is one data-gradient
        # and one weight-gradient GEMM instead of three of each plus the accumulation of three d_enc tensors)
        atts = (self.attn, self.upper_decoder.attn, self.lower_decoder.attn)
        W_e = torch.cat([a_.attn.weight[:, D:] for a_ in atts])
        b_e = torch.cat([a_.attn.bias for a_ in atts])
        Ep_bar, Ep_up, Ep_lo = (e_.view(B, T, -1) for e_ in ops.SplitColsFn.apply(ops.linear(enc2, W_e, b_e), 3))

        grad = torch.is_grad_enabled() and (enc.requires_grad or any(w.requires_grad for d in decs for w in d._weights()))
        main = torch.cuda.current_stream()
        if self._side_streams is None:
            self._make_streams()
        sides = self._side_streams if self.parallel_staves else (None, None)
        runs = []
        for si, (dec, Ep) in enumerate(zip(decs, (Ep_up, Ep_lo))):
            gt_staff = (upper_gt, lower_gt)[si] if have_gt else None
            runs.append(ops.StaffRun(dec._weights(), enc, Ep, nb, dec.max_steps, steps[si], inference or not have_gt, grad, sides[si],
                                     SOS, EOS, gt=gt_staff, tf_bits=tf_bits[si], mask=note_masks[si]))

        lins, self._bar_lins = self._bar_lins, None
        if lins is None:
            lins = self.bar_linears(D)
        token = self.get_SOS_token(B)[0].squeeze(1)                      # (B, 4*staff+ts+key)
        h = hidden[0]
        # one autograd node per bar for the bar-level chain (attention query + attention step + GRU cell), when the Linear maps have
        # deferred weight gradients (training); otherwise the individual ops
        chain = ops.BarChain(lins, self.attn.v.weight) if (ops.BAR_CHAIN and all(l.defer for l in lins)) else None
        tok_gt = None
        if tf_bars and any(bar_tf[:nb - 1]):
            # tokens built from the targets (models.py:290-299), for all bars in ONE staff-summariser call per staff: row (b, bar)
            us, ls = self._staff_summary_pair(upper_gt.reshape(B * nb, -1), upper_len_gt.reshape(-1),
                                              lower_gt.reshape(B * nb, -1), lower_len_gt.reshape(-1))
            us, ls = us.view(B, nb, -1), ls.view(B, nb, -1)
            tok_gt = torch.cat([us, ls, self.time_sig_emb(time_sig_gt), self.key_emb(key_gt)], dim=-1)       # [:, k] = token of bar k+1
        summaries, contexts = [], []
        done = []
        for (k0, k1) in segs:
            seg_h, seg_ctx = [], []
            for bar in range(k0, k1):
                if bar > 0 and tf_bars and bar_tf[bar - 1]:              # teacher-forced: token from the targets of bar-1
                    token = tok_gt[:, bar - 1]
                elif bar > 0:                                            # token from bar-1's predictions (first bar of a segment)
                    assert bar == k0
                    for ev in done:
                        main.wait_event(ev)
                    done = []
                    us, ls = self._staff_summary_pair(torch.argmax(runs[0].logp[:, bar - 1], dim=-1), runs[0].lengths[bar - 1],
                                                      torch.argmax(runs[1].logp[:, bar - 1], dim=-1), runs[1].lengths[bar - 1])
                    with torch.no_grad():                                # (the differentiable heads of all bars are formed at the end)
                        head_in = torch.cat([summaries[-1], contexts[-1]], dim=1)
                        ts_pred = torch.argmax(self._heads(self.time_sig_out, head_in), dim=-1)
                        key_pred = torch.argmax(self._heads(self.key_out, head_in), dim=-1)
                    token = torch.cat([us, ls, self.time_sig_emb(ts_pred), self.key_emb(key_pred)], dim=-1)
                if training:
                    token = token * bar_masks[bar]
                h, context = chain.step(token, h, enc, Ep_bar) if chain is not None else self._bar_step(token, h, enc, Ep_bar, lins)
                seg_h.append(h)
                seg_ctx.append(context)
            summaries += seg_h
            contexts += seg_ctx
            h0 = torch.stack([x.detach() for x in seg_h])               # (nq, B, D)
            ready = main.record_event()
            for si, run in enumerate(runs):
                side = sides[si]
                if side is not None:
                    side.wait_event(ready)
                    h0.record_stream(side)
                with ops._on_stream(side):
                    run.launch(k0, k1 - k0, h0)
                    if side is not None:
                        done.append(side.record_event())
        for ev in done:
            main.wait_event(ev)
        counters = [c for run in runs for c in run.counters]
        if not have_gt and self.consume_python_rng:
            # the reference draws one coin per executed note step (models.py:404); only the count matters here
            src.coins(int(torch.stack(counters)[:, 1].sum().item()))
        self.last_step_counters = counters
        # Order of creation = reverse order of execution in backward.  h0_all (whose backward feeds the bar-level GRU chain and has to wait
        # for the note decoders' reverse pass) first; then the time-signature / key heads (models.py:281-286) of all bars in one pass,
        # rows (b, bar) -- they are off the chain that feeds the note decoders; LAST the node of all note decoding, so that its backward
        # is the first thing autograd runs: the two reverse kernels are enqueued at once and the heads' backward overlaps them.
        h0_all = torch.stack(summaries) if grad else None
        head_in = torch.cat([torch.stack(summaries, 1), torch.stack(contexts, 1)], dim=-1).view(B * nb, -1)
        ts_all = self._heads(self.time_sig_out, head_in).view(B, nb, -1)
        key_all = self._heads(self.key_out, head_in).view(B, nb, -1)
        if grad:
            w = [x for d in decs for x in d._weights()]
            # autograd runs a node's backward on the stream its forward was issued on and orders it against the producers / consumers
            # of its gradients: issued on the upper staff's stream, the reverse pass does not hold up the main stream
            with ops._on_stream(sides[0]):
                up_all, lo_all = ops.DecodersFn.apply(tuple(runs), enc, Ep_up, Ep_lo, h0_all, *w)
        else:
            up_all, lo_all = runs[0].logp, runs[1].logp
        return (ts_all, key_all, up_all, lo_all)

    def _decode_bars_per_bar(self, encoder_outputs, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0):
        enc = encoder_outputs
        B, T, D = enc.shape
        dev = enc.device
        training = self.training
        src = rng.source()
        if inference:
            assert teacher_forcing_ratio == 0
            assert ground_truth is None
        have_gt = ground_truth is not None
        if have_gt:
            time_sig_gt, key_gt, upper_gt, upper_len_gt, lower_gt, lower_len_gt = ground_truth
            pend, self._steps_pending = self._steps_pending, None
            hint = (getattr(upper_gt, "_pa2s_steps", None), getattr(lower_gt, "_pa2s_steps", None))
            if hint[0] is not None and hint[1] is not None:
                steps = [list(hint[0]), list(hint[1])]                   # counted on the host by the loader (train.targets_to_device)
            elif pend is not None and pend[1] is upper_gt and pend[2] is lower_gt:
                pend[0].synchronize()                                    # waits for the tiny side-stream read only
                steps = self._steps_host.tolist()
            else:
                steps = torch.stack([self._steps_from_gt(upper_gt), self._steps_from_gt(lower_gt)]).cpu().tolist()   # one sync
        else:
            steps = [[self.max_length[0]] * self.max_bars, [self.max_length[1]] * self.max_bars]

        # step-invariant encoder half of the three attention layers (models.py:458): Ep = enc W_e^T + b
        enc2 = enc.reshape(B * T, D)
        def ep(att):
            return ops.linear(enc2, att.attn.weight[:, D:], att.attn.bias).view(B, T, -1)
        Ep_bar, Ep_up, Ep_lo = ep(self.attn), ep(self.upper_decoder.attn), ep(self.lower_decoder.attn)

        # weight gradients of the two note decoders: one set of contractions per backward pass over all bars (ops.DecoderWeightSinkFn)
        sunk, self._presunk = self._presunk, None
        if sunk is None:
            sunk = self.weight_sinks() if enc.is_cuda else [None, None]

        # ... and their data gradients: every (bar, staff) backward is enqueued as soon as the loss gradient exists (ops.StackLogpFn)
        early = ops.DecoderEarlyBackward() if (torch.is_grad_enabled() and enc.is_cuda and self.parallel_staves) else None
        if early is not None:
            if self._side_streams is None:
                self._make_streams()
            early.sinks = [(sk[0], self._side_streams[si]) for si, sk in enumerate(sunk) if sk is not None]

        token = self.get_SOS_token(B)[0].squeeze(1)                      # (B, 4*staff+ts+key)
        h = hidden[0]
        ts_outs, key_outs, up_outs, lo_outs, counters, pending = [], [], [], [], [], []
        for bar in range(self.max_bars):
            if training:
                token = token * src.dropout_mask((B, 1, token.shape[-1]), 0.1, dev, "bar_token").view(B, -1)
            q = ops.linear(h, self.attn.attn.weight[:, :D], None)
            context = ops.AttnStepFn.apply(q, Ep_bar, enc, self.attn.v.weight)
            g = self.gru
            h = ops.gru_cell(torch.cat([token, context], dim=1), h, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
            bar_summary = h
            # The two staves only depend on bar_summary (models.py:261-275), and the bar-level chain itself only depends on
            # their OUTPUT when the next bar token is built from predictions (models.py:289-311).  So the note decoders run on
            # two side streams, alternating so that the long (upper) and short (lower) staves balance, while the bar chain
            # (heads, staff summaries, next bar's attention + GRU) runs ahead on the current stream; it waits for the decoders
            # only when it has to read their predictions.
            res = []
            main = torch.cuda.current_stream() if enc.is_cuda else None
            use_streams = self.parallel_staves and main is not None
            if use_streams and self._side_streams is None:
                self._make_streams()
            ready = main.record_event() if use_streams else None
            done = []
            for si, (dec, Ep) in enumerate(((self.upper_decoder, Ep_up), (self.lower_decoder, Ep_lo))):
                gt_staff = (upper_gt, lower_gt)[si][:, bar, :] if have_gt else None
                tf_in = teacher_forcing_ratio if have_gt else 0.
                if use_streams:
                    side = self._side_streams[(bar + si) % 2]
                    side.wait_event(ready)
                    for t_ in (enc, Ep, bar_summary) + ((gt_staff,) if gt_staff is not None else ()):
                        t_.record_stream(side)
                    out = dec._decode(enc, Ep, bar_summary, inference, gt_staff, tf_in, steps[si][bar], src, sunk[si],
                                      side=side, early=(early, (si, bar)) if early is not None else None,
                                      defer_stream=self._defer_stream if early is not None else None)
                    done.append(side.record_event())
                    for t_ in out:
                        t_.record_stream(main)
                    res.append(out)
                else:
                    res.append(dec._decode(enc, Ep, bar_summary, inference, gt_staff, tf_in, steps[si][bar], src, sunk[si]))
            (up_p, up_len, up_cnt), (lo_p, lo_len, lo_cnt) = res
            pending += done
            counters += [up_cnt, lo_cnt]
            up_outs.append(up_p)
            lo_outs.append(lo_p)
            head_in = torch.cat([bar_summary, context], dim=1)
            ts_lp = self._heads(self.time_sig_out, head_in)
            key_lp = self._heads(self.key_out, head_in)
            ts_outs.append(ts_lp)
            key_outs.append(key_lp)
            teacher_force = src.coin() < teacher_forcing_ratio
            if bar == self.max_bars - 1:
                break                                                    # the token built after the last bar is never used
            if teacher_force and not inference and have_gt:
                us = self._staff_summary(upper_gt[:, bar, :], upper_len_gt[:, bar])
                ls = self._staff_summary(lower_gt[:, bar, :], lower_len_gt[:, bar])
                tst = self.time_sig_emb(time_sig_gt[:, bar])
                kyt = self.key_emb(key_gt[:, bar])
            else:
                for ev in pending:                                       # predictions of this bar feed the next bar token
                    main.wait_event(ev)
                pending = []
                us = self._staff_summary(torch.argmax(up_p, dim=-1), up_len)
                ls = self._staff_summary(torch.argmax(lo_p, dim=-1), lo_len)
                tst = self.time_sig_emb(torch.argmax(ts_lp, dim=-1))
                kyt = self.key_emb(torch.argmax(key_lp, dim=-1))
            token = torch.cat([us, ls, tst, kyt], dim=-1)
        for ev in pending:
            main.wait_event(ev)
        if not have_gt and self.consume_python_rng:
            # the reference draws one coin per executed note step (models.py:404); only the count matters here
            src.coins(int(torch.stack(counters)[:, 1].sum().item()))
        self.last_step_counters = counters
        if early is not None:
            up_all, lo_all = ops.StackLogpFn.apply(early, len(up_outs), *up_outs, *lo_outs)
        else:
            up_all, lo_all = torch.stack(up_outs, 1), torch.stack(lo_outs, 1)
        return (torch.stack(ts_outs, 1), torch.stack(key_outs, 1), up_all, lo_all)

    def forward(self, encoder_outputs, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0,
                device=torch.device("cuda" if torch.cuda.is_available() else "cpu")):
        self.device = device
        with ops.module_precision(self):
            if inference:
                assert teacher_forcing_ratio == 0
                assert ground_truth is None
                return self.decode_bars(encoder_outputs, hidden, True, None, 0.)
            return self.decode_bars(encoder_outputs, hidden, False, ground_truth, teacher_forcing_ratio)


class NoteDecoder(nn.Module):
    def __init__(self, max_steps=25, note_emb_size=128, hidden_size=400):
        super().__init__()
        self.max_steps = max_steps
        self.note_emb_size = note_emb_size
        self.hidden_size = hidden_size
        self.embedding = nn.Embedding(vocab_size, note_emb_size)
        self.attn = AttentionLayer(hidden_size)
        self.gru = nn.GRU(note_emb_size + hidden_size * 2, hidden_size * 2, num_layers=1, batch_first=True)
        self.out = nn.Linear(hidden_size * 4, vocab_size)
        self.init_weight()

    def init_weight(self):
        init_gru(self.gru)
        init_layer(self.out)

    def _weights(self):
        g = self.gru
        return (self.attn.attn.weight, self.attn.v.weight, self.embedding.weight, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0,
                g.bias_hh_l0, self.out.weight, self.out.bias)

    def _decode(self, enc, Ep, h0, inference, gt, tf_ratio, S, src, sunk=None, side=None, early=None, defer_stream=None):
        """All steps of one (bar, staff).  Coins/masks are pre-drawn in the reference's order (models.py:391,404).
        `sunk` = (DecoderGradSink, weight aliases from DecoderWeightSinkFn) when the caller defers the weight gradients;
        `side` = the CUDA stream the call is enqueued on (the caller orders it against the current stream with events);
        `early` = (ops.DecoderEarlyBackward, key) when the caller launches the backward of all its calls at once;
        `defer_stream` = stream for the parallel dEp / dv kernel of the backward (default: the call's own stream)."""
        B = enc.shape[0]
        dev = enc.device
        training = self.training
        have_gt = gt is not None
        use_gt = mask = None
        with ops._on_stream(side):
            if have_gt:
                coins = src.coins(S)
                if not inference:
                    use_gt = torch.tensor([1 if c < tf_ratio else 0 for c in coins], dtype=torch.int32).to(dev, non_blocking=True)
                gt = gt.contiguous()
            if training:
                mask = src.dropout_mask((S, B, self.note_emb_size), 0.1, dev, "note_steps").contiguous()
        cfg = dict(S=S, max_steps=self.max_steps, inference=inference or not have_gt, gt=gt if have_gt else None,
                   use_gt=use_gt, mask=mask, sos=SOS, eos=EOS, stream=side, early=early, defer_stream=defer_stream)
        wts = self._weights()
        if sunk is not None:
            cfg["sink"], wts = sunk
        return ops.NoteDecoderFn.apply(enc, Ep, h0, *wts, cfg)

    def decode_notes(self, encoder_outputs, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0):
        if inference:
            assert teacher_forcing_ratio == 0
            assert ground_truth is None
        enc = encoder_outputs
        B, T, D = enc.shape
        Ep = ops.linear(enc.reshape(B * T, D), self.attn.attn.weight[:, D:], self.attn.attn.bias).view(B, T, -1)
        src = rng.source()
        if ground_truth is not None:
            S = int(HierarchicalDecoder._steps_from_gt(ground_truth.unsqueeze(1))[0].item())
        else:
            S = self.max_steps
        logp, lengths, counters = self._decode(enc, Ep, hidden[0], inference, ground_truth, teacher_forcing_ratio, S, src)
        if ground_truth is None:
            src.coins(int(counters[1].item()))
        return logp, lengths.cpu()

    def forward(self, encoder_outputs, hidden, inference=True, ground_truth=None, teacher_forcing_ratio=0,
                device=torch.device("cuda" if torch.cuda.is_available() else "cpu")):
        self.device = device
        with ops.module_precision(self):
            if inference:
                assert teacher_forcing_ratio == 0
                assert ground_truth is None
                return self.decode_notes(encoder_outputs, hidden, True, None, 0.)
            return self.decode_notes(encoder_outputs, hidden, False, ground_truth, teacher_forcing_ratio)


class AttentionLayer(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.attn = nn.Linear(hidden_size * 4, hidden_size)
        self.v = nn.Linear(hidden_size, 1, bias=False)
        self.init_weight()

    def init_weight(self):
        init_layer(self.attn)
        init_layer(self.v)

    def forward(self, hidden, encoder_output):
        """(1,B,2H), (B,T,2H) -> softmax attention weights (B,T) (models.py:452-461)."""
        B, T, D = encoder_output.shape
        with ops.module_precision(self):
            Ep = ops.linear(encoder_output.reshape(B * T, D), self.attn.weight[:, D:], self.attn.bias).view(B, T, -1)
            q = ops.linear(hidden[0], self.attn.weight[:, :D], None)
            energy = torch.tanh(q.unsqueeze(1) + Ep)
            return F.softmax(ops.linear(energy, self.v.weight, None).squeeze(2), dim=1)


class ConvStack(nn.Module):
    def __init__(self, in_channels=1, freq_bins=480, output_size=200):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 20, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv2 = nn.Conv2d(20, 20, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv3 = nn.Conv2d(20, 40, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv4 = nn.Conv2d(40, 40, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(20)
        self.bn2 = nn.BatchNorm2d(20)
        self.bn3 = nn.BatchNorm2d(40)
        self.bn4 = nn.BatchNorm2d(40)
        self.out = nn.Linear(freq_bins * 40, output_size, bias=False)
        self.out_bn = nn.BatchNorm1d(output_size)
        self.init_weight()

    def init_weight(self):
        for c in (self.conv1, self.conv2, self.conv3, self.conv4):
            init_layer(c)
        for b in (self.bn1, self.bn2, self.bn3, self.bn4):
            init_bn(b)
        init_layer(self.out)
        init_bn(self.out_bn)

    def forward(self, x):
        _dev(x.device)
        B, _, T, _ = x.shape
        bns = (self.bn1, self.bn2, self.bn3, self.bn4, self.out_bn)
        training = self.training
        mask = None
        if training:
            mask = rng.source().dropout_mask((B, T, self.out.weight.shape[0]), 0.2, x.device, "conv")
        # speechbrain wraps the model with SyncBatchNorm.convert_sync_batchnorm under DDP: honour cross-rank statistics
        sync = isinstance(self.bn1, nn.SyncBatchNorm) or getattr(self, "sync_batchnorm", False)
        bufs = [(b.running_mean, b.running_var) for b in bns]
        params = []
        for c, b in zip((self.conv1, self.conv2, self.conv3, self.conv4), bns[:4]):
            params += [c.weight, b.weight, b.bias]
        params += [self.out.weight, self.out_bn.weight, self.out_bn.bias]
        with ops.module_precision(self):
            y = ops.ConvStackFn.apply(x, mask, bufs, training, sync, float(self.bn1.eps), float(self.bn1.momentum), *params)
        if training:
            with torch.no_grad():
                for b in bns:
                    b.num_batches_tracked += 1
        return y


# ---------------------------------------------------------------------------------------------------------------
# Initialisers (models.py:548-585)
# ---------------------------------------------------------------------------------------------------------------
def init_layer(layer):
    """Xavier-uniform weight, zero bias."""
    nn.init.xavier_uniform_(layer.weight)
    if hasattr(layer, "bias") and layer.bias is not None:
        layer.bias.data.fill_(0.)


def init_bn(bn):
    bn.bias.data.fill_(0.)
    bn.weight.data.fill_(1.)


def init_gru(rnn):
    """Per-gate uniform(+-sqrt(3/fan_in)) for W_ih and the r,z blocks of W_hh, orthogonal n block, zero biases.
    (Like the reference, only the forward-direction parameters of each layer are touched.)"""
    def _blocks(tensor, fns):
        n = tensor.shape[0] // len(fns)
        for i, fn in enumerate(fns):
            fn(tensor[i * n:(i + 1) * n, :])

    def _uniform(t):
        fan_in = nn.init._calculate_correct_fan(t, "fan_in")
        nn.init.uniform_(t, -math.sqrt(3 / fan_in), math.sqrt(3 / fan_in))

    for i in range(rnn.num_layers):
        _blocks(getattr(rnn, f"weight_ih_l{i}"), [_uniform, _uniform, _uniform])
        torch.nn.init.constant_(getattr(rnn, f"bias_ih_l{i}"), 0)
        _blocks(getattr(rnn, f"weight_hh_l{i}"), [_uniform, _uniform, nn.init.orthogonal_])
        torch.nn.init.constant_(getattr(rnn, f"bias_hh_l{i}"), 0)
