"""Per-phase timing of the persistent note decoder at full size (one upper-staff call, B clips, S steps): CTA 0 accumulates
globaltimer deltas per phase (compute A/B/C and the three grid barriers) into ops.PROF buffers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import models  # noqa: E402
from piano_a2s_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, S = int(os.environ.get("PB", 16)), 1201, int(os.environ.get("PS", 80))
torch.manual_seed(0)
dec = models.NoteDecoder(398, 16, 256).to(dev).train()
enc = torch.randn(B, T, 512, device=dev, requires_grad=True)
Ep = torch.randn(B, T, 256, device=dev, requires_grad=True)
h0 = torch.randn(B, 512, device=dev, requires_grad=True)
gt = torch.randint(0, 144, (B, 398), device=dev)
use_gt = torch.ones(S, dtype=torch.int32, device=dev)
mask = (torch.rand(S, B, 16, device=dev) > 0.1).float() / 0.9
cfg = dict(S=S, max_steps=398, inference=False, gt=gt, use_gt=use_gt, mask=mask, sos=models.SOS, eos=models.EOS)
g = dec.gru


def run():
    with ops.use_precision(os.environ.get("PPREC", "bf16x3")):
        logp, _, _ = ops.NoteDecoderFn.apply(enc, Ep, h0, dec.attn.attn.weight, dec.attn.v.weight, dec.embedding.weight,
                                             g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0, dec.out.weight, dec.out.bias, cfg)
        logp.square().sum().backward()


for _ in range(2):
    run()
torch.cuda.synchronize()
ops.PROF["fwd"] = torch.zeros(16, dtype=torch.int64, device=dev)
ops.PROF["bwd"] = torch.zeros(16, dtype=torch.int64, device=dev)
ops.KernelTimers.reset(True)
run()
torch.cuda.synchronize()
ops.check_sync_flags()
kt = ops.KernelTimers.summary()
for k, names in (("fwd", ["A attention(+D)", "barrier", "B gru", "barrier", "C logits+q", "barrier", "prologue"]),
                 ("bwd", ["P1 gates", "barrier", "P2 gemv", "barrier", "P3 attention", "barrier"])):
    v = ops.PROF[k].cpu().tolist()
    print(f"{k}: B={B} S={S} total {sum(v[:7]) / 1e3:.1f} us  ({sum(v[:7]) / 1e3 / S:.2f} us/step)   call {kt.get('note_decoder_' + k, (0, 0))[1]:.3f} ms")
    for n, x in zip(names, v):
        print(f"    {n:18s} {x / 1e3 / S:8.2f} us/step")
    for n, x in zip(["attn: stage q/dc", "attn: frame loop", "attn: combine+ticket", "attn: last-arriver"], v[8:12]):
        print(f"      {n:22s} {x / 1e3 / S:8.2f} us/step")
