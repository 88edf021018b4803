"""Per-phase timing of the multi-sequence note decoder (dec_multi.cu) at full size: one staff, B clips, NQ bars per launch, S steps.
CTA 0 accumulates globaltimer deltas per phase into ops.PROF buffers.   PB / PS / PQ / PPREC environment variables."""
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
enc = torch.randn(B, T, 512, device=dev)
Ep = torch.randn(B, T, 256, device=dev)
for NQ in [int(x) for x in os.environ.get("PQ", "1,2,3,5").split(",")]:
    bars = NQ
    gt = torch.randint(0, 144, (B, bars, 398), device=dev)
    use_gt = [(1 << S) - 1] * bars
    mask = (torch.rand(S, bars * B, 16, device=dev) > 0.1).float() / 0.9
    h0 = torch.randn(bars, B, 512, device=dev)
    dl = torch.randn(B, bars, 398, 173, device=dev) * 1e-3

    def run():
        with ops.use_precision(os.environ.get("PPREC", "bf16x3")):
            r = ops.StaffRun(dec._weights(), enc, Ep, bars, 398, [S] * bars, False, True, None, models.SOS, models.EOS, gt=gt, tf_bits=use_gt, mask=mask)
            r.launch(0, NQ, h0)
            return r.backward(dl)

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
    for k, names, sub in (("fwd", ["A attention(+D)", "barrier", "B gru", "barrier", "C logits+q", "barrier", "prologue"],
                           ["Eq load", "pass1 scores", "max", "softmax p", "pass2 context", "partials+ticket"]),
                          ("bwd", ["P1 gates", "barrier", "P2 gemv", "barrier", "P3 attention", "barrier"],
                           ["dc/Eq/c0", "pass1 da", "ds", "pass2 dq", "dq reduce+ticket"])):
        v = ops.PROF[k].cpu().tolist()
        print(f"{k}: B={B} NQ={NQ} S={S} total {sum(v[:7]) / 1e3:.1f} us  ({sum(v[:7]) / 1e3 / S:.2f} us/step, {sum(v[:7]) / 1e3 / S / NQ:.2f} us/bar-step)"
              f"   call {kt.get('note_decoder_' + k, (0, 0))[1]:.3f} ms")
        for n, x in zip(names, v):
            print(f"    {n:18s} {x / 1e3 / S:8.2f} us/step")
        for n, x in zip(sub, v[8:8 + len(sub)]):
            print(f"      {n:22s} {x / 1e3 / S:8.2f} us/step")
    ops.PROF.clear()
