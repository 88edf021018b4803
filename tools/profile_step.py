"""Per-stage CUDA-event timing of the full-size training step (config 2: pretrain.yaml model, B=16, fp32)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--frames", type=int, default=1201)
    args = ap.parse_args()
    import models
    from helpers import make_ground_truth
    from piano_a2s_b200 import train
    from piano_a2s_b200._lib import lib
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    m = models.ScoreTranscription(max_length=(398, 189)).to(dev).train()
    opt = train.FlatAdadelta(m)
    B = args.batch
    x = torch.rand(B, 1, args.frames, 480, device=dev)
    gt = [g.to(dev) for g in make_ground_truth(B, 5, 398, 189, seed=1234)]

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    for it in range(args.iters):
        torch.cuda.synchronize()
        l0 = lib.pa2s_launch_count()
        t0 = time.time()
        e0 = ev()
        conv = m.convstack(x)
        e1 = ev()
        enc, hid = m.encoder(conv)
        e2 = ev()
        outs = m.decoder(enc, hid, False, gt, 0.7, dev)
        e3 = ev()
        loss, _ = train.compute_objectives(outs, gt)
        e4 = ev()
        loss.backward()
        e5 = ev()
        opt.step()
        opt.zero_grad()
        e6 = ev()
        torch.cuda.synchronize()
        wall = time.time() - t0
        names = ["convstack_fwd", "encoder_fwd", "decoder_fwd", "loss", "backward", "optimizer"]
        es = [e0, e1, e2, e3, e4, e5, e6]
        print(f"iter {it}: wall {wall*1e3:.1f} ms  loss {loss.item():.4f}  launches {lib.pa2s_launch_count()-l0}  " +
              "  ".join(f"{n} {es[i].elapsed_time(es[i+1]):.2f}" for i, n in enumerate(names)), flush=True)
    print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)


if __name__ == "__main__":
    main()
