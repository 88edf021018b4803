"""Kernel-level timeline of one full-size training step (torch.profiler / CUPTI): per-kernel GPU time, GPU busy vs wall time of
the step, and the idle gaps between kernels (host-launch-bound stretches).  Not a bench: profiler overhead is included."""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--top", type=int, default=45)
    args = ap.parse_args()
    import models
    from piano_a2s_b200 import train
    from piano_a2s_b200.synthetic import make_audio, make_ground_truth
    from piano_a2s_b200.vqt import VQT
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    m = models.ScoreTranscription(max_length=(398, 189)).to(dev).train()
    opt = train.FlatAdadelta(m)
    vqt = VQT().to(dev)
    B = args.batch
    audio = make_audio(B, 192000, seed=1234).to(dev)
    gt = train.targets_to_device([g.pin_memory() for g in make_ground_truth(B, 5, 398, 189, seed=1234)], dev)

    def step():
        spec = vqt(audio).unsqueeze(1)
        return train.fit_batch(m, opt, spec, gt, 0.7)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if os.environ.get("PA2S_HOSTLAUNCH"):
        # host launch time next to the device start time of every kernel >= 100 us (chrome trace, matched by correlation id):
        # tells a device-side dependency (launched long before it starts) from a launch-bound stretch (starts when launched)
        import json
        import tempfile
        path = os.path.join(tempfile.mkdtemp(), "trace.json")
        prof.export_chrome_trace(path)
        tr = json.load(open(path))["traceEvents"]
        launch = {e["args"]["correlation"]: e for e in tr if e.get("cat") == "cuda_runtime" and "correlation" in e.get("args", {})}
        kern = sorted((e for e in tr if e.get("cat") == "kernel"), key=lambda e: e["ts"])
        t00 = kern[0]["ts"]
        print("  device start | dur | host launch (ms rel. to first kernel) | stream | kernel")
        for e in kern:
            if e["dur"] >= 100:
                l = launch.get(e["args"].get("correlation"))
                lt = f"{1e-3 * (l['ts'] - t00):8.3f}" if l else "    n/a"
                print(f"  t={1e-3 * (e['ts'] - t00):8.3f}  dur={1e-3 * e['dur']:7.3f}  host={lt}  stream={e['args'].get('stream')}  {e['name'][:50]}")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if os.environ.get("PA2S_TIMELINE"):
        # coarse timeline: per stream (device_index of a CUDA event = stream id in kineto), kernels >= 100 us, and per-stream busy time
        import re
        tl = sorted(((e.time_range.start, e.time_range.end, getattr(e, "device_resource_id", -1), e.name) for e in evs), key=lambda x: x[0])
        t00 = tl[0][0]
        busy = collections.defaultdict(float)
        for s_, e_, st, n in tl:
            busy[st] += e_ - s_
        print("per-stream kernel time (ms):", {k: round(v * 1e-3, 2) for k, v in busy.items()})
        wins = [[float(x) for x in w.split(",")] for w in os.environ.get("PA2S_TIMELINE_WINDOW", "0,0").split(";")]
        for s_, e_, st, n in tl:
            if e_ - s_ >= 100 or any(lo <= 1e-3 * (s_ - t00) <= hi for lo, hi in wins):
                short = re.sub(r"\(anonymous namespace\)::|void ", "", n)[:60]
                print(f"  t={1e-3 * (s_ - t00):8.3f}  dur={1e-3 * (e_ - s_):7.3f}  stream={st}  {short}")
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda x: x[0])
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    # union of busy intervals
    busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    gaps = []
    for s, e, n in ks[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f"step span {1e-3 * (t1 - t0):.2f} ms   GPU busy (union) {1e-3 * busy:.2f} ms   idle {1e-3 * (t1 - t0 - busy):.2f} ms   kernels {len(ks)}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in ks:
        agg[n][0] += 1
        agg[n][1] += e - s
    print(f"{'kernel':90s} {'n':>6s} {'total ms':>10s} {'avg us':>10s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"{n[:90]:90s} {c:6d} {1e-3 * t:10.3f} {t / c:10.2f}")
    gagg = collections.defaultdict(lambda: [0, 0.0])
    for g, n in gaps:
        gagg[n][0] += 1
        gagg[n][1] += g
    print("\nidle gaps, by the kernel that ended them")
    for n, (c, t) in sorted(gagg.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"{n[:90]:90s} {c:6d} {1e-3 * t:10.3f} {t / c:10.2f}")


if __name__ == "__main__":
    main()
