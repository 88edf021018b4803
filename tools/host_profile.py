"""Host-side (Python / ctypes / autograd) cost of ENQUEUEING one training step of the bench workload: cProfile over a few
steps, sorted by own time and by cumulative time.  The device runs asynchronously; what is measured is launch overhead."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import models  # noqa: E402
from piano_a2s_b200 import train  # noqa: E402
from piano_a2s_b200.synthetic import make_audio, make_ground_truth  # noqa: E402
from piano_a2s_b200.vqt import VQT  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("PB", 16))
torch.manual_seed(1234)
m = models.ScoreTranscription(max_length=(398, 189)).to(dev).train()
opt = train.FlatAdadelta(m)
vqt = VQT().to(dev)
audio = make_audio(B, 192000, seed=1234).to(dev)
gt = train.targets_to_device([t.pin_memory() for t in make_ground_truth(B, 5, 398, 189, seed=1234)], dev)


def step():
    spec = vqt(audio).unsqueeze(1)
    return train.fit_batch(m, opt, spec, gt, 0.7)


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 3
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
for _ in range(n):
    step()
pr.disable()
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print(f"host enqueue {1e3 * (t1 - t0) / n:.1f} ms/step (under cProfile), device drained {1e3 * (t2 - t1):.1f} ms later")
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(70)
    print(s.getvalue())
