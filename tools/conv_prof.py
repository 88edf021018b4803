"""Development tool: per-tile clock64 stamps of the conv2 forward launch of conv_tma3_kernel (library built with
PA2S_NVCC_DEFS=-DPA2S_CONV_PROF python -m piano_a2s_b200.build --force; rebuild without it afterwards).
Columns (cycles relative to tile 0): MMA warp [before tempty wait, after, after window waits, after issuing the tile],
epilogue [before tfull wait, after, after the tile's stores + tempty arrive]."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import models  # noqa: E402
from piano_a2s_b200 import train  # noqa: E402
from piano_a2s_b200._lib import lib  # noqa: E402
from piano_a2s_b200.synthetic import make_audio, make_ground_truth  # noqa: E402
from piano_a2s_b200.vqt import VQT  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = models.ScoreTranscription(max_length=(398, 189)).to(dev).train()
opt = train.FlatAdadelta(m)
vqt = VQT().to(dev)
audio = make_audio(16, 192000, seed=1234).to(dev)
gt = train.targets_to_device([t.pin_memory() for t in make_ground_truth(16, 5, 398, 189, seed=1234)], dev)
for _ in range(3):
    train.fit_batch(m, opt, vqt(audio).unsqueeze(1), gt, 0.7)
torch.cuda.synchronize()
raw = ctypes.CDLL(os.path.join(ROOT, "piano_a2s_b200", "libpa2s.so"))
buf = (ctypes.c_ulonglong * 512)()
raw.pa2s_conv_tma_prof_read.argtypes = [ctypes.c_void_p]
rc = raw.pa2s_conv_tma_prof_read(buf)
a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
t0 = a[0, 0]
print("rc", rc)
print("tile  m:wait_tempty> m:tempty_ok m:windows_ok m:issued | e:wait_tfull> e:tfull_ok e:done   (cycles since tile 0)")
for i in range(4, 40):
    r = a[i] - t0
    print(f"{i:3d}  {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d}")
d = np.diff(a[8:60, 3])
print("tile period (issue end to issue end): mean", d.mean(), "min", d.min(), "max", d.max())
print("mean m:tempty wait", (a[8:60, 1] - a[8:60, 0]).mean(), "window wait", (a[8:60, 2] - a[8:60, 1]).mean(), "issue", (a[8:60, 3] - a[8:60, 2]).mean(),
      "loop overhead", (a[9:60, 0] - a[8:59, 3]).mean())
print("mean e:tfull wait", (a[8:60, 5] - a[8:60, 4]).mean(), "epilogue work", (a[8:60, 6] - a[8:60, 5]).mean(), "e loop (same group, tile k+2)", (a[10:60, 4] - a[8:58, 6]).mean())
print("tfull_ok(k) - issued(k)", (a[8:60, 5] - a[8:60, 3]).mean(), " tempty_ok(k+2) - e:done(k)", (a[10:60, 1] - a[8:58, 6]).mean())
