"""Runs the dominant ConvStack kernels alone at full size (for `ncu --set full`): tcgen05 conv4 forward, data gradient,
weight gradient and the `out` Linear forward.  PB = clips (default 16 = bench batch), PITERS = launches of each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from piano_a2s_b200 import ops  # noqa: E402
from piano_a2s_b200._lib import lib, ptr, stream  # noqa: E402

dev = torch.device("cuda:0")
B, T, Fq = int(os.environ.get("PB", 16)), 1201, 480
ITERS = int(os.environ.get("PITERS", 3))
NSPLIT = int(os.environ.get("PNSPLIT", 3))
torch.manual_seed(0)
x = torch.randn(B, T, Fq, 40, device=dev)
g = torch.randn(B, T, Fq, 40, device=dev)
W = torch.randn(40, 40, 3, 3, device=dev) * 0.05
c = lambda: torch.rand(40, device=dev) + 0.5
sc, sh, mean, invstd, k1, k2, k3 = c(), c() * 0.1, c() * 0.1, c(), c(), c() * 0.01, c() * 0.01
y = torch.empty(B, T, Fq, 40, device=dev)
Wpk = ops._tc_pack(W, 40, 40, 0)
Wpk2 = ops._tc_pack(W, 40, 40, 1)
npart = lib.pa2s_tc_conv_num_partials(B, T, Fq)
partial = torch.zeros(npart, 80, device=dev)
nwp = lib.pa2s_tc_conv_wgrad_num_partials(B, T, Fq)
wpartial = torch.empty(nwp, 40 * 40 * 9, device=dev)


def timed(name, fn, flop):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    print(f"{name} B={B}: {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)


conv_flop = 2.0 * T * Fq * 9 * 40 * 40 * B
timed("tc conv4 fwd", lambda: lib.pa2s_tc_conv3x3(stream(), 0, B, T, Fq, 40, 40, ptr(x), ptr(Wpk), ptr(y), ptr(partial), NSPLIT, ptr(sc), ptr(sh), 1,
                                                   None, None, None, None, None, None, None, None), conv_flop)
timed("tc conv4 dgrad", lambda: lib.pa2s_tc_conv3x3(stream(), 1, B, T, Fq, 40, 40, ptr(g), ptr(Wpk2), ptr(y), None, NSPLIT, None, None, 1,
                                                     ptr(x), ptr(sc), ptr(sh), ptr(mean), ptr(invstd), ptr(k1), ptr(k2), ptr(k3)), conv_flop)
timed("tc conv4 wgrad", lambda: lib.pa2s_tc_conv3x3_wgrad(stream(), B, T, Fq, 40, 40, ptr(x), ptr(g), ptr(wpartial), NSPLIT, ptr(sc), ptr(sh), 1,
                                                           ptr(x), ptr(sc), ptr(sh), ptr(mean), ptr(invstd), ptr(k1), ptr(k2), ptr(k3)), conv_flop)
M, K, N = B * T, Fq * 40, 256
Wl = torch.randn(N, K, device=dev) * 0.01
z = torch.zeros(M, N, device=dev)
prec = "bf16x3" if NSPLIT == 3 else "bf16"
timed("tc linear fwd", lambda: ops.gemm(x, Wl, z, M, N, K, transB=True, lda=K, ldb=K, ldc=N, t_scale=sc, t_shift=sh, t_period=40, t_relu=True,
                                         precision=prec, zeroed=True), 2.0 * M * N * K)
