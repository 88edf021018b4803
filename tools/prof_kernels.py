"""Runs the dominant kernels alone (for `ncu --set full`): tc conv4 forward, tc out-Linear forward, one note-decoder call."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from piano_a2s_b200 import ops  # noqa: E402
from piano_a2s_b200._lib import lib, ptr, stream  # noqa: E402

dev = torch.device("cuda:0")
B, T, Fq = int(os.environ.get("PB", 4)), 1201, 480
torch.manual_seed(0)
x = torch.randn(B, T, Fq, 40, device=dev)
W = torch.randn(40, 40, 3, 3, device=dev) * 0.05
sc = torch.rand(40, device=dev) + 0.5
sh = torch.randn(40, device=dev) * 0.1
y = torch.empty(B, T, Fq, 40, device=dev)
Wpk = ops._tc_pack(W, 40, 40, 0)
npart = lib.pa2s_tc_conv_num_partials(B, T, Fq)
partial = torch.zeros(npart, 80, device=dev)
for _ in range(3):
    lib.pa2s_tc_conv3x3(stream(), 0, B, T, Fq, 40, 40, ptr(x), ptr(Wpk), ptr(y), ptr(partial), 3, ptr(sc), ptr(sh), 1,
                        None, None, None, None, None, None, None, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.pa2s_tc_conv3x3(stream(), 0, B, T, Fq, 40, 40, ptr(x), ptr(Wpk), ptr(y), ptr(partial), 3, ptr(sc), ptr(sh), 1,
                    None, None, None, None, None, None, None, None)
e1.record()
torch.cuda.synchronize()
print("tc conv4 fwd B=%d: %.3f ms" % (B, e0.elapsed_time(e1)))
M, K, N = B * T, Fq * 40, 256
Wl = torch.randn(N, K, device=dev) * 0.01
z = torch.empty(M, N, device=dev)
for _ in range(2):
    ops.gemm(y, Wl, z, M, N, K, transB=True, lda=K, ldb=K, ldc=N, t_scale=sc, t_shift=sh, t_period=40, t_relu=True, precision="bf16x3")
torch.cuda.synchronize()
e0.record()
ops.gemm(y, Wl, z, M, N, K, transB=True, lda=K, ldb=K, ldc=N, t_scale=sc, t_shift=sh, t_period=40, t_relu=True, precision="bf16x3")
e1.record()
torch.cuda.synchronize()
print("tc linear fwd M=%d: %.3f ms" % (M, e0.elapsed_time(e1)))
