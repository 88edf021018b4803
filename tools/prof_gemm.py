"""Times the TMA-fed tcgen05 GEMM (tc_gemm_tma.cu) and the bf16 split on the contraction shapes of the training step at B clips."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from piano_a2s_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, Fq = int(os.environ.get("PB", 16)), 1201, 480
ITERS = int(os.environ.get("PITERS", 5))
PIECES = int(os.environ.get("PPIECES", 2))
TERMS = {1: 1, 2: 3, 3: 6}[PIECES]
torch.manual_seed(0)


def timed(name, fn, flop=0.0, nbytes=0.0):
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
    extra = ""
    if flop:
        extra += f"  {flop / ms / 1e9:7.1f} TFLOP/s algorithmic ({TERMS * flop / ms / 1e9:7.1f} issued)"
    if nbytes:
        extra += f"  {nbytes / ms / 1e6:7.1f} GB/s"
    print(f"{name:44s} {ms:8.3f} ms{extra}", flush=True)
    return ms


M, K, N = B * T, Fq * 40, 256
y4 = torch.randn(M, K, device=dev)
sc, sh = torch.rand(40, device=dev) + 0.5, torch.randn(40, device=dev) * 0.1
Wl = torch.randn(N, K, device=dev) * 0.01
dz = torch.randn(M, N, device=dev)
z = torch.zeros(M, N, device=dev)
G = torch.empty(M, K, device=dev)
dW = torch.zeros(N, K, device=dev)

timed("split a4 = relu(bn(y4)) (M x 19200)", lambda: ops.split_operand(y4, M, K, K, npieces=PIECES, t_scale=sc, t_shift=sh, t_period=40, t_relu=True),
      nbytes=M * K * (4 + 2 * PIECES))
a4 = ops.split_operand(y4, M, K, K, npieces=PIECES, t_scale=sc, t_shift=sh, t_period=40, t_relu=True)
timed("split W_out (256 x 19200)", lambda: ops.split_operand(Wl, N, K, K, npieces=PIECES), nbytes=N * K * (4 + 2 * PIECES))
Wop = ops.split_operand(Wl, N, K, K, npieces=PIECES)
timed("split dz (M x 256)", lambda: ops.split_operand(dz, M, N, N, npieces=PIECES), nbytes=M * N * (4 + 2 * PIECES))
dzop = ops.split_operand(dz, M, N, N, npieces=PIECES)
for sk in (1, 2, 4, 7, 8):
    timed(f"linear fwd   M={M} N=256 K=19200 splitk={sk}", lambda: ops.gemm_bf16(a4, False, Wop, False, z, M, N, K, ldc=N, splitk=sk), 2.0 * M * N * K)
timed(f"linear dgrad M={M} N=19200 K=256", lambda: ops.gemm_bf16(dzop, False, Wop, True, G, M, K, N, ldc=K), 2.0 * M * N * K, nbytes=M * K * 4)
for sk in (1, 2, 4):
    timed(f"linear wgrad M=256 N=19200 K={M} splitk={sk}", lambda: ops.gemm_bf16(dzop, True, a4, True, dW, N, K, M, ldc=K, splitk=sk), 2.0 * M * N * K)
del y4, a4, G, dW

# encoder input projections / attention memory projections
for (n, k) in ((1536, 256), (1536, 512), (256, 512)):
    x = torch.randn(M, k, device=dev)
    W = torch.randn(n, k, device=dev) * 0.05
    out = torch.empty(M, n, device=dev)
    xo, wo = ops.split_operand(x, M, k, k, npieces=PIECES), ops.split_operand(W, n, k, k, npieces=PIECES)
    timed(f"proj M={M} N={n} K={k} (pre-split)", lambda: ops.gemm_bf16(xo, False, wo, False, out, M, n, k, ldc=n), 2.0 * M * n * k, nbytes=M * n * 4)
    timed(f"proj M={M} N={n} K={k} (split + gemm)", lambda: ops.gemm(x, W, out, M, n, k, transB=True, lda=k, ldb=k, ldc=n,
                                                                     precision={1: "bf16", 2: "bf16x3", 3: "bf16x6"}[PIECES]), 2.0 * M * n * k)

# VQT filterbank: overlapping frames of the padded clip (row pitch = hop), 960 filters x 1008 taps
from piano_a2s_b200.vqt import VQT  # noqa: E402
v = VQT().to(dev)
audio = torch.clamp(0.25 * torch.randn(B, 192000, device=dev), -1, 1)
Kv = v.filters.shape[1]
timed("VQT module (split + contraction + dB epilogue)", lambda: v(audio), 2.0 * T * B * 960 * Kv)
