"""GPU: the tcgen05 tensor-core contraction (tc_gemm.cu) against float64 matmul, all four operand layouts."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_tensor_core_path():
    """Route even tiny contractions through tc_gemm.cu (the product only does so above ops.TC_MIN_FLOP)."""
    from piano_a2s_b200 import ops
    old = ops.TC_MIN_FLOP
    ops.TC_MIN_FLOP = 0.0
    yield
    ops.TC_MIN_FLOP = old


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (128, 256, 64), (300, 173, 515), (1000, 960, 800), (37, 40, 19)])
def test_tc_gemm_bf16x3_matches_fp64(cuda, ta, tb, M, N, K):
    from piano_a2s_b200 import ops
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    Bm = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.t() if ta else A).double() @ (Bm.t() if tb else Bm).double() + bias.double()
    C = torch.empty(M, N, device=cuda)
    ops.gemm(A.to(cuda), Bm.to(cuda), C, M, N, K, transA=bool(ta), transB=bool(tb), lda=A.shape[1], ldb=Bm.shape[1], ldc=N,
             bias=bias.to(cuda), precision="bf16x3")
    e = rel_err(C, ref)
    print(f"tc bf16x3 ta={ta} tb={tb} {M}x{N}x{K}: rel err {e:.2e}")
    assert e < 2e-5
    C1 = torch.empty(M, N, device=cuda)
    ops.gemm(A.to(cuda), Bm.to(cuda), C1, M, N, K, transA=bool(ta), transB=bool(tb), lda=A.shape[1], ldb=Bm.shape[1], ldc=N,
             bias=bias.to(cuda), precision="bf16")
    e1 = rel_err(C1, ref)
    print(f"tc bf16   rel err {e1:.2e}")
    assert e1 < 2e-2


def test_tc_gemm_batched_splitk_transform(cuda):
    from piano_a2s_b200 import ops
    ops_tc_min = ops.TC_MIN_FLOP
    ops.TC_MIN_FLOP = 0.0
    try:
        g = torch.Generator().manual_seed(3)
        M, N, K, P = 200, 96, 320, 40
        A = torch.randn(M, K, generator=g); Bm = torch.randn(N, K, generator=g)
        s = torch.randn(P, generator=g); t = torch.randn(P, generator=g)
        At = F.relu(A * s.repeat(K // P) + t.repeat(K // P))
        C = torch.zeros(M, N, device=cuda)
        ops.gemm(A.to(cuda), Bm.to(cuda), C, M, N, K, transB=True, lda=K, ldb=K, ldc=N, t_scale=s.to(cuda), t_shift=t.to(cuda),
                 t_period=P, t_relu=True, splitk=3, precision="bf16x3")
        assert rel_err(C, At.double() @ Bm.double().t()) < 2e-5
        # transform on a non-transposed B (weight-gradient form), transposed A
        Bn = torch.randn(K, 80, generator=g)
        Bt = F.relu(Bn * s.repeat(2) + t.repeat(2))
        A2 = torch.randn(K, M, generator=g)
        C = torch.empty(M, 80, device=cuda)
        ops.gemm(A2.to(cuda), Bn.to(cuda), C, M, 80, K, transA=True, lda=M, ldb=80, ldc=80, t_scale=s.to(cuda), t_shift=t.to(cuda),
                 t_period=P, t_relu=True, t_on_b=True, precision="bf16x3")
        assert rel_err(C, A2.double().t() @ Bt.double()) < 2e-5
        # batched, reduction over the batch with atomics (dW_hh form), overlapping rows (VQT form)
        y = torch.randn(3, 4000, generator=g)
        W = torch.randn(64, 512, generator=g)
        Tn, hop = 20, 160
        C = torch.empty(3, Tn, 64, device=cuda)
        ops.gemm(y.to(cuda), W.to(cuda), C, Tn, 64, 512, transB=True, lda=hop, ldb=512, ldc=64, batch=3, strideA=4000, strideB=0,
                 strideC=Tn * 64, a_off=8, precision="bf16x3")
        fr = torch.stack([torch.stack([y[b, 8 + i * hop: 8 + i * hop + 512] for i in range(Tn)]) for b in range(3)])
        assert rel_err(C, fr.double() @ W.double().t()) < 2e-5
        X = torch.randn(4, 50, 96, generator=g); D = torch.randn(4, 50, 72, generator=g)
        C = torch.zeros(72, 96, device=cuda)
        ops.gemm(D.to(cuda), X.to(cuda), C, 72, 96, 50, transA=True, lda=72, ldb=96, ldc=96, atomic=True, batch=4, strideA=50 * 72,
                 strideB=50 * 96, strideC=0, precision="bf16x3")
        assert rel_err(C, torch.einsum("bkm,bkn->mn", D.double(), X.double())) < 2e-5
    finally:
        ops.TC_MIN_FLOP = ops_tc_min


@pytest.mark.parametrize("Cin,Cout", [(20, 20), (20, 40), (40, 40)])
@pytest.mark.parametrize("B,T,Fq", [(2, 9, 30), (1, 70, 481)])
def test_tc_conv_forward_and_dgrad_match_torch(cuda, Cin, Cout, B, T, Fq):
    """tc_conv.cu (mode 0 with the fused BatchNorm-apply+ReLU operand transform and batch-statistics epilogue; mode 1 with the
    fused BatchNorm/ReLU backward transform) against float64 torch ops."""
    from piano_a2s_b200 import ops
    from piano_a2s_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(Cin * 100 + Cout + T)
    xraw = torch.randn(B, T, Fq, Cin, generator=g)
    sc = torch.rand(Cin, generator=g) + 0.5
    sh = torch.randn(Cin, generator=g) * 0.3
    W = torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))
    a_in = F.relu(xraw.double() * sc.double() + sh.double())                                   # (B,T,F,Cin)
    ref = F.conv2d(a_in.permute(0, 3, 1, 2), W.double(), None, 1, 1).permute(0, 2, 3, 1)     # (B,T,F,Cout)
    xd, scd, shd, Wd = xraw.to(cuda), sc.to(cuda), sh.to(cuda), W.to(cuda)
    Wpk = ops._tc_pack(Wd, Cout, Cin, 0)
    y = torch.empty(B, T, Fq, Cout, device=cuda)
    npart = lib.pa2s_tc_conv_num_partials(B, T, Fq)
    partial = torch.zeros(npart, 2 * Cout, device=cuda)
    lib.pa2s_tc_conv3x3(stream(), 0, B, T, Fq, Cin, Cout, ptr(xd), ptr(Wpk), ptr(y), ptr(partial), 3, ptr(scd), ptr(shd), 1,
                        None, None, None, None, None, None, None, None)
    e = rel_err(y, ref)
    print(f"tc conv fwd {Cin}->{Cout} B{B} T{T} F{Fq}: rel err {e:.2e}")
    assert e < 2e-5
    sums = partial.double().sum(0).cpu()
    assert rel_err(sums[:Cout], ref.sum((0, 1, 2))) < 1e-4 or ref.sum((0, 1, 2)).abs().max() < 1e-3
    assert rel_err(sums[Cout:], (ref * ref).sum((0, 1, 2))) < 1e-4
    # data gradient: dy = BN/ReLU backward transform of (G, yraw); d a_in = conv_transpose(dy, W)
    yraw = ref.float()
    G = torch.randn(B, T, Fq, Cout, generator=g)
    zs = torch.rand(Cout, generator=g) + 0.5; zb = torch.randn(Cout, generator=g) * 0.2
    mean = torch.randn(Cout, generator=g) * 0.1; invstd = torch.rand(Cout, generator=g) + 0.5
    k1 = torch.rand(Cout, generator=g) + 0.5; k2 = torch.randn(Cout, generator=g) * 0.1; k3 = torch.randn(Cout, generator=g) * 0.1
    z = yraw.double() * zs.double() + zb.double()
    gg = torch.where(z > 0, G.double(), torch.zeros_like(z))
    dy = k1.double() * (gg - k2.double() - (yraw.double() - mean.double()) * invstd.double() * k3.double())
    ref_dx = F.conv_transpose2d(dy.permute(0, 3, 1, 2), W.double(), None, 1, 1).permute(0, 2, 3, 1)
    W2 = ops._tc_pack(Wd, Cout, Cin, 1)
    dx = torch.empty(B, T, Fq, Cin, device=cuda)
    dev = lambda t: t.to(cuda)
    cs = [dev(t) for t in (yraw, zs, zb, mean, invstd, k1, k2, k3)]
    lib.pa2s_tc_conv3x3(stream(), 1, B, T, Fq, Cout, Cin, ptr(dev(G)), ptr(W2), ptr(dx), None, 3, None, None, 1, *[ptr(t) for t in cs])
    e = rel_err(dx, ref_dx)
    print(f"tc conv dgrad: rel err {e:.2e}")
    assert e < 2e-5
    # weight gradient: dW = conv2d_weight(a_in, dy)
    ref_dw = torch.nn.grad.conv2d_weight(a_in.permute(0, 3, 1, 2), W.shape, dy.permute(0, 3, 1, 2), 1, 1)
    nwp = lib.pa2s_tc_conv_wgrad_num_partials(B, T, Fq)
    part = torch.zeros(nwp, Cout * Cin * 9, device=cuda)
    lib.pa2s_tc_conv3x3_wgrad(stream(), B, T, Fq, Cin, Cout, ptr(xd), ptr(dev(G)), ptr(part), 3, ptr(scd), ptr(shd), 1, *[ptr(t) for t in cs])
    dW = part.double().sum(0).view(Cout, Cin, 3, 3)
    e = rel_err(dW, ref_dw)
    print(f"tc conv wgrad: rel err {e:.2e}")
    assert e < 2e-5


@pytest.mark.parametrize("Cin,Cout", [(20, 20), (20, 40), (40, 40)])
# (F = 126, 252: the 126-output strips of conv_tma3_kernel need more strips per row than the 128-position strips of the planes / wgrad)
@pytest.mark.parametrize("B,T,Fq,npieces", [(2, 9, 30, 2), (1, 70, 481, 2), (3, 5, 130, 1), (2, 20, 32, 2), (3, 21, 37, 2), (1, 300, 480, 2),
                                            (2, 6, 126, 2), (1, 5, 252, 2), (1, 4, 125, 1)])
def test_conv_tma_planes_forward_dgrad_wgrad_match_torch(cuda, Cin, Cout, B, T, Fq, npieces):
    """tc_conv_tma.cu: plane producers (BatchNorm-apply+ReLU / BatchNorm-ReLU backward, fused with the bf16 split) feeding the
    bulk-copy-fed tcgen05 convolution (forward with batch-statistics epilogue, data gradient, weight gradient) vs float64 torch."""
    from piano_a2s_b200 import ops
    from piano_a2s_b200._lib import lib, ptr, stream
    tol = 2e-5 if npieces == 2 else 2e-2
    g = torch.Generator().manual_seed(Cin * 100 + Cout + T)
    xraw = torch.randn(B, T, Fq, Cin, generator=g)
    sc = torch.rand(Cin, generator=g) + 0.5
    sh = torch.randn(Cin, generator=g) * 0.3
    W = torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))
    a_in = F.relu(xraw.double() * sc.double() + sh.double())
    ref = F.conv2d(a_in.permute(0, 3, 1, 2), W.double(), None, 1, 1).permute(0, 2, 3, 1)
    xd, scd, shd, Wd = xraw.to(cuda), sc.to(cuda), sh.to(cuda), W.to(cuda)
    Pin = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cin, npieces), device=cuda, dtype=torch.uint8)
    Pin.fill_(0x7f)                                     # the producer must write every halo byte itself
    lib.pa2s_planes_fwd(stream(), B, T, Fq, Cin, ptr(xd), ptr(scd), ptr(shd), 1, ptr(Pin), npieces)
    Wpk = ops._tc_pack(Wd, Cout, Cin, 0)
    y = torch.empty(B, T, Fq, Cout, device=cuda)
    partial = torch.zeros(lib.pa2s_conv_tma_num_partials(B, T, Fq), 2 * Cout, device=cuda)
    lib.pa2s_conv_tma(stream(), B, T, Fq, Cin, Cout, ptr(Pin), npieces, ptr(Wpk), ptr(y), ptr(partial))
    e = rel_err(y, ref)
    print(f"conv_tma fwd {Cin}->{Cout} B{B} T{T} F{Fq} pieces {npieces}: rel err {e:.2e}")
    assert e < tol
    sums = partial.double().sum(0).cpu()
    assert rel_err(sums[Cout:], (y.double().cpu() ** 2).sum((0, 1, 2))) < 1e-4
    assert (sums[:Cout] - y.double().cpu().sum((0, 1, 2))).abs().max() < 1e-3 * (y.double().cpu().abs().sum((0, 1, 2)).max())
    # backward: dy = BN/ReLU backward transform of (G, yraw)
    yraw = ref.float()
    G = torch.randn(B, T, Fq, Cout, generator=g)
    zs = torch.rand(Cout, generator=g) + 0.5; zb = torch.randn(Cout, generator=g) * 0.2
    mean = torch.randn(Cout, generator=g) * 0.1; invstd = torch.rand(Cout, generator=g) + 0.5
    k1 = torch.rand(Cout, generator=g) + 0.5; k2 = torch.randn(Cout, generator=g) * 0.1; k3 = torch.randn(Cout, generator=g) * 0.1
    z = yraw.double() * zs.double() + zb.double()
    gg = torch.where(z > 0, G.double(), torch.zeros_like(z))
    dy = k1.double() * (gg - k2.double() - (yraw.double() - mean.double()) * invstd.double() * k3.double())
    cs = [t.to(cuda) for t in (yraw, zs, zb, mean, invstd, k1, k2, k3)]
    Pdy = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cout, npieces), device=cuda, dtype=torch.uint8)
    Pdy.fill_(0x7f)
    lib.pa2s_planes_bwd(stream(), B, T, Fq, Cout, ptr(G.to(cuda)), *[ptr(t) for t in cs], ptr(Pdy), npieces)
    ref_dx = F.conv_transpose2d(dy.permute(0, 3, 1, 2), W.double(), None, 1, 1).permute(0, 2, 3, 1)
    W2 = ops._tc_pack(Wd, Cout, Cin, 1)
    dx = torch.empty(B, T, Fq, Cin, device=cuda)
    lib.pa2s_conv_tma(stream(), B, T, Fq, Cout, Cin, ptr(Pdy), npieces, ptr(W2), ptr(dx), None)
    e = rel_err(dx, ref_dx)
    d = (dx.double().cpu() - ref_dx).abs()
    print(f"conv_tma dgrad: rel err {e:.2e}  worst at (b,t,f,c) = {tuple(int(i) for i in torch.unravel_index(d.argmax(), d.shape))}")
    assert e < tol
    # the same data gradient with the fused statistics pass of the layer below: [sum g, sum g*xhat], g = dx * (x*s + b > 0)
    s2 = torch.rand(Cin, generator=g) + 0.5; b2 = torch.randn(Cin, generator=g) * 0.2
    mu2 = torch.randn(Cin, generator=g) * 0.1; is2 = torch.rand(Cin, generator=g) + 0.5
    dx2 = torch.empty(B, T, Fq, Cin, device=cuda)
    sp = torch.full((lib.pa2s_conv_tma_num_partials(B, T, Fq), 2 * Cin), float("nan"), device=cuda)
    cs2 = [t.to(cuda) for t in (s2, b2, mu2, is2)]
    lib.pa2s_conv_tma_dgrad_stats(stream(), B, T, Fq, Cout, Cin, ptr(Pdy), npieces, ptr(W2), ptr(dx2), ptr(xd),
                                  *[ptr(t) for t in cs2], ptr(sp))
    assert torch.equal(dx2, dx)
    gm = torch.where(xraw.double() * s2.double() + b2.double() > 0, ref_dx, torch.zeros_like(ref_dx))
    ref_s0 = gm.sum((0, 1, 2))
    ref_s1 = (gm * (xraw.double() - mu2.double()) * is2.double()).sum((0, 1, 2))
    got = sp.double().sum(0).cpu()
    scale0 = gm.abs().sum((0, 1, 2)).max()
    print(f"conv_tma dgrad stats: {float((got[:Cin] - ref_s0).abs().max() / scale0):.2e} {float((got[Cin:] - ref_s1).abs().max() / scale0):.2e}")
    assert (got[:Cin] - ref_s0).abs().max() < max(tol, 1e-4) * scale0
    assert (got[Cin:] - ref_s1).abs().max() < max(tol, 1e-4) * scale0 * 2
    ref_dw = torch.nn.grad.conv2d_weight(a_in.permute(0, 3, 1, 2), W.shape, dy.permute(0, 3, 1, 2), 1, 1)
    part = torch.zeros(lib.pa2s_conv_tma_wgrad_num_partials(B, T, Fq), Cout * Cin * 9, device=cuda)
    lib.pa2s_conv_tma_wgrad(stream(), B, T, Fq, Cin, Cout, ptr(Pin), ptr(Pdy), npieces, ptr(part))
    dW = part.double().sum(0).view(Cout, Cin, 3, 3)
    e = rel_err(dW, ref_dw)
    print(f"conv_tma wgrad: rel err {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("Cin,Cout", [(20, 20), (20, 40), (40, 40)])
@pytest.mark.parametrize("B,T,Fq", [(2, 9, 30), (1, 70, 481), (2, 33, 480)])
def test_three_piece_convolution_is_fp32_level(cuda, Cin, Cout, B, T, Fq):
    """The eval-mode convolution: three tensor-core passes over the bf16 pieces of a = a1+a2+a3 and W = W1+W2+W3 (pa2s_conv_tma,
    2 x pa2s_conv_tma_acc with pa2s_tc_conv_pack3 / pa2s_planes_fwd_low) vs float64 torch: at least as accurate as an fp32 FFMA conv."""
    from piano_a2s_b200._lib import lib, ptr, stream
    g = torch.Generator().manual_seed(Cin * 7 + Cout + T)
    xraw = torch.randn(B, T, Fq, Cin, generator=g)
    sc = torch.rand(Cin, generator=g) + 0.5
    sh = torch.randn(Cin, generator=g) * 0.3
    W = torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))
    a_in = F.relu(xraw.double() * sc.double() + sh.double())
    ref = F.conv2d(a_in.permute(0, 3, 1, 2), W.double(), None, 1, 1).permute(0, 2, 3, 1)
    ref32 = F.conv2d(a_in.float().permute(0, 3, 1, 2), W, None, 1, 1).permute(0, 2, 3, 1)          # what plain fp32 gives (CPU)
    xd, scd, shd, Wd = xraw.to(cuda), sc.to(cuda), sh.to(cuda), W.to(cuda).contiguous()
    P = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cin, 2), device=cuda, dtype=torch.uint8)
    packs = []
    for sel in range(3):
        buf = torch.empty(lib.pa2s_tc_conv_pack_bytes(Cin, Cout), device=cuda, dtype=torch.uint8)
        lib.pa2s_tc_conv_pack3(stream(), ptr(Wd), Cout, Cin, 0, sel, ptr(buf))
        packs.append(buf)
    y = torch.full((B, T, Fq, Cout), float("nan"), device=cuda)
    lib.pa2s_planes_fwd(stream(), B, T, Fq, Cin, ptr(xd), ptr(scd), ptr(shd), 1, ptr(P), 2)
    lib.pa2s_conv_tma(stream(), B, T, Fq, Cin, Cout, ptr(P), 2, ptr(packs[0]), ptr(y), None)
    e2 = rel_err(y, ref)
    lib.pa2s_conv_tma_acc(stream(), B, T, Fq, Cin, Cout, ptr(P), ptr(packs[1]), ptr(y))
    lib.pa2s_planes_fwd_low(stream(), B, T, Fq, Cin, ptr(xd), ptr(scd), ptr(shd), 1, ptr(P))
    lib.pa2s_conv_tma_acc(stream(), B, T, Fq, Cin, Cout, ptr(P), ptr(packs[2]), ptr(y))
    e3, e32 = rel_err(y, ref), rel_err(ref32, ref)
    print(f"conv {Cin}->{Cout} B{B} T{T} F{Fq}: two pieces {e2:.2e}, three pieces {e3:.2e}, plain fp32 {e32:.2e}")
    assert e3 < 1e-6 and e3 <= 3 * e32 and e3 < e2 / 4
