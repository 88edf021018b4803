"""Host-side operators: torch.autograd.Functions whose forward AND backward are libpa2s kernels.

PyTorch here is plumbing only (device memory, streams, autograd bookkeeping between fused groups).  Every function
raises if its inputs are not CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import collections
import contextlib
import ctypes
import os
import sys
from typing import Optional

import torch
import torch.distributed as dist

from ._lib import lib, ptr, stream, make_dec_args, make_decm_args

F32 = torch.float32
N_SM = 148


def _f(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("piano_a2s_b200 operators run on CUDA (sm_100a) only; got a CPU tensor")
    if t.dtype != F32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class KernelTimers:
    """Optional CUDA-event brackets around named kernel launches (bench.py roofline leg).  Disabled by default."""
    enabled = False
    events = {}

    @classmethod
    def reset(cls, enabled):
        cls.enabled = enabled
        cls.events = {}

    @classmethod
    def summary(cls):
        """-> {name: (launches, mean_ms)}; call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in cls.events.items() if v}


class ktime:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KernelTimers.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if KernelTimers.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            KernelTimers.events.setdefault(self.name, []).append((self.a, b))


def _on_stream(s):
    return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ----------------------------------------------------------------------------------------------------------------
# GEMM wrapper
# ----------------------------------------------------------------------------------------------------------------
# Contraction precision: "fp32"   exact-fp32 FFMA kernel (gemm.cu);
#                        "bf16x3" tcgen05, operands split into bf16 hi/lo, 3 MMAs per product (~fp32 accuracy);
#                        "bf16"   tcgen05, bf16 operands, fp32 accumulation (BASELINE configs[2..3]).
# Small contractions stay on the FFMA kernel (a 128-row UMMA tile would be mostly padding).
PRECISION = {"train": "bf16x3", "eval": "fp32"}
EVAL_CONV = os.environ.get("PA2S_EVAL_CONV", "bf16x6")         # conv2-4 in eval(): "bf16x6" (three tensor-core passes over bf16 pieces, fp32-level) or "fp32" (FFMA)
EVAL_LINEAR = os.environ.get("PA2S_EVAL_LINEAR", "bf16x6")     # the 19200 -> 256 projection in eval(): "bf16x6" (tensor cores, fp32-level) or "fp32" (FFMA)
TC_MIN_FLOP = 2.0e8
_CURRENT = ["bf16x3"]
_DEPTH = [0]                # > 0 inside a use_precision scope


def set_precision(train=None, eval=None):
    """Contraction precision used by modules in train() / eval() mode."""
    for k, v in (("train", train), ("eval", eval)):
        if v is not None:
            assert v in ("fp32", "bf16x3", "bf16", "bf16x6")
            PRECISION[k] = v


def current_precision():
    return _CURRENT[0]


class use_precision:
    """Scope set by the module forwards (`PRECISION["train" if module.training else "eval"]`); autograd Functions read it in
    forward and remember it for their backward."""

    def __init__(self, prec):
        assert prec in ("fp32", "bf16x3", "bf16", "bf16x6")
        self.prec = prec

    def __enter__(self):
        self.old = _CURRENT[0]
        _CURRENT[0] = self.prec
        _DEPTH[0] += 1

    def __exit__(self, *exc):
        _CURRENT[0] = self.old
        _DEPTH[0] -= 1


def _bwd_precision(fn):
    """backward runs under the precision its forward ran with"""
    def wrapper(ctx, *grads):
        with use_precision(ctx.prec):
            return fn(ctx, *grads)
    return wrapper


def module_precision(module):
    """Scope entered by EVERY public module forward (ScoreTranscription and the sub-modules the reference exposes: ConvStack, Encoder,
    HierarchicalDecoder, NoteDecoder, AttentionLayer): `PRECISION["train" | "eval"]` by the module's mode, so that a sub-module called
    on its own in eval() runs the exact-fp32 kernels too.  An enclosing scope (the parent module's, or an explicit `use_precision`)
    wins."""
    if _DEPTH[0] > 0:
        return contextlib.nullcontext()
    return use_precision(PRECISION["train" if module.training else "eval"])


class Bf16Operand:
    """An fp32 matrix [rows][cols] (x batch) held as 1..3 bf16 pieces for the TMA-fed tensor-core GEMM (tc_gemm_tma.cu)."""
    __slots__ = ("buf", "rows", "cols", "ld", "npieces", "batch", "piece_stride", "batch_stride")

    def __init__(self, buf, rows, cols, ld, npieces, batch, piece_stride, batch_stride):
        self.buf, self.rows, self.cols, self.ld, self.npieces, self.batch = buf, rows, cols, ld, npieces, batch
        self.piece_stride, self.batch_stride = piece_stride, batch_stride


def npieces_for(prec):
    return {"bf16": 1, "bf16x3": 2, "bf16x6": 3}[prec]


def split_operand(src, rows, cols, ld_src, *, off=0, batch=1, batch_stride=0, npieces=2, t_scale=None, t_shift=None, t_period=1,
                  t_relu=False):
    """fp32 storage `src` (+ element offset) viewed as `batch` matrices [rows][cols] of row pitch ld_src -> Bf16Operand."""
    ld = (cols + 7) // 8 * 8
    shared = batch > 1 and batch_stride == 0
    nb = 1 if shared else batch
    buf = torch.empty(nb, npieces, rows, ld, device=src.device, dtype=torch.bfloat16)
    lib.pa2s_split_bf16(stream(), ctypes.c_void_p(src.data_ptr() + off * 4), rows, cols, ld_src, batch_stride, ptr(buf), ld, rows * ld,
                        npieces * rows * ld, npieces, nb, ptr(t_scale), ptr(t_shift), t_period, int(t_relu))
    return Bf16Operand(buf, rows, cols, ld, npieces, batch, rows * ld, 0 if (shared or batch == 1) else npieces * rows * ld)


def gemm_bf16(Aop, a_mn, Bop, b_mn, C, M, N, K, *, ldc, bias=None, atomic=False, batch=1, strideC=0, splitk=1, c_off=0,
              a_ld=None, b_ld=None):
    """C (+)= sum of piece products of two Bf16Operands (see include/pa2s.h: pa2s_gemm_bf16_tma)."""
    lib.pa2s_gemm_bf16_tma(stream(), M, N, K,
                           ptr(Aop.buf), a_ld or Aop.ld, Aop.piece_stride, Aop.batch_stride, Aop.npieces, int(a_mn),
                           ptr(Bop.buf), b_ld or Bop.ld, Bop.piece_stride, Bop.batch_stride, Bop.npieces, int(b_mn),
                           ctypes.c_void_p(C.data_ptr() + c_off * 4), ldc, strideC, ptr(bias), int(atomic), batch, splitk)
    return C


def _operand(X, trans, mn, k, ld, off, batch, stride, npieces, tf):
    """Storage X described BLAS-style -> (Bf16Operand, mn_major, row pitch override).  Non-transposed A / transposed B are stored
    [mn][k] (K-major); the other two [k][mn].  ld < row length (overlapping rows, the VQT framing) splits the underlying
    1-D signal once and lets the tensor map stride over it."""
    if isinstance(X, Bf16Operand):
        return X, trans, None
    rows, cols = (k, mn) if trans else (mn, k)
    if ld < cols:
        span = (rows - 1) * ld + cols
        op = split_operand(X, 1, span, span, off=off, batch=batch, batch_stride=stride, npieces=npieces, **tf)
        assert ld % 8 == 0
        op.rows, op.cols = rows, cols
        return op, trans, ld
    return split_operand(X, rows, cols, ld, off=off, batch=batch, batch_stride=stride, npieces=npieces, **tf), trans, None


def gemm(A, B, C, M, N, K, *, transA=False, transB=False, lda=None, ldb=None, ldc=None, bias=None, accumulate=False,
         atomic=False, batch=1, strideA=0, strideB=0, strideC=0, t_scale=None, t_shift=None, t_period=1, t_relu=False,
         t_on_b=False, splitk=1, a_off=0, b_off=0, c_off=0, precision=None, zeroed=False):
    """C = op(A) op(B); A/B/C are tensors used as raw storage (+ element offsets), see include/pa2s.h.  A / B may also be
    pre-split Bf16Operands (tensor-core precisions only).
    `zeroed=True`: the caller guarantees C is zero-filled, so the wrapper may split K (atomic accumulation) to fill the GPU."""
    es = 4
    prec = precision or current_precision()
    presplit = isinstance(A, Bf16Operand) or isinstance(B, Bf16Operand)
    use_tc = prec != "fp32" and K > 0 and (presplit or 2.0 * M * N * K * batch >= TC_MIN_FLOP)
    if zeroed and splitk == 1 and batch == 1:
        splitk = _tc_splitk(M, N, K) if use_tc else _auto_splitk(M, N, K)
    if use_tc:
        tf = dict(t_scale=t_scale, t_shift=t_shift, t_period=t_period, t_relu=t_relu)
        none = dict(t_scale=None, t_shift=None, t_period=1, t_relu=False)
        npc = npieces_for(prec)
        Aop, a_mn, a_ld = _operand(A, transA, M, K, lda, a_off, batch, strideA, npc, none if (t_on_b or t_scale is None) else tf)
        Bop, b_mn, b_ld = _operand(B, not transB, N, K, ldb, b_off, batch, strideB, npc, tf if (t_on_b and t_scale is not None) else none)
        return gemm_bf16(Aop, a_mn, Bop, b_mn, C, M, N, K, ldc=ldc, bias=bias, atomic=atomic or accumulate, batch=batch,
                         strideC=strideC, splitk=splitk, c_off=c_off, a_ld=a_ld, b_ld=b_ld)
    assert not isinstance(A, Bf16Operand) and not isinstance(B, Bf16Operand)
    pa = ctypes.c_void_p(A.data_ptr() + a_off * es)
    pb = ctypes.c_void_p(B.data_ptr() + b_off * es)
    pc = ctypes.c_void_p(C.data_ptr() + c_off * es)
    lib.pa2s_gemm_f32(stream(), int(transA), int(transB), M, N, K, pa, lda, pb, ldb, pc, ldc, ptr(bias), int(accumulate),
                      int(atomic), batch, strideA, strideB, strideC, ptr(t_scale), ptr(t_shift), t_period, int(t_relu),
                      int(t_on_b), splitk)
    return C


def _auto_splitk(M, N, K):
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if tiles >= N_SM or K < 2048:
        return 1
    return max(1, min(K // 512, (2 * N_SM) // tiles))


def _tc_splitk(M, N, K):
    """split-K factor for the persistent 128 x min(N,256) tensor-core tiles: minimise the makespan ceil(items/148)/split"""
    bn = 256 if N >= 256 else (N + 31) // 32 * 32
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    best, best_cost = 1, None
    for sk in range(1, 9):
        if sk > 1 and K // sk < 256:
            break
        cost = -(-tiles * sk // N_SM) / sk
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = sk, cost
    return best


def colsum(X2d: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate=False) -> torch.Tensor:
    """Column sums of a (R, N) fp32 matrix (fp64 accumulation)."""
    R, N = X2d.shape
    if out is None:
        out = torch.empty(N, device=X2d.device, dtype=F32)
    if R >= 512:
        # tall: spread the rows over ~2 CTAs per SM (two deterministic stages) instead of ceil(N/32) CTAs walking all rows
        nchunks = max(1, min((R + 63) // 64, (2 * N_SM) // ((N + 31) // 32)))
        scratch = torch.empty(nchunks, N, device=X2d.device, dtype=torch.float64)
        lib.pa2s_colsum(stream(), ptr(X2d), R, N, ptr(scratch), nchunks, ptr(out), int(accumulate))
    else:
        lib.pa2s_reduce_rows(stream(), ptr(X2d), R, N, None, ptr(out), int(accumulate))
    return out


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (F.linear).  W may be a column-sliced view (stride(1) == 1)."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.prec = current_precision()
        x2 = _f(x).reshape(-1, x.shape[-1])
        assert W.stride(1) == 1 and W.dtype == F32
        M, K = x2.shape
        N = W.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=F32)
        gemm(x2, W, y, M, N, K, transB=True, lda=K, ldb=W.stride(0), ldc=N, bias=b)
        ctx.save_for_backward(x2, W)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    @_bwd_precision
    def backward(ctx, dy):
        x2, W = ctx.saved_tensors
        M, K = x2.shape
        N = W.shape[0]
        dy2 = _f(dy).reshape(M, N)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=dy.device, dtype=F32)
            gemm(dy2, W, dx, M, K, N, lda=N, ldb=W.stride(0), ldc=K)
            dx = dx.reshape(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dW = torch.zeros(N, K, device=dy.device, dtype=F32)
            gemm(dy2, x2, dW, N, K, M, transA=True, lda=N, ldb=K, ldc=K, zeroed=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dW, db


def linear(x, W, b=None):
    return LinearFn.apply(x, W, b)


class SplitColsFn(torch.autograd.Function):
    """(M, n*w) -> n contiguous (M, w) blocks; backward writes the n gradients side by side into ONE (M, n*w) buffer (autograd's own
    slicing would zero-fill and add a full-size tensor per block)."""

    @staticmethod
    def forward(ctx, x, n):
        M, N = x.shape
        ctx.n = n
        w = N // n
        return tuple(x[:, i * w:(i + 1) * w].contiguous() for i in range(n))

    @staticmethod
    def backward(ctx, *gs):
        M, w = gs[0].shape
        out = torch.empty(M, ctx.n * w, device=gs[0].device, dtype=gs[0].dtype)
        for i, g in enumerate(gs):
            out[:, i * w:(i + 1) * w].copy_(g)
        return out, None


# ---- Linear layers that are applied once per bar (bar-level GRU cell, attention query): deferred weight gradients ------------------
class LinearSink:
    """Rows (x, dy) left behind by the backward of every use of one Linear in a forward pass."""
    def __init__(self):
        self.rows = []


class LinearSinkFn(torch.autograd.Function):
    """Identity on (W[, b]).  Every LinearDeferFn call that uses the returned aliases only forms its INPUT gradient (the part that is on
    the bar-level chain) and leaves (x, dy) in the sink; autograd runs this node after the last of them, and the weight / bias gradient
    is then ONE contraction over the rows of all bars: 5x fewer small GEMMs, zero fills and gradient-accumulation adds between the
    note decoders' reverse pass and the encoder's."""

    @staticmethod
    def forward(ctx, sink, W, b):
        ctx.sink, ctx.prec, ctx.has_bias = sink, current_precision(), b is not None
        ctx.wshape = W.shape
        ctx.set_materialize_grads(False)
        return (W.view_as(W), b.view_as(b)) if b is not None else (W.view_as(W), None)

    @staticmethod
    @_bwd_precision
    def backward(ctx, gW, gb):
        rows, ctx.sink.rows = ctx.sink.rows, []
        dW, db = gW, gb
        if rows:
            X = rows[0][0] if len(rows) == 1 else torch.cat([r[0] for r in rows])
            DY = rows[0][1] if len(rows) == 1 else torch.cat([r[1] for r in rows])
            M, K = X.shape
            N = DY.shape[1]
            d = torch.zeros(N, K, device=X.device, dtype=F32)
            gemm(DY, X, d, N, K, M, transA=True, lda=N, ldb=K, ldc=K, zeroed=True)
            dW = d if dW is None else dW + d
            if ctx.has_bias:
                c = colsum(DY)
                db = c if db is None else db + c
        return None, dW, (db if ctx.has_bias else None)


class LinearDeferFn(torch.autograd.Function):
    """y = x W^T + b like LinearFn, with W / b aliases from LinearSinkFn: backward returns dx only and hands (x, dy) to the sink."""

    @staticmethod
    def forward(ctx, x, W, b, sink):
        ctx.prec = current_precision()
        x2 = _f(x).reshape(-1, x.shape[-1])
        assert W.stride(1) == 1 and W.dtype == F32
        M, K = x2.shape
        N = W.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=F32)
        gemm(x2, W, y, M, N, K, transB=True, lda=K, ldb=W.stride(0), ldc=N, bias=b)
        ctx.save_for_backward(x2, W)
        ctx.sink, ctx.xshape = sink, x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    @_bwd_precision
    def backward(ctx, dy):
        x2, W = ctx.saved_tensors
        M, K = x2.shape
        N = W.shape[0]
        dy2 = _f(dy).reshape(M, N)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=dy.device, dtype=F32)
            gemm(dy2, W, dx, M, K, N, lda=N, ldb=W.stride(0), ldc=K)
            dx = dx.reshape(ctx.xshape)
        ctx.sink.rows.append((x2, dy2))
        return dx, None, None, None


class DeferredLinear:
    """One Linear of the bar-level chain for one forward pass: `lin(x)` = F.linear(x, W, b) with the weight gradient deferred."""

    def __init__(self, W, b=None):
        self.sink = LinearSink()
        self.defer = torch.is_grad_enabled() and W.requires_grad
        if self.defer:
            self.W, self.b = LinearSinkFn.apply(self.sink, W, b)
        else:
            self.W, self.b = W, b

    def __call__(self, x):
        if self.defer:
            return LinearDeferFn.apply(x, self.W, self.b, self.sink)
        return LinearFn.apply(x, self.W, self.b)


# ----------------------------------------------------------------------------------------------------------------
# ConvStack (models.py:463-543) as one fused group
# ----------------------------------------------------------------------------------------------------------------
def _nsplit(prec):
    return 3 if prec == "bf16x3" else 1


def _tc_pack(W, Cout, Cin, dgrad):
    """fp32 (Cout,Cin,3,3) -> bf16 hi/lo UMMA weight blocks for tc_conv.cu"""
    kin, nout = (Cout, Cin) if dgrad else (Cin, Cout)
    buf = torch.empty(lib.pa2s_tc_conv_pack_bytes(kin, nout), device=W.device, dtype=torch.uint8)
    lib.pa2s_tc_conv_pack(stream(), ptr(W.detach().contiguous()), Cout, Cin, dgrad, ptr(buf))
    return buf


def _bn_sums(partial, C, count=None):
    """fp64 column sums of the per-CTA partials; with `count` the vector gets a third part [local element count], so that one
    all-reduce carries sums AND count (torch.nn.SyncBatchNorm gathers per-rank counts: ranks may hold different batch sizes)."""
    sums = torch.empty(2 * C + (1 if count is not None else 0), device=partial.device, dtype=torch.float64)
    if count is not None:
        sums[2 * C:].fill_(float(count))
    lib.pa2s_reduce_rows(stream(), ptr(partial), partial.shape[0], 2 * C, ptr(sums), None, 0)
    return sums


class ConvStackFn(torch.autograd.Function):
    """spectrogram (B,1,T,F) -> relu(out_bn(out(flatten(relu(bn4(conv4(...)))))))*mask  (B,T,O).

    Argument order: spec, mask(or None), then for i=1..4: conv_i.weight, bn_i.weight, bn_i.bias, then out.weight,
    out_bn.weight, out_bn.bias; `bufs` = list of 5 (running_mean, running_var) pairs (updated in place when training);
    `training`, `sync` (cross-rank BatchNorm statistics, SyncBatchNorm semantics), eps, momentum."""

    @staticmethod
    def forward(ctx, spec, mask, bufs, training, sync, eps, momentum, *params):
        ctx.prec = current_precision()
        spec = _f(spec)
        B, Cin0, T, Fq = spec.shape
        assert Cin0 == 1, "in_channels must be 1 (pretrain.yaml:85)"
        dev = spec.device
        st = stream()
        conv_w = [params[3 * i] for i in range(4)]
        gam = [params[3 * i + 1] for i in range(4)] + [params[13]]
        bet = [params[3 * i + 2] for i in range(4)] + [params[14]]
        Wout = params[12]
        O = Wout.shape[0]
        world = dist.get_world_size() if (sync and _dist_on()) else 1
        ntile = 4
        ys, affs, planes = [], [], [None]
        xin, in_scale, in_shift = spec, None, None           # (B,1,T,F) == (B,T,F,1) channels-last
        for i in range(4):
            W = conv_w[i]
            Cout, Cin = W.shape[0], W.shape[1]
            y = torch.empty(B, T, Fq, Cout, device=dev, dtype=F32)
            use_tc = ctx.prec != "fp32" and Cin >= 16
            if (ctx.prec == "fp32" and not training and EVAL_CONV == "bf16x6" and Cin >= 16 and lib.pa2s_conv_tma_get_impl()
                    and 2.0 * 9 * Cin * Cout * B * T * Fq >= TC_MIN_FLOP):
                # eval / greedy decode: the exact-fp32 convolution as three tensor-core passes over the bf16 pieces of a = a1 + a2 + a3
                # and W = W1 + W2 + W3 (every piece product except a3*W3: ~2^-24 relative, the accuracy of an fp32 FMA chain)
                # instead of 7-27 ms of FFMA per layer at B = 32
                Wc = W.detach().contiguous()

                def pack3(sel):
                    buf = torch.empty(lib.pa2s_tc_conv_pack_bytes(Cin, Cout), device=dev, dtype=torch.uint8)
                    lib.pa2s_tc_conv_pack3(st, ptr(Wc), Cout, Cin, 0, sel, ptr(buf))
                    return buf
                P = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cin, 2), device=dev, dtype=torch.uint8)
                with ktime(f"conv{i + 1}_planes"):
                    lib.pa2s_planes_fwd(st, B, T, Fq, Cin, ptr(xin), ptr(in_scale), ptr(in_shift), 1, ptr(P), 2)
                with ktime(f"conv{i + 1}_fwd"):
                    lib.pa2s_conv_tma(st, B, T, Fq, Cin, Cout, ptr(P), 2, ptr(pack3(0)), ptr(y), None)
                    lib.pa2s_conv_tma_acc(st, B, T, Fq, Cin, Cout, ptr(P), ptr(pack3(1)), ptr(y))
                with ktime(f"conv{i + 1}_planes"):
                    lib.pa2s_planes_fwd_low(st, B, T, Fq, Cin, ptr(xin), ptr(in_scale), ptr(in_shift), 1, ptr(P))
                with ktime(f"conv{i + 1}_fwd"):
                    lib.pa2s_conv_tma_acc(st, B, T, Fq, Cin, Cout, ptr(P), ptr(pack3(2)), ptr(y))
                del P
                partial = None
            elif use_tc:
                # a_{i-1} = relu(bn(y_{i-1})) as bf16 planes: written once, read by this convolution and by its weight gradient
                npc = min(npieces_for(ctx.prec), 2)
                Pin = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cin, npc), device=dev, dtype=torch.uint8)
                with ktime(f"conv{i + 1}_planes"):
                    lib.pa2s_planes_fwd(st, B, T, Fq, Cin, ptr(xin), ptr(in_scale), ptr(in_shift), 1, ptr(Pin), npc)
                planes.append(Pin)
                Wpk = _tc_pack(W, Cout, Cin, 0)
                nparts = lib.pa2s_conv_tma_num_partials(B, T, Fq)
                partial = torch.empty(nparts, 2 * Cout, device=dev, dtype=F32) if training else None
                with ktime(f"conv{i + 1}_fwd"):
                    lib.pa2s_conv_tma(st, B, T, Fq, Cin, Cout, ptr(Pin), npc, ptr(Wpk), ptr(y), ptr(partial))
            elif Cin == 1 and Cout == 20:
                # conv1: one input channel, nothing to contract on the tensor cores -- a coalesced fp32 stream (conv1.cu), every mode
                nparts = 2 * N_SM
                partial = torch.empty(nparts, 2 * Cout, device=dev, dtype=F32) if training else None
                with ktime(f"conv{i + 1}_fwd"):
                    lib.pa2s_conv1_fwd(st, B, T, Fq, ptr(xin), ptr(W.detach().contiguous()), ptr(y), ptr(partial), nparts)
            else:
                Wp = W.detach().permute(2, 3, 1, 0).contiguous()
                nparts = lib.pa2s_conv3x3_num_partials(B, T, Fq, ntile)
                partial = torch.empty(nparts, 2 * Cout, device=dev, dtype=F32) if training else None
                with ktime(f"conv{i + 1}_fwd"):
                    lib.pa2s_conv3x3(st, 0, B, T, Fq, Cin, Cout, ptr(xin), ptr(Wp), ptr(y), ptr(partial), ntile,
                                     ptr(in_scale), ptr(in_shift), 1, None, None, None, None, None, None, None, None)
            aff = torch.empty(4, Cout, device=dev, dtype=F32)      # scale, shift, mean, invstd
            if training:
                sums = _bn_sums(partial, Cout, B * T * Fq if world > 1 else None)
                if world > 1:
                    dist.all_reduce(sums)
                rm, rv = bufs[i]
                lib.pa2s_bn_finalize(st, ptr(sums), -1.0 if world > 1 else float(B * T * Fq), Cout, ptr(gam[i]), ptr(bet[i]), eps, momentum,
                                     ptr(rm), ptr(rv), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]))
            else:
                rm, rv = bufs[i]
                lib.pa2s_bn_eval_affine(st, Cout, ptr(gam[i]), ptr(bet[i]), ptr(rm), ptr(rv), eps,
                                        ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]))
            ys.append(y)
            affs.append(aff)
            xin, in_scale, in_shift = y, aff[0], aff[1]
        C4 = conv_w[3].shape[0]
        Kf = Fq * C4
        # reference feature index is c*F+f (models.py:537); ours is f*C+c
        Wp_out = Wout.detach().view(O, C4, Fq).permute(0, 2, 1).reshape(O, Kf).contiguous()
        M = B * T
        z = torch.zeros(M, O, device=dev, dtype=F32)
        ctx.lin_ops = None
        if ctx.prec == "fp32" and not training and EVAL_LINEAR == "bf16x6" and 2.0 * M * O * Kf >= TC_MIN_FLOP:
            # eval / greedy decode: the one large contraction of the exact-fp32 mode on the tensor cores with THREE bf16 pieces per
            # operand (3 x 8 mantissa bits = an fp32 value exactly; the six leading piece products, fp32 TMEM accumulation: ~2^-24
            # relative, the accuracy of an fp32 FMA chain) instead of 15 ms of FFMA at B = 32
            with ktime("out_linear_split"):
                a4op = split_operand(ys[3], M, Kf, Kf, npieces=3, t_scale=affs[3][0], t_shift=affs[3][1], t_period=C4, t_relu=True)
                Wop = split_operand(Wp_out, O, Kf, Kf, npieces=3)
            with ktime("out_linear_fwd"):
                gemm(a4op, Wop, z, M, O, Kf, transB=True, ldc=O, zeroed=True, precision="bf16x6")
        elif ctx.prec == "fp32":
            with ktime("out_linear_fwd"):
                gemm(ys[3], Wp_out, z, M, O, Kf, transB=True, lda=Kf, ldb=Kf, ldc=O,
                     t_scale=affs[3][0], t_shift=affs[3][1], t_period=C4, t_relu=True, zeroed=True)
        else:
            # a4 = relu(bn4(y4)) and the permuted weight are split into bf16 pieces ONCE; the forward contraction, the weight
            # gradient (a4 as MN-major operand) and the data gradient (W as MN-major operand) all read these pieces through TMA
            npc = npieces_for(ctx.prec)
            with ktime("out_linear_split"):
                a4op = split_operand(ys[3], M, Kf, Kf, npieces=npc, t_scale=affs[3][0], t_shift=affs[3][1], t_period=C4, t_relu=True)
                Wop = split_operand(Wp_out, O, Kf, Kf, npieces=npc)
            with ktime("out_linear_fwd"):
                gemm(a4op, Wop, z, M, O, Kf, transB=True, ldc=O, zeroed=True)
            if training:
                ctx.lin_ops = (a4op, Wop)
        ctx.planes = planes if training else None
        aff5 = torch.empty(4, O, device=dev, dtype=F32)
        if training:
            nct = 4 * N_SM
            partial = torch.empty(nct, 2 * O, device=dev, dtype=F32)
            lib.pa2s_colstats(st, 0, ptr(z), None, None, M, O, None, None, None, None, ptr(partial), nct)
            sums = _bn_sums(partial, O, M if world > 1 else None)
            if world > 1:
                dist.all_reduce(sums)
            rm, rv = bufs[4]
            lib.pa2s_bn_finalize(st, ptr(sums), -1.0 if world > 1 else float(M), O, ptr(gam[4]), ptr(bet[4]), eps, momentum,
                                 ptr(rm), ptr(rv), ptr(aff5[0]), ptr(aff5[1]), ptr(aff5[2]), ptr(aff5[3]))
        else:
            rm, rv = bufs[4]
            lib.pa2s_bn_eval_affine(st, O, ptr(gam[4]), ptr(bet[4]), ptr(rm), ptr(rv), eps,
                                    ptr(aff5[0]), ptr(aff5[1]), ptr(aff5[2]), ptr(aff5[3]))
        out = torch.empty(B, T, O, device=dev, dtype=F32)
        mk = _f(mask) if mask is not None else None
        lib.pa2s_bn_relu_mask(st, ptr(z), ptr(aff5[0]), ptr(aff5[1]), ptr(mk), ptr(out), M, O)
        ctx.save_for_backward(spec, mk, z, aff5, Wp_out, *ys, *affs, *conv_w, *gam)
        ctx.dims = (B, T, Fq, O, world, training)
        return out

    @staticmethod
    @_bwd_precision
    def backward(ctx, dout):
        sv = ctx.saved_tensors
        spec, mk, z, aff5, Wp_out = sv[:5]
        ys, affs, conv_w, gam = sv[5:9], sv[9:13], sv[13:17], sv[17:22]
        B, T, Fq, O, world, training = ctx.dims
        assert training, "backward through ConvStack requires train-mode BatchNorm statistics"
        dev = dout.device
        st = stream()
        dout = _f(dout)
        M = B * T
        nct = 4 * N_SM
        grads = [None] * 15

        def bn_bwd_consts(Y, G, mask, aff, gamma, npix, C, partial=None):
            if partial is None:                   # (the data gradient of the layer above may have formed the sums in its epilogue)
                partial = torch.empty(nct, 2 * C, device=dev, dtype=F32)
                lib.pa2s_colstats(st, 1, ptr(Y), ptr(G), ptr(mask), npix, C, ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]),
                                  ptr(partial), nct)
            sums = _bn_sums(partial, C, npix if world > 1 else None)
            local = None
            if world > 1:
                # SyncBatchNorm: the input gradient uses the sums over ALL ranks, dgamma / dbeta stay this rank's sums (the gradient
                # all-reduce averages them afterwards, like torch.nn.SyncBatchNorm under DDP): keep them before the all-reduce
                local = sums[:2 * C].to(F32)
                dist.all_reduce(sums)
            dg = torch.zeros(C, device=dev, dtype=F32)
            db = torch.zeros(C, device=dev, dtype=F32)
            k = torch.empty(3, C, device=dev, dtype=F32)
            lib.pa2s_bn_bwd_finalize(st, ptr(sums), -1.0 if world > 1 else float(npix), C, ptr(gamma), ptr(aff[3]), ptr(dg), ptr(db),
                                     ptr(k[0]), ptr(k[1]), ptr(k[2]))
            if local is not None:
                dg, db = local[C:].contiguous(), local[:C].contiguous()
                if os.environ.get("PA2S_CHECK_LOCAL_BN") == "1":
                    # debug: a second pass over the tensors must give the same numbers (up to the order of colstats' shared-memory atomics)
                    dg2, db2 = _local_bn_param_grads(Y, G, mask, aff, npix, C, nct, dev)
                    err = max(float((dg - dg2).abs().max() / dg2.abs().max().clamp_min(1e-30)),
                              float((db - db2).abs().max() / db2.abs().max().clamp_min(1e-30)))
                    print(f"PA2S_CHECK_LOCAL_BN C={C}: max relative difference {err:.2e}", file=sys.stderr)
                    if not err < 1e-5:
                        raise RuntimeError("SyncBatchNorm backward: local dgamma/dbeta differ from the recomputed sums")
            return dg, db, k

        # out_bn + relu + dropout backward
        dg5, db5, k5 = bn_bwd_consts(z, dout, mk, aff5, gam[4], M, O)
        grads[13], grads[14] = dg5, db5
        dz = torch.empty(M, O, device=dev, dtype=F32)
        lib.pa2s_bn_bwd_apply(st, ptr(dout), ptr(z), ptr(mk), M, O, ptr(aff5[0]), ptr(aff5[1]), ptr(aff5[2]), ptr(aff5[3]),
                              ptr(k5[0]), ptr(k5[1]), ptr(k5[2]), ptr(dz))
        C4 = conv_w[3].shape[0]
        Kf = Fq * C4
        # dW_out[n,k] = sum_m dz[m,n] * relu(bn4(y4))[m,k]
        dWp = torch.zeros(O, Kf, device=dev, dtype=F32)
        G = torch.empty(B, T, Fq, C4, device=dev, dtype=F32)           # G4 = dL/d relu(bn4(y4))
        if ctx.lin_ops is None:
            with ktime("out_linear_wgrad"):
                gemm(dz, ys[3], dWp, O, Kf, M, transA=True, lda=O, ldb=Kf, ldc=Kf,
                     t_scale=affs[3][0], t_shift=affs[3][1], t_period=C4, t_relu=True, t_on_b=True, zeroed=True)
            with ktime("out_linear_dgrad"):
                gemm(dz, Wp_out, G, M, Kf, O, lda=O, ldb=Kf, ldc=Kf)
        else:
            a4op, Wop = ctx.lin_ops
            ctx.lin_ops = None
            dzop = split_operand(dz, M, O, O, npieces=a4op.npieces)
            with ktime("out_linear_wgrad"):
                gemm(dzop, a4op, dWp, O, Kf, M, transA=True, ldc=Kf, zeroed=True)
            del a4op
            with ktime("out_linear_dgrad"):
                gemm(dzop, Wop, G, M, Kf, O, ldc=Kf)
        grads[12] = dWp.view(O, Fq, C4).permute(0, 2, 1).reshape(O, Kf).contiguous()
        nw = 8 * N_SM
        stats_below = None
        for i in (3, 2, 1, 0):
            W = conv_w[i]
            Cout, Cin = W.shape[0], W.shape[1]
            y, aff = ys[i], affs[i]
            npix = B * T * Fq
            dg, db, k = bn_bwd_consts(y, G, None, aff, gam[i], npix, Cout, partial=stats_below)
            stats_below = None
            grads[3 * i + 1], grads[3 * i + 2] = dg, db
            xin = ys[i - 1] if i > 0 else spec
            isc = affs[i - 1][0] if i > 0 else None
            ish = affs[i - 1][1] if i > 0 else None
            tc = ctx.prec != "fp32" and Cin >= 16
            if tc:
                # dy_i = BatchNorm/ReLU backward of (G_i, y_i) as bf16 planes: read by the weight gradient and the data gradient
                npc = min(npieces_for(ctx.prec), 2)
                Pdy = torch.empty(lib.pa2s_planes_bytes(B, T, Fq, Cout, npc), device=dev, dtype=torch.uint8)
                with ktime(f"conv{i + 1}_dy_planes"):
                    lib.pa2s_planes_bwd(st, B, T, Fq, Cout, ptr(G), ptr(y), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]),
                                        ptr(k[0]), ptr(k[1]), ptr(k[2]), ptr(Pdy), npc)
                nwp = lib.pa2s_conv_tma_wgrad_num_partials(B, T, Fq)
                partial = torch.empty(nwp, Cout * Cin * 9, device=dev, dtype=F32)
                with ktime(f"conv{i + 1}_wgrad"):
                    lib.pa2s_conv_tma_wgrad(st, B, T, Fq, Cin, Cout, ptr(ctx.planes[i]), ptr(Pdy), npc, ptr(partial))
                ctx.planes[i] = None
            elif Cin == 1 and Cout == 20:
                nwp = 2 * N_SM
                partial = torch.empty(nwp, Cout * Cin * 9, device=dev, dtype=F32)
                with ktime(f"conv{i + 1}_wgrad"):
                    lib.pa2s_conv1_wgrad(st, B, T, Fq, ptr(xin), ptr(G), ptr(y), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]),
                                         ptr(k[0]), ptr(k[1]), ptr(k[2]), ptr(partial), nwp)
            else:
                nwp = nw
                partial = torch.empty(nw, Cout * Cin * 9, device=dev, dtype=F32)
                with ktime(f"conv{i + 1}_wgrad"):
                    lib.pa2s_conv3x3_wgrad(st, B, T, Fq, Cin, Cout, ptr(xin), ptr(G), ptr(partial), nw, ptr(isc), ptr(ish), 1,
                                           ptr(y), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), ptr(k[0]), ptr(k[1]), ptr(k[2]))
            dW = torch.empty(Cout, Cin, 3, 3, device=dev, dtype=F32)
            lib.pa2s_reduce_rows(st, ptr(partial), nwp, Cout * Cin * 9, None, ptr(dW), 0)
            grads[3 * i] = dW
            if i > 0:
                Gp = torch.empty(B, T, Fq, Cin, device=dev, dtype=F32)
                if tc:
                    W2 = _tc_pack(W, Cout, Cin, 1)
                    # the epilogue also forms the sums of the BatchNorm/ReLU backward of the layer below (sum g, sum g*xhat)
                    affb = affs[i - 1]
                    with ktime(f"conv{i + 1}_dgrad"):
                        if lib.pa2s_conv_tma_get_impl():
                            stats_below = torch.empty(lib.pa2s_conv_tma_num_partials(B, T, Fq), 2 * Cin, device=dev, dtype=F32)
                            lib.pa2s_conv_tma_dgrad_stats(st, B, T, Fq, Cout, Cin, ptr(Pdy), npc, ptr(W2), ptr(Gp), ptr(ys[i - 1]),
                                                          ptr(affb[0]), ptr(affb[1]), ptr(affb[2]), ptr(affb[3]), ptr(stats_below))
                        else:                 # the comparator kernel has no statistics epilogue: separate colstats pass
                            lib.pa2s_conv_tma(st, B, T, Fq, Cout, Cin, ptr(Pdy), npc, ptr(W2), ptr(Gp), None)
                    del Pdy
                else:
                    W2 = W.detach().flip(2, 3).permute(2, 3, 0, 1).contiguous()        # [tap][co][ci]
                    with ktime(f"conv{i + 1}_dgrad"):
                        lib.pa2s_conv3x3(st, 1, B, T, Fq, Cout, Cin, ptr(G), ptr(W2), ptr(Gp), None, 4, None, None, 1,
                                         ptr(y), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), ptr(k[0]), ptr(k[1]), ptr(k[2]))
                G = Gp
        return (None, None, None, None, None, None, None, *grads)


def _local_bn_param_grads(Y, G, mask, aff, npix, C, nct, dev):
    """SyncBatchNorm: dgamma/dbeta are LOCAL sums (DDP averages them later), while the input gradient uses the global sums."""
    partial = torch.empty(nct, 2 * C, device=dev, dtype=F32)
    lib.pa2s_colstats(stream(), 1, ptr(Y), ptr(G), ptr(mask), npix, C, ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]),
                      ptr(partial), nct)
    s = _bn_sums(partial, C).float()
    return s[C:].contiguous(), s[:C].contiguous()


_GRU_EXCHANGE_SET = [False]


def _gru_exchange():
    """PA2S_GRU_EXCHANGE=barrier selects round 1's DSMEM stores + cluster barrier for the encoder recurrences (default: st.async + mbarrier)"""
    if not _GRU_EXCHANGE_SET[0]:
        lib.pa2s_gru_seq_set_exchange({"barrier": 0, "async": 1, "async_cols": 3}[os.environ.get("PA2S_GRU_EXCHANGE", "async")])
        _GRU_EXCHANGE_SET[0] = True


# ----------------------------------------------------------------------------------------------------------------
# Encoder BiGRU layer (models.py:63-67,77)
# ----------------------------------------------------------------------------------------------------------------
class BiGRULayerFn(torch.autograd.Function):
    """One bidirectional GRU layer over (B,T,I): returns out (B,T,2H) and h_n (2,B,H)."""

    @staticmethod
    def forward(ctx, x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_b, w_hh_b, b_ih_b, b_hh_b, bg):
        ctx.prec = current_precision()
        x = _f(x)
        B, T, I = x.shape
        H = w_hh_f.shape[1]
        dev = x.device
        Wih = torch.cat([w_ih_f, w_ih_b], 0).detach().contiguous()       # (6H, I)
        bih = torch.cat([b_ih_f, b_ih_b], 0).detach().contiguous()
        Whh = torch.stack([w_hh_f, w_hh_b], 0).detach().contiguous()     # (2,3H,H)
        bhh = torch.stack([b_hh_f, b_hh_b], 0).detach().contiguous()
        gi = torch.empty(B * T, 6 * H, device=dev, dtype=F32)
        gemm(x, Wih, gi, B * T, 6 * H, I, transB=True, lda=I, ldb=I, ldc=6 * H, bias=bih)
        out = torch.empty(B, T, 2 * H, device=dev, dtype=F32)
        need = any(ctx.needs_input_grad)
        gates = torch.empty(B, T, 2, 4 * H, device=dev, dtype=F32) if need else None
        hN = torch.empty(2, B, H, device=dev, dtype=F32)
        _gru_exchange()
        with ktime("encoder_gru_fwd"):
            lib.pa2s_gru_seq_fwd(stream(), B, T, 2, H, bg, ptr(gi), ptr(Whh), ptr(bhh), ptr(out), ptr(gates), ptr(hN))
        if need:
            ctx.save_for_backward(x, Wih, Whh, out, gates)
        ctx.bg = bg
        return out, hN

    @staticmethod
    @_bwd_precision
    def backward(ctx, dout, dhN):
        x, Wih, Whh, out, gates = ctx.saved_tensors
        B, T, I = x.shape
        H = Whh.shape[2]
        dev = x.device
        dout = _f(dout) if dout is not None else torch.zeros_like(out)
        dhN = _f(dhN) if dhN is not None else None
        dgi = torch.empty(B, T, 6 * H, device=dev, dtype=F32)
        dgh = torch.empty(B, T, 6 * H, device=dev, dtype=F32)
        with ktime("encoder_gru_bwd"):
            lib.pa2s_gru_seq_bwd(stream(), B, T, 2, H, ctx.bg, ptr(Whh), ptr(out), ptr(gates), ptr(dout), ptr(dhN), ptr(dgi), ptr(dgh))
        M = B * T
        dx = torch.empty(B, T, I, device=dev, dtype=F32)
        gemm(dgi, Wih, dx, M, I, 6 * H, lda=6 * H, ldb=I, ldc=I)
        dWih = torch.zeros(6 * H, I, device=dev, dtype=F32)
        gemm(dgi, x, dWih, 6 * H, I, M, transA=True, lda=6 * H, ldb=I, ldc=I, zeroed=True)
        dbih = colsum(dgi.view(M, 6 * H))
        dbhh = colsum(dgh.view(M, 6 * H))
        # dW_hh[dir] = sum_{b,t} dgh[b,t,dir]^T h_prev[b,t,dir];  h_prev = out at the previous step of that direction
        dWhh = torch.zeros(2, 3 * H, H, device=dev, dtype=F32)
        if T > 1:
            gemm(dgh, out, dWhh, 3 * H, H, T - 1, transA=True, lda=6 * H, ldb=2 * H, ldc=H, atomic=True, batch=B,
                 strideA=T * 6 * H, strideB=T * 2 * H, strideC=0, a_off=6 * H, b_off=0, c_off=0)
            gemm(dgh, out, dWhh, 3 * H, H, T - 1, transA=True, lda=6 * H, ldb=2 * H, ldc=H, atomic=True, batch=B,
                 strideA=T * 6 * H, strideB=T * 2 * H, strideC=0, a_off=3 * H, b_off=2 * H + H, c_off=3 * H * H)
        return (dx, dWih[:3 * H], dWhh[0], dbih[:3 * H], dbhh[:3 * H], dWih[3 * H:], dWhh[1], dbih[3 * H:], dbhh[3 * H:], None)


# ----------------------------------------------------------------------------------------------------------------
# Staff summariser (models.py:164-189)
# ----------------------------------------------------------------------------------------------------------------
class StaffGRUFn(torch.autograd.Function):
    """tokens (B,L) int64, lengths (B) int64 -> (B, 2*S) = [h_n forward | h_n reverse] of the packed BiGRU."""

    @staticmethod
    def forward(ctx, tokens, lengths, emb, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_b, w_hh_b, b_ih_b, b_hh_b):
        tokens = tokens.contiguous()
        lengths = lengths.to(device=tokens.device, dtype=torch.int64).contiguous()
        B, L = tokens.shape
        dev = tokens.device
        H, I = w_hh_f.shape[1], w_ih_f.shape[1]
        wih = torch.stack([w_ih_f, w_ih_b]).detach().contiguous()
        whh = torch.stack([w_hh_f, w_hh_b]).detach().contiguous()
        bih = torch.stack([b_ih_f, b_ih_b]).detach().contiguous()
        bhh = torch.stack([b_hh_f, b_hh_b]).detach().contiguous()
        need = any(ctx.needs_input_grad)
        hN = torch.empty(B, 2 * H, device=dev, dtype=F32)
        hs = torch.empty(B, 2, L, H, device=dev, dtype=F32) if need else None
        gates = torch.empty(B, 2, L, 4 * H, device=dev, dtype=F32) if need else None
        lib.pa2s_staff_gru_fwd(stream(), B, L, I, H, ptr(tokens), ptr(lengths), ptr(emb), ptr(wih), ptr(whh), ptr(bih), ptr(bhh),
                               ptr(hN), ptr(hs), ptr(gates))
        if need:
            ctx.save_for_backward(tokens, lengths, emb, wih, whh, hs, gates)
        return hN

    @staticmethod
    def backward(ctx, dhN):
        tokens, lengths, emb, wih, whh, hs, gates = ctx.saved_tensors
        B, L = tokens.shape
        H, I = whh.shape[2], wih.shape[2]
        dev = tokens.device
        d_emb = torch.zeros_like(emb)
        dwih = torch.zeros_like(wih)
        dwhh = torch.zeros_like(whh)
        dbih = torch.zeros(2, 3 * H, device=dev, dtype=F32)
        dbhh = torch.zeros(2, 3 * H, device=dev, dtype=F32)
        lib.pa2s_staff_gru_bwd(stream(), B, L, I, H, ptr(tokens), ptr(lengths), ptr(emb), ptr(wih), ptr(whh), ptr(hs), ptr(gates),
                               ptr(_f(dhN)), ptr(d_emb), ptr(dwih), ptr(dwhh), ptr(dbih), ptr(dbhh))
        return (None, None, d_emb, dwih[0], dwhh[0], dbih[0], dbhh[0], dwih[1], dwhh[1], dbih[1], dbhh[1])


# ----------------------------------------------------------------------------------------------------------------
# GRU cell gates (bar-level GRU)
# ----------------------------------------------------------------------------------------------------------------
class GRUGatesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, gh, hprev):
        gi, gh, hprev = _f(gi), _f(gh), _f(hprev)
        B, H = hprev.shape
        hnew = torch.empty_like(hprev)
        save = torch.empty(B, 4 * H, device=gi.device, dtype=F32)
        lib.pa2s_gru_gates_fwd(stream(), B, H, ptr(gi), ptr(gh), ptr(hprev), ptr(hnew), ptr(save))
        ctx.save_for_backward(save, hprev)
        return hnew

    @staticmethod
    def backward(ctx, dh):
        save, hprev = ctx.saved_tensors
        B, H = hprev.shape
        dgi = torch.empty(B, 3 * H, device=dh.device, dtype=F32)
        dgh = torch.empty(B, 3 * H, device=dh.device, dtype=F32)
        dhp = torch.empty_like(hprev)
        lib.pa2s_gru_gates_bwd(stream(), B, H, ptr(_f(dh)), ptr(save), ptr(hprev), ptr(dgi), ptr(dgh), ptr(dhp))
        return dgi, dgh, dhp


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    return GRUGatesFn.apply(linear(x, w_ih, b_ih), linear(h, w_hh, b_hh), h)


# ----------------------------------------------------------------------------------------------------------------
# Attention (models.py:440-461) -- one step, used by the bar-level decoder
# ----------------------------------------------------------------------------------------------------------------
def attn_split(B, T, ctas=N_SM):
    """frames of a clip are split over `ns` CTAs (ns * tile >= T) so that B * ns items fill `ctas` CTAs"""
    ns = max(1, min(16, ctas // max(B, 1)))
    tile = (T + ns - 1) // ns
    ns = (T + tile - 1) // tile
    return ns, tile


class AttnStepFn(torch.autograd.Function):
    """context = softmax_T(v . tanh(q + Ep)) @ enc.   q (B,A) = W_h h;  Ep (B,T,A) = enc W_e^T + b."""

    @staticmethod
    def forward(ctx, q, Ep, enc, v):
        ctx.prec = current_precision()
        q, Ep, enc = _f(q), _f(Ep), _f(enc)
        vv = _f(v).reshape(-1)
        B, T, D = enc.shape
        A = Ep.shape[2]
        assert D == 512 and A == 256, "attention kernels are specialised for hidden_size=256"
        dev = q.device
        NS, tile = attn_split(B, T)
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        e = lambda *s, dt=F32: torch.empty(*s, device=dev, dtype=dt)
        bufs = dict(attn=e(1, B, T), ctxs=e(1, B, D), xbuf=e(B, 16 + D), hc=e(B, 2 * D), pm=e(B, NS), pl=e(B, NS), pc=e(B, NS, D),
                    tickets=z(B, dt=torch.int32), counters=z(2, dt=torch.int32))
        qs = q.reshape(1, B, A).contiguous()
        args = make_dec_args(B=B, T=T, V=0, VP=0, S=1, max_steps=1, NS=NS, tile=tile, inference=0, save=1, enc=enc, Ep=Ep, v=vv, qs=qs,
                             **bufs)
        lib.pa2s_attn_step_fwd(stream(), ctypes.byref(args))
        ctx.save_for_backward(qs, Ep, enc, vv, bufs["attn"], bufs["ctxs"])
        ctx.split = (NS, tile)
        ctx.vshape = v.shape
        return bufs["ctxs"].reshape(B, D)

    @staticmethod
    @_bwd_precision
    def backward(ctx, dctx):
        qs, Ep, enc, vv, attn, ctxs = ctx.saved_tensors
        B, T, D = enc.shape
        A = Ep.shape[2]
        dev = enc.device
        NS, tile = ctx.split
        d_hc = torch.zeros(B, 2 * D, device=dev, dtype=F32)
        d_hc[:, D:] = dctx
        dx = torch.zeros(B, 16 + D, device=dev, dtype=F32)
        dEp = torch.zeros(B, T, A, device=dev, dtype=F32)
        dq_part = torch.empty(B, NS, A, device=dev, dtype=F32)
        dv_part = torch.zeros(B * NS, A, device=dev, dtype=F32)
        dctx_all = torch.empty(1, B, D, device=dev, dtype=F32)
        args = make_dec_args(B=B, T=T, V=0, VP=0, S=1, max_steps=1, NS=NS, tile=tile, inference=0, save=1, enc=enc, Ep=Ep, v=vv, qs=qs,
                             attn=attn, ctxs=ctxs, d_hc=d_hc, dx=dx, dEp=dEp, dq_part=dq_part, dv_part=dv_part, dctx_all=dctx_all)
        lib.pa2s_attn_step_bwd(stream(), ctypes.byref(args))
        dq = dq_part.sum(1)
        dv = colsum(dv_part).reshape(ctx.vshape)
        denc = context_grad_enc(attn, dctx_all, B, T, D, 1)
        return dq, dEp, denc, dv


BAR_CHAIN = os.environ.get("PA2S_BAR_CHAIN", "1") != "0"          # 0: the bar-level chain as individual autograd ops (round-2 start)


class BarChain:
    """The bar-level chain of one forward pass (models.py:239-247 per bar): attention query, attention step, GRU cell -- one autograd
    node per bar (BarStepFn) instead of ~15 (Linear x3, cat, attention, gates and the gradient-accumulation adds between them).
    The gradients that every bar adds to the SAME tensors (d Ep_bar, d enc, d v) are accumulated in place across the bars'
    backward calls and handed to autograd once, by the node that runs last (the first bar's); the weight gradients of the three Linear
    maps go through their DeferredLinear sinks (one contraction per map at the end of the backward pass)."""

    def __init__(self, lins, v):
        self.lin_q, self.lin_ih, self.lin_hh = lins
        self.v = v
        self.count = 0
        self.dEp = self.dv_part = self.zero_dx = None
        self.attn_rows, self.dctx_rows = [], []

    def step(self, token, h, enc, Ep):
        lq, li, lh = self.lin_q, self.lin_ih, self.lin_hh
        self.count += 1
        return BarStepFn.apply(token, h, enc, Ep, self.v, lq.W, li.W, li.b, lh.W, lh.b, self, self.count - 1)


class BarStepFn(torch.autograd.Function):
    """(token, h) -> (h', context) of one bar: q = W_q h;  context = attention(q, Ep, enc);  h' = GRUCell([token | context], h)."""

    @staticmethod
    def forward(ctx, token, h, enc, Ep, v, Wq, Wih, bih, Whh, bhh, chain, index):
        ctx.prec = current_precision()
        token, h, enc, Ep = _f(token), _f(h), _f(enc), _f(Ep)
        vv = _f(v).reshape(-1)
        B, T, D = enc.shape
        A = Ep.shape[2]
        H = h.shape[1]
        K = token.shape[1]
        assert D == 512 and A == 256, "attention kernels are specialised for hidden_size=256"
        dev = h.device
        e = lambda *s_, dt=F32: torch.empty(*s_, device=dev, dtype=dt)
        qs = e(1, B, A)
        gemm(h, Wq, qs, B, A, H, transB=True, lda=H, ldb=Wq.stride(0), ldc=A)
        NS, tile = attn_split(B, T)
        bufs = dict(attn=e(1, B, T), ctxs=e(1, B, D), xbuf=e(B, 16 + D), hc=e(B, 2 * D), pm=e(B, NS), pl=e(B, NS), pc=e(B, NS, D),
                    tickets=torch.zeros(B, device=dev, dtype=torch.int32), counters=torch.zeros(2, device=dev, dtype=torch.int32))
        args = make_dec_args(B=B, T=T, V=0, VP=0, S=1, max_steps=1, NS=NS, tile=tile, inference=0, save=1, enc=enc, Ep=Ep, v=vv, qs=qs,
                             **bufs)
        lib.pa2s_attn_step_fwd(stream(), ctypes.byref(args))
        context = bufs["ctxs"].reshape(B, D)
        x = torch.cat([token, context], dim=1)
        gi, gh = e(B, 3 * H), e(B, 3 * H)
        gemm(x, Wih, gi, B, 3 * H, K + D, transB=True, lda=K + D, ldb=Wih.stride(0), ldc=3 * H, bias=bih)
        gemm(h, Whh, gh, B, 3 * H, H, transB=True, lda=H, ldb=Whh.stride(0), ldc=3 * H, bias=bhh)
        hnew = e(B, H)
        save = e(B, 4 * H)
        lib.pa2s_gru_gates_fwd(stream(), B, H, ptr(gi), ptr(gh), ptr(h), ptr(hnew), ptr(save))
        ctx.save_for_backward(x, h, qs, Ep, enc, vv, bufs["attn"], bufs["ctxs"], save, Wq, Wih, Whh)
        ctx.chain, ctx.split, ctx.vshape, ctx.K, ctx.index = chain, (NS, tile), v.shape, K, index
        return hnew, context

    @staticmethod
    @_bwd_precision
    def backward(ctx, dh, dctx):
        x, h, qs, Ep, enc, vv, attn, ctxs, save, Wq, Wih, Whh = ctx.saved_tensors
        chain, (NS, tile), K = ctx.chain, ctx.split, ctx.K
        B, T, D = enc.shape
        A = Ep.shape[2]
        H = h.shape[1]
        dev = enc.device
        st = stream()
        e = lambda *s_: torch.empty(*s_, device=dev, dtype=F32)
        dgi, dgh, dhp = e(B, 3 * H), e(B, 3 * H), e(B, H)
        lib.pa2s_gru_gates_bwd(st, B, H, ptr(_f(dh)), ptr(save), ptr(h), ptr(dgi), ptr(dgh), ptr(dhp))
        dx = e(B, K + D)
        gemm(dgi, Wih, dx, B, K + D, 3 * H, lda=3 * H, ldb=Wih.stride(0), ldc=K + D)
        gemm(dgh, Whh, dhp, B, H, 3 * H, lda=3 * H, ldb=Whh.stride(0), ldc=H, accumulate=True)
        # attention step: d context = (consumers of the context outside the cell) + (GRU input part)
        d_hc = e(B, 2 * D)
        torch.add(_f(dctx), dx[:, K:], out=d_hc[:, D:])
        if chain.dEp is None:
            chain.dEp = torch.zeros(B, T, A, device=dev, dtype=F32)
            chain.dv_part = torch.zeros(B * NS, A, device=dev, dtype=F32)
            chain.zero_dx = torch.zeros(B, 16 + D, device=dev, dtype=F32)
        dq_part = e(B, NS, A)
        dctx_all = e(1, B, D)
        args = make_dec_args(B=B, T=T, V=0, VP=0, S=1, max_steps=1, NS=NS, tile=tile, inference=0, save=1, enc=enc, Ep=Ep, v=vv, qs=qs,
                             attn=attn, ctxs=ctxs, d_hc=d_hc, dx=chain.zero_dx, dEp=chain.dEp, dq_part=dq_part, dv_part=chain.dv_part,
                             dctx_all=dctx_all)
        lib.pa2s_attn_step_bwd(st, ctypes.byref(args))
        dq = dq_part.sum(1)
        gemm(dq, Wq, dhp, B, H, A, lda=A, ldb=Wq.stride(0), ldc=H, accumulate=True)
        # d enc = sum over the bars of attn^T dctx: ONE batched GEMM with K = bars, by the node that runs last
        chain.attn_rows.append(attn)
        chain.dctx_rows.append(dctx_all)
        # weight gradients: rows for the three sinks
        chain.lin_q.sink.rows.append((h, dq))
        chain.lin_ih.sink.rows.append((x, dgi))
        chain.lin_hh.sink.rows.append((h, dgh))
        # every later bar's state derives from this bar's (h is chained), so of the nodes that run at all the one of the FIRST bar
        # runs last: it hands the accumulated gradients to autograd (also when the loss does not reach some later bars)
        dEp = denc = dv = None
        if ctx.index == 0:
            dEp = chain.dEp
            S = len(chain.attn_rows)
            denc = context_grad_enc(torch.cat(chain.attn_rows), torch.cat(chain.dctx_rows), B, T, D, S)
            dv = colsum(chain.dv_part).reshape(ctx.vshape)
            chain.dEp = chain.dv_part = None
            chain.attn_rows, chain.dctx_rows = [], []
        return dx[:, :K], dhp, denc, dEp, dv, None, None, None, None, None, None, None


def context_grad_enc(attn, dctx_all, B, T, D, S):
    """d_enc[b] = attn[:, b, :]^T (T x S) @ dctx_all[:, b, :] (S x D): the context read-out's gradient, one GEMM per clip."""
    denc = torch.empty(B, T, D, device=attn.device, dtype=F32)
    gemm(attn, dctx_all, denc, T, D, S, transA=True, lda=B * T, ldb=B * D, ldc=D, batch=B, strideA=T, strideB=D, strideC=T * D)
    return denc


# ----------------------------------------------------------------------------------------------------------------
# Note decoder (models.py:366-420) -- all steps of one (bar, staff) in one call
# ----------------------------------------------------------------------------------------------------------------
# "multi" (default): HierarchicalDecoder batches every run of teacher-forced bars of a staff into one cooperative launch and runs
# the reverse pass of all bars of a staff in one launch (dec_multi.cu, StaffRun / DecodersFn below);
# "persistent": one cooperative kernel per (bar, staff) call (dec_persist.cu); "steps": one launch per phase per step (decoder.cu).
# The stand-alone NoteDecoder module always uses the per-call kernels ("multi" -> "persistent" there).
DECODER_IMPL = os.environ.get("PA2S_DECODER", "multi")
PROF = {}              # optional {"fwd": uint64[8] tensor, "bwd": ...}: per-phase ns of CTA 0 (tools/prof_decoder.py)
SYNC_FLAGS = collections.deque(maxlen=256)        # [arrivals, watchdog flag] of recent persistent launches (tests / bench check flag == 0)


def check_sync_flags():
    """Raises if a grid-barrier watchdog fired in any persistent decoder launch since the last call (host sync)."""
    flags = list(SYNC_FLAGS)
    SYNC_FLAGS.clear()
    if flags and int(torch.stack([f[1] for f in flags]).max().item()) != 0:
        raise RuntimeError("persistent decoder: grid barrier watchdog fired (a CTA of the cooperative grid was lost)")


class NoteDecoderFn(torch.autograd.Function):
    """Returns (logp (B,max_steps,V), lengths (B) int64 device, counters (2) int32 device = [eos_count, steps])."""

    @staticmethod
    def forward(ctx, enc, Ep, h0, attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out, cfg):
        # cfg["stream"]: the call is ENQUEUED on that side stream (which the caller has made wait for the inputs) while
        # autograd sees a node of the current stream; the caller waits for the side stream before it reads the outputs.
        side = cfg.get("stream")
        ctx.side, ctx.pre = side, None
        ctx.defer_stream, ctx.defer_done = cfg.get("defer_stream"), None     # stream of the parallel dEp / dv kernel of the backward
        with _on_stream(side):
            return NoteDecoderFn._forward(ctx, enc, Ep, h0, attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out, cfg)

    @staticmethod
    def _forward(ctx, enc, Ep, h0, attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out, cfg):
        ctx.prec = current_precision()
        enc, Ep, h0 = _f(enc), _f(Ep), _f(h0)
        B, T, D = enc.shape
        A = Ep.shape[2]
        V, E = emb.shape
        assert D == 512 and A == 256 and E == 16 and V <= 256, "decoder kernels are specialised for hidden_size=256, note_emb_size=16"
        dev = enc.device
        S, max_steps = int(cfg["S"]), int(cfg["max_steps"])
        inference = bool(cfg["inference"])
        gt, use_gt, mask = cfg.get("gt"), cfg.get("use_gt"), cfg.get("mask")
        save = any(ctx.needs_input_grad)
        VP = (V + 3) // 4 * 4
        persist = DECODER_IMPL != "steps"
        NS, tile = attn_split(B, T, lib.pa2s_dec_persist_grid() if persist else N_SM)
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        e = lambda *s, dt=F32: torch.empty(*s, device=dev, dtype=dt)
        SS = S if save else 1
        logp = z(B, max_steps, V)
        lengths = torch.full((B,), max_steps, device=dev, dtype=torch.int64)
        eos = z(B, dt=torch.int32)
        counters = z(2, dt=torch.int32)
        hs = e(SS + 1, B, D)
        hs[0] = h0
        sv = dict(hs=hs, ctxs=e(SS, B, D), attn=e(SS, B, T), gates=e(SS, B, 4 * D) if save else None, qs=e(SS + 1, B, A),
                  xtok=e(SS + 1, B, E) if save else None, toks=e(SS + 1, B, dt=torch.int32) if save else None)
        scratch = dict(xbuf=e(B, E + D), hc=e(B, 2 * D), logits=z(B, VP), pm=e(B, NS), pl=e(B, NS), pc=e(B, NS, D),
                       tickets=z(B, dt=torch.int32), sync=z(2, dt=torch.int32))
        if gt is not None:
            gt = gt.contiguous()
            assert gt.shape == (B, max_steps) and gt.dtype == torch.int64
        wts = dict(Wattn=attn_w, v=_f(attn_v).reshape(-1), emb=emb, W_ih=W_ih, W_hh=W_hh, b_ih=b_ih, b_hh=b_hh, W_out=W_out, b_out=b_out)
        for k, w in wts.items():
            assert w.is_contiguous() and w.dtype == F32 and w.data_ptr() % 16 == 0, f"{k}: decoder weights must be contiguous, 16-byte aligned fp32"

        args = make_dec_args(B=B, T=T, V=V, VP=VP, S=S, max_steps=max_steps, NS=NS, tile=tile, inference=int(inference), save=int(save),
                             enc=enc, Ep=Ep, gt=gt, use_gt=use_gt, mask=mask, logp=logp, lengths=lengths, eos=eos, counters=counters,
                             prof=PROF.get("fwd"), **wts, **sv, **scratch)
        with ktime("note_decoder_fwd"):
            if persist:
                lib.pa2s_note_decoder_fwd_persist(stream(), ctypes.byref(args), int(cfg["sos"]), int(cfg["eos"]))
            else:
                lib.pa2s_note_decoder_fwd(stream(), ctypes.byref(args), int(cfg["sos"]), int(cfg["eos"]))
        SYNC_FLAGS.append(scratch["sync"])
        ctx.mark_non_differentiable(lengths, counters)
        if save:
            ctx.save_for_backward(enc, Ep, attn_w, wts["v"], emb, W_ih, W_hh, W_out, logp, sv["hs"], sv["ctxs"], sv["attn"], sv["gates"],
                                  sv["qs"], sv["xtok"], sv["toks"], mask if mask is not None else torch.empty(0, device=dev))
            ctx.meta = (B, T, D, A, V, E, VP, S, max_steps, NS, tile, mask is not None, attn_v.shape)
            ctx.sink = cfg.get("sink")
            early = cfg.get("early")
            if early is not None:
                early[0].calls[early[1]] = ctx
        return logp, lengths, counters

    @staticmethod
    def backward(ctx, dlogp, _dl, _dc):
        pre, ctx.pre = ctx.pre, None
        if pre is None:
            pre = NoteDecoderFn.launch_backward(ctx, dlogp)
        grads, done = pre
        for ev in done:
            torch.cuda.current_stream().wait_event(ev)
        return grads

    @staticmethod
    def launch_backward(ctx, dlogp):
        """Enqueues the whole backward of this call on the stream its forward ran on -> (input gradients, [completion events]).
        Called by backward(), or ahead of it by StackLogpFn.backward as soon as the loss gradient exists."""
        side = ctx.side
        with use_precision(ctx.prec):
            if side is None:
                grads = NoteDecoderFn._backward(ctx, dlogp)
                return grads, [ev for ev in (ctx.defer_done,) if ev is not None]
            cur = torch.cuda.current_stream()
            side.wait_event(cur.record_event())
            dlogp.record_stream(side)
            with torch.cuda.stream(side):
                grads = NoteDecoderFn._backward(ctx, dlogp)
                done = [side.record_event()]
            if ctx.defer_done is not None:
                done.append(ctx.defer_done)
            for g in grads:
                if torch.is_tensor(g):
                    g.record_stream(cur)
            return grads, done

    @staticmethod
    def _backward(ctx, dlogp):
        (enc, Ep, attn_w, v, emb, W_ih, W_hh, W_out, logp, hs, ctxs, attn, gates, qs, xtok, toks, mask) = ctx.saved_tensors
        B, T, D, A, V, E, VP, S, max_steps, NS, tile, has_mask, vshape = ctx.meta
        dev = enc.device
        st = stream()
        dlogp = _f(dlogp)
        X = E + D
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)
        e = lambda *s, dt=F32: torch.empty(*s, device=dev, dtype=dt)
        persist = DECODER_IMPL != "steps"
        W_hT = attn_w.detach()[:, :D].t().contiguous()
        W_ihT = W_ih.detach().t().contiguous()
        W_hhT = W_hh.detach().t().contiguous()
        if persist:
            # off-chain work first: log-softmax backward of every step and its out-projection gradient as one GEMM
            nblk = lib.pa2s_dec_deferred_blocks(T)
            bw = dict(dlogits_all=e(S, B, VP), dgi_all=e(S, B, 3 * D), dgh_all=e(S, B, 3 * D), dq_all=e(S + 1, B, A), dctx_all=e(S, B, D),
                      dxtok_all=e(S, B, E), dEp=e(B, T, A), dv_part=e(B * nblk, A), d_hc=e(B, 2 * D), dhq=e(B, D), dx=e(B, X),
                      dq_part=e(B, NS, A), dh_carry=e(2, B, D), dhc_all=e(S * B, 2 * D), ds_all=e(S, B, T),
                      tickets=z(B, dt=torch.int32), sync=z(2, dt=torch.int32))
            args = make_dec_args(B=B, T=T, V=V, VP=VP, S=S, max_steps=max_steps, NS=NS, tile=tile, inference=0, save=1,
                                 enc=enc, Ep=Ep, Wattn=attn_w, v=v, emb=emb, W_ih=W_ih, W_hh=W_hh, W_out=W_out,
                                 W_hT=W_hT, W_ihT=W_ihT, W_hhT=W_hhT, logp=logp, hs=hs, ctxs=ctxs, attn=attn, gates=gates, qs=qs,
                                 dlogp=dlogp, prof=PROF.get("bwd"), **bw)
            with ktime("note_decoder_bwd"):
                lib.pa2s_dec_dlogits(st, ctypes.byref(args))
                gemm(bw["dlogits_all"], W_out, bw["dhc_all"], S * B, 2 * D, V, lda=VP, ldb=2 * D, ldc=2 * D)
                aux = ctx.defer_stream
                if aux is None:
                    lib.pa2s_note_decoder_bwd_persist(st, ctypes.byref(args))
                else:
                    # the sequential chain here; the parallel dEp / dv kernel (reads ds_all, qs, Ep, v) on `aux`, so that the next
                    # persistent kernel of this stream starts right behind the chain
                    lib.pa2s_note_decoder_bwd_chain(st, ctypes.byref(args))
                    cur = torch.cuda.current_stream()
                    aux.wait_event(cur.record_event())
                    with torch.cuda.stream(aux):
                        lib.pa2s_note_decoder_bwd_deferred(stream(), ctypes.byref(args))
                        ctx.defer_done = aux.record_event()
                    for t_ in (bw["ds_all"], bw["dEp"], bw["dv_part"], qs, Ep, v):
                        t_.record_stream(aux)
            SYNC_FLAGS.append(bw["sync"])
            dh0 = bw["dhq"]
        else:
            W_outT = z(2 * D, VP)
            W_outT[:, :V] = W_out.detach().t()
            bw = dict(dlogits_all=e(S, B, VP), dgi_all=e(S, B, 3 * D), dgh_all=e(S, B, 3 * D), dq_all=z(S + 1, B, A), dctx_all=e(S, B, D),
                      dxtok_all=e(S, B, E), dEp=z(B, T, A), dv_part=z(B * NS, A), d_hc=z(B, 2 * D), dhq=z(B, D), dx=z(B, X),
                      dq_part=z(B, NS, A), dh_carry=z(2, B, D))
            args = make_dec_args(B=B, T=T, V=V, VP=VP, S=S, max_steps=max_steps, NS=NS, tile=tile, inference=0, save=1,
                                 enc=enc, Ep=Ep, Wattn=attn_w, v=v, emb=emb, W_ih=W_ih, W_hh=W_hh, W_out=W_out,
                                 W_outT=W_outT, W_hT=W_hT, W_ihT=W_ihT, W_hhT=W_hhT, logp=logp, hs=hs, ctxs=ctxs, attn=attn, gates=gates, qs=qs,
                                 dlogp=dlogp, **bw)
            with ktime("note_decoder_bwd"):
                lib.pa2s_note_decoder_bwd(st, ctypes.byref(args))
            dh0 = bw["dh_carry"][0] + bw["dhq"]
        dxt = bw["dxtok_all"]
        if has_mask:
            dxt = dxt * mask[:S]
        rec = dict(S=S, B=B, dlogits=bw["dlogits_all"], hs=hs, ctxs=ctxs, dgi=bw["dgi_all"], dgh=bw["dgh_all"], dq=bw["dq_all"],
                   xtok=xtok, dv_part=bw["dv_part"], dxt=dxt, toks=toks)
        dims = (D, A, V, E, VP, vshape)
        sink = ctx.sink
        if sink is not None:
            # weight gradients of this module are formed once per backward pass, over the rows of ALL its calls (one per bar)
            rec["event"] = torch.cuda.current_stream().record_event()
            rec["stream"] = torch.cuda.current_stream()
            rec["event2"] = ctx.defer_done                  # dv_part comes from the deferred kernel's stream
            sink.records.append(rec)
            sink.dims = dims
            wg = (None,) * 9
        else:
            wg = decoder_weight_grads([rec], dims)
        denc = context_grad_enc(attn, bw["dctx_all"], B, T, D, S)
        return (denc, bw["dEp"], dh0) + tuple(wg) + (None,)


def decoder_weight_grads(recs, dims):
    """Deferred weight gradients of a NoteDecoder: contractions over the (step, clip) rows saved by one or several calls.
    -> (d_attn_w, dv, d_emb, dW_ih, dW_hh, db_ih, db_hh, dW_out, db_out), the parameter order of NoteDecoderFn."""
    D, A, V, E, VP, vshape = dims
    X = E + D
    dev = recs[0]["hs"].device
    z = lambda *s_: torch.zeros(*s_, device=dev, dtype=F32)

    def rows(key, lo=0):
        parts = [r[key][lo:lo + r["S"]].reshape(r["S"] * r["B"], -1) for r in recs]
        return parts[0] if len(parts) == 1 else torch.cat(parts)
    SB = sum(r["S"] * r["B"] for r in recs)
    dlogits, dgi, dgh, dq = rows("dlogits"), rows("dgi"), rows("dgh"), rows("dq")
    h_prev, h_new, ctxs, xtok = rows("hs"), rows("hs", 1), rows("ctxs"), rows("xtok")
    dW_out = z(V, 2 * D)
    gemm(dlogits, h_new, dW_out, V, D, SB, transA=True, lda=VP, ldb=D, ldc=2 * D, zeroed=True)                # h' part
    gemm(dlogits, ctxs, dW_out, V, D, SB, transA=True, lda=VP, ldb=D, ldc=2 * D, c_off=D, zeroed=True)        # ctx part
    db_out = colsum(dlogits)[:V].contiguous()
    dW_ih = z(3 * D, X)
    gemm(dgi, xtok, dW_ih, 3 * D, E, SB, transA=True, lda=3 * D, ldb=E, ldc=X, zeroed=True)
    gemm(dgi, ctxs, dW_ih, 3 * D, D, SB, transA=True, lda=3 * D, ldb=D, ldc=X, c_off=E, zeroed=True)
    db_ih = colsum(dgi)
    dW_hh = z(3 * D, D)
    gemm(dgh, h_prev, dW_hh, 3 * D, D, SB, transA=True, lda=3 * D, ldb=D, ldc=D, zeroed=True)
    db_hh = colsum(dgh)
    d_attn_w = z(A, 2 * D)
    gemm(dq, h_prev, d_attn_w, A, D, SB, transA=True, lda=A, ldb=D, ldc=2 * D, zeroed=True)                  # W_h half only
    dvp = [r["dv_part"] for r in recs]
    # (dv_part None: the caller forms dv itself once the kernel that produces the partial sums has finished on its own stream)
    dv = None if dvp[0] is None else colsum(dvp[0] if len(dvp) == 1 else torch.cat(dvp)).reshape(vshape)
    # embedding: scatter-add of the (masked) token-input gradients
    d_emb = z(V, E)
    toks = torch.cat([r["toks"][:r["S"]].reshape(-1) for r in recs]).long()
    d_emb.index_add_(0, toks, rows("dxt"))
    return d_attn_w, dv, d_emb, dW_ih, dW_hh, db_ih, db_hh, dW_out, db_out


class DecoderGradSink:
    """Per NoteDecoder module and forward pass: the step buffers its calls (one per bar) leave behind in backward, and --
    when StackLogpFn has already enqueued the contractions over them on a side stream -- their result `pre`."""
    def __init__(self):
        self.records, self.dims, self.pre = [], None, None

    def take(self, cur):
        """Weight gradients over all records, enqueued on stream `cur` (which is made to wait for the records' streams)."""
        recs, self.records = self.records, []
        if not recs:
            return None
        for r in recs:
            if r.get("event2") is not None:
                cur.wait_event(r["event2"])
            if r["stream"] != cur:
                cur.wait_event(r["event"])
            if r["stream"] != cur or r.get("event2") is not None:
                for v in r.values():
                    if torch.is_tensor(v):
                        v.record_stream(cur)
        return list(decoder_weight_grads(recs, self.dims))

    def launch(self, side):
        """take() on `side`, ahead of the sink node -> self.pre = (gradients, completion event)."""
        with torch.cuda.stream(side):
            grads = self.take(side)
            if grads is not None:
                self.pre = (grads, side.record_event())


class DecoderWeightSinkFn(torch.autograd.Function):
    """Identity on the nine NoteDecoder weights.  Every NoteDecoderFn call that uses the returned aliases hands its weight
    gradients here as saved rows instead of computing them: autograd runs this node after the last of those calls, and the
    gradients are then ONE set of contractions over all bars (K = sum of S*B) -- enqueued on a side stream by StackLogpFn right
    behind the last decoder backward kernel (DecoderGradSink.launch), or here on the current stream when nothing was enqueued
    ahead.  Besides the 5x fewer small GEMMs this keeps parameter-gradient accumulation off the two decoder streams, where
    autograd's AccumulateGrad (pinned to the stream of a parameter's first use) serialised the two staves of every other bar."""

    @staticmethod
    def forward(ctx, sink, *weights):
        ctx.prec = current_precision()
        ctx.sink = sink
        ctx.set_materialize_grads(False)
        return tuple(w.view_as(w) for w in weights)

    @staticmethod
    @_bwd_precision
    def backward(ctx, *grads):
        sink = ctx.sink
        cur = torch.cuda.current_stream()
        pre, sink.pre = sink.pre, None
        if pre is not None:
            out, done = pre
            cur.wait_event(done)
            for g in out:
                g.record_stream(cur)
            assert not sink.records
        else:
            out = sink.take(cur) or [None] * 9
        for i, g in enumerate(grads):            # uses of the aliases outside NoteDecoderFn (none on the hot path)
            if g is not None:
                out[i] = g if out[i] is None else out[i] + g
        return (None,) + tuple(out)



class DecoderEarlyBackward:
    """Per forward pass of the bar loop: the NoteDecoderFn calls, keyed (staff, bar), whose backward StackLogpFn launches."""
    def __init__(self):
        self.calls = {}
        self.sinks = []         # (DecoderGradSink, side stream): weight-gradient contractions enqueued right behind the last call


class StackLogpFn(torch.autograd.Function):
    """(torch.stack(upper per-bar log-probs, 1), torch.stack(lower ..., 1)) of models.py:313-316.  A note decoder's backward
    depends on nothing but the loss gradient of its own log-probabilities (its predictions reach the next bar through argmax
    only), so this node -- the first of the decoder to run in a backward pass -- enqueues ALL (bar, staff) backward kernels on
    their side streams at once, last bar first.  The two 64-CTA kernels that fit the GPU side by side then run back to back,
    instead of each bar's pair waiting for the bar chain of the bar after it (autograd's node order)."""

    @staticmethod
    def forward(ctx, early, n_upper, *logps):
        ctx.early, ctx.n = early, (n_upper, len(logps) - n_upper)
        ctx.prec = current_precision()
        ctx.set_materialize_grads(False)
        return torch.stack(logps[:n_upper], 1), torch.stack(logps[n_upper:], 1)

    @staticmethod
    def backward(ctx, d_up, d_lo):
        nu, nl = ctx.n
        parts = [[None if d is None else d[:, bar] for bar in range(n)] for d, n in ((d_up, nu), (d_lo, nl))]
        calls = ctx.early.calls
        for bar in reversed(range(max(nu, nl))):
            for si in (0, 1):
                c = calls.pop((si, bar), None)
                if c is not None and bar < len(parts[si]) and parts[si][bar] is not None:
                    c.pre = NoteDecoderFn.launch_backward(c, parts[si][bar])
        if not calls:               # every call has left its rows: the weight gradients follow on the side streams
            with use_precision(ctx.prec):
                for sink, side in ctx.early.sinks:
                    sink.launch(side)
        return (None, None) + tuple(parts[0]) + tuple(parts[1])


# ----------------------------------------------------------------------------------------------------------------
# Multi-sequence note decoder (dec_multi.cu): NQ bars of one staff x B clips per launch, all bars in one reverse launch
# ----------------------------------------------------------------------------------------------------------------
def decm_split(B, T, nq=None):
    """frames of a clip are split over NS items so that B * NS items fill the grid; an item holds at most pa2s_decm_tile_max() frames
    (x 5 // nq in forward-only runs, whose launches decode nq bars: the score table of the kernel is shared by the queries)"""
    tmax, pg = lib.pa2s_decm_tile_max(), lib.pa2s_decm_grid()
    if nq is not None:
        tmax *= lib.pa2s_decm_max_queries() // nq
    ns = max(-(-T // tmax), min(16, max(1, pg // max(B, 1))))
    tile = -(-T // ns)
    return -(-T // tile), tile


def zeros_multi(dev, **specs):
    """name -> zero-filled tensor for every name = (shape, dtype) in `specs`, carved out of ONE allocation zeroed by ONE memset (each
    tensor starts on a 256-byte boundary).  A `torch.zeros` per buffer was ~60 fill launches per training step."""
    offs, total = {}, 0
    for k, (shape, dt) in specs.items():
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dt).element_size()
        offs[k] = (total, nbytes)
        total += (nbytes + 255) // 256 * 256
    buf = torch.zeros(max(total, 1), device=dev, dtype=torch.uint8)
    return {k: buf[o:o + nb].view(specs[k][1]).view(*specs[k][0]) for k, (o, nb) in offs.items()}


_AUX_STREAMS = {}          # one auxiliary stream per staff stream (deferred attention gradients of the reverse pass)


class StaffRun:
    """All bars of ONE staff in one forward pass of HierarchicalDecoder.decode_bars.

    Holds the staff's output log-probabilities (B, bars, max_steps, V) and, when a backward pass will follow, the step-major
    saved state of every (step, bar, clip) row (row = bar * B + clip).  `launch(k0, nq, h0)` decodes bars k0 .. k0+nq-1 (whose
    bar summaries `h0` are known) in one cooperative kernel on `self.stream`; `backward(dlogp)` runs the reverse pass over ALL
    rows in one launch (per group of pa2s_decm_max_queries() bars) followed by the weight-gradient contractions."""

    def __init__(self, weights, enc, Ep, bars, max_steps, steps, inference, save, side, sos, eos, gt=None, tf_bits=None, mask=None):
        self.w = tuple(weights)
        attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out = self.w
        for k, w in zip(("attn_w", "attn_v", "emb", "W_ih", "W_hh", "b_ih", "b_hh", "W_out", "b_out"), self.w):
            assert w.is_contiguous() and w.dtype == F32 and w.data_ptr() % 16 == 0, f"{k}: decoder weights must be contiguous, 16-byte aligned fp32"
        enc, Ep = _f(enc.detach()), _f(Ep.detach())
        B, T, D = enc.shape
        A = Ep.shape[2]
        V, E = emb.shape
        assert D == 512 and A == 256 and E == 16 and V <= 256, "decoder kernels are specialised for hidden_size=256, note_emb_size=16"
        self.prec = current_precision()
        # weight-stationary products of the kernels: tensor cores (bf16 hi/lo split, ~5e-6) in the training precisions, exact FFMA in fp32
        self.tc = int(self.prec != "fp32" and os.environ.get("PA2S_DECM_TC", "1") == "1")
        self.enc, self.B, self.T, self.D, self.A, self.V, self.E = enc, B, T, D, A, V, E
        self.VP = (V + 3) // 4 * 4
        self.bars, self.max_steps, self.steps = bars, max_steps, [int(x) for x in steps]
        self.inference, self.save, self.stream, self.sos, self.eos = bool(inference), bool(save), side, int(sos), int(eos)
        self.Rtot = bars * B
        self.Smax = max(self.steps)
        # (greedy inference decodes bar by bar -- one query per launch -- and has no reverse pass: longer frame ranges, fewer items)
        self.NS, self.tile = decm_split(B, T, 1 if (self.inference and not self.save) else None)
        dev = enc.device
        z = lambda *s_, dt=F32: torch.zeros(*s_, device=dev, dtype=dt)
        self.Ee = torch.empty_like(Ep)
        lib.pa2s_exp2x(stream(), ptr(Ep), ptr(self.Ee), Ep.numel())
        self.logp = None                                             # (allocated with the saved-state buffers below)
        self.lengths = torch.full((bars, B), max_steps, device=dev, dtype=torch.int64)
        self.gt = gt.contiguous() if gt is not None else None
        if self.gt is not None:
            assert self.gt.shape == (B, bars, max_steps) and self.gt.dtype == torch.int64
        # tf_bits: per bar, a python int whose bit s says 'step s takes its next token from the targets' (the pre-drawn coins of
        # models.py:404; they travel in the kernel's argument block: no device tensor, no H2D copy); mask: (Smax, Rtot, E) fp32 or None
        self.tf_bits, self.mask = tf_bits, mask
        # buffers are sized for Smax rounded up to a multiple of 16 steps: the sizes then repeat from step to step and the caching
        # allocator serves them from its pool (a fresh cudaMalloc in the middle of a step stalls the launch thread for tens of ms)
        self.Salloc = (self.Smax + 15) // 16 * 16
        S, R = self.Salloc, self.Rtot
        self.sv = None
        if self.save:
            # zero-filled: rows of bars with fewer steps than Smax are never written but are read (times zero) by the contractions
            zb = zeros_multi(dev, logp=((B, bars, max_steps, V), F32), hs=((S + 1, R, D), F32), ctxs=((S, R, D), F32), attn=((S, R, T), F32),
                             qs=((S + 1, R, A), F32), eqs=((S, R, A), F32), xtok=((S + 1, R, E), F32), toks=((S + 1, R), torch.int32),
                             ml=((S, R, 2), F32))
            self.logp = zb.pop("logp")
            self.sv = dict(gates=torch.empty(S, R, 4 * D, device=dev, dtype=F32), **zb)
        else:
            self.logp = z(B, bars, max_steps, V)
        self.counters = []
        self.shared = [self.Ee, self.logp, self.lengths, self.enc] + ([self.gt] if self.gt is not None else []) + \
                      ([self.mask] if self.mask is not None else []) + \
                      (list(self.sv.values()) if self.sv else [])
        if side is not None:
            for t_ in self.shared:
                t_.record_stream(side)

    def _wargs(self):
        attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out = self.w
        return dict(Wattn=attn_w, v=attn_v, emb=emb, W_ih=W_ih, W_hh=W_hh, b_ih=b_ih, b_hh=b_hh, W_out=W_out, b_out=b_out)

    def launch(self, k0, nq, h0):
        """Bars k0 .. k0+nq-1; h0 (nq, B, D) = their bar summaries.  Call with `self.stream` current and ordered after h0."""
        B, D, A, E, V, VP, NS = self.B, self.D, self.A, self.E, self.V, self.VP, self.NS
        dev = self.enc.device
        R = nq * B
        z = lambda *s_, dt=F32: torch.zeros(*s_, device=dev, dtype=dt)
        e = lambda *s_, dt=F32: torch.empty(*s_, device=dev, dtype=dt)
        Sq = self.steps[k0:k0 + nq]
        h0 = _f(h0.detach()).reshape(R, D)
        mask = self.mask
        if self.save:
            sv, Rtot, r0 = self.sv, self.Rtot, k0 * B
            sv["hs"][0, r0:r0 + R] = h0
            lengths = self.lengths
        else:
            hs = e(2, R, D)
            hs[0] = h0
            sv = dict(hs=hs, ctxs=e(1, R, D), attn=None, gates=None, qs=e(2, R, A), eqs=None, xtok=None, toks=None, ml=None)
            Rtot, r0 = R, 0
            lengths = self.lengths.data_ptr() + k0 * B * 8
            if mask is not None:
                mask = mask[:, k0 * B:k0 * B + R].contiguous()
        zb = zeros_multi(dev, counters=((2,), torch.int32), logits=((R, VP), F32), tickets=((B,), torch.int32), sync=((2,), torch.int32),
                         eos=((R,), torch.int32))
        counters = zb["counters"]
        scratch = dict(xbuf=e(R, E + D), pm=e(R, NS), pl=e(R, NS), pc=e(R, NS, D), **zb)
        bits = None
        if self.tf_bits is not None:
            bits = 0
            for q in range(nq):
                bits |= self.tf_bits[k0 + q] << (q * self.Smax)
        args = make_decm_args(Sq, bits, B=B, NQ=nq, T=self.T, V=V, VP=VP, S=max(Sq), max_steps=self.max_steps, NS=NS, tile=self.tile,
                              inference=int(self.inference), save=int(self.save), Rtot=Rtot, r0=r0, bars=self.bars, k0=k0, Spitch=self.Smax, tc=self.tc,
                              enc=self.enc, Ee=self.Ee, gt=self.gt,
                              mask=mask, logp=self.logp, lengths=lengths, prof=PROF.get("fwd"), **self._wargs(), **sv, **scratch)
        with ktime("note_decoder_fwd"):
            lib.pa2s_decm_fwd(stream(), ctypes.byref(args), self.sos, self.eos)
        SYNC_FLAGS.append(scratch["sync"])
        self.counters.append(counters)
        return counters

    def backward(self, dlogp):
        """Reverse pass over every saved row -> dict(denc, dEp, dh0 (bars,B,D), wgrads (9, parameter order of NoteDecoder._weights)).
        Call with `self.stream` current and ordered after dlogp."""
        assert self.save
        B, T, D, A, E, V, VP, NS, S, R = self.B, self.T, self.D, self.A, self.E, self.V, self.VP, self.NS, self.Smax, self.Rtot
        attn_w, attn_v, emb, W_ih, W_hh, b_ih, b_hh, W_out, b_out = self.w
        dev = self.enc.device
        sv = self.sv
        z = lambda *s_, dt=F32: torch.zeros(*s_, device=dev, dtype=dt)
        e = lambda *s_, dt=F32: torch.empty(*s_, device=dev, dtype=dt)
        dlogp = _f(dlogp)
        assert dlogp.shape == self.logp.shape
        W_hT = attn_w.detach()[:, :D].t().contiguous()
        W_ihT = W_ih.detach().t().contiguous()
        W_hhT = W_hh.detach().t().contiguous()
        nqmax = lib.pa2s_decm_max_queries()
        groups = [(k0, min(nqmax, self.bars - k0)) for k0 in range(0, self.bars, nqmax)]
        nblk = lib.pa2s_decm_deferred_blocks(T)
        Sa = self.Salloc
        bw = dict(dlogits_all=e(Sa, R, VP),
                  **zeros_multi(dev, dgi_all=((Sa, R, 3 * D), F32), dgh_all=((Sa, R, 3 * D), F32), dq_all=((Sa + 1, R, A), F32),
                                dctx_all=((Sa, R, D), F32), dxtok_all=((Sa, R, E), F32), ds_all=((Sa, R, T), F32)))
        dhc_all = e(Sa * R, 2 * D)
        dhq = e(R, D)
        st = stream()

        def gargs(k0, nq, **extra):
            return make_decm_args(self.steps[k0:k0 + nq], None, B=B, NQ=nq, T=T, V=V, VP=VP, S=S, max_steps=self.max_steps, NS=NS, tile=self.tile,
                                  inference=0, save=1, Rtot=R, r0=k0 * B, bars=self.bars, k0=k0, Spitch=S, tc=self.tc, enc=self.enc, Ee=self.Ee,
                                  logp=self.logp, dlogp=dlogp, dhc_all=dhc_all, prof=PROF.get("bwd"), W_hT=W_hT, W_ihT=W_ihT, W_hhT=W_hhT,
                                  **self._wargs(), **sv, **bw, **extra)
        with ktime("note_decoder_bwd"):
            for k0, nq in groups:
                a0 = gargs(k0, nq)
                lib.pa2s_decm_dlogits(st, ctypes.byref(a0))
            gemm(bw["dlogits_all"], W_out, dhc_all, S * R, 2 * D, V, lda=VP, ldb=2 * D, ldc=2 * D)
            dEps, dv_parts = [], []
            for k0, nq in groups:
                Rg = nq * B
                scr = dict(d_hc=e(Rg, 2 * D), dx=e(Rg, E + D), dq_part=e(Rg, NS, A), dh_carry=e(Rg, D), dEp=e(B, T, A), dv_part=e(B * nblk, A),
                           **zeros_multi(dev, tickets=((B,), torch.int32), sync=((2,), torch.int32)))
                a1 = gargs(k0, nq, dhq=dhq.data_ptr() + k0 * B * D * 4, **scr)
                lib.pa2s_decm_bwd_chain(st, ctypes.byref(a1))
                # the deferred attention gradients (dEp, dv: 0.7 ms) run on an auxiliary stream, next to the weight-gradient
                # contractions and the d_enc GEMM below, which only read what the chain kernel left behind
                aux = self._aux_stream()
                aux.wait_event(torch.cuda.current_stream().record_event())
                for t_ in (scr["dEp"], scr["dv_part"]):
                    t_.record_stream(aux)
                lib.pa2s_decm_bwd_deferred(ctypes.c_void_p(aux.cuda_stream), ctypes.byref(a1))
                SYNC_FLAGS.append(scr["sync"])
                dEps.append(scr["dEp"])
                dv_parts.append(scr["dv_part"])
        dxt = bw["dxtok_all"][:S]
        if self.mask is not None:
            dxt = dxt * self.mask[:S]
        rec = dict(S=S, B=R, dlogits=bw["dlogits_all"], hs=sv["hs"], ctxs=sv["ctxs"], dgi=bw["dgi_all"], dgh=bw["dgh_all"], dq=bw["dq_all"],
                   xtok=sv["xtok"], dv_part=None, dxt=dxt, toks=sv["toks"])
        wg = list(decoder_weight_grads([rec], (D, A, V, E, VP, attn_v.shape)))
        # d_enc[b] = sum over (step, bar) of attn^T dctx: rows (s, bar, b) -> one batched GEMM with K = S * bars
        denc = context_grad_enc(sv["attn"], bw["dctx_all"], B, T, D, S * self.bars)
        torch.cuda.current_stream().wait_event(aux.record_event())
        dEp = dEps[0]
        for x_ in dEps[1:]:
            dEp = dEp + x_
        wg[1] = colsum(dv_parts[0] if len(dv_parts) == 1 else torch.cat(dv_parts)).reshape(attn_v.shape)
        return dict(denc=denc, dEp=dEp, dh0=dhq.view(self.bars, B, D), wgrads=wg)

    def _aux_stream(self):
        if getattr(self, "_aux", None) is None:
            key = self.stream.cuda_stream if self.stream is not None else 0
            if key not in _AUX_STREAMS:
                _AUX_STREAMS[key] = torch.cuda.Stream()
            self._aux = _AUX_STREAMS[key]
        return self._aux


class DecodersFn(torch.autograd.Function):
    """The autograd node of ALL note decoding of a forward pass (both staves, every bar): the forward kernels were launched by the
    StaffRuns while the bar loop ran; this node only hands out their log-probabilities and, in backward, runs one reverse pass
    per staff concurrently on the two staff streams.
    Inputs: runs (StaffRun upper, StaffRun lower), enc, Ep_upper, Ep_lower, h0 (bars,B,D) stacked bar summaries, then the nine
    weights of the upper and of the lower NoteDecoder."""

    @staticmethod
    def forward(ctx, runs, enc, Ep_up, Ep_lo, h0, *weights):
        ctx.runs = runs
        ctx.prec = current_precision()
        return runs[0].logp, runs[1].logp

    @staticmethod
    def backward(ctx, d_up, d_lo):
        runs = ctx.runs
        cur = torch.cuda.current_stream()
        ev = cur.record_event()
        res = []
        with use_precision(ctx.prec):
            for run, d in zip(runs, (d_up, d_lo)):
                if d is None:
                    d = torch.zeros_like(run.logp)
                side = run.stream or cur
                if side != cur:
                    side.wait_event(ev)
                    d.record_stream(side)
                with torch.cuda.stream(side):
                    r = run.backward(d)
                    done = side.record_event()
                res.append((r, done, side))
        for r, done, side in res:
            if side != cur:
                cur.wait_event(done)
                for t_ in [r["denc"], r["dEp"], r["dh0"]] + r["wgrads"]:
                    t_.record_stream(cur)
        (ru, _, _), (rl, _, _) = res
        denc = ru["denc"] + rl["denc"]
        dh0 = ru["dh0"] + rl["dh0"]
        ctx.runs = None
        return (None, denc, ru["dEp"], rl["dEp"], dh0, *ru["wgrads"], *rl["wgrads"])


# ----------------------------------------------------------------------------------------------------------------
# Loss / optimiser (pretrain.py:56-93, 125-128)
# ----------------------------------------------------------------------------------------------------------------
class NLLFn(torch.autograd.Function):
    """mean over rows with target != ignore of -logp[row, target] (torch.nn.NLLLoss)."""

    @staticmethod
    def forward(ctx, logp, target, ignore):
        logp = _f(logp)
        V = logp.shape[-1]
        rows = logp.numel() // V
        tgt = target.reshape(-1).contiguous()
        assert tgt.numel() == rows and tgt.dtype == torch.int64
        acc = torch.empty(2, device=logp.device, dtype=F32)
        lib.pa2s_nll_fwd(stream(), ptr(logp), ptr(tgt), rows, V, ignore, ptr(acc))
        ctx.save_for_backward(tgt, acc)
        ctx.meta = (logp.shape, rows, V, ignore)
        return -(acc[0] / acc[1])

    @staticmethod
    def backward(ctx, g):
        tgt, acc = ctx.saved_tensors
        shape, rows, V, ignore = ctx.meta
        grad = torch.empty(shape, device=tgt.device, dtype=F32)
        lib.pa2s_nll_bwd(stream(), ptr(grad), ptr(tgt), rows, V, ignore, ptr(acc), ptr(_f(g).reshape(1)))
        return grad, None, None


def nll_loss(logp, target, ignore_index=-100):
    return NLLFn.apply(logp, target, int(ignore_index))
