"""Training-step glue restated from the reference trainer (`ASR.compute_objectives` / `fit_batch`,
pretrain.py:56-129; finetune.py:55-127) without speechbrain: 4 NLL losses, backward, data-parallel gradient
averaging, gradient clipping and the Adadelta update (pretrain.yaml:44-47), all on libpa2s kernels.

Data parallelism: one process per GPU, batch-sharded; BatchNorm statistics are reduced across
ranks inside ConvStackFn (SyncBatchNorm semantics) and parameter gradients are averaged by NCCL all-reduces over
contiguous buckets of a flat fp32 gradient buffer (16.36 M elements = 65.4 MB).  A bucket is reduced as soon as
autograd has accumulated its last gradient (post-accumulate hooks), on NCCL's own stream, so the bar-level decoder,
encoder and `convstack.out` buckets travel over NVLink while the ConvStack backward is still running (what torch DDP
does for the reference under speechbrain); the note decoders' weights, whose gradients autograd hands over last
(ops.DecoderWeightSinkFn), form buckets of their own that are reduced at the end of backward.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from ._lib import lib, ptr, stream
from .models import PAD


def compute_objectives(predictions, ground_truth):
    """loss = NLL(time_sig) + NLL(key) + NLL_ignore<pad>(upper) + NLL_ignore<pad>(lower); returns (loss, parts)."""
    time_sig_outs, key_outs, upper_outs, lower_outs = predictions
    time_sig_gt, key_gt, upper_gt, _, lower_gt, _ = ground_truth
    parts = (ops.nll_loss(time_sig_outs, time_sig_gt), ops.nll_loss(key_outs, key_gt),
             ops.nll_loss(upper_outs, upper_gt, PAD), ops.nll_loss(lower_outs, lower_gt, PAD))
    return parts[0] + parts[1] + parts[2] + parts[3], parts


def targets_to_device(ground_truth, device):
    """`batch.to(device)` of the reference trainer (pretrain.py:39) for the six target tensors, asynchronous when they are
    pinned.  The host copies are still at hand here, so the number of decoder steps each (staff, bar) executes -- a function
    of the targets alone (HierarchicalDecoder._steps_from_gt), which the host needs to size and launch the decoder kernels --
    is counted on the host and travels with the device tensors (attribute `_pa2s_steps`), instead of being read back from the
    device inside forward().  Plain device tensors without the attribute work too (models.HierarchicalDecoder.prefetch_steps)."""
    from .models import HierarchicalDecoder
    out = [t.to(device, non_blocking=True) for t in ground_truth]
    for i in (2, 4):
        if not ground_truth[i].is_cuda:
            out[i]._pa2s_steps = HierarchicalDecoder._steps_from_gt(ground_truth[i]).tolist()
    return out


class FlatAdadelta:
    """All parameters (and their .grad) re-pointed into two flat fp32 buffers; clip_grad_norm_(max_grad_norm) +
    torch.optim.Adadelta(lr, rho, eps) semantics in two kernels (sum of squares, fused update).

    Gradients: `zero_grad()` zeroes the flat gradient buffer and sets every `p.grad` to None, so that autograd's AccumulateGrad
    just keeps the tensor a backward node hands it (with `p.grad` pointing into the flat buffer it launched one `add` per
    parameter: ~100 tiny kernels per step); `gather()` then moves a whole bucket (multi-GPU: as soon as its last gradient has
    arrived, right before its all-reduce) or everything (single GPU: in `allreduce_mean` / `step`) into the flat buffer with ONE
    multi-tensor copy and re-points `p.grad` at the flat views."""

    def __init__(self, model, lr=1.0, rho=0.95, eps=1e-8, max_grad_norm=5.0, bucket_bytes=24 << 20, overlap=True):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        # gradients of the two NoteDecoder modules are handed over by ops.DecoderWeightSinkFn, the LAST node of a backward pass
        # (models.HierarchicalDecoder.weight_sinks): they get buckets of their own, so that every other bucket -- bar-level decoder,
        # encoder, convstack.out -- is reduced while the ConvStack backward is still running
        late = [("upper_decoder." in n or "lower_decoder." in n) for n, _ in named]
        # every parameter starts on a 256-byte boundary: the kernels read weight rows with 128-bit loads
        al = lambda k: (k + 63) // 64 * 64
        n = sum(al(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.square_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.acc_delta = torch.zeros(n, device=dev, dtype=torch.float32)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.norm = torch.zeros(1, device=dev, dtype=torch.float32)
        off = 0
        self.offsets = []
        self.gviews = []
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            self.gviews.append(self.grad[off:off + k].view_as(p.data))
            p.grad = self.gviews[-1]
            self.offsets.append(off)
            off += al(k)
        self.n = n
        self.lr, self.rho, self.eps, self.max_grad_norm = lr, rho, eps, max_grad_norm
        # gradient buckets: contiguous parameter ranges of <= bucket_bytes, formed from the LAST parameter backwards because
        # backward produces gradients in roughly reverse registration order (decoder, encoder, ConvStack)
        self.buckets = []            # [lo, hi) element ranges of the flat buffer
        hi, members = n, []
        for i in range(len(self.params) - 1, -1, -1):
            members.append(i)
            if (hi - self.offsets[i]) * 4 >= bucket_bytes or i == 0 or late[i] != late[i - 1]:
                self.buckets.append((self.offsets[i], hi, members))
                hi, members = self.offsets[i], []
        self.bucket_of = [0] * len(self.params)
        for b, (_, _, idx) in enumerate(self.buckets):
            for i in idx:
                self.bucket_of[i] = b
        self._pending = [len(idx) for _, _, idx in self.buckets]
        self._works = [None] * len(self.buckets)
        self.overlap = overlap
        self.launched_early = 0      # buckets whose all-reduce started from a hook (i.e. before backward returned)
        if overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    @staticmethod
    def _world():
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _reduce_bucket(self, b):
        lo, hi, _ = self.buckets[b]
        view = self.grad[lo:hi]
        if dist.get_backend() == "nccl":
            self._works[b] = (dist.all_reduce(view, op=dist.ReduceOp.AVG, async_op=True), None)
        else:
            self._works[b] = (dist.all_reduce(view, async_op=True), view)

    def _make_hook(self, i):
        def hook(_p):
            if self._world() == 1:
                return
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0 and self._works[b] is None:
                self.gather(self.buckets[b][2])
                self._reduce_bucket(b)
                self.launched_early += 1
        return hook

    def gather(self, members=None):
        """Gradients autograd left in tensors of their own -> their places in the flat buffer (one multi-tensor copy); afterwards
        `p.grad` is the flat view again.  Parameters without a gradient keep the zeros of `zero_grad()`."""
        idx = range(len(self.params)) if members is None else members
        src, dst = [], []
        for i in idx:
            p, v = self.params[i], self.gviews[i]
            g = p.grad
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                src.append(g.detach().reshape(v.shape) if g.shape != v.shape else g.detach())
                dst.append(v)
                p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:
            p.grad = None

    def allreduce_mean(self):
        """Completes the gradient mean over ranks: buckets not yet started by a hook (parameters without a gradient this
        step, or overlap=False) are reduced now, then the compute stream waits for every bucket."""
        world = self._world()
        self.gather()
        if world > 1:
            for b in range(len(self.buckets)):
                if self._works[b] is None:
                    self._reduce_bucket(b)
            for b, (work, view) in enumerate(self._works):
                work.wait()
                if view is not None:
                    view.div_(world)
        self._pending = [len(idx) for _, _, idx in self.buckets]
        self._works = [None] * len(self.buckets)

    def step(self):
        st = stream()
        self.gather()
        lib.pa2s_sumsq(st, ptr(self.grad), self.n, ptr(self.sumsq), 1)
        lib.pa2s_adadelta(st, ptr(self.flat), ptr(self.grad), ptr(self.square_avg), ptr(self.acc_delta), self.n, ptr(self.sumsq),
                          self.max_grad_norm, self.lr, self.rho, self.eps, ptr(self.norm))
        return self.norm


def reserve_memory(gigabytes: float, device=None, streams=(), stream_gigabytes: float = 4.0, small_megabytes: int = 128):
    """Pre-fills the caching allocator's pools (allocate + free), so that a training loop never calls `cudaMalloc` in the middle of a
    step.  The launch thread runs ahead of the device; blocks that were used on the decoder's side streams stay unavailable until
    those streams have passed the point of their last use, and without spare cached memory the allocator answers the next request
    with a `cudaMalloc` -- which can wait for the device to drain (seen as 100 ms outlier steps when the host is two steps ahead).
    The allocator keeps separate pools per stream and per size class, hence: ONE large segment on the current stream (`gigabytes`),
    one per stream in `streams` (`stream_gigabytes`: the decoder's two staff streams, `model.decoder.streams()`), and
    `small_megabytes` of 2 MB segments for the < 1 MB requests of each.  Call once after the first step (the side streams exist then)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    free, _ = torch.cuda.mem_get_info(dev)
    budget = int(free * 0.5)
    for st, gb in [(None, gigabytes)] + [(s_, stream_gigabytes) for s_ in streams if s_ is not None]:
        ctx = torch.cuda.stream(st) if st is not None else torch.cuda.stream(torch.cuda.current_stream(dev))
        with ctx:
            n = min(int(gb * (1 << 30)), budget)
            keep = []
            if n > 0:
                keep.append(torch.empty(n, dtype=torch.uint8, device=dev))
                budget -= n
            keep += [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _ in range(2 * small_megabytes)]
            del keep


def teacher_forcing_schedule(ratio: float, decay: float, epoch: int, training: bool = True) -> float:
    """`on_stage_start` of pretrain.py:149-153: the ratio decays exponentially with the epoch in training
    (`teacher_forcing_ratio * teacher_forcing_decay ** epoch`, pretrain.yaml:41-42: 0.7, 0.99) and is 0 in validation / test.
    finetune.py uses the constant `hparams.teacher_forcing_ratio` (0.6): pass decay = 1."""
    return ratio * decay ** epoch if training else 0.


def fit_batch(model, optimizer: FlatAdadelta, spectrogram, ground_truth, teacher_forcing_ratio):
    """One training step (pretrain.py:121-129).  Returns the detached loss tensor (no host sync)."""
    preds = model(spectrogram=spectrogram, inference=False, ground_truth=ground_truth,
                  teacher_forcing_ratio=teacher_forcing_ratio, device=spectrogram.device)
    loss, _ = compute_objectives(preds, ground_truth)
    loss.backward()
    optimizer.allreduce_mean()
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach()
