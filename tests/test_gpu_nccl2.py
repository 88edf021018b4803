"""Two B200s, NCCL: the multi-GPU code paths EXECUTED and compared with the single-process oracle on the union batch
(skipped on a one-GPU box -- run with `gpurun --gpus 2`).

* `ops.ConvStackFn` with `sync=True` (SyncBatchNorm semantics, models.py:523-543 under speechbrain's
  `convert_sync_batchnorm`) on UNEVEN shards (3 + 2 clips): per-rank outputs, rank-averaged parameter gradients and the
  BatchNorm running statistics equal the oracle's ConvStack on the 5-clip union; running statistics bit-identical across ranks.
* the whole `ScoreTranscription`, wrapped the way the reference's trainer wraps it --
  `torch.nn.SyncBatchNorm.convert_sync_batchnorm` + `torch.nn.parallel.DistributedDataParallel` -- one teacher-forced training
  step per rank: DDP's averaged gradients equal the oracle's gradients of (loss(shard 0) + loss(shard 1)) / 2 with BatchNorm
  over the union.
* `train.FlatAdadelta` bucketed all-reduce + step on both ranks: parameters stay bit-identical across ranks.
"""
import os
import socket

import pytest
import torch

from helpers import lcg_uniform, make_ground_truth, rel_err, synth_state_dict
from oracle import a2s_oracle as O

pytestmark = pytest.mark.gpu

SMALL = dict(freq_bins=32, max_bars=2, max_length=(14, 9))
SHARDS = (3, 2)


class OnesSource(O.RandomSource):
    """no dropout, coins that always teacher-force: removes every cross-clip coupling except BatchNorm"""

    def coin(self):
        return 0.0

    def dropout_mask(self, shape, p):
        return torch.ones(shape)


class OnesDeviceSource:
    def coin(self):
        return 0.0

    def coins(self, n):
        return [0.0] * n

    def dropout_mask(self, shape, p, device, kind):
        return torch.ones(shape, device=device)


def _to_np(o):
    """tensors -> numpy (pickled by value: torch's shared-memory hand-over can outlive the worker that owns the segment)"""
    if torch.is_tensor(o):
        return o.detach().cpu().numpy()
    if isinstance(o, dict):
        return {k: _to_np(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_to_np(v) for v in o]
    return o


def _to_torch(o):
    import numpy as np
    if isinstance(o, np.ndarray):
        return torch.from_numpy(o)
    if isinstance(o, dict):
        return {k: _to_torch(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_to_torch(v) for v in o]
    return o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    import models
    torch.manual_seed(1234)
    m = models.ScoreTranscription(**SMALL)
    sd = synth_state_dict(m)
    n = sum(SHARDS)
    x = lcg_uniform((n, 1, 21, 32), seed=11)
    x[1, :, 15:, :] = 0.                                         # a shorter clip (zero-padded frames)
    gt = make_ground_truth(n, 2, 14, 9, seed=3, lo_up=(3, 13), lo_lo=(2, 9))
    w = lcg_uniform((n, 21, 256), seed=5) - 0.5
    return sd, x, gt, w


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import models
        from piano_a2s_b200 import ops, rng, train
        sd, x, gt, w = _inputs()
        lo = sum(SHARDS[:rank])
        hi = lo + SHARDS[rank]
        out = {}
        # ---- (1) ConvStackFn, sync=True, uneven shards
        cs = models.ConvStack(1, 32, 256)
        cs.load_state_dict({k[len("convstack."):]: v for k, v in sd.items() if k.startswith("convstack.")})
        cs = cs.to(dev).train()
        cs.sync_batchnorm = True
        with rng.use_source(OnesDeviceSource()):
            y = cs(x[lo:hi].to(dev))
        (y * w[lo:hi].to(dev)).sum().backward()
        grads = {}
        for k, p in cs.named_parameters():
            g = p.grad.clone()
            dist.all_reduce(g)
            grads[k] = (g / world).cpu()
        out["cs_y"] = y.detach().cpu()
        out["cs_grads"] = grads
        out["cs_stats"] = {k: v.detach().cpu() for k, v in cs.state_dict().items() if "running" in k or "tracked" in k}
        # the same in the exact-fp32 precision mode: isolates the SyncBatchNorm arithmetic from the split-bf16 rounding of the convolutions
        cs32 = models.ConvStack(1, 32, 256)
        cs32.load_state_dict({k[len("convstack."):]: v for k, v in sd.items() if k.startswith("convstack.")})
        cs32 = cs32.to(dev).train()
        cs32.sync_batchnorm = True
        with rng.use_source(OnesDeviceSource()), ops.use_precision("fp32"):
            y32 = cs32(x[lo:hi].to(dev))
            (y32 * w[lo:hi].to(dev)).sum().backward()
        g32 = {}
        for k, p in cs32.named_parameters():
            g = p.grad.clone()
            dist.all_reduce(g)
            g32[k] = (g / world).cpu()
        out["cs32_y"], out["cs32_grads"] = y32.detach().cpu(), g32
        # ---- (2) the reference trainer's wrapping: SyncBatchNorm.convert_sync_batchnorm + DistributedDataParallel
        m = models.ScoreTranscription(**SMALL)
        m.load_state_dict(sd)
        m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m.to(dev).train())
        assert isinstance(m.convstack.bn1, torch.nn.SyncBatchNorm)
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank], find_unused_parameters=False)
        gtd = train.targets_to_device([t[lo:hi] for t in gt], dev)
        with rng.use_source(OnesDeviceSource()):
            outs = ddp(x[lo:hi].to(dev), inference=False, ground_truth=gtd, teacher_forcing_ratio=1.0, device=dev)
        loss, _ = train.compute_objectives(outs, gtd)
        loss.backward()
        torch.cuda.synchronize()
        ops.check_sync_flags()
        out["ddp_loss"] = float(loss)
        out["ddp_outs"] = [o_.detach().cpu() for o_ in outs]
        out["ddp_grads"] = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
        # ---- (3) FlatAdadelta: bucketed all-reduce (hooks) + fused update; parameters must stay identical across ranks
        m2 = models.ScoreTranscription(**SMALL)
        m2.load_state_dict(sd)
        m2 = m2.to(dev).train()
        m2.convstack.sync_batchnorm = True
        opt = train.FlatAdadelta(m2, bucket_bytes=1 << 20)
        for _ in range(2):
            with rng.use_source(OnesDeviceSource()):
                train.fit_batch(m2, opt, x[lo:hi].to(dev), gtd, 1.0)
        torch.cuda.synchronize()
        bn = torch.cat([b.reshape(-1).float() for n_, b in m2.named_buffers() if "running" in n_])
        sums = torch.stack([t.contiguous().view(torch.int32).to(torch.int64).sum() for t in (opt.flat, opt.square_avg, opt.acc_delta, bn)])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        out["flat_consistent"] = all(bool((a == allsums[0]).all().item()) for a in allsums)
        out["flat_launched_early"] = (opt.launched_early, len(opt.buckets))
        q.put((rank, _to_np(out)))
    except Exception as e:                                        # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, {"error": traceback.format_exc() + repr(e)}))
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_results():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: _to_torch(o) for r, o in (q.get(timeout=600) for _ in procs)}
    for p in procs:
        p.join(timeout=120)
    for r in (0, 1):
        assert "error" not in res[r], res[r]["error"]
    return res


def test_convstack_syncbn_uneven_shards_match_union_oracle(two_rank_results):
    res = two_rank_results
    sd, x, gt, w = _inputs()
    sdo = {k: v.clone() for k, v in sd.items() if k.startswith("convstack.")}
    sdg = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in sdo.items()}
    ns = {}
    ref = O.conv_stack(x, sdg, True, OnesSource(), new_stats=ns)
    ((ref * w).sum() / 2).backward()                            # DDP semantics: mean over ranks of the per-rank (sum) losses
    lo = 0
    for r, n in enumerate(SHARDS):
        e = rel_err(res[r]["cs_y"], ref[lo:lo + n].detach())
        print("rank", r, "ConvStack (sync) output rel err %.2e" % e)
        assert e < 1e-4
        lo += n
    lo = 0
    for r, n in enumerate(SHARDS):
        assert rel_err(res[r]["cs32_y"], ref[lo:lo + n].detach()) < 1e-5
        lo += n
    for k, g in res[0]["cs32_grads"].items():                   # exact-fp32 kernels: the cross-rank statistics arithmetic itself
        e = rel_err(g, sdg["convstack." + k].grad)
        print("  fp32 grad", k, "%.2e" % e)
        assert e < 2e-3, k                                      # gradient bound of the fp32 mode (tests/test_gpu_parity.py)
    for k, g in res[0]["cs_grads"].items():
        e = rel_err(g, sdg["convstack." + k].grad)
        print("  bf16x3 grad", k, "%.2e" % e)
        # split-bf16 convolutions on a ragged batch: a ReLU pre-activation within ~1e-5 of zero may take the other branch than in the
        # oracle, which moves a weight-gradient entry by a few 1e-2 of the tensor's largest (3e-2 in tests/test_gpu_paths.py)
        assert e < 5e-2, k
    for k, v in ns.items():
        kk = k[len("convstack."):]
        a, b = res[0]["cs_stats"][kk], res[1]["cs_stats"][kk]
        assert torch.equal(a, b), kk                            # identical on both ranks
        assert rel_err(a.float(), v.float()) < 1e-5, kk         # and equal to the union's statistics (uneven counts: 3 + 2 clips)


def test_ddp_wrapped_model_matches_union_oracle(two_rank_results):
    res = two_rank_results
    sd, x, gt, _ = _inputs()
    sdg = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in sd.items()}
    ref = O.score_transcription(sdg, x, SMALL, False, gt, 1.0, True, OnesSource())
    losses, lo = [], 0
    for n in SHARDS:
        losses.append(O.training_loss([t[lo:lo + n] for t in ref], [t[lo:lo + n] for t in gt]))
        lo += n
    (sum(losses) / len(SHARDS)).backward()
    lo = 0
    for r, n in enumerate(SHARDS):
        assert abs(res[r]["ddp_loss"] - float(losses[r])) < 1e-4 * abs(float(losses[r]))
        for a, b in zip(res[r]["ddp_outs"], ref):
            b = b[lo:lo + n].detach()
            # a rank decodes max over ITS clips of the target lengths, the union oracle max over all five: rows the rank did not run
            # stay zero (models.py:385) and belong to <pad> targets -- compare the rows the rank executed
            ran = (a != 0).any(-1, keepdim=True).expand_as(a)
            assert ran.any() and rel_err(torch.where(ran, a, torch.zeros_like(a)), torch.where(ran, b, torch.zeros_like(b))) < 5e-4
        lo += n
    num = den = 0.0
    for k, g in res[0]["ddp_grads"].items():
        want = sdg[k].grad
        assert torch.equal(g, res[1]["ddp_grads"][k]), k        # DDP left the same averaged gradient on both ranks
        e = rel_err(g, want)
        num += (g.double() - want.double()).pow(2).sum().item()
        den += want.double().pow(2).sum().item()
        print("  ddp grad", k, "%.2e" % e)
        assert e < 3e-2, k                                      # bf16x3 bound of tests/test_gpu_paths.py (ragged batch)
    assert (num / den) ** 0.5 < 1e-2


def test_flat_adadelta_keeps_ranks_identical(two_rank_results):
    for r in (0, 1):
        assert two_rank_results[r]["flat_consistent"]
        early, nb = two_rank_results[r]["flat_launched_early"]
        assert early >= nb                                      # every bucket's all-reduce started from a backward hook
