"""CPU, world_size 2, gloo: the host-side data-parallel logic (flat gradient averaging, SyncBatchNorm statistic reduction
arithmetic, per-rank synthetic shards).  The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from piano_a2s_b200.synthetic import make_ground_truth
        from piano_a2s_b200.train import FlatAdadelta
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(7, 173), torch.nn.Linear(173, 3))       # 173-element bias: exercises the padding
        opt = FlatAdadelta(lin)
        assert all(p.data_ptr() % 256 == opt.flat.data_ptr() % 256 for p in lin.parameters())
        for i, p in enumerate(lin.parameters()):
            p.grad.fill_(float(rank + 1) * (i + 1))
        opt.allreduce_mean()
        ok = all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5 * (i + 1))) for i, p in enumerate(lin.parameters()))
        # overlapped path: buckets are reduced from post-accumulate hooks during backward; result = mean of the per-rank gradients
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(7, 173), torch.nn.Tanh(), torch.nn.Linear(173, 3))
        opt2 = FlatAdadelta(net, bucket_bytes=1024)
        assert len(opt2.buckets) >= 2 and sorted(i for b in opt2.buckets for i in b[2]) == list(range(4))
        assert all(b[0] < b[1] for b in opt2.buckets) and opt2.buckets[0][1] == opt2.n and opt2.buckets[-1][0] == 0
        xs_all = [torch.randn(5, 7, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        expect = []
        for r in range(world):
            ref = torch.nn.Sequential(torch.nn.Linear(7, 173), torch.nn.Tanh(), torch.nn.Linear(173, 3))
            ref.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
            ref(xs_all[r]).square().sum().backward()
            expect.append([q_.grad.clone() for q_ in ref.parameters()])
        for _ in range(2):                      # twice: the bucket bookkeeping must reset between steps
            opt2.zero_grad()
            opt2.launched_early = 0
            net(xs_all[rank]).square().sum().backward()
            ok = ok and opt2.launched_early == len(opt2.buckets)
            opt2.allreduce_mean()
            for i, q_ in enumerate(net.parameters()):
                ok = ok and torch.allclose(q_.grad, sum(e[i] for e in expect) / world, rtol=1e-5, atol=1e-6)
        # SyncBatchNorm statistics: global mean/var from all-reduced fp64 (sum, sumsq) equals the statistics of the union
        x = torch.randn(5 + rank, 4, dtype=torch.float64)
        sums = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(x.shape[0])], dtype=torch.float64)])
        dist.all_reduce(sums)
        n = sums[-1]
        mean, var = sums[:4] / n, sums[4:8] / n - (sums[:4] / n) ** 2
        xs = [torch.zeros(5 + r, 4, dtype=torch.float64) for r in range(world)]
        # gather the shards to check against the union directly
        gathered = [torch.zeros(6, 4, dtype=torch.float64) for _ in range(world)]
        pad = torch.zeros(6, 4, dtype=torch.float64)
        pad[: x.shape[0]] = x
        dist.all_gather(gathered, pad)
        union = torch.cat([gathered[r][: 5 + r] for r in range(world)])
        ok = ok and torch.allclose(mean, union.mean(0)) and torch.allclose(var, union.var(0, unbiased=False))
        # SyncBatchNorm BACKWARD arithmetic as ops.ConvStackFn does it: every rank forms (sum g, sum g*xhat) with the GLOBAL statistics,
        # keeps them as ITS dbeta / dgamma (the gradient all-reduce averages them afterwards), all-reduces a copy and uses the global
        # sums for the input gradient.  Checked against autograd through a plain BatchNorm over the union with loss = mean of the
        # per-rank losses (DDP semantics): averaged dgamma/dbeta equal the union's, and dx equals world * the union's dx.
        eps = 1e-5
        gamma = torch.tensor([0.7, -1.3, 2.0, 0.4], dtype=torch.float64)
        beta = torch.tensor([0.1, 0.0, -0.2, 0.3], dtype=torch.float64)
        wgt = torch.randn(6, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(9))       # loss_r = sum(relu(bn(x_r)) * wgt_r)
        invstd = 1.0 / torch.sqrt(var + eps)
        xhat = (x - mean) * invstd
        ypre = xhat * gamma + beta
        g = wgt[: x.shape[0]] * (ypre > 0)                                       # dloss_r / d bn output (ReLU mask folded in)
        local = torch.cat([g.sum(0), (g * xhat).sum(0)])
        glob = local.clone()
        dist.all_reduce(glob)
        dx = gamma * invstd * (g - glob[:4] / n - xhat * (glob[4:] / n))
        grads = local.clone()
        dist.all_reduce(grads)
        grads /= world                                                           # what FlatAdadelta.allreduce_mean does to dbeta, dgamma
        xu = union.clone().requires_grad_(True)
        gu, bu = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        yu = torch.relu(torch.nn.functional.batch_norm(xu, None, None, gu, bu, True, 0.0, eps))
        wu = torch.cat([wgt[: 5 + r] for r in range(world)])
        ((yu * wu).sum() / world).backward()
        lo = sum(5 + r for r in range(rank))
        ok = ok and torch.allclose(grads[:4], bu.grad) and torch.allclose(grads[4:], gu.grad)
        ok = ok and torch.allclose(dx, world * xu.grad[lo: lo + x.shape[0]])
        # per-rank synthetic shards differ (seed = base + rank, as bench.py does)
        g = make_ground_truth(2, 2, 14, 9, seed=1234 + rank, lo_up=(3, 13), lo_lo=(2, 9))
        tok = [torch.zeros_like(g[2]) for _ in range(world)]
        dist.all_gather(tok, g[2])
        ok = ok and not torch.equal(tok[0], tok[1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_mean_and_syncbn_sums_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
