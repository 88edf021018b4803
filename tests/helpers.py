"""Shared test helpers: synthetic targets, deterministic weights, RNG replay into the CUDA path."""
import numpy as np
import torch

from oracle import a2s_oracle as O


from piano_a2s_b200.synthetic import make_ground_truth  # noqa: E402,F401  (same generator the bench uses)


def synth_state_dict(model, seed=7):
    """Deterministic, machine-independent weights: an integer LCG mapped to (-a, a) with a = xavier-like bound per tensor
    (BatchNorm weight/var in (0.5,1.5)).  Used for golden vectors so they do not depend on torch's RNG or LAPACK."""
    sd = {}
    state = np.uint64(seed * 2654435761 + 12345)
    for k, v in model.state_dict().items():
        n = v.numel()
        if v.dtype != torch.float32:
            sd[k] = v.clone()
            continue
        idx = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            x = (idx * np.uint64(6364136223846793005) + state) ^ ((idx * np.uint64(1442695040888963407)) >> np.uint64(29))
            x = (x * np.uint64(2862933555777941757) + np.uint64(3037000493))
            state = state * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
        u = ((x >> np.uint64(40)).astype(np.float64) / float(1 << 24))          # [0,1)
        if "running_var" in k or (("bn" in k) and k.endswith(".weight")):
            t = 0.5 + u
        elif "running_mean" in k or (("bn" in k) and k.endswith(".bias")):
            t = (u - 0.5) * 0.2
        else:
            fan_in = v.shape[1] * (v[0][0].numel() if v.dim() > 2 else 1) if v.dim() > 1 else max(n, 1)
            a = (3.0 / fan_in) ** 0.5 if v.dim() > 1 else 0.05
            t = (2 * u - 1) * a
        sd[k] = torch.from_numpy(t.astype(np.float32)).reshape(v.shape)
    return sd


def lcg_uniform(shape, seed=1):
    """Deterministic U[0,1) float32 tensor from integer arithmetic only (bit-exact on any machine)."""
    n = int(np.prod(shape))
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (idx * np.uint64(6364136223846793005) + np.uint64(seed * 1000003 + 7)) ^ ((idx * np.uint64(1442695040888963407)) >> np.uint64(31))
        x = x * np.uint64(2862933555777941757) + np.uint64(3037000493)
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return torch.from_numpy(u.astype(np.float32)).reshape(shape)


class ReplayDeviceSource:
    """Feeds coins/masks recorded by oracle.RecordingSource into the CUDA path (piano_a2s_b200.rng interface)."""

    def __init__(self, coins, masks):
        self.coins_ = list(coins)
        self.masks = list(masks)
        self.ci = 0
        self.mi = 0

    def coin(self):
        c = self.coins_[self.ci]
        self.ci += 1
        return c

    def coins(self, n):
        out = self.coins_[self.ci:self.ci + n]
        assert len(out) == n, "coin stream exhausted"
        self.ci += n
        return out

    def dropout_mask(self, shape, p, device, kind):
        if kind == "conv":
            m = self.masks[self.mi].transpose(1, 2).contiguous()       # recorded as (B,C,T)
            self.mi += 1
        elif kind == "bar_token":
            m = self.masks[self.mi]
            self.mi += 1
        elif kind == "note_steps":
            S = shape[0]
            m = torch.stack([x[:, 0, :] for x in self.masks[self.mi:self.mi + S]], 0)
            self.mi += S
        else:
            raise KeyError(kind)
        assert tuple(m.shape) == tuple(shape), (kind, m.shape, shape)
        return m.to(device)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
