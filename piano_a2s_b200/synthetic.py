"""Synthetic workload of SURVEY section 8(d): 12-s clips (192 000 samples @ 16 kHz) and teacher-forcing targets shaped
like the reference's collated batches (datasets/syn.py:46-74,113-121)."""
import torch

from .models import EOS, PAD


def make_audio(B, n_samples=192000, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(0.25 * torch.randn(B, n_samples, generator=g), -1.0, 1.0)


def make_ground_truth(B, bars, L_up, L_lo, seed=1, lo_up=(40, 80), lo_lo=(20, 50), n_ts=7, n_key=14):
    """[time_sig (B,bars), key (B,bars), upper (B,bars,L_up), upper_len (B,bars), lower (B,bars,L_lo), lower_len] int64 CPU.
    tokens randint(0,144), <eos> at index len, <pad> after; lengths exclude <eos> (datasets/syn.py:60-74)."""
    g = torch.Generator().manual_seed(seed)
    ts = torch.randint(0, n_ts, (B, bars), generator=g)
    key = torch.randint(0, n_key, (B, bars), generator=g)

    def staff(L, lo, hi):
        tok = torch.full((B, bars, L), PAD, dtype=torch.long)
        ln = torch.zeros(B, bars, dtype=torch.long)
        for b in range(B):
            for k in range(bars):
                n = int(torch.randint(min(lo, L - 1), min(hi, L), (1,), generator=g))
                tok[b, k, :n] = torch.randint(0, 144, (n,), generator=g)
                if n < L:
                    tok[b, k, n] = EOS
                ln[b, k] = n
        return tok, ln
    up, ul = staff(L_up, *lo_up)
    lo, ll = staff(L_lo, *lo_lo)
    return [ts, key, up, ul, lo, ll]


def executed_steps(gt):
    """Decoder steps one forward executes for these targets: sum over bars of (max_b len + 1) per staff."""
    return int((gt[3].max(0).values + 1).clamp(max=gt[2].shape[-1]).sum() + (gt[5].max(0).values + 1).clamp(max=gt[4].shape[-1]).sum())
