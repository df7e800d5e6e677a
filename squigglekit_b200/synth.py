"""Deterministic synthetic squiggle: piecewise-constant levels ~ N(510, 80)
DAC units with dwell 1+Geometric(1/8) samples, additive noise N(0, 8), rounded to int16;
0.05 % of samples forced out of the (0, 1200) outlier window so compaction is exercised.

MotifSeq sets plant the motif (10 levels ~ N(0,1), dwell 8 -> 80 float64 points) in half of
the reads; segmenter sets plant a stall plateau near the start of every read and a
homopolymer plateau in half of them.  ``*_np`` build small sets with numpy (tests, oracle
parity); ``*_torch`` build the bench-sized sets directly in HBM.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 1234
LEVEL_MEAN, LEVEL_SD, NOISE_SD = 510.0, 80.0, 8.0
CHANGE_P = 1.0 / 9.0          # mean dwell 9 samples (4 kHz / 450 bases/s)
OUTLIER_P = 5e-4


def make_motif(n_levels: int = 10, dwell: int = 8, seed: int = BASE_SEED) -> np.ndarray:
    """Expanded motif model, float64, len n_levels*dwell (z-score units, like a scrappie .model)."""
    rng = np.random.default_rng(seed)
    return np.repeat(rng.standard_normal(n_levels), dwell).astype(np.float64)


def _base_np(rng, n_reads: int, n_samples: int) -> np.ndarray:
    change = rng.random((n_reads, n_samples)) < CHANGE_P
    idx = np.cumsum(change, axis=1)
    levels = rng.standard_normal((n_reads, int(idx.max()) + 1)) * LEVEL_SD + LEVEL_MEAN
    sig = np.take_along_axis(levels, idx, axis=1)
    sig += rng.standard_normal((n_reads, n_samples)) * NOISE_SD
    return sig


def _finish_np(rng, sig: np.ndarray) -> np.ndarray:
    out = rng.random(sig.shape) < OUTLIER_P
    hi = rng.random(sig.shape) < 0.5
    sig = np.where(out & hi, 1200.0 + rng.random(sig.shape) * 800.0, sig)
    sig = np.where(out & ~hi, -rng.random(sig.shape) * 50.0, sig)
    return np.clip(np.rint(sig), -32768, 32767).astype(np.int16)


def motifseq_reads_np(n_reads: int, n_samples: int, motif: np.ndarray | None = None,
                      seed: int = BASE_SEED, plant_frac: float = 0.5):
    """-> (signals int16 [n_reads*n_samples], offsets int64 [n_reads+1], planted_at int64 [n_reads], -1 = none)."""
    rng = np.random.default_rng([seed, n_reads, n_samples])
    sig = _base_np(rng, n_reads, n_samples)
    planted = np.full(n_reads, -1, dtype=np.int64)
    if motif is not None and n_samples > motif.size:
        for r in np.nonzero(rng.random(n_reads) < plant_frac)[0]:
            at = int(rng.integers(0, n_samples - motif.size))
            sig[r, at:at + motif.size] = LEVEL_MEAN + LEVEL_SD * motif + rng.standard_normal(motif.size) * NOISE_SD
            planted[r] = at
    sig = _finish_np(rng, sig)
    offsets = np.arange(n_reads + 1, dtype=np.int64) * n_samples
    return sig.reshape(-1), offsets, planted


def segmenter_reads_np(n_reads: int, n_samples: int, seed: int = BASE_SEED):
    """Stall plateau (len U(200,600), sd 5, near the read median) starting in the first 40 samples
    of every read; homopolymer plateau (len U(200,500)) at offset >= 1500 in half of the reads."""
    rng = np.random.default_rng([seed, n_reads, n_samples, 7])
    sig = _base_np(rng, n_reads, n_samples)
    for r in range(n_reads):
        med = np.median(sig[r])
        sd = sig[r].std()
        s0 = int(rng.integers(0, 40))
        ln = int(rng.integers(200, 600))
        ln = min(ln, n_samples - s0)
        sig[r, s0:s0 + ln] = med + rng.uniform(-0.5, 0.5) * sd * 0.75 + rng.standard_normal(ln) * 5.0
        if rng.random() < 0.5 and n_samples > 2100:
            ln2 = int(rng.integers(200, 500))
            s1 = int(rng.integers(1500, n_samples - ln2))
            sig[r, s1:s1 + ln2] = med + rng.uniform(-0.5, 0.5) * sd * 0.75 + rng.standard_normal(ln2) * 5.0
    sig = _finish_np(rng, sig)
    offsets = np.arange(n_reads + 1, dtype=np.int64) * n_samples
    return sig.reshape(-1), offsets


def ragged_reads_np(lengths, motif: np.ndarray | None = None, seed: int = BASE_SEED):
    """Reads of the given lengths (0 allowed), concatenated.  -> (signals, offsets)."""
    rng = np.random.default_rng([seed, len(lengths), 99])
    parts = []
    for ln in lengths:
        if ln == 0:
            parts.append(np.zeros(0, dtype=np.int16))
            continue
        s = _base_np(rng, 1, int(ln))
        if motif is not None and ln > motif.size + 1 and rng.random() < 0.5:
            at = int(rng.integers(0, ln - motif.size))
            s[0, at:at + motif.size] = LEVEL_MEAN + LEVEL_SD * motif + rng.standard_normal(motif.size) * NOISE_SD
        parts.append(_finish_np(rng, s).reshape(-1))
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=offsets[1:])
    return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.int16)), offsets


def motifseq_reads_torch(n_reads: int, n_samples: int, motif: np.ndarray, device, seed: int = BASE_SEED,
                         plant_frac: float = 0.5, chunk_reads: int = 16384):
    """Bench-sized set generated in HBM with torch's Philox generator.  -> int16 tensor [n_reads, n_samples]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_reads, n_samples), dtype=torch.int16, device=device)
    mt = torch.as_tensor(motif, dtype=torch.float32, device=device)
    n_motif = mt.numel()
    max_levels = int(n_samples * CHANGE_P * 1.5) + 64
    for r0 in range(0, n_reads, chunk_reads):
        r1 = min(n_reads, r0 + chunk_reads)
        nr = r1 - r0
        change = torch.rand((nr, n_samples), generator=g, device=device) < CHANGE_P
        idx = torch.cumsum(change, dim=1).clamp_(max=max_levels - 1)
        levels = torch.randn((nr, max_levels), generator=g, device=device) * LEVEL_SD + LEVEL_MEAN
        sig = torch.gather(levels, 1, idx)
        sig += torch.randn((nr, n_samples), generator=g, device=device) * NOISE_SD
        if n_samples > n_motif:
            plant = torch.rand(nr, generator=g, device=device) < plant_frac
            at = (torch.rand(nr, generator=g, device=device) * (n_samples - n_motif)).long()
            cols = at[:, None] + torch.arange(n_motif, device=device)[None, :]
            val = LEVEL_MEAN + LEVEL_SD * mt[None, :] + torch.randn((nr, n_motif), generator=g, device=device) * NOISE_SD
            cur = torch.gather(sig, 1, cols)
            sig.scatter_(1, cols, torch.where(plant[:, None], val, cur))
        u = torch.rand((nr, n_samples), generator=g, device=device)
        sig = torch.where(u < OUTLIER_P / 2, torch.full_like(sig, 1500.0), sig)
        sig = torch.where(u > 1.0 - OUTLIER_P / 2, torch.full_like(sig, -20.0), sig)
        out[r0:r1] = sig.round_().clamp_(-32768, 32767).to(torch.int16)
    return out


def segmenter_reads_torch(n_reads: int, n_samples: int, device, seed: int = BASE_SEED, chunk_reads: int = 16384):
    """Bench-sized segmenter set in HBM (stall plateau in every read, homopolymer in half)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed + 7)
    out = torch.empty((n_reads, n_samples), dtype=torch.int16, device=device)
    max_levels = int(n_samples * CHANGE_P * 1.5) + 64
    ar = torch.arange(n_samples, device=device)[None, :]
    for r0 in range(0, n_reads, chunk_reads):
        r1 = min(n_reads, r0 + chunk_reads)
        nr = r1 - r0
        change = torch.rand((nr, n_samples), generator=g, device=device) < CHANGE_P
        idx = torch.cumsum(change, dim=1).clamp_(max=max_levels - 1)
        levels = torch.randn((nr, max_levels), generator=g, device=device) * LEVEL_SD + LEVEL_MEAN
        sig = torch.gather(levels, 1, idx)
        sig += torch.randn((nr, n_samples), generator=g, device=device) * NOISE_SD
        med = sig.median(dim=1).values[:, None]
        sd = sig.std(dim=1)[:, None]

        def plateau(start, length, on):
            lvl = med + (torch.rand((nr, 1), generator=g, device=device) - 0.5) * sd * 0.75
            noise = torch.randn((nr, n_samples), generator=g, device=device) * 5.0
            m = (ar >= start) & (ar < start + length) & on
            return torch.where(m, lvl + noise, sig)

        s0 = (torch.rand((nr, 1), generator=g, device=device) * 40).long()
        l0 = 200 + (torch.rand((nr, 1), generator=g, device=device) * 400).long()
        sig = plateau(s0, l0, torch.ones((nr, 1), dtype=torch.bool, device=device))
        if n_samples > 2100:
            l1 = 200 + (torch.rand((nr, 1), generator=g, device=device) * 300).long()
            s1 = 1500 + (torch.rand((nr, 1), generator=g, device=device) * (n_samples - 1500 - 500)).long()
            sig = plateau(s1, l1, torch.rand((nr, 1), generator=g, device=device) < 0.5)
        u = torch.rand((nr, n_samples), generator=g, device=device)
        sig = torch.where(u < OUTLIER_P / 2, torch.full_like(sig, 1500.0), sig)
        sig = torch.where(u > 1.0 - OUTLIER_P / 2, torch.full_like(sig, -20.0), sig)
        out[r0:r1] = sig.round_().clamp_(-32768, 32767).to(torch.int16)
    return out
