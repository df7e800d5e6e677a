"""Seeded inputs for the rolling-mean adapter finder (dRNA_segmenter.py TSV branch, :272-326): synthetic dRNA-like
reads (a low adapter plateau, then the RNA signal), generated -- not stored -- so that the generator script and the tests
see the same bytes (numpy's PCG64 stream is stable across versions)."""
from __future__ import annotations

import numpy as np

SEED = 20251017


def reads():
    rng = np.random.default_rng(SEED)
    out = []
    for i in range(36):
        n = int(rng.integers(9000, 42000))
        lead = int(rng.integers(0, 3000)) if i % 3 else 0          # open-pore / noise in front of the adapter
        adapter = int(rng.integers(2500, 9000))
        lo_lvl, hi_lvl = rng.uniform(380, 470), rng.uniform(560, 720)
        x = np.arange(n)
        sig = np.where((x >= lead) & (x < lead + adapter), lo_lvl, hi_lvl) + rng.normal(0, rng.uniform(5, 30), n)
        lv = np.repeat(rng.normal(0, 40, n // 12 + 1), 12)[:n]      # dwell structure on the RNA part
        sig[lead + adapter:] += lv[lead + adapter:]
        if i % 4 == 0:                                              # a second low stretch: merged or separate segment
            at = min(n - 10, lead + adapter + int(rng.integers(300, 6000))); ln = int(rng.integers(500, 5000))
            sig[at:at + ln] = lo_lvl + rng.normal(0, 10, min(ln, n - at))
        if i % 5 == 0:                                              # short dips that do not last lo_thresh samples
            for _ in range(4):
                at = int(rng.integers(0, n - 2500)); ln = int(rng.integers(100, 2400))
                sig[at:at + ln] -= rng.uniform(60, 200)
        if i % 6 == 0:
            sig[rng.integers(0, n, 40)] = rng.choice([-5, 0, 1200, 1500, 3000], 40)   # outliers
        out.append(np.clip(np.rint(sig), -32768, 32767).astype(np.int16))
    out.append(rng.integers(300, 700, 500).astype(np.int16))        # shorter than the window: every mean is NaN
    out.append(rng.integers(300, 700, 2000).astype(np.int16))       # exactly one window: count == 1, std NaN
    out.append(rng.integers(300, 700, 2001).astype(np.int16))       # two means
    out.append(np.full(9000, 500, np.int16))                        # constant: nothing is < bot
    out.append(np.zeros(0, np.int16))
    out.append(np.full(5000, 2000, np.int16))                       # all outliers
    out.append(np.r_[np.full(6000, 400), np.full(9000, 650)].astype(np.int16))      # noiseless step
    out.append(np.r_[np.full(5000, 650), np.full(4000, 400), np.full(6000, 650)].astype(np.int16))
    long_low = np.r_[rng.normal(420, 12, 230000), rng.normal(640, 30, 260000)]      # low stretch longer than hi_thresh ...
    long_low[300000:304000] = rng.normal(420, 10, 4000)                              # ... then one that qualifies
    out.append(np.clip(np.rint(long_low), 1, 1199).astype(np.int16))
    # levels that make the rolling mean sit exactly on / chatter around the threshold
    out.append(np.tile(np.r_[np.full(3000, 400), np.full(3000, 600)], 5).astype(np.int16))
    out.append((500 + 100 * np.sign(np.sin(np.arange(40000) / 700.0))).astype(np.int16))
    return out


def concatenated():
    rs = reads()
    offsets = np.zeros(len(rs) + 1, np.int64)
    np.cumsum([r.size for r in rs], out=offsets[1:])
    return np.concatenate(rs), offsets
