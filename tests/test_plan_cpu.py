"""The exact two-pass DTW plan (squigglekit_b200/csrc/sqk_dtw_plan.cuh) replayed on the CPU.

tests/plan_harness.cpp includes the header the CUDA kernels use and re-runs, per read, the float32 lower-bound
scan, the candidate clusters, the window-start search, the tainted exact windows and the final decision.  Checked
here: (i) the lower bound never exceeds mlpy's last row, (ii) a result the plan calls proven equals the full
float64 recurrence bit for bit, (iii) that recurrence equals the oracle, (iv) the fallback triggers when the
window is made too small or the clusters overflow, (v) on benchmark-like reads nearly every read is proven from
a window a small fraction of the read long.  No GPU involved."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from squigglekit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("plan") / "libplan_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wall",
                           os.path.join(ROOT, "tests", "plan_harness.cpp"), "-o", out])
    lib = C.CDLL(out)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.plan_two_pass.argtypes = [dp, C.c_int, dp, dp, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_int, C.c_int, ip, dp, C.POINTER(C.c_int64)]
    lib.plan_exact.argtypes = [dp, C.c_int, dp, C.c_int, ip, dp]
    lib.stats_tree_sum.argtypes = [dp, C.c_int]
    lib.stats_tree_sum.restype = C.c_double
    return lib


def _prep(sig, lo, hi, scale):
    keep = (sig > lo) & (sig < hi)
    kept = sig[keep].astype(np.float64)
    if scale == "zscale":
        center, sc = kept.mean(), kept.std()
        sc = sc if sc != 0 else 1.0
    elif scale == "medmad":
        center = np.median(kept)
        sc = np.median(np.abs(kept - center)) * 1.4826
    else:
        center, sc = 0.0, 1.0
    y = (kept - center) / sc
    return np.ascontiguousarray(y), np.ascontiguousarray(kept), keep.astype(np.uint8), float(center), float(sc)


def _run(lib, motif, sig, lo=0, hi=1200, scale="zscale", W=0, lanes=8, align_off=0, W2=-1):
    y, kept, keep, center, sc = _prep(sig, lo, hi, scale)
    x = np.ascontiguousarray(motif, dtype=np.float64)
    out = np.zeros(2, np.int32); dist = C.c_double(); diag = np.zeros(9, np.int64)
    rc = lib.plan_two_pass(x.ctypes.data_as(C.POINTER(C.c_double)), x.size, y.ctypes.data_as(C.POINTER(C.c_double)),
                           kept.ctypes.data_as(C.POINTER(C.c_double)), y.size,
                           keep.ctypes.data_as(C.POINTER(C.c_uint8)), keep.size, align_off, 8 * lanes, lo, hi, center, sc, W, W2,
                           out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(dist), diag.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0, f"harness rc {rc}"
    return (int(out[0]), int(out[1]), dist.value), diag, y


def test_exact_recurrence_matches_oracle(harness):
    rng = np.random.default_rng(5)
    for _ in range(20):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 300))
        x = rng.integers(-3, 4, n).astype(np.float64) if rng.random() < 0.5 else rng.standard_normal(n)
        y = rng.integers(-3, 4, m).astype(np.float64) if rng.random() < 0.5 else rng.standard_normal(m)
        out = np.zeros(2, np.int32); dist = C.c_double()
        harness.plan_exact(x.ctypes.data_as(C.POINTER(C.c_double)), n, y.ctypes.data_as(C.POINTER(C.c_double)), m,
                           out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(dist))
        d, _, (px, py) = oracle.dtw_subsequence(x, y, want_cost=False)
        assert (int(out[0]), int(out[1]), dist.value) == (int(py[0]), int(py[-1]), d)


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_benchmark_like_reads_are_proven_from_small_windows(harness, scale):
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(120, 4096, motif)
    want, kept = oracle.motifseq_batch(sig, off, motif, scale=scale, full_matrix=False)
    fallbacks = cols = 0
    for r in range(120):
        got, diag, y = _run(harness, motif, sig[off[r]:off[r + 1]], scale=scale, align_off=int(off[r] % 8))
        assert diag[4] == 0, "lower bound exceeded the exact last row"
        assert got == (int(want["start"][r]), int(want["end"][r]), float(want["dist"][r]))
        fallbacks += int(diag[2]); cols += int(diag[6])
    assert fallbacks <= 2
    assert cols < 0.12 * 120 * 4096          # exact work is a small fraction of the read


def test_tie_heavy_and_odd_shapes(harness):
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.choice([5, 8, 17, 40, 80, 163]))
        m = int(rng.integers(n + 1, 3000))
        levels = rng.integers(-2, 3, n // 4 + 1).astype(np.float64)
        motif = np.repeat(levels, 4)[:n] if trial % 2 else rng.standard_normal(n)
        # integer plateaus -> exact ties in the last row and in the predecessor choice
        sig = np.repeat(rng.integers(480, 540, m // 6 + 1), 6)[:m].astype(np.int16)
        if trial % 3 == 0:
            sig[rng.integers(0, m, 5)] = 3000       # outliers: raw and post-outlier positions differ
        lanes = int(rng.choice([4, 8, 16, 32]))
        got, diag, y = _run(harness, motif, sig, scale=["zscale", "medmad", "none"][trial % 3] if sig.std() > 0 else "none",
                            lanes=lanes, align_off=int(rng.integers(0, 8)))
        assert diag[4] == 0
        d, _, (px, py) = oracle.dtw_subsequence(np.ascontiguousarray(motif, dtype=np.float64), y, want_cost=False)
        assert got == (int(py[0]), int(py[-1]), d), (trial, diag)


def test_small_window_taints_and_falls_back(harness):
    motif = synth.make_motif()
    sig, off, planted = synth.motifseq_reads_np(40, 4096, motif)
    want, _ = oracle.motifseq_batch(sig, off, motif, scale="zscale", full_matrix=False)
    tainted = rescued = 0
    for r in range(40):
        # far too small a window: the path starts before it; without a second attempt the read falls back ...
        got, diag, _ = _run(harness, motif, sig[off[r]:off[r + 1]], W=20, W2=0)
        assert got == (int(want["start"][r]), int(want["end"][r]), float(want["dist"][r]))
        tainted += int(diag[5] > 0)
        if diag[5] > 0:
            assert diag[2] == 1          # a tainted minimum is never accepted
        # ... with the second attempt (default W2) the wider windows settle it
        got, diag2, _ = _run(harness, motif, sig[off[r]:off[r + 1]], W=20)
        assert got == (int(want["start"][r]), int(want["end"][r]), float(want["dist"][r]))
        if diag[5] > 0:
            assert diag2[8] == 1
            rescued += int(diag2[2] == 0)
        # ... and a second attempt that is too small as well falls back, too
        got, diag3, _ = _run(harness, motif, sig[off[r]:off[r + 1]], W=20, W2=30)
        assert got == (int(want["start"][r]), int(want["end"][r]), float(want["dist"][r]))
    assert tainted >= 20 and rescued >= tainted - 2


def test_constant_read_overflows_or_proves(harness):
    # a constant read makes every column a candidate (one huge cluster): the window start falls out of the checkpoint
    # ring or the window is the whole read; either way the answer is the exact one
    motif = synth.make_motif()
    sig = np.full(5000, 500, np.int16)
    sig[::2] += 1
    got, diag, y = _run(harness, motif, sig, scale="zscale")
    d, _, (px, py) = oracle.dtw_subsequence(motif, y, want_cost=False)
    assert got == (int(py[0]), int(py[-1]), d)


def test_stats_tree_order_is_numpys(harness):
    """The stats kernel sums fl((x-mean)^2) per leaf slot and folds the slots pairwise (sqk_stats_plan.cuh): for every
    n <= 8192 that must be np.sum's own association, bit for bit (np.std / sklearn's scale depend on it)."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal(8192) ** 2 * 1e3 + rng.random(8192)
    for n in range(1, 8193):
        if n > 300 and n % 7 and n not in (2047, 2048, 2049, 4095, 4096, 4097, 8191, 8192):
            continue
        a = np.ascontiguousarray(base[:n])
        got = harness.stats_tree_sum(a.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert got == float(np.sum(a)), n


def test_adversarial_shapes_never_violate_the_bound(harness):
    """Motifs of 7..1024 points in z-score, scaled and raw units, reads of up to 30 k samples with dwell structure,
    integer plateaus (exact ties) or three-level noise, every lane layout: the lower bound never exceeds mlpy's last
    row, the cheap per-step test never misses a candidate, and (inside the harness) every result the plan calls
    proven equals the full float64 recurrence.  Fallbacks are allowed -- they are the plan's answer to such reads."""
    rng = np.random.default_rng(99)
    proven = 0
    for trial in range(100):
        n = int(rng.choice([7, 16, 33, 80, 163, 400, 1024]))
        m = int(rng.integers(max(4 * n, 300), 30000 if n < 400 else 12000))
        kind = trial % 5
        if kind == 0:
            motif = rng.standard_normal(n)
        elif kind == 1:
            motif = np.repeat(rng.standard_normal(n // 6 + 1), 6)[:n]
        elif kind == 2:
            motif = np.repeat(rng.integers(400, 620, n // 8 + 1), 8)[:n].astype(float)      # raw units, scale "none"
        elif kind == 3:
            motif = rng.standard_normal(n) * 5
        else:
            motif = np.zeros(n)
        if kind == 2:
            sig, scale = np.repeat(rng.integers(380, 640, m // 5 + 1), 5)[:m].astype(np.int16), "none"
        else:
            sig, _, _ = synth.motifseq_reads_np(1, m, motif if n < 200 and kind < 2 else None, seed=trial)
            scale = ["zscale", "medmad"][trial % 2]
            if kind == 4:
                sig = (np.full(m, 500) + rng.integers(-1, 2, m)).astype(np.int16)
        got, diag, y = _run(harness, np.ascontiguousarray(motif, dtype=np.float64), sig, scale=scale,
                            lanes=int(rng.choice([4, 8, 16, 32])), align_off=int(rng.integers(0, 8)))
        assert diag[4] == 0, (trial, n, m, kind)
        proven += int(diag[2] == 0)
    assert proven >= 50


def test_branch_free_split_rule_equals_the_loops(harness):
    """sqk_tree_depth7 / sqk_tree_leaf7 (sqk_stats3.cuh's form of numpy's split rule) == the loops, every n <= 8192."""
    harness.stats_tree_check7.argtypes = [C.c_int]
    harness.stats_tree_check7.restype = C.c_int
    assert sum(harness.stats_tree_check7(n) for n in range(1, 8193)) == 0
