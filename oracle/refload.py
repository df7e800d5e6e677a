"""Import the reference's own MotifSeq.py / segmenter.py UNMODIFIED, for pinning the oracle.

TEST INFRASTRUCTURE ONLY, and only usable where ``/root/reference`` exists (the build
container).  Nothing that runs on the GPU box may call this: ``tests/golden/make_golden.py``
uses it here to write fixtures, and the ``not gpu`` tests that use it skip when the tree is
absent.

The scripts import packages that are not installed (h5py, scrappy, mlpy, matplotlib with a
Tk backend).  Those names are pre-seeded in ``sys.modules`` with inert stubs; ``mlpy`` gets
``dtw_subsequence`` = the oracle's restatement, so the reference's ``get_region_multi``
(MotifSeq.py:431-456) runs end to end on top of it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SQK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "segmenter.py"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is a callable no-op / nested stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return None


def _seed_stubs(dtw_subsequence):
    saved = {}
    names = ["h5py", "scrappy", "mlpy", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
             "matplotlib.cm"]
    for n in names:
        saved[n] = sys.modules.get(n)
    mpl = _Anything("matplotlib")
    mpl.rcParams = {}
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    for sub in ("pyplot", "patches", "cm"):
        m = _Anything(f"matplotlib.{sub}")
        setattr(mpl, sub, m)
        sys.modules[f"matplotlib.{sub}"] = m
    sys.modules["h5py"] = _Anything("h5py")
    sys.modules["scrappy"] = _Anything("scrappy")
    ml = types.ModuleType("mlpy")
    ml.dtw_subsequence = dtw_subsequence
    sys.modules["mlpy"] = ml
    return saved


def _restore(saved):
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


def load(script: str, dtw_subsequence=None):
    """Load ``/root/reference/<script>.py`` as a module object (not registered in sys.modules)."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    if dtw_subsequence is None:
        from .cpu import dtw_subsequence as _d
        dtw_subsequence = _d
    saved = _seed_stubs(dtw_subsequence)
    try:
        path = os.path.join(REFERENCE_ROOT, f"{script}.py")
        spec = importlib.util.spec_from_file_location(f"_sqk_reference_{script}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        _restore(saved)
    return mod


class Args:
    """argparse.Namespace stand-in carrying the reference defaults (segmenter.py:57-96,
    MotifSeq.py:90-125)."""

    def __init__(self, **kw):
        d = dict(error=5, corrector=50, window=150, seg_dist=50, std_scale=0.75, stall_len=0.25,
                 stall_start=300, gap_dist=3000, lim_hi=900, lim_low=0, stall=False, gap=False,
                 test=False, Num=-1, raw_signal=False,
                 scale_hi=1200, scale_low=0, sig_extract=False, view=False, save=None,
                 slope=2.90, intercept=-9.6, std_const=0.08468, scale="medmad")
        d.update(kw)
        self.__dict__.update(d)
