"""squigglekit_b200 -- B200-native MotifSeq / segmenter hot paths of SquiggleKit.

The compute lives in ``libsqk.so`` (hand-written sm_100a CUDA, C ABI in include/sqk.h); this
package is the thin Python host layer that mirrors the reference's per-read functions, batched.
There is no CPU fallback: importing the package is cheap, using it needs the built library and a
B200.
"""
from .core import (  # noqa: F401
    AdapterConfig,
    Context,
    RollmeanConfig,
    SegConfig,
    SqkError,
    HIT_DTYPE,
    hits_from_torch,
    pinned_empty,
    pinned_free,
    segs_to_lists,
    test_segs,
)
from .models import read_bait_model, read_model, read_synth_model  # noqa: F401

__version__ = "0.1.0"
