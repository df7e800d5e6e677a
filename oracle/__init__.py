"""CPU oracle for the MotifSeq / segmenter hot paths -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``squigglekit_b200``)
never does; it fails loudly when its CUDA library is missing instead of falling back here.

Parity status (see ``sqk_oracle.c`` header and DESIGN.md): normalisation and get_segs are
pinned against sklearn / numpy / the reference's own ``segmenter.py`` run in the build
container; the DTW is a restatement of un-vendored mlpy 3.5.0 -> "parity unpinned" there.
"""
from .cpu import (  # noqa: F401
    AdapterCfg,
    RollmeanCfg,
    SegCfg,
    adapter_batch,
    adapter_seg,
    build,
    convert_to_pa,
    dtw_subsequence,
    dtw_subsequence_rolling,
    get_segs,
    lib,
    max_threads,
    medmad,
    motifseq_batch,
    motifseq_batch_f64,
    np_median,
    np_sum,
    rollmean_batch,
    rollmean_seg,
    segmenter_batch,
    segmenter_batch_pa,
    segmenter_batch_f64,
    zscale,
)
