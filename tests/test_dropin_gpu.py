"""Literal drop-in: the reference's OWN get_region_multi (MotifSeq.py:431-456), imported unmodified from
/root/reference, runs with ``mlpy.dtw_subsequence`` replaced by the libsqk shim and must print the golden rows.
Needs the reference tree, so it only runs where that exists AND a GPU is visible; elsewhere the same function
restated in squigglekit_b200.cli_motifseq.format_row is checked against the golden rows instead."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from oracle import numpy_ref, refload
from squigglekit_b200 import mlpy_shim

pytestmark = pytest.mark.gpu


def test_shim_matches_oracle_dtw(ctx):
    import oracle
    rng = np.random.default_rng(5)
    shim = mlpy_shim.make_dtw_subsequence(ctx)
    for n, m in ((80, 3000), (163, 5000), (7, 50), (1, 9), (33, 33)):
        x, y = rng.standard_normal(n), rng.standard_normal(m) * 1.3
        dist, cost, path = shim(x, y)
        d, s, e = oracle.dtw_subsequence_rolling(x, y)
        assert (dist, path[1][0], path[1][-1]) == (d, s, e)
        with pytest.raises(NotImplementedError):
            cost[-1, ]


@pytest.mark.skipif(not refload.available(), reason="reference tree not present on this box")
def test_reference_get_region_multi_runs_on_libsqk(ctx, golden_dir):
    ms = refload.load("MotifSeq", dtw_subsequence=mlpy_shim.make_dtw_subsequence(ctx))
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    model = {str(ex["name"]): ex["model"]}
    for scale in ("zscale", "medmad"):
        args = refload.Args(scale=scale)
        sig = ms.scale_outliers(np.array(ex["raw"], dtype=int), args)
        sig = numpy_ref.zscale_np(sig) if scale == "zscale" else numpy_ref.medmad_np(sig)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ms.get_region_multi(args, sig, model, [str(ex["name"])], "test.fast5", str(ex["read_id"]), args.slope,
                                args.intercept, args.std_const, [20])
        assert buf.getvalue() == exp["tsv"][scale]
