"""Golden rows for the SquigglePull.py drop-in.  Runs ONLY in the build container (needs /root/reference).

The reference's own ``extract_f5_all`` and ``print_data`` (SquigglePull.py:130-253) run UNMODIFIED; the one thing replaced is
``h5py`` (not installed): a stand-in with the handful of calls those two functions make (``File`` as a context manager,
``keys()``, ``[...]`` paths, ``.attrs[...]``, ``dataset[()]``), answered by this repo's HDF5 reader
(squigglekit_b200/fast5.py, itself pinned on the example files' known contents).  So the conversion to pA, the rounding,
``range`` through ``"{0:.2f}"``, the ``str()`` of every field and the column order are the reference's.

Inputs: example/test.fast5 (single-read) and the first files of example/example_fast5s.tar.  Rows are long (one field per
sample): the fixture keeps the first 12 and last 3 fields of each row, its field count and a CRC of the whole line.

usage:  python tests/golden/make_squigglepull_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tarfile
import tempfile
import types
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from squigglekit_b200 import fast5  # noqa: E402


class _Attrs:
    def __init__(self, d):
        self.d = d

    def __getitem__(self, k):
        v = self.d[k]
        return bytes(v) if isinstance(v, (bytes, bytearray)) else v


class _Obj:
    """h5py Group / Dataset as far as SquigglePull.py uses them."""

    def __init__(self, node):
        self.node = node

    def keys(self):
        return self.node.keys()

    @property
    def attrs(self):
        return _Attrs(self.node.attrs)

    def __getitem__(self, k):
        if k == ():
            return self.node.read()
        return _Obj(self.node[k.strip("/")])


class _File(_Obj):
    def __init__(self, filename, mode="r"):
        f = fast5.Fast5File(filename)
        super().__init__(f)
        self.f = f

    def keys(self):
        return self.f.keys()

    def __getitem__(self, k):
        return _Obj(self.f[k.strip("/")])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def summarise(line: str):
    f = line.split("\t")
    return {"n_fields": len(f), "head": f[:12], "tail": f[-3:], "crc32": zlib.crc32(line.encode())}


def main():
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sp = refload.load("SquigglePull")
    sp.h5py = h5                                            # the module imported refload's inert stub: swap in the stand-in
    files = [os.path.join(refload.REFERENCE_ROOT, "example", "test.fast5")]
    tmp = tempfile.mkdtemp(prefix="sqk_sp_")
    with tarfile.open(os.path.join(refload.REFERENCE_ROOT, "example", "example_fast5s.tar")) as tf:
        members = sorted(m.name for m in tf.getmembers() if m.name.endswith(".fast5"))[:3]
        for m in members:
            tf.extract(m, tmp, filter="data")
            files.append(os.path.join(tmp, m))
    out = {"files": [os.path.basename(p) for p in files], "tar_members": members, "cases": []}
    for raw_signal in (True, False):
        for extra_info in (False, True):
            args = types.SimpleNamespace(type="auto", verbose=False, raw_signal=raw_signal, extra_info=extra_info)
            rows = []
            for p in files:
                data, multi = sp.extract_f5_all(p, args)
                assert data and not multi, p
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    sp.print_data(data, args, os.path.basename(p))
                line = buf.getvalue()
                assert line.endswith("\n") and line.count("\n") == 1
                rows.append(summarise(line[:-1]))
            out["cases"].append({"raw_signal": raw_signal, "extra_info": extra_info, "rows": rows})
    with open(os.path.join(HERE, "squigglepull_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote squigglepull_golden.json:", [(c["raw_signal"], c["extra_info"], [r["n_fields"] for r in c["rows"]]) for c in out["cases"]])


if __name__ == "__main__":
    main()
