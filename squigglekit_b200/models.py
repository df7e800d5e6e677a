"""Motif model files (the `.model` format MotifSeq consumes).  Host-side, once per run.

* ``read_synth_model``  scrappie CLI output: ``#name`` line, optional ``pos base current sd dwell``
  header, then one row per base; each row contributes ``current`` repeated ``round(dwell)``
  times (MotifSeq.py:354-379).  This is the format of example/CATCTATCCAGGGTTAAATT.model.
* ``read_bait_model``   one motif per line: ``name <tab> kmer_length <tab> ? <tab> sig...`` with the
  signal from column 3 (MotifSeq.py:408-428).  The reference forgets to fill ``m_order`` and
  ``L_list`` there, so its search loop (MotifSeq.py:436) never runs; here they are
  filled -- a disclosed fix-forward.
* ``read_model``        sniffs which of the two a file is (the reference's ``-m`` hard-wires the
  broken bait reader, MotifSeq.py:155-157, and raises IndexError on its own example file).

All return ``(model: dict name -> float64 array, m_order: list of names, L: list of k-mer lengths)``
exactly like the reference functions.
"""
from __future__ import annotations

import gzip

import numpy as np


def _open(path):
    return gzip.open(path, "rt") if str(path).endswith(".gz") else open(path, "rt")


def read_synth_model(filename):
    model, m_order, L_list = {}, [], []
    L = 0
    name = None
    with _open(filename) as fh:
        for line in fh:
            line = line.strip("\n")
            if not line:
                continue
            if line[0] == "#":
                if L != 0:
                    L_list.append(L)
                L = 0
                name = line[1:]
                model[name] = []
                m_order.append(name)
            elif line[:3] == "pos":
                continue
            else:
                if name is None:
                    raise ValueError(f"{filename}: data row before any '#name' line")
                L += 1
                f = line.split()
                model[name] = model[name] + [float(f[2])] * int(round(float(f[4])))
        L_list.append(L)
    return {k: np.asarray(v, dtype=np.float64) for k, v in model.items()}, m_order, L_list


def read_bait_model(filename):
    model, m_order, L_list = {}, [], []
    with _open(filename) as fh:
        for line in fh:
            f = line.strip("\n").split("\t")
            if len(f) < 4:
                continue
            name = f[0]
            model[name] = np.array([float(v) for v in f[3:]], dtype=np.float64)
            m_order.append(name)
            L_list.append(int(f[1]))
    return model, m_order, L_list


def read_model(filename):
    with _open(filename) as fh:
        first = fh.readline()
    return read_synth_model(filename) if first.startswith("#") else read_bait_model(filename)
