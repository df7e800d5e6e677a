"""The C-ABI library loads and exports every symbol include/sqk.h declares; structs have the layout
the header promises; without a GPU the product fails loudly (no CPU fallback, no oracle import)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sqk.h")).read()
    return sorted(set(re.findall(r"SQK_API\s+[\w\s\*]+?\b(sqk_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("sqk_motifseq", "sqk_segmenter", "sqk_ctx_create", "sqk_ctx_destroy", "sqk_last_error", "sqk_host_alloc"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from squigglekit_b200 import _cabi
    lib = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"libsqk.so does not export {n}"
    assert sorted(_cabi.EXPORTS) == names
    assert lib.sqk_version() == 100


def test_struct_layouts_match_header():
    from squigglekit_b200 import _cabi
    assert C.sizeof(_cabi.MotifParams) == 16
    assert C.sizeof(_cabi.SegParams) == 48
    assert C.sizeof(_cabi.AdapterParams) == 48 and _cabi.AdapterParams.std_scale.offset == 32
    assert _cabi.HIT_DTYPE.itemsize == 16 and _cabi.HIT_DTYPE.fields["dist"][1] == 8
    assert C.sizeof(_cabi.Timing) == 8 * _cabi.K_COUNT + 8 * _cabi.K_COUNT and _cabi.K_COUNT == 5   # enum sqk_kernel_id


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sqk.h"\nint main(void){ sqk_hit h; sqk_seg_params p; sqk_adapter_params a = SQK_ADAPTER_DEFAULTS; (void)h; (void)p;\n'
                   '  return sizeof(sqk_hit) == 16 && sizeof(sqk_adapter_params) == 48 && a.t_end == 5000 && a.lim_hi == 1200 && SQK_K_COUNT == 5 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import squigglekit_b200 as sqk
    with pytest.raises(sqk.SqkError) as ei:
        sqk.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """squigglekit_b200 must not reach into oracle/ (import, ctypes load or subprocess)."""
    pkg = os.path.join(ROOT, "squigglekit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "libsqk_oracle" not in text, f
    code = "import sys; import squigglekit_b200, squigglekit_b200.core, squigglekit_b200.models, squigglekit_b200.dist; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
