"""The drop-in boundary without a GPU: the C-ABI shared library loads, exports every symbol that
include/trajopt_b200.h declares (and nothing is bound from Python that the header does not
declare), validates its arguments before touching CUDA, and the Python `trajopt_params` mirror has
the C struct's layout.  No compute calls here (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "trajopt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trajopt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from trajectory_optimization_matrix_lie_groups_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (trajopt_[a-z0-9_]+)\b", nm))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    # the ctypes table binds exactly the declared functions
    assert sorted(_lib.SYMBOLS) == names


def test_params_struct_layout_matches_header(tmp_path):
    """Compile a 10-line C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from trajectory_optimization_matrix_lie_groups_b200 import _lib
    fields = [f[0] for f in _lib.Params._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"trajopt_b200.h\"\nint main(void){\n"
    prog += 'printf("%zu\\n", sizeof(trajopt_params));\n'
    for f in fields:
        prog += f'printf("%zu\\n", offsetof(trajopt_params, {f}));\n'
    prog += "return 0;}\n"
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == C.sizeof(_lib.Params)
    assert vals[1:] == [getattr(_lib.Params, f).offset for f in fields]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "trajopt_b200.h"\nint main(void){return TRAJOPT_E_INVALID == -1 ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_argument_validation_needs_no_gpu():
    from trajectory_optimization_matrix_lie_groups_b200 import _lib
    lib = _lib.lib
    assert lib.trajopt_version() >= 100
    h = C.c_void_p()
    assert lib.trajopt_create(7, 0, 10, 1, 0, C.byref(h)) == -1 and b"kind" in lib.trajopt_last_error()
    assert lib.trajopt_create(1, 9, 10, 1, 0, C.byref(h)) == -1 and b"method" in lib.trajopt_last_error()
    assert lib.trajopt_create(1, 1, 0, 1, 0, C.byref(h)) == -1
    assert lib.trajopt_create(1, 1, 10, 0, 0, C.byref(h)) == -1
    assert lib.trajopt_create(0, 2, 10, 1, 0, C.byref(h)) == -1 and b"SE3" in lib.trajopt_last_error()   # no AL for SO3
    assert lib.trajopt_create(1, 1, 10, 1, 0, None) == -1
    assert h.value is None
    for fn, args in (("trajopt_begin", (None, None, None, 0, None)), ("trajopt_iterate", (None, 1, None, None)),
                     ("trajopt_set_params", (None, None)), ("trajopt_set_reference", (None, None, None)),
                     ("trajopt_export", (None,) * 9), ("trajopt_solve_host", (None,) * 3 + (0,) + (None,) * 8),
                     ("trajopt_set_sweep", (None, 6, 1)), ("trajopt_set_line_search_batch", (None, 256)),
                     ("trajopt_set_compaction", (None, 1024, 4))):
        assert getattr(lib, fn)(*args) == -1, fn
    assert lib.trajopt_destroy(None) == 0
    assert lib.trajopt_launch_count(1) >= 0


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the product path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from trajectory_optimization_matrix_lie_groups_b200 import BatchSolver, TrajoptError, _lib
    with pytest.raises(TrajoptError):
        BatchSolver("se3", "ms", 10, 4)
    h = C.c_void_p()
    rc = _lib.lib.trajopt_create(1, 1, 10, 4, 0, C.byref(h))
    assert rc in (-1, -2) and h.value is None


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "trajectory_optimization_matrix_lie_groups_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
