"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/clipk.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "clipk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clipk_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from easynlp_b200 import build as B
    lib_path = B.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clipk.h but not exported by libclipk.so"
    lib.clipk_version.restype = ctypes.c_int
    assert lib.clipk_version() >= 100
    lib.clipk_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.clipk_last_error(), bytes)


def test_ctypes_binding_covers_the_header():
    from easynlp_b200 import _lib
    L = _lib.lib()
    for n in _declared():
        fn = getattr(L, n)
        if n not in ("clipk_last_error", "clipk_version", "clipk_launch_count"):
            assert fn.argtypes is not None, f"{n}: argtypes not declared in easynlp_b200/_lib.py"


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md 'what proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from easynlp_b200 import build as B
    if shutil.which("cuobjdump") is None:
        return
    sass = subprocess.run(["cuobjdump", "-sass", B.build()], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, mnem
    assert "HMMA.16816" not in sass      # no legacy mma.sync path


def test_product_path_has_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        return
    from easynlp_b200.appzoo.clip.model import CLIPApp
    with pytest.raises(RuntimeError):
        CLIPApp("/nonexistent-dir")
