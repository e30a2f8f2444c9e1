"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made here (there is no GPU on the build box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "visinger_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from visinger_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in the header but not exported"


def test_abi_version_and_error_string():
    from visinger_b200 import _lib
    L = _lib.lib()
    assert L.vsg_abi_version() == 1
    assert isinstance(L.vsg_last_error(), bytes)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The bf16 path must really be tcgen05 + TMA (UTC*MMA / UTMALDG in SASS), not mma.sync."""
    import shutil
    import subprocess
    from visinger_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma (UTC*MMA) in SASS"
    assert "UTMALDG" in sass, "no TMA loads (UTMALDG) in SASS"
    assert "LDTM" in sass, "no tcgen05.ld (LDTM) in SASS"
    assert not re.search(r"(?<!UTC)HMMA", sass), "legacy mma.sync path present"


def test_pack_create_rejects_bad_arguments_without_gpu():
    from visinger_b200 import _lib
    L = _lib.lib()
    out = ctypes.c_void_p()
    rc = L.vsg_pack_create(None, None, 0, b"", b"", 0, ctypes.byref(out))
    assert rc != 0 and not out.value
    assert L.vsg_last_error()
    assert L.vsg_workspace_bytes(None, 1, 1, 0) == 0
    assert L.vsg_hop_size(None) == 0
