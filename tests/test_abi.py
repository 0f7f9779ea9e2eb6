"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/svs_b200.h
declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "svs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from svs_unet_pytorch_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in svs_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_version_and_error_string():
    from svs_unet_pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.svs_version() == _lib.ABI_VERSION == 3
    assert isinstance(lib.svs_last_error(), (bytes, type(None)))


def test_invalid_arguments_are_errors_not_crashes():
    from svs_unet_pytorch_b200 import _lib
    lib = _lib.load()
    rc = lib.svs_stft_mag_phase(None, None, None, 1, 1, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.svs_last_error()
    assert lib.svs_unet_workspace_bytes(None, 4) == 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    from svs_unet_pytorch_b200 import _lib
    _lib.load()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass        # tcgen05.mma
    assert "UTMALDG" in sass        # TMA tensor loads
    assert "LDTM" in sass           # tcgen05.ld
    assert "sm_100a" in sass
