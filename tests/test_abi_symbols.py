"""CPU: libgsx.so loads without a GPU and exports every function include/gsx.h declares."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    src = (ROOT / "include" / "gsx.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(gsx_lib):
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(gsx_lib, name), f"{name} declared in include/gsx.h but not exported by libgsx.so"


def test_binding_covers_header(gsx_lib):
    from gsx import _abi
    assert set(declared_functions()) == set(_abi._SIGS), "gsx/_abi.py and include/gsx.h disagree"


def test_no_torch_types_in_abi():
    src = (ROOT / "include" / "gsx.h").read_text()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def test_host_helpers_run_without_gpu(gsx_lib):
    assert gsx_lib.gsx_version() >= 100
    assert isinstance(gsx_lib.gsx_last_error(), bytes)
    assert gsx_lib.gsx_mean_std_workspace_bytes(10_000_000) > 0


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under 3dgsconverter_b200/ may reference it."""
    for p in (ROOT / "3dgsconverter_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") and p.is_file():
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "liborc" not in txt, p
