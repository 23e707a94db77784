"""The C-ABI library loads on a CPU-only box and exports every symbol include/tbg.h declares
(no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "tbg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tbg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from textboxgan_b200 import build, lib

    path = build.build()
    assert path.exists()
    cdll = ctypes.CDLL(str(path))
    declared = _declared_symbols()
    assert "tbg_conv2d_igemm" in declared and "tbg_upfirdn2d" in declared
    for sym in declared:
        assert hasattr(cdll, sym), f"{sym} declared in include/tbg.h but not exported"
    # the Python binding covers exactly the declared symbols
    assert sorted(lib.exported_symbols()) == declared
    handle = lib.load()
    assert handle.tbg_version() >= 1
    assert handle.tbg_last_error() is not None


def test_invalid_arguments_return_status_not_exceptions():
    """Argument validation happens before any CUDA call, so it is checkable without a GPU
    (mirrors the OP_REQUIRES checks of upfirdn_2d.cu:241-256)."""
    from textboxgan_b200 import lib

    h = lib.load()
    st = h.tbg_conv2d_igemm(None, None)
    assert st == -1 and b"null" in h.tbg_last_error()
    a = lib.ConvArgs(x=1 << 20, w=1 << 20, out=1 << 20, B=1, H=4, W=4, Cin=48, Ho=4, Wo=4, n_total=64, cout=64,
                     taps_h=3, taps_w=3, pad_h=1, pad_w=1, stride_h=1, stride_w=1)
    st = h.tbg_conv2d_igemm(ctypes.byref(a), None)
    assert st == -1 and b"Cin" in h.tbg_last_error()
    st = h.tbg_upfirdn2d(1 << 20, 1 << 20, 1 << 20, 0, 1, 4, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, None)
    assert st == -1 and b"upx" in h.tbg_last_error()


def test_product_does_not_import_the_oracle():
    for f in (ROOT / "textboxgan_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
