"""CPU: the C-ABI library loads here (no GPU) and exports every symbol include/adpst.h declares; the ctypes table
lists exactly those symbols; compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes
import importlib
import os
import re

import pytest

from conftest import PKG_NAME, ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "adpst.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adpst_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    b = importlib.import_module(PKG_NAME + ".build")
    b.build()
    return importlib.import_module(PKG_NAME + "._lib")


def test_header_and_ctypes_table_agree(lib):
    assert _declared() == sorted(lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    L = ctypes.CDLL(lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    assert lib.missing_symbols() == []
    assert lib.lib().adpst_version() >= 100


def test_argument_validation_without_a_gpu(lib):
    """Pure host-side checks (no kernel launch): bad arguments return error codes with a message."""
    L = lib.lib()
    h = ctypes.c_void_p()
    rc = L.adpst_laplacian_create(7, 4, 4, 1, 1e-7, None, 0, 0, None, ctypes.byref(h))
    assert rc == lib.ERR_INVALID and b"mode" in L.adpst_last_error()
    rc = L.adpst_laplacian_create(lib.LAP_V2, 4, 4, 9, 1e-7, ctypes.c_void_p(16), 0, 0, None, ctypes.byref(h))
    assert rc == lib.ERR_UNSUPPORTED and b"radius" in L.adpst_last_error()
    hh, ww, cc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.adpst_vgg_conv_shape(12, 1024, 1024, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc)) == 0
    assert (hh.value, ww.value, cc.value) == (64, 64, 512)
    assert L.adpst_vgg_conv_shape(9, 96, 138, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc)) == 0
    assert (hh.value, ww.value, cc.value) == (12, 17, 512)
    assert L.adpst_vgg_pool_shape(0, 37, 50, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc)) == 0
    assert (hh.value, ww.value, cc.value) == (18, 25, 64)
    assert L.adpst_vgg_conv_shape(13, 8, 8, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc)) == lib.ERR_INVALID
    assert L.adpst_gram_workspace_bytes(1024 * 1024, 64, 8) > 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    v2 = importlib.import_module(PKG_NAME + ".components.matting_v2")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v2.MattingLaplacian(torch.zeros(4, 4, 3, dtype=torch.float64))
    vgg = importlib.import_module(PKG_NAME + ".components.VGG19.model")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vgg.StyleContentModel(["block4_conv2"], ["block1_conv1"], weights={})


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, PKG_NAME)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
