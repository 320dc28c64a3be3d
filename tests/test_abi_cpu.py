"""CPU: the C-ABI library builds, loads and exports every symbol include/ivlm_b200.h declares
(no compute calls without a GPU) and refuses to run without a CUDA device instead of falling back."""
import ctypes

import pytest


def test_library_builds_and_exports_declared_symbols():
    from interactvlm_b200 import _lib, build
    path = build.build()
    assert path.exists()
    names = _lib.declared_symbols()
    assert len(names) >= 30 and "ivlm_gemm_bf16" in names and "ivlm_lift" in names
    lib = ctypes.CDLL(str(path))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.ivlm_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interactvlm_b200.ops import Context
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Context(0)


def test_create_reports_error_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interactvlm_b200 import _lib
    lib = _lib.lib()
    h = ctypes.c_void_p()
    assert lib.ivlm_create(ctypes.byref(h), 0) != 0
    assert len(lib.ivlm_last_error()) > 0
