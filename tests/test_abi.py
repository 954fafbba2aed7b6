"""The C-ABI boundary: libb200cls.so loads on a CPU-only box and exports every symbol include/b200cls.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200cls.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from deeplearning_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200cls.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in deeplearning_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_compute_calls_without_gpu_but_metadata_works():
    from deeplearning_b200 import _lib

    lib = _lib.load()
    assert lib.b200_abi_version() >= 1
    # pure host-side planners are usable without a device
    # BN statistics rows: (persistent CTAs / channel blocks) x 4 TMEM quadrants (x2 when the warp pair alternates tiles)
    assert lib.b200_conv2d_fwd_stats_rows(256, 56, 56, 64, 3, 1) == 148 * 8
    assert lib.b200_conv2d_fwd_stats_rows(256, 56, 56, 256, 1, 1) == 148 * 4
    assert lib.b200_conv2d_fwd_stats_rows(256, 7, 7, 2048, 1, 1) == 144 // 8 * 4
    assert lib.b200_conv2d_fwd_stats_rows(1, 8, 8, 64, 1, 1) == 8
    assert lib.b200_conv2d_wgrad_workspace_bytes(256, 56, 56, 64, 64, 3, 1) > 0
    assert lib.b200_bn_bwd_blocks(256 * 56 * 56, 64) > 0
    assert lib.b200_bn_bwd_blocks(100, 96) == -1  # unsupported channel count is reported, not guessed


def test_error_convention():
    from deeplearning_b200 import _lib

    lib = _lib.load()
    rc = lib.b200_conv2d_fwd(None, None, None, 1, 8, 8, 64, 64, 5, 1, None, None, 0, None, None, 0, None)
    assert rc == -1  # B200_EINVAL: 5x5 is not supported, and nothing was launched
    assert "ksize" in _lib.last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from deeplearning_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200cls.so")
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA extension is missing")
