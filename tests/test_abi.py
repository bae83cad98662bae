"""CPU: the C-ABI library loads and exports exactly what include/difformer_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "difformer_b200.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^DIF_API\s+[\w\s\*]+?\b(dif_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 20
    for must in ("dif_simple_reduce", "dif_simple_apply", "dif_simple_bwd_reduce", "dif_simple_bwd_apply",
                 "dif_sigmoid_fwd", "dif_sigmoid_bwd", "dif_csr_build", "dif_gcn_spmm", "dif_segmented_simple_fwd"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from difformer_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_ctypes_signatures_cover_the_header():
    from difformer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_pure_host_queries():
    from difformer_b200 import _lib
    lib = _lib.lib
    assert lib.dif_version() >= 100
    # partials = [S | z | u | sum q^2 | sum k^2]: 16 898 floats at H=4, D=64
    assert lib.dif_simple_partials_len(4, 4, 64, 64) == 4 * 64 * 64 + 4 * 64 + 4 * 64 + 2 == 16898
    assert lib.dif_simple_partials_len(4, 1, 64, 64) == 4 * 64 * 64 + 4 * 64 + 64 + 2
    assert lib.dif_simple_bwd_partials_len(4, 64, 64) == 4 * 64 * 64 + 4 * 64 + 4 * 64 + 2
    assert lib.dif_csr_workspace_bytes(1000, 5000) > 0
    assert lib.dif_csr_workspace_bytes(1 << 31, 10) == -1     # int32 index range
    assert lib.dif_sigmoid_bwd_workspace_bytes(100, 100, 2, 2, 64, 64) >= 100 * 2 * 4
    # 'sigmoid' forward: tcgen05 shapes (M == D == 64) also need the bf16 hi|lo operand images of K and V, 32 KB per
    # 128-key tile and head (sigmoid_sm100.cu); other shapes only the key-split partials
    tiles = (10000 + 127) // 128
    assert lib.dif_sigmoid_fwd_workspace_bytes(10000, 10000, 1, 1, 64, 64) >= tiles * 2 * 32768
    assert lib.dif_sigmoid_fwd_workspace_bytes(10000, 10000, 4, 1, 64, 64) >= tiles * 5 * 32768
    assert lib.dif_sigmoid_fwd_workspace_bytes(100, 100, 1, 1, 32, 32) < 32768
    # impl selector: AUTO / GENERIC / TCGEN05 accepted, anything else is a bad argument (and says why)
    for impl in (_lib.DIF_IMPL_GENERIC, _lib.DIF_IMPL_TCGEN05, _lib.DIF_IMPL_AUTO):
        assert lib.dif_sigmoid_set_impl(impl) == 0
    assert lib.dif_sigmoid_set_impl(7) == -1 and b"unknown impl" in lib.dif_last_error()
    # exchange buffer = [128-byte header | LL words u64 [2 slots][16 source ranks][len padded to 64]]
    n = 16898
    slot = (n + 63) // 64 * 64
    assert lib.dif_comm_buffer_bytes(n) == 128 + 2 * 16 * slot * 8
    # pass-1 workspace: one record per CTA + ready flags + the generation word
    assert lib.dif_simple_workspace_bytes(132534, 4, 4, 64, 64) >= 148 * 16898 * 4 + 149 * 8


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    from difformer_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU/PyTorch fallback"):
        _lib._load()
    importlib.reload(_lib)
