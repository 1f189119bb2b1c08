"""The C-ABI library loads on a CPU-only box and exports every symbol include/nnue_b200.h declares
(no compute calls here: there is no GPU in the build container)."""
import ctypes
import re

from util import ROOT

HEADER = (ROOT / "include" / "nnue_b200.h").read_text()


def declared_functions():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(nnue_[a-z0-9_]+)\s*\(", code)))


def test_header_declares_the_scope_table_entry_points():
    names = declared_functions()
    for required in ("nnue_extract_fwd", "nnue_extract_bwd", "nnue_ft_fwd", "nnue_ft_bwd_dw", "nnue_ft_bwd_dval",
                     "nnue_head_fwd", "nnue_head_bwd", "nnue_q_load", "nnue_q_infer", "nnue_workspace_bytes"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from nnue_vision_b200 import _lib
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in nnue_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared_functions()), "python binding table out of sync with the header"
    assert _lib.lib().nnue_b200_abi_version() == 4
    assert _lib.lib().nnue_error_string(-5) == b"malformed .nnue file"


def test_no_torch_types_cross_the_boundary():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    assert "torch" not in code and "at::" not in code and "std::" not in code
    assert 'extern "C"' in HEADER


def test_shape_derivation_matches_reference_rules():
    from nnue_vision_b200 import _lib
    # config D: 32x32 into grid 10 -> stride 3 -> 11x11 raster, 968 positions over an 800-row table
    s = _lib.make_shape(16384, 32, 32, 8, 10, 64, 32, 8, 10, 3)
    assert (s.F, s.Gh, s.Gw, s.P, s.CW, s.NW, s.PP, s.BW) == (800, 11, 11, 968, 4, 32, 1024, 512)
    assert _lib.workspace_bytes(s) > 0
    s = _lib.make_shape(7, 224, 224, 64, 32, 1024, 128, 32, 1000, 7)
    assert (s.F, s.Gh, s.P, s.CW) == (65536, 32, 65536, 32)
    import pytest
    with pytest.raises(_lib.NnueError):
        _lib.make_shape(0, 32, 32, 8, 10, 64, 32, 8, 10, 3)
    with pytest.raises(_lib.NnueError):
        _lib.make_shape(4, 32, 32, 8, 10, 63, 32, 8, 10, 3)  # odd L1: torch.split would give 3 chunks


def test_truncated_nnue_with_huge_header_sizes_is_refused_without_allocating():
    """A header that promises gigabytes of payload the file does not hold answers NNUE_ERR_FORMAT (load_model -> false,
    nnue_engine.cpp:544-657) before anything is allocated for it, and nothing unwinds through the C boundary."""
    import struct
    from nnue_vision_b200 import _lib
    good = (ROOT / "tests" / "golden" / "parity_small.nnue").read_bytes()
    F, L1, L2, L3 = struct.unpack_from("<4I", good, 8)
    cases = []
    huge_table = bytearray(good)                      # F x L1 = 2^31 int16 elements, file ends long before
    struct.pack_into("<2I", huge_table, 8, 1 << 20, 1 << 11)
    cases.append(bytes(huge_table))
    huge_l2 = bytearray(good)                         # an absurd L2
    struct.pack_into("<I", huge_l2, 16, 0xFFFFFFF0)
    cases.append(bytes(huge_l2))
    cases.append(good[: len(good) // 2])              # plain truncation
    for blob in cases:
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = _lib.lib().nnue_q_load_memory(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(h))
        assert rc in (-5, -2) and not h.value, rc


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    from nnue_vision_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NnueError, match="no CPU or PyTorch fallback"):
        _lib.lib()
