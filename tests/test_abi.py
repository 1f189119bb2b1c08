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


def test_conv_bound_is_the_engine_threshold_rule():
    """The fixed-shape integer conv kernel fires a feature iff `a >= a_min` instead of computing
    `(float)clamp(a / scale, -127, 127) > threshold` (C++ truncating division; nnue_engine.cpp:93-103, 199).  The host-side
    bound (nnue_q_conv_bound, no GPU involved) is checked against that rule exhaustively around every boundary: thresholds
    on and between integers, negative (truncation toward zero makes the two sides of zero asymmetric), at and beyond the
    clamp ends, NaN; several scales."""
    import numpy as np
    from nnue_vision_b200 import _lib
    L = _lib.lib()
    thresholds = [-200.0, -127.5, -127.0, -126.99, -126.5, -2.0, -1.5, -1.0, -0.5, -0.0, 0.0, 0.25, 0.5, 1.0, 1.5, 63.0,
                  126.0, 126.5, 126.999, 127.0, 300.0, float("nan")]
    for scale in (1, 3, 64, 1000):
        a = np.arange(-130 * scale - 5, 130 * scale + 6, dtype=np.int64)
        q = np.where(a >= 0, a // scale, -((-a) // scale))          # truncation toward zero
        v = np.clip(q, -127, 127).astype(np.float32)
        for thr in thresholds:
            want = v > np.float32(thr)                               # NaN compares false
            a_min = ctypes.c_int32(0)
            mode = L.nnue_q_conv_bound(ctypes.c_float(thr), scale, ctypes.byref(a_min))
            assert mode in (0, 1, 2)
            got = (a >= a_min.value) if mode == 0 else np.full(a.shape, mode == 2)
            assert np.array_equal(got, want), (scale, thr, mode, a_min.value)
    assert L.nnue_q_conv_bound(ctypes.c_float(0.0), 0, ctypes.byref(ctypes.c_int32(0))) < 0


def test_scratch_size_queries_answer_without_a_gpu():
    """Pure host arithmetic: the sizes a caller must allocate before the calls that need them."""
    from nnue_vision_b200 import _lib
    L = _lib.lib()
    wide = _lib.make_shape(16384, 32, 32, 8, 10, 1024, 128, 32, 10, 3)   # the reference's "real" stack: has a side chain
    small = _lib.make_shape(16384, 32, 32, 8, 10, 64, 32, 8, 10, 3)      # config D: the one-kernel head, no side chain
    assert int(L.nnue_head_side_workspace_bytes(ctypes.byref(wide))) > 16384 * 128 * 4
    assert int(L.nnue_head_side_workspace_bytes(ctypes.byref(small))) == 0
    assert int(L.nnue_head_is_fused(ctypes.byref(small))) == 1 and int(L.nnue_head_uses_umma(ctypes.byref(wide))) == 1
    # counting sort in front of the indexed weight gradient: chunks x (F + 1) counters + two (F + 1) arrays
    assert int(L.nnue_ft_sort_pairs_workspace_bytes(9, 23, 800)) >= (1 + 2) * 801 * 4
    assert int(L.nnue_ft_sort_pairs_workspace_bytes(0, 23, 800)) == 0
    assert int(L.nnue_allreduce_ll_max_floats()) == 65536 and int(L.nnue_allreduce_recv_floats(8, 256)) == 2 * 8 * 256 * 2
