"""Repository contract checks: the product never touches the oracle or the reference checkout."""
import re

from util import ROOT

PKG = ROOT / "nnue-vision_b200"


def product_sources():
    return [p for p in PKG.rglob("*") if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h")] + [ROOT / "nnue_vision_b200.py"]


def test_product_never_imports_the_oracle():
    for p in product_sources():
        text = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), p
        assert "oracle/" not in text and "liboracle" not in text and "libnnue_ref" not in text, p


def test_nothing_reads_the_reference_checkout_at_run_time():
    for p in product_sources() + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]:
        if p.exists():
            assert "/root/reference" not in p.read_text().replace("/root/reference/", "REFDOC/"), p


def test_no_compatibility_layers_in_the_product():
    for p in product_sources():
        text = p.read_text()
        assert "import triton" not in text and "torch.compile" not in text and "tilelang" not in text, p


def test_required_layout_exists():
    for rel in ("include/nnue_b200.h", "oracle/nnue_int_oracle.c", "oracle/float_oracle.py", "oracle/Makefile",
                "tests/golden/make_golden.py", "nnue-vision_b200/csrc/ft.cu", "nnue-vision_b200/nnue.py"):
        assert (ROOT / rel).exists(), rel
