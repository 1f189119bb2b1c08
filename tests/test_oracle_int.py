"""The integer-path oracle is pinned before anything is compared against it.

(a) golden vectors: logits + density produced by the reference's serialize.py +
    C++ engine in the build container (tests/golden/make_golden.py);
(b) the reference engine itself (oracle/_ref), when that build is present, on
    randomized .nnue files that use the format's full integer ranges.
Bit-exact in both cases.
"""
import numpy as np
import pytest

from util import GOLDEN, GOLDEN_CASES, INT_ARCHS, load_golden, write_random_nnue


def chw_bytes_as_hwc(images_chw):
    """evaluate.py:154-168 dumps CHW floats and the engine reads them as HWC: a reinterpretation."""
    B, C, H, W = images_chw.shape
    return np.ascontiguousarray(images_chw).reshape(B, H, W, C)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_oracle_matches_reference_golden(oracle_built, name):
    rec = load_golden(name)
    n = int(rec["int.n_images"])
    orc = oracle_built.IntOracle(GOLDEN / f"{name}.nnue")
    logits, dens = orc.eval_batch(chw_bytes_as_hwc(rec["images"][:n]))
    assert logits.dtype == np.float32
    np.testing.assert_array_equal(logits, rec["int.logits"])
    np.testing.assert_array_equal(dens, rec["int.density"])


def test_c_oracle_header_fields(oracle_built):
    rec = load_golden("default_cfg")
    orc = oracle_built.IntOracle(GOLDEN / "default_cfg.nnue")
    c = rec["cfg"]
    assert (orc.F, orc.L1, orc.L2, orc.L3, orc.NC, orc.OC, orc.G, orc.n_buckets) == (
        c["grid"] ** 2 * c["C"], c["L1"], c["L2"], c["L3"], c["NC"], c["C"], c["grid"], 1)
    # engine stride rule ceil((H-1)/(G-1)) (nnue_engine.cpp:710-718), not the Python floor rule
    assert orc.stride(32) == 4 and orc.stride(96) == 11 and orc.stride(10) == 1


def test_c_oracle_rejects_malformed(oracle_built, tmp_path):
    good = (GOLDEN / "parity_small.nnue").read_bytes()
    for bad in (b"XXXX" + good[4:], good[:4] + b"\x03\x00\x00\x00" + good[8:], good[:200], good[:-3]):
        p = tmp_path / "bad.nnue"
        p.write_bytes(bad)
        with pytest.raises(ValueError):
            oracle_built.IntOracle(p)


@pytest.mark.parametrize("arch", INT_ARCHS[:7], ids=lambda a: "x".join(map(str, a)))
@pytest.mark.parametrize("wild", [False, True])
def test_c_oracle_matches_reference_engine(oracle_built, tmp_path, arch, wild):
    if not oracle_built.RefEngine.available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    G, C, L1, L2, L3, NC, H = arch
    rng = np.random.default_rng(hash(arch) % (2**32) + wild)
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=wild)
    imgs = (rng.standard_normal((6, H, H, 3)) * (3.0 if wild else 1.0)).astype(np.float32)
    imgs[0] = 0.0
    ref = oracle_built.RefEngine(path)
    orc = oracle_built.IntOracle(path)
    rl, rd = ref.eval_batch(imgs, threads=2)
    ol, od = orc.eval_batch(imgs, threads=2)
    np.testing.assert_array_equal(ol, rl)
    np.testing.assert_array_equal(od, rd)
    assert (ref.F, ref.L1, ref.L2, ref.L3, ref.OC, ref.G) == (orc.F, orc.L1, orc.L2, orc.L3, orc.OC, orc.G)


def test_legacy_score_restatement_is_pinned_to_the_reference_engine(oracle_built, tmp_path):
    """evaluate_incremental's single score (nnue_engine.cpp:739-787, LayerStack::forward :382-478): the numpy
    restatement equals the reference engine itself (oracle/_ref, an AVX2 build: the eight-fold bias of
    simd_avx2.cpp:119 is part of what it computes) on random files, full int16 rows included."""
    from util import write_random_nnue
    from nnue_vision_b200 import serialize
    if not oracle_built.RefEngine.available():
        pytest.skip("oracle/_ref not built (no reference checkout)")
    rng = np.random.default_rng(3)
    for (G, C, L1, L2, L3, NC) in [(8, 4, 64, 4, 8, 10), (10, 8, 64, 32, 8, 10), (5, 3, 30, 5, 7, 3), (6, 16, 128, 16, 32, 1000)]:
        for wild in (False, True):
            p = tmp_path / "m.nnue"
            write_random_nnue(p, rng, G, C, L1, L2, L3, NC, wild=wild)
            q = serialize.read_nnue(p)
            ref = oracle_built.RefEngine(p)
            F = G * G * C
            for _ in range(6):
                feats = rng.choice(F, size=int(rng.integers(0, min(F, 150))), replace=False).tolist()
                ref.mark_dirty()
                assert ref.eval_incremental(feats) == oracle_built.legacy_score(q, feats)
            # incremental walk: the engine diffs against its previous call; the restatement recomputes from scratch
            feats = rng.choice(F, size=min(F, 40), replace=False).tolist()
            ref.mark_dirty()
            for _ in range(5):
                assert ref.eval_incremental(feats) == oracle_built.legacy_score(q, feats)
                drop = set(rng.choice(feats, size=min(len(feats), 5), replace=False).tolist())
                feats = [f for f in feats if f not in drop] + [int(f) for f in rng.choice(F, size=5) if f not in feats]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_legacy_score_restatement_matches_reference_engine_golden(name):
    """The same pin without the compiled engine at hand: committed scores the reference engine produced
    (tests/golden/incremental.npz, generator make_incremental_golden.py)."""
    from oracle import int_oracle
    from util import load_incremental_golden
    from nnue_vision_b200 import serialize
    q = serialize.read_nnue(GOLDEN / f"{name}.nnue")
    g = load_incremental_golden(name)
    for feats, want in zip(g["fresh"] + g["walk"], list(g["fresh_score"]) + list(g["walk_score"])):
        assert np.float32(int_oracle.legacy_score(q, feats)) == want
