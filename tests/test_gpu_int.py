"""-m gpu: the quantized integer inference path must be BIT-EXACT against the reference's
serialize.py + C++ engine (golden vectors), against the C oracle, and -- where oracle/_ref was
built -- against the reference engine itself, through both the device and the host entry points."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util import GOLDEN, GOLDEN_CASES, INT_ARCHS, load_golden, write_random_nnue


def _engine():
    from nnue_vision_b200 import engine
    return engine


def chw_bytes_as_hwc(images_chw):
    B, C, H, W = images_chw.shape
    return np.ascontiguousarray(images_chw).reshape(B, H, W, C)


def gpu_eval(ev, imgs_bhwc):
    logits, dens = ev.evaluate_logits(torch.as_tensor(imgs_bhwc).cuda().contiguous())
    torch.cuda.synchronize()
    return logits.cpu().numpy(), dens.cpu().numpy()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_matches_reference_engine_golden(name):
    rec = load_golden(name)
    n = int(rec["int.n_images"])
    ev = _engine().NNUEEvaluator(GOLDEN / f"{name}.nnue")
    c = rec["cfg"]
    assert (ev.num_features, ev.l1_size, ev.l2_size, ev.l3_size, ev.num_classes, ev.grid_size) == (
        c["grid"] ** 2 * c["C"], c["L1"], c["L2"], c["L3"], c["NC"], c["grid"])
    imgs = chw_bytes_as_hwc(rec["images"][:n])
    logits, dens = gpu_eval(ev, imgs)
    np.testing.assert_array_equal(logits, rec["int.logits"])
    np.testing.assert_array_equal(dens, rec["int.density"])
    hl, hd = ev.evaluate_logits_host(imgs)
    np.testing.assert_array_equal(hl, rec["int.logits"])
    np.testing.assert_array_equal(hd, rec["int.density"])


@pytest.mark.parametrize("arch", INT_ARCHS, ids=lambda a: "x".join(map(str, a)))
@pytest.mark.parametrize("wild", [False, True])
@pytest.mark.parametrize("form", ["cta", "warp"])
def test_random_models_match_oracle(oracle_built, tmp_path, arch, wild, form):
    """Random .nnue payloads over the format's full integer ranges: int16 wraparound, every clamp,
    negative thresholds (zero-filled tail of the conv buffer becomes active), > 64 channels.
    Both one-kernel forms: a CTA per sample (the default up to 1024 samples) and a warp per sample."""
    from nnue_vision_b200 import _lib
    G, C, L1, L2, L3, NC, H = arch
    rng = np.random.default_rng(abs(hash(arch)) % (2**32) + 17 * wild)
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=wild)
    B = 5 if H > 100 else 37
    imgs = (rng.standard_normal((B, H, H, 3)) * (3.0 if wild else 1.0)).astype(np.float32)
    imgs[0] = 0.0
    orc = oracle_built.IntOracle(path)
    ol, od = orc.eval_batch(imgs, threads=4)
    ev = _engine().NNUEEvaluator(path)
    _lib.set_option("q_cta_max_batch", 1024 if form == "cta" else 0)
    try:
        gl, gd = gpu_eval(ev, imgs)
    finally:
        _lib.set_option("q_cta_max_batch", 1024)
    np.testing.assert_array_equal(gl, ol)
    np.testing.assert_array_equal(gd, od)
    if oracle_built.RefEngine.available():
        rl, rd = oracle_built.RefEngine(path).eval_batch(imgs, threads=4)
        np.testing.assert_array_equal(gl, rl)
        np.testing.assert_array_equal(gd, rd)


@pytest.mark.parametrize("thr", [-0.5, 0.0, 126.5, 127.0])
def test_threshold_edges(oracle_built, tmp_path, thr):
    rng = np.random.default_rng(3)
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, 6, 8, 64, 8, 8, 10, wild=True, threshold=thr)
    imgs = (rng.standard_normal((16, 40, 40, 3)) * 4).astype(np.float32)
    ol, od = oracle_built.IntOracle(path).eval_batch(imgs)
    gl, gd = gpu_eval(_engine().NNUEEvaluator(path), imgs)
    np.testing.assert_array_equal(gl, ol)
    np.testing.assert_array_equal(gd, od)


def test_batch_sweep_is_batch_invariant_and_exact(oracle_built):
    """BASELINE config: batch sweep 1..65536 on the default model.  EVERY one of the 65536 samples is checked against
    the compiled reference engine (oracle/_ref; the C restatement when it is absent), and per-sample results must not
    depend on the batch they ran in."""
    path = GOLDEN / "default_cfg.nnue"
    rng = np.random.default_rng(11)
    full = rng.standard_normal((65536, 32, 32, 3)).astype(np.float32)
    ev = _engine().NNUEEvaluator(path)
    big_l, big_d = gpu_eval(ev, full)
    cpu = oracle_built.RefEngine(path) if oracle_built.RefEngine.available() else oracle_built.IntOracle(path)
    ol, od = cpu.eval_batch(full, threads=oracle_built.host_threads())
    np.testing.assert_array_equal(big_l, ol)
    np.testing.assert_array_equal(big_d, od)
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 4096, 16384):
        l, d = gpu_eval(ev, full[:B])
        np.testing.assert_array_equal(l, big_l[:B])
        np.testing.assert_array_equal(d, big_d[:B])
    # logits are multiples of 1/64 (nnue_engine.cpp:532)
    assert np.all(big_l * 64 == np.round(big_l * 64))


def test_serialized_model_round_trip(oracle_built, tmp_path):
    """float model -> our serializer -> our GPU engine == C oracle on the same file."""
    from nnue_vision_b200 import nnue, serialize
    torch.manual_seed(5)
    model = nnue.NNUE(nnue.GridFeatureSet(10, 8), 64, 32, 8, num_classes=10, input_size=32)
    with torch.no_grad():
        model.input.weight.mul_(3.0)
        model.input.bias.normal_(0, 0.5)
    path = tmp_path / "m.nnue"
    serialize.serialize_model(model, path)
    rng = np.random.default_rng(0)
    chw = rng.standard_normal((64, 3, 32, 32)).astype(np.float32)
    imgs = chw_bytes_as_hwc(chw)
    ol, od = oracle_built.IntOracle(path).eval_batch(imgs)
    gl, gd = gpu_eval(_engine().NNUEEvaluator(path), imgs)
    np.testing.assert_array_equal(gl, ol)
    np.testing.assert_array_equal(gd, od)


def test_load_errors(tmp_path):
    eng = _engine()
    ev = eng.NNUEEvaluator()
    assert ev.load_model(tmp_path / "missing.nnue") is False
    good = (GOLDEN / "parity_small.nnue").read_bytes()
    for bad in (b"XXXX" + good[4:], good[:4] + b"\x03\x00\x00\x00" + good[8:], good[:300], good[:-2]):
        p = tmp_path / "bad.nnue"
        p.write_bytes(bad)
        assert ev.load_model(p) is False
    assert ev.load_model(GOLDEN / "parity_small.nnue") is True
    from nnue_vision_b200 import _lib
    with pytest.raises(_lib.NnueError):  # 64x48 image: raster would overrun the feature buffer
        ev.evaluate_logits(torch.zeros(1, 48, 200, 3).cuda())


def test_evaluate_compiled_model_matches_per_sample_engine(oracle_built):
    """evaluate.evaluate_compiled_model (evaluate.py:90-400 of the reference): same metrics and density as
    evaluating every sample on its own through the CPU engine, as the reference's subprocess loop does."""
    from pathlib import Path
    from nnue_vision_b200 import evaluate, nnue, serialize
    from oracle import int_oracle
    torch.manual_seed(11)
    model = nnue.NNUE(nnue.GridFeatureSet(10, 8), 64, 32, 8, num_classes=10, input_size=32)
    g = torch.Generator().manual_seed(2)
    loader = [(torch.randn(n, 3, 32, 32, generator=g), torch.randint(0, 10, (n,), generator=g)) for n in (37, 64, 5)]
    metrics = evaluate.evaluate_compiled_model(model, loader, "nnue")
    assert set(metrics) == {"acc", "f1", "precision", "recall", "ms_per_sample", "latent_density"}
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "m.nnue"
        serialize.serialize_model(model, path)
        cpu = int_oracle.RefEngine(path) if int_oracle.RefEngine.available() else int_oracle.IntOracle(path)
        logits, dens = [], []
        for images, _ in loader:
            l, d = cpu.eval_batch(images.numpy().reshape(images.shape[0], 32, 32, 3))
            logits.append(torch.as_tensor(l)); dens.append(torch.as_tensor(d))
    ref = evaluate.compute_metrics(torch.cat(logits), torch.cat([t for _, t in loader]))
    for k in ("acc", "f1", "precision", "recall"):
        assert metrics[k] == ref[k], k
    assert abs(metrics["latent_density"] - float(torch.cat(dens).mean())) < 1e-7
    assert metrics["ms_per_sample"] > 0
    with pytest.raises(ValueError):
        evaluate.evaluate_compiled_model(model, loader, "etinynet")


def test_evaluate_model_float_path():
    from nnue_vision_b200 import evaluate, nnue
    torch.manual_seed(12)
    model = nnue.NNUE(nnue.GridFeatureSet(8, 4), 64, 4, 8, num_classes=10, input_size=32).cuda()
    g = torch.Generator().manual_seed(3)
    loader = [(torch.randn(16, 3, 32, 32, generator=g), torch.randint(0, 10, (16,), generator=g)) for _ in range(3)]
    loss, metrics = evaluate.evaluate_model(model, loader)
    with torch.no_grad():
        ref = sum(float(torch.nn.functional.cross_entropy(model(x.cuda()), y.cuda())) for x, y in loader) / 3
    assert abs(loss - ref) < 1e-6 and 0.0 <= metrics["acc"] <= 1.0


def test_layer_stack_buckets_select_their_own_stack(oracle_built, tmp_path):
    """N-bucket files (num_ls_buckets > 1, nnue_engine.cpp:619-635): `layer_stack_index` picks the stack, an
    index past the end falls back to stack 0 (nnue_engine.cpp:705-707).  Each bucket is checked bit-exactly
    against the oracle on a one-bucket file that holds only that stack (rebuilt with serialize.read_nnue)."""
    import struct
    from nnue_vision_b200 import serialize
    G, C, L1, L2, L3, NC, H = 8, 4, 64, 8, 8, 10, 32
    rng = np.random.default_rng(77)
    path = tmp_path / "three.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=True, n_buckets=3)
    q = serialize.read_nnue(path)
    assert q["metadata"]["num_ls_buckets"] == 3 and q["trailing_bytes"] == 0
    blob = path.read_bytes()
    stack_bytes = 16 + sum(8 + w.size + 4 + 4 * b.size for w, b in (
        (q["layer_stacks"][0][k + "_weight"], q["layer_stacks"][0][k + "_bias"]) for k in ("l1", "l1_fact", "l2", "output")))
    head_end = len(blob) - 3 * stack_bytes
    imgs = rng.standard_normal((33, H, H, 3)).astype(np.float32)
    ev = _engine().NNUEEvaluator(path)
    assert ev.num_layer_stacks == 3
    outs = []
    for bucket in range(3):
        single = tmp_path / f"only{bucket}.nnue"
        single.write_bytes(blob[:24] + struct.pack("<I", 1) + blob[28:head_end] +
                           blob[head_end + bucket * stack_bytes: head_end + (bucket + 1) * stack_bytes])
        ol, od = oracle_built.IntOracle(single).eval_batch(imgs, threads=2)
        logits, dens = ev.evaluate_logits(torch.as_tensor(imgs).cuda(), layer_stack_index=bucket)
        np.testing.assert_array_equal(logits.cpu().numpy(), ol)
        np.testing.assert_array_equal(dens.cpu().numpy(), od)
        outs.append(ol)
    assert not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[1], outs[2])
    past_end, _ = ev.evaluate_logits(torch.as_tensor(imgs).cuda(), layer_stack_index=7)
    np.testing.assert_array_equal(past_end.cpu().numpy(), outs[0])


TC_ARCHS = [a for a in INT_ARCHS if a[2] % 64 == 0]


@pytest.mark.parametrize("arch", TC_ARCHS, ids=lambda a: "x".join(map(str, a)))
@pytest.mark.parametrize("wild", [False, True])
@pytest.mark.parametrize("thr", [None, -0.5, 127.0])
def test_tensor_core_accumulate_matches_oracle(oracle_built, tmp_path, arch, wild, thr):
    """The large-batch form (bitmask -> tcgen05 accumulate of the high/low bytes of the int16 rows -> layer
    stack) forced onto small batches: full int16 rows and biases (wrap-around), negative thresholds (the
    cells past the conv raster become active), ragged batches that do not fill a 128-sample tile."""
    from nnue_vision_b200 import _lib
    G, C, L1, L2, L3, NC, H = arch
    rng = np.random.default_rng(abs(hash((arch, wild, thr))) % (2**32))
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=wild, threshold=thr)
    B = 7 if H > 100 else 300
    imgs = (rng.standard_normal((B, H, H, 3)) * (3.0 if wild else 1.0)).astype(np.float32)
    ol, od = oracle_built.IntOracle(path).eval_batch(imgs, threads=4)
    ev = _engine().NNUEEvaluator(path)
    fl, fd = gpu_eval(ev, imgs)  # fused kernel (batch below the switch-over)
    _lib.set_option("q_tc_min_batch", 1)
    _lib.set_option("q_stack_fused", 1)
    try:
        tl, td = gpu_eval(ev, imgs)  # (L1 = 64 with a small stack: the layer stack runs in the accumulate kernel's epilogue)
        tl1, td1 = gpu_eval(ev, imgs[:1])
        _lib.set_option("q_stack_fused", 0)
        sl, sd = gpu_eval(ev, imgs)  # three launches: bitmask, accumulate to int16, layer stack
    finally:
        _lib.set_option("q_tc_min_batch", 2048)
        _lib.set_option("q_stack_fused", 8192)
    np.testing.assert_array_equal(fl, ol)
    np.testing.assert_array_equal(tl, ol)
    np.testing.assert_array_equal(td, od)
    np.testing.assert_array_equal(tl1, ol[:1])
    np.testing.assert_array_equal(sl, ol)
    np.testing.assert_array_equal(sd, od)


# 32 x 32 images at conv stride 4 (grids of 9..11 squares): the shapes q_conv_bits32_kernel serves, every channel count it is built for
FIXED_CONV_ARCHS = [(10, 8, 64, 32, 8, 10, 32), (9, 4, 64, 8, 8, 10, 32), (11, 16, 64, 16, 8, 10, 32), (10, 32, 128, 16, 16, 10, 32)]


@pytest.mark.parametrize("arch", FIXED_CONV_ARCHS, ids=lambda a: "x".join(map(str, a)))
@pytest.mark.parametrize("thr", [-127.5, -127.0, -126.5, -1.0, -0.5, 0.0, 0.5, 1.0, 63.0, 126.5, 127.0])
def test_fixed_conv_bitmask_matches_oracle(oracle_built, tmp_path, arch, thr):
    """The TMA-staged conv + bitmask kernel of the large-batch form replaces the engine's truncating divide, clamp and
    float compare by one integer compare against a host-computed bound: every threshold regime (never, always, the
    truncation asymmetry around zero, the clamp ends) against the oracle and against the general conv kernel, on
    saturating inputs, with a ragged batch (persistent CTAs, 14 consumer warps, two half-sample stages per sample)."""
    from nnue_vision_b200 import _lib
    G, C, L1, L2, L3, NC, H = arch
    rng = np.random.default_rng(abs(hash((arch, thr))) % (2**32))
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=True, threshold=thr)
    B = 2500
    imgs = (rng.standard_normal((B, H, H, 3)) * rng.choice([0.05, 0.5, 3.0], size=(B, 1, 1, 1))).astype(np.float32)
    imgs[0] = 0.0
    ol, od = oracle_built.IntOracle(path).eval_batch(imgs, threads=oracle_built.host_threads())
    ev = _engine().NNUEEvaluator(path)
    _lib.set_option("q_stack_fused", 1)
    try:
        tl, td = gpu_eval(ev, imgs)                  # batch >= 2048: the tensor-core form with the fixed conv kernel
        _lib.set_option("q_conv_fixed", 0)
        _lib.set_option("q_stack_fused", 0)
        gl, gd = gpu_eval(ev, imgs)                  # the general conv kernel, accumulate and layer stack as separate launches
    finally:
        _lib.set_option("q_conv_fixed", 1)
        _lib.set_option("q_stack_fused", 8192)
    np.testing.assert_array_equal(td, od)
    np.testing.assert_array_equal(tl, ol)
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gl, ol)
    assert 0.0 < float(od.mean()) < 1.0 or thr in (-127.5, -127.0, 127.0)  # (the middle thresholds do exercise both outcomes)


@pytest.mark.parametrize("arch", [(8, 4, 64, 4, 8, 10), (10, 8, 64, 32, 8, 10), (5, 3, 30, 5, 7, 3), (6, 16, 128, 16, 32, 1000)],
                         ids=lambda a: "x".join(map(str, a)))
@pytest.mark.parametrize("wild", [False, True])
def test_incremental_accumulator_api_matches_engine(oracle_built, tmp_path, arch, wild):
    """refresh_accumulator / update_features / evaluate_incremental / save / restore (nnue_engine.cpp:739-821),
    one stream (the reference's interface) and 33 streams at once, against the restatement that
    tests/test_oracle_int.py pins to the reference engine -- and against the engine itself where oracle/_ref exists."""
    from nnue_vision_b200 import serialize
    G, C, L1, L2, L3, NC = arch
    rng = np.random.default_rng(abs(hash((arch, wild))) % (2**32))
    path = tmp_path / "m.nnue"
    write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=wild)
    q = serialize.read_nnue(path)
    F = G * G * C
    ev = _engine().NNUEEvaluator(path)
    ref = oracle_built.RefEngine(path) if oracle_built.RefEngine.available() else None

    def draw(n):
        return rng.choice(F, size=min(F, n), replace=False).tolist()

    # one stream: a walk of small changes, each step scored incrementally
    feats = draw(60)
    if ref:
        ref.mark_dirty()
    for step in range(6):
        got = ev.evaluate_incremental(feats)
        assert got == oracle_built.legacy_score(q, feats), f"step {step}"
        if ref:
            assert got == ref.eval_incremental(feats)
        drop = set(feats[:4])
        feats = [f for f in feats if f not in drop] + [f for f in draw(6) if f not in feats]
    # the accumulator itself: int16 wrap-around sums
    acc = ev.get_accumulator().cpu().numpy()[0].astype(np.int64)
    want = q["feature_transformer"]["bias"].astype(np.int64)
    last = ev._last_features[0]
    want = want + q["feature_transformer"]["weight"][last].astype(np.int64).sum(0)
    np.testing.assert_array_equal(acc, (want + 32768) % 65536 - 32768)
    # save / restore / mark_dirty / disable
    ev.save_accumulator()
    ev.update_features([1, 2, 3], [])
    ev.restore_accumulator()
    np.testing.assert_array_equal(ev.get_accumulator().cpu().numpy()[0], acc)
    ev.update_features([F + 5, -1], [F])  # out-of-range indices are ignored (nnue_engine.cpp:215, 234)
    np.testing.assert_array_equal(ev.get_accumulator().cpu().numpy()[0], acc)
    ev.mark_dirty()
    assert ev.evaluate_incremental(last) == oracle_built.legacy_score(q, last)
    ev.enable_incremental(False)
    other = draw(30)
    assert ev.evaluate_incremental(other) == oracle_built.legacy_score(q, other)
    ev.enable_incremental(True)
    # 33 streams at once
    streams = [draw(int(rng.integers(0, 80))) for _ in range(33)]
    ev.mark_dirty()
    for step in range(3):
        got = ev.evaluate_incremental(streams).cpu().numpy()
        want = np.array([oracle_built.legacy_score(q, s) for s in streams], np.float32)
        np.testing.assert_array_equal(got, want)
        streams = [[f for f in s[2:]] + [f for f in draw(3) if f not in s] for s in streams]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_incremental_api_matches_reference_engine_golden(name):
    """evaluate_incremental against scores the reference engine itself produced on the golden models
    (tests/golden/incremental.npz): fresh refreshes, an incremental walk, and all fresh lists as one batch of streams."""
    from util import load_incremental_golden
    g = load_incremental_golden(name)
    ev = _engine().NNUEEvaluator(GOLDEN / f"{name}.nnue")
    for feats, want in zip(g["fresh"], g["fresh_score"]):
        ev.mark_dirty()
        assert np.float32(ev.evaluate_incremental(feats)) == want
    ev.mark_dirty()
    for feats, want in zip(g["walk"], g["walk_score"]):
        assert np.float32(ev.evaluate_incremental(feats)) == want
    ev.mark_dirty()
    batch = ev.evaluate_incremental([list(f) for f in g["fresh"]]).cpu().numpy()
    np.testing.assert_array_equal(batch, g["fresh_score"])


def test_command_line_twin_prints_the_reference_csv(tmp_path):
    """nnue_inference_b200 <model> <image.bin> H W (host C++ over the C ABI): the same CSV line the reference's
    engine/nnue_inference.cpp:56-62 prints -- logits at setprecision(10), then the density -- for a golden model whose
    expected numbers came from the compiled reference engine."""
    import subprocess
    from util import ROOT, load_golden
    exe = ROOT / "nnue-vision_b200" / "nnue_inference_b200"
    assert exe.exists(), "build.py builds the CLI next to the library"
    rec = load_golden("default_cfg")
    n = int(rec["int.n_images"])
    chw = np.asarray(rec["images"][:n], np.float32)
    for i in range(min(n, 3)):
        img = tmp_path / f"img{i}.bin"
        chw[i].tofile(img)  # evaluate.py:154-158 dumps the CHW floats; the engine reads the bytes as HWC
        H, W = chw.shape[2], chw.shape[3]
        out = subprocess.run([str(exe), str(GOLDEN / "default_cfg.nnue"), str(img), str(H), str(W)], capture_output=True,
                             text=True, check=True).stdout.strip()
        expect = ",".join(f"{v:.10f}" for v in list(rec["int.logits"][i]) + [float(rec["int.density"][i])])
        assert out == expect
    bad = subprocess.run([str(exe), str(tmp_path / "missing.nnue"), str(img), "32", "32"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Failed to load model" in bad.stderr
