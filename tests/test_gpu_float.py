"""-m gpu: parity of the float training path (CUDA kernels through the C ABI) against the oracle
and against golden vectors written by the reference itself.

Bar (BASELINE north_star): forward, backward and loss within 1e-5 relative of the PyTorch reference:
|got - ref| <= 1e-5*|ref| + 1e-5*max|ref| per tensor; the hard threshold bit-for-bit.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import (RTOL, ambiguous_samples, assert_close, build_model, deambiguate, golden_model, model_state_numpy,
                      oracle_step, shape_of, unpack_bits)
from oracle import float_oracle as fo
from util import GOLDEN_CASES


def _lib():
    from nnue_vision_b200 import _lib
    return _lib


# ---------------------------------------------------------------- golden vectors (reference outputs)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_backward_matches_reference_golden(name):
    rec, model = golden_model(name)
    images = torch.as_tensor(rec["images"]).cuda()
    labels = torch.as_tensor(rec["labels"]).cuda()
    model.train()
    logits = model(images)
    loss = torch.nn.functional.cross_entropy(logits, labels.long())
    loss.backward()
    assert_close(logits, rec["float.logits"], "logits")
    assert_close(loss, rec["float.loss"], "loss")
    grads = dict(model.named_parameters())
    assert grads["nnue2score"].grad is None  # tests/test_model.py:179-182
    for k, p in grads.items():
        if k == "nnue2score":
            continue
        assert p.grad is not None, f"no gradient for {k}"
        assert_close(p.grad, rec["float.grad." + k], "grad " + k)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fused_loss_matches_reference_golden(name):
    rec, model = golden_model(name)
    images = torch.as_tensor(rec["images"]).cuda()
    labels = torch.as_tensor(rec["labels"]).cuda()
    loss = model.loss(images, labels)
    (loss * 1.0).backward()
    assert_close(loss, rec["float.loss"], "loss")
    for k, p in model.named_parameters():
        if k != "nnue2score":
            assert_close(p.grad, rec["float.grad." + k], "grad " + k)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_extraction_bits_match_reference_golden(name):
    rec, model = golden_model(name)
    images = torch.as_tensor(rec["images"]).cuda()
    shape, bits = model.extract_bits(images)
    got = unpack_bits(shape, bits).reshape(images.shape[0], -1)
    ref = np.unpackbits(rec["float.bits"], axis=1, bitorder="little")[:, : got.shape[1]].astype(bool)
    amb = ambiguous_samples(rec["float.conv_out"], rec["state.visual_threshold"], eps=1e-6)
    assert amb.sum() == 0, "golden inputs were chosen away from the threshold"
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------- per-kernel checks against the oracle
CFGS = {
    # name: cfg, image size, batch
    "T": (dict(grid=8, C=4, L1=64, L2=4, L3=8, NC=10, model_input=32), 32, 16),
    "D": (dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32), 32, 300),
    "D_staged": (dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32), 32, 1500),
    "D1k": (dict(grid=10, C=8, L1=1024, L2=128, L3=32, NC=10, model_input=32), 32, 96),
    "I_small": (dict(grid=16, C=32, L1=256, L2=16, L3=32, NC=1000, model_input=224), 224, 24),
    "big_into_small": (dict(grid=4, C=8, L1=32, L2=4, L3=4, NC=10, model_input=32), 96, 33),
    "small_into_big": (dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=64), 20, 40),
    "L1_16": (dict(grid=5, C=6, L1=16, L2=8, L3=8, NC=7, model_input=32), 32, 70),
    "odd_L1_generic": (dict(grid=6, C=5, L1=50, L2=6, L3=5, NC=3, model_input=30), 30, 37),
    "L1_128": (dict(grid=6, C=16, L1=128, L2=16, L3=32, NC=1000, model_input=64), 64, 50),
    "one_sample": (dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32), 32, 1),
    # wide stacks at a batch where layer 1 runs as the split-bf16 tcgen05 GEMM (gemm_umma.cu)
    "D1k_tensor_head": (dict(grid=10, C=8, L1=1024, L2=128, L3=32, NC=10, model_input=32), 32, 300),
    "wide_head_ragged": (dict(grid=6, C=16, L1=256, L2=48, L3=16, NC=10, model_input=64), 64, 257),
    # 1000 classes at a batch where the output layer's two gradients run as split-bf16 tcgen05 GEMMs (head.cu / gemm_umma.cu)
    "many_classes_tensor_head": (dict(grid=6, C=16, L1=128, L2=16, L3=32, NC=1000, model_input=64), 64, 300),
    "many_classes_ragged": (dict(grid=6, C=8, L1=64, L2=24, L3=20, NC=333, model_input=32), 32, 515),
    # the table gradient (side stream) and the conv gradient's partials (main stream) share one workspace: shapes where
    # the partials are larger than the value gradient's operands (L1 = 32 / small grids, B >= 148)
    "L1_32_overlap": (dict(grid=8, C=8, L1=32, L2=8, L3=8, NC=10, model_input=32), 32, 300),
    "umma_small_grid_overlap": (dict(grid=5, C=32, L1=64, L2=8, L3=8, NC=10, model_input=32), 32, 200),
    # SURVEY config I: the largest grid the integer engine supports (F = 65536 rows x 1024 columns = 268 MB, beyond L2)
    "I_large": (dict(grid=32, C=64, L1=1024, L2=128, L3=32, NC=1000, model_input=224), 224, 6),
}


def _make(name, seed=0):
    cfg, isize, B = CFGS[name]
    torch.manual_seed(42 + seed)
    model = build_model(cfg)
    g = torch.Generator().manual_seed(100 + seed)
    images = torch.randn(B, 3, isize, isize, generator=g).cuda()
    labels = torch.randint(0, cfg["NC"], (B,), generator=g).cuda()
    # give the biases / thresholds some spread so nothing is accidentally symmetric
    with torch.no_grad():
        model.input.bias.copy_(torch.randn(cfg["L1"], generator=g).cuda() * 0.05)
        model.visual_threshold.copy_((torch.rand(cfg["C"], generator=g) * 0.3 - 0.05).cuda())
    if cfg["grid"] * cfg["grid"] * cfg["C"] >= 16384:  # large feature sets: see gpu_util.deambiguate
        images = deambiguate(model, images)
    return cfg, model, images, labels


@pytest.mark.parametrize("name", list(CFGS))
def test_full_step_against_oracle(name):
    """Every intermediate and every gradient of one training step vs the fp64 oracle."""
    cfg, model, images, labels = _make(name)
    lib = _lib()
    ref = oracle_step(model, images, labels, dtype=torch.float64)
    B = images.shape[0]
    shape = shape_of(model, images)

    # stage 1: extraction (conv + threshold)
    bits_s = torch.empty((B, shape.NW), dtype=torch.int32, device="cuda")
    bits_t = torch.empty((shape.PP, shape.BW), dtype=torch.int32, device="cuda")
    conv_out = torch.empty((B, shape.C, shape.Gh, shape.Gw), dtype=torch.float32, device="cuda")
    nnz = torch.empty((B,), dtype=torch.int32, device="cuda")
    xpad = torch.full((B, shape.PP), float("nan"), dtype=torch.float32, device="cuda")
    lib.check(lib.lib().nnue_extract_fwd(
        ctypes.byref(shape), lib.dptr(images), lib.dptr(model.conv.weight.detach().contiguous()),
        lib.dptr(model.visual_threshold.detach().contiguous()), lib.dptr(bits_s), lib.dptr(bits_t),
        lib.dptr(xpad), lib.dptr(conv_out), lib.dptr(nnz), lib.stream_ptr()))
    torch.cuda.synchronize()
    assert tuple(conv_out.shape) == tuple(ref["conv_out"].shape)
    assert_close(conv_out, ref["conv_out"], "conv_out")
    # the padded-layout copy kept for the threshold gradient holds the same activations
    cells = shape.Gh * shape.Gw
    xp = xpad.view(B, shape.C, shape.CW * 32)[:, :, :cells].reshape(B, shape.C, shape.Gh, shape.Gw)
    assert torch.equal(xp, conv_out)
    got_bits = unpack_bits(shape, bits_s)
    ref_bits = ref["bits"].reshape(B, shape.C, -1).numpy()
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy())
    assert amb.mean() < 0.2
    np.testing.assert_array_equal(got_bits[~amb], ref_bits[~amb])
    np.testing.assert_array_equal(nnz.cpu().numpy(), got_bits.reshape(B, -1).sum(1))
    # transposed copy is consistent with the sample-major one
    t = bits_t.cpu().numpy().astype(np.uint32).reshape(shape.C, shape.CW * 32, shape.BW)
    tb = ((t[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(shape.C, shape.CW * 32, -1)
    np.testing.assert_array_equal(tb[:, : shape.Gh * shape.Gw, :B].transpose(2, 0, 1), got_bits)
    assert not tb[:, :, B:].any() and not tb[:, shape.Gh * shape.Gw:, :].any()
    if amb.any():  # the rest of the test needs identical active sets: drop ambiguous samples
        keep = torch.as_tensor(~amb).cuda()
        images, labels = images[keep].contiguous(), labels[keep].contiguous()
        ref = oracle_step(model, images, labels, dtype=torch.float64)

    # stage 2: sparse index lists as _to_sparse_features lays them out
    idx_ref, _ = fo.sparse_features(ref["bits"])
    shape, bits = model.extract_bits(images)
    K = idx_ref.shape[1]
    idx = torch.empty((images.shape[0], K), dtype=torch.int64, device="cuda")
    val = torch.empty((images.shape[0], K), dtype=torch.float32, device="cuda")
    lib.check(lib.lib().nnue_sparse_from_bits(ctypes.byref(shape), lib.dptr(bits), K, lib.dptr(idx), lib.dptr(val),
                                              lib.stream_ptr()))
    np.testing.assert_array_equal(idx.cpu().numpy(), idx_ref.numpy())
    np.testing.assert_array_equal(val.cpu().numpy(), (idx_ref >= 0).float().numpy())

    # stage 3: whole forward + backward through the module
    model.zero_grad()
    logits = model(images)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    assert_close(logits, ref["logits"], "logits")
    assert_close(loss, ref["loss"], "loss")
    for k, g in ref["grads"].items():
        p = dict(model.named_parameters())[k]
        assert_close(p.grad, g, "grad " + k, rtol=RTOL)

    # stage 4: the fused-loss entry gives the same numbers
    model.zero_grad()
    loss2 = model.loss(images, labels)
    loss2.backward()
    assert_close(loss2, ref["loss"], "fused loss")
    for k, g in ref["grads"].items():
        assert_close(dict(model.named_parameters())[k].grad, g, "fused grad " + k)


@pytest.mark.parametrize("mode", [0, 2])
def test_ft_forward_staging_modes_agree(mode):
    """Table rows gathered from global/L2 (0) and from TMA-staged shared memory (2) give identical sums."""
    lib = _lib()
    cfg, model, images, labels = _make("D")
    ref = oracle_step(model, images, labels)
    shape, bits = model.extract_bits(images)
    out = torch.empty((images.shape[0], cfg["L1"]), dtype=torch.float32, device="cuda")
    lib.set_option("ft_fwd_staging", mode)
    try:
        lib.check(lib.lib().nnue_ft_fwd(ctypes.byref(shape), lib.dptr(bits),
                                        lib.dptr(model.input.weight.detach().contiguous()),
                                        lib.dptr(model.input.bias.detach().contiguous()), lib.dptr(out), None, 0,
                                        lib.stream_ptr()))
        torch.cuda.synchronize()
    finally:
        lib.set_option("ft_fwd_staging", 1)
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy())
    assert_close(out[torch.as_tensor(~amb).cuda()], ref["ft_out"][torch.as_tensor(~amb)], "ft_out")


OPTION_DEFAULTS = {"extract_tma": 0, "ft_form": 0, "input_bwd_onchip": 0, "gemm_inline_a": 0}


@pytest.mark.parametrize("options", [dict(extract_tma=1), dict(extract_fixed=0), dict(ft_umma=0), dict(ft_umma=0, ft_mma=0),
                                     dict(ft_umma=0, ft_mma=0, ft_bwd_both=0), dict(ft_umma=0, ft_mma=0, ft_bwd_dw_owner=0),
                                     dict(ft_umma=0, ft_mma=0, input_bwd_fused=0), dict(input_bwd_fused=0), dict(input_bwd_onchip=1), dict(ft_form=2), dict(ft_form=2, input_bwd_fused=0), dict(input_bwd_variant=0), dict(input_bwd_swizzle=0), dict(input_bwd_swizzle=0, input_bwd_variant=0), dict(head_fused=0), dict(head_umma=0), dict(head_mid=0), dict(head_mid=0, head_umma=0), dict(head_mid=2), dict(head_pair_epilogue=0), dict(gemm_inline_a=1), dict(conv_bwd_packed=0),
                                     dict(ft_umma=0, ft_mma=0, ft_bwd_dw_owner=0, input_bwd_fused=0, head_fused=0)],
                         ids=str)
@pytest.mark.parametrize("name", ["D", "T", "big_into_small", "D1k_tensor_head"])
def test_backward_kernel_variants_agree(name, options):
    """The general kernels (transposed-bitmask segment reduction; value-gradient + conv-gradient pair)
    stay correct on the shapes where the row-owner / fused kernels normally run."""
    lib = _lib()
    cfg, model, images, labels = _make(name)
    ref = oracle_step(model, images, labels, dtype=torch.float64)
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy())
    if amb.any():
        keep = torch.as_tensor(~amb).cuda()
        images, labels = images[keep].contiguous(), labels[keep].contiguous()
        ref = oracle_step(model, images, labels, dtype=torch.float64)
    for k, v in options.items():
        lib.set_option(k, v)
    try:
        model.zero_grad()
        loss = model.loss(images, labels)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        for k in options:
            lib.set_option(k, OPTION_DEFAULTS.get(k, 1))
    assert_close(loss, ref["loss"], "loss")
    for k, g in ref["grads"].items():
        assert_close(dict(model.named_parameters())[k].grad, g, "grad " + k)


@pytest.mark.parametrize("form", [1, 2], ids=["dense_tcgen05", "gather_tma"])
@pytest.mark.parametrize("name", ["I_small", "I_large", "D1k"])
def test_large_table_formulations_against_oracle(name, form):
    """Both formulations of the feature transformer (plan.cuh cost model: dense bit-GEMMs on the tensor cores, or the
    TMA-staged row gather / segment reduction) forced on the large-table shapes, loss and every gradient against the
    fp64 oracle -- whichever the model would pick at this batch, the other one is covered too."""
    lib = _lib()
    cfg, model, images, labels = _make(name)
    ref = oracle_step(model, images, labels, dtype=torch.float64)
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy())
    if amb.any():
        keep = torch.as_tensor(~amb).cuda()
        images, labels = images[keep].contiguous(), labels[keep].contiguous()
        ref = oracle_step(model, images, labels, dtype=torch.float64)
    lib.set_option("ft_form", form)
    try:
        shape = shape_of(model, images)
        assert bool(lib.lib().nnue_ft_uses_umma(ctypes.byref(shape))) == (form == 1)
        model.zero_grad()
        loss = model.loss(images, labels)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        lib.set_option("ft_form", 0)
    assert_close(loss, ref["loss"], "loss")
    for k, g in ref["grads"].items():
        assert_close(dict(model.named_parameters())[k].grad, g, "grad " + k)


def test_cost_model_picks_the_measured_winner():
    """plan.cuh: at SURVEY config I the gather form serves single-digit batches (one sample: 0.04 ms against 0.63 ms),
    the dense form everything from a few dozen samples up (batch 512: 0.6 ms against 9.5 ms); a density hint of 1 %
    turns large batches over to the gather form; the CIFAR tables are always dense."""
    lib = _lib()
    def dense(B, cfg=(224, 64, 32, 1024, 128, 32, 1000, 7)):
        H, C, G, L1, L2, L3, NC, s = cfg
        return bool(lib.lib().nnue_ft_uses_umma(ctypes.byref(lib.make_shape(B, H, H, C, G, L1, L2, L3, NC, s))))
    assert not dense(1) and not dense(4) and dense(64) and dense(4096)
    assert dense(1, (32, 8, 10, 64, 32, 8, 10, 3)) and dense(16384, (32, 8, 10, 64, 32, 8, 10, 3))
    lib.set_option("ft_density_permille", 10)
    try:
        assert not dense(4096)
    finally:
        lib.set_option("ft_density_permille", 400)


def test_feature_transformer_indexed_interface():
    """model.input(idx, val) with repeated / unsorted / out-of-range / -1 indices and float values
    (tests/test_model.py:1026-1064 use this sub-interface), forward and all three gradients."""
    cfg, model, _, _ = _make("D")
    ft = model.input
    g = torch.Generator().manual_seed(5)
    B, K = 9, 23
    idx = torch.randint(-1, ft.num_features + 40, (B, K), generator=g)
    idx[3] = -1  # a sample with no active feature at all
    val = torch.randn(B, K, generator=g).cuda().requires_grad_(True)
    out = ft(idx.cuda(), val)
    go = torch.randn(B, cfg["L1"], generator=g).cuda()
    out.backward(go)
    W, b = ft.weight.detach().double().cpu(), ft.bias.detach().double().cpu()
    ref = fo.ft_forward(idx, val.detach().double().cpu(), W, b)
    gw, gb, gv = fo.ft_backward(idx, val.detach().double().cpu(), W, go.double().cpu())
    assert_close(out, ref, "ft indexed out")
    assert_close(ft.weight.grad, gw, "ft indexed dW")
    assert_close(ft.bias.grad, gb, "ft indexed dbias")
    assert_close(val.grad, gv, "ft indexed dval")


@pytest.mark.parametrize("B,K,F", [(300, 40, 97), (7, 5, 3), (1, 1, 1), (513, 129, 70000)])
def test_sort_pairs_is_the_stable_sort_by_row(B, K, F):
    """nnue_ft_sort_pairs (the chunked counting sort in front of the indexed weight gradient) against torch's stable sort:
    keys min(idx, F - 1), negative indices behind every row with value 0, (sample, slot) order kept inside a row; several
    chunks of pairs, keys that repeat inside a 32-pair group, tables larger than the pair count."""
    lib = _lib()
    g = torch.Generator().manual_seed(B * 1000 + K)
    idx = torch.randint(-3, F + 10, (B, K), generator=g).cuda()
    val = torch.randn(B, K, generator=g).cuda()
    n = B * K
    rows = torch.empty(n, dtype=torch.int32, device="cuda")
    samples = torch.empty(n, dtype=torch.int32, device="cuda")
    vals = torch.empty(n, dtype=torch.float32, device="cuda")
    nb = int(lib.lib().nnue_ft_sort_pairs_workspace_bytes(B, K, F))
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    lib.check(lib.lib().nnue_ft_sort_pairs(B, K, F, lib.dptr(idx), lib.dptr(val), lib.dptr(rows), lib.dptr(samples), lib.dptr(vals),
                                           lib.dptr(ws), nb, lib.stream_ptr()))
    torch.cuda.synchronize()
    key = torch.where(idx < 0, torch.full_like(idx, F), idx.clamp(max=F - 1)).reshape(-1)
    order = torch.sort(key, stable=True).indices
    assert torch.equal(rows.long(), key[order])
    assert torch.equal(samples.long(), (order // K))
    assert torch.equal(vals, torch.where(idx < 0, torch.zeros_like(val), val).reshape(-1)[order])


def test_sparse_feature_values_keep_autograd_edge():
    cfg, model, images, _ = _make("T")
    x = model.conv(images)
    from nnue_vision_b200.nnue import binary_activation_ste
    binm = binary_activation_ste(x, model.visual_threshold.view(1, -1, 1, 1))
    idx, val = model._to_sparse_features(binm)
    assert idx.dtype == torch.int64 and val.dtype == torch.float32 and idx.shape == val.shape
    assert val.requires_grad  # nnue.py:602-603
    ref_idx, _ = fo.sparse_features(binm.detach().cpu() > 0.5)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx.numpy())


def test_gradient_flow_contract():
    """tests/test_model.py:142-230 of the reference: which parameters receive gradients."""
    cfg, model, images, labels = _make("D")
    images.requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(model(images), labels)
    loss.backward()
    named = dict(model.named_parameters())
    assert named["nnue2score"].grad is None
    for k in ("conv.weight", "input.weight", "input.bias", "classifier.classifier.0.weight",
              "classifier.classifier.0.bias", "visual_threshold"):
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all(), k
    assert named["input.weight"].grad.abs().sum() > 0 and named["classifier.classifier.0.weight"].grad.abs().sum() > 0
    assert named["conv.weight"].grad.abs().sum() > 0


def test_model_can_learn():
    """tests/test_model.py:232-296: a few Adam steps must not make things worse."""
    cfg, model, images, labels = _make("T")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(20):
        opt.zero_grad()
        loss = model.loss(images, labels)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


def test_cpu_tensors_fail_loudly():
    cfg, model, images, _ = _make("T")
    with pytest.raises(_lib().NnueError):
        model(images.cpu())


def test_large_batch_properties():
    """BASELINE batch (16384, config D): checks that do not need the oracle at full size --
    determinism, per-sample independence against the same samples run in small batches, and
    linearity of the gradient in the upstream gradient."""
    cfg, isize, _ = CFGS["D"]
    torch.manual_seed(1)
    model = build_model(cfg)
    g = torch.Generator().manual_seed(9)
    B = 16384
    images = torch.randn(B, 3, isize, isize, generator=g).cuda()
    labels = torch.randint(0, cfg["NC"], (B,), generator=g).cuda()
    with torch.no_grad():
        logits = model(images)
        again = model(images)
        assert torch.equal(logits, again), "forward must be run-to-run deterministic"
        sub = torch.randperm(B, generator=g)[:512].cuda()
        small = model(images[sub].contiguous())
    assert_close(small, logits[sub], "sample independence")
    # oracle on a slice of the big batch
    ref = oracle_step(model, images[:256], labels[:256])
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy())
    assert_close(logits[:256][torch.as_tensor(~amb).cuda()], ref["logits"][torch.as_tensor(~amb)], "big-batch logits")
    # gradients: deterministic and linear in the upstream gradient
    def grads(scale):
        model.zero_grad()
        (model.loss(images, labels) * scale).backward()
        return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    g1, g1b, g3 = grads(1.0), grads(1.0), grads(3.0)
    for k in g1:
        assert torch.equal(g1[k], g1b[k]), f"{k}: backward must be deterministic"
        assert_close(g3[k], 3.0 * g1[k].double(), "linearity " + k)
    # mean over the batch: summing 64 shard gradients (each scaled by 1/B) reproduces the full one
    step_grads = None
    for i in range(0, B, 4096):
        model.zero_grad()
        model.loss(images[i:i + 4096].contiguous(), labels[i:i + 4096].contiguous(), global_batch=B).backward()
        cur = {k: p.grad.double() for k, p in model.named_parameters() if p.grad is not None}
        step_grads = cur if step_grads is None else {k: step_grads[k] + cur[k] for k in cur}
    for k in g1:
        assert_close(g1[k], step_grads[k], "shard additivity " + k, rtol=2e-5)


def test_benchmark_batch_against_the_fp64_oracle():
    """Config D at the BENCHMARK batch (16384): loss and every gradient of the fused step (the 296-CTA split-K weight
    gradient, the chunked folds) against the vectorised fp64 oracle -- not only against themselves."""
    cfg, isize, _ = CFGS["D"]
    torch.manual_seed(1)
    model = build_model(cfg)
    g = torch.Generator().manual_seed(9)
    B = 16384
    images = torch.randn(B, 3, isize, isize, generator=g)
    labels = torch.randint(0, cfg["NC"], (B,), generator=g)
    ref = oracle_step(model, images, labels, dtype=torch.float64)
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy(), eps=2e-6)
    if amb.any():  # (a handful of the 15.9 M activations lie within rounding of their threshold: drop those samples)
        keep = torch.as_tensor(~amb)
        images, labels = images[keep].contiguous(), labels[keep].contiguous()
        ref = oracle_step(model, images, labels, dtype=torch.float64)
    assert images.shape[0] > 16000
    model.zero_grad()
    loss = model.loss(images.cuda(), labels.cuda())
    loss.backward()
    assert_close(loss, ref["loss"], "loss")
    for k, gr in ref["grads"].items():
        assert_close(dict(model.named_parameters())[k].grad, gr, "grad " + k)


def test_wide_stack_split_k_batch_against_the_fp64_oracle():
    """The reference's "real" config (L1 = 1024, 128/32 stack) at a batch that engages the split-K weight gradient of
    gemm_umma.cu and several K chunks of the table gradient."""
    cfg, isize, _ = CFGS["D1k"]
    torch.manual_seed(2)
    model = build_model(cfg)
    g = torch.Generator().manual_seed(10)
    B = 4096
    images = torch.randn(B, 3, isize, isize, generator=g)
    labels = torch.randint(0, cfg["NC"], (B,), generator=g)
    ref = oracle_step(model, images, labels, dtype=torch.float64)
    amb = ambiguous_samples(ref["conv_out"].numpy(), model.visual_threshold.detach().cpu().numpy(), eps=2e-6)
    if amb.any():
        keep = torch.as_tensor(~amb)
        images, labels = images[keep].contiguous(), labels[keep].contiguous()
        ref = oracle_step(model, images, labels, dtype=torch.float64)
    model.zero_grad()
    loss = model.loss(images.cuda(), labels.cuda())
    loss.backward()
    assert_close(loss, ref["loss"], "loss")
    for k, gr in ref["grads"].items():
        assert_close(dict(model.named_parameters())[k].grad, gr, "grad " + k)


@pytest.mark.parametrize("name", ["D1k_tensor_head", "wide_head_ragged"])
def test_head_side_stream_chain_is_bit_identical(name):
    """Wide stacks hand the layer-1 weight-gradient chain to the side stream (nnue_head_train_overlapped): the same
    kernels on the same data, so loss and every gradient must be bit-identical to the one-stream step, eager and replayed
    as a CUDA graph (the event fork / join is captured)."""
    from nnue_vision_b200 import nnue as _n
    lib = _lib()
    cfg, model, images, labels = _make(name)
    shape = shape_of(model, images)
    assert int(lib.lib().nnue_head_side_workspace_bytes(ctypes.byref(shape))) > 0
    out = {}
    for flag in (True, False):
        _n.HEAD_SIDE_STREAM = flag
        try:
            runs = []
            for _ in range(3):  # third call: replayed graph
                model.zero_grad()
                loss = model.loss(images, labels)
                loss.backward()
                torch.cuda.synchronize()
                runs.append([loss.detach().clone()] + [p.grad.detach().clone() for p in model.parameters() if p.grad is not None])
            for r in runs[1:]:
                for a, b in zip(runs[0], r):
                    assert torch.equal(a, b)
            out[flag] = runs[0]
        finally:
            _n.HEAD_SIDE_STREAM = False
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["D1k_tensor_head", "wide_head_ragged", "I_small", "D1k"])
def test_in_kernel_operand_split_is_bit_identical(name):
    """Option gemm_inline_a (measured no faster, off by default): wide shapes build the A operand of the layer-1 forward GEMM
    (l0, pairwise transform included) and of the value-gradient GEMM (g_ft) inside the kernel, by the epilogue warps, instead
    of reading tiles a formatting kernel wrote: the same exact
    split feeds the same UMMAs in the same order, so loss and every gradient are bit-identical to the formatter path
    (ragged batches: rows past the batch are zero in both)."""
    lib = _lib()
    cfg, model, images, labels = _make(name)
    out = {}
    for flag in (1, 0):
        lib.set_option("gemm_inline_a", flag)
        try:
            model.zero_grad()
            loss = model.loss(images, labels)
            loss.backward()
            torch.cuda.synchronize()
            out[flag] = [loss.detach().clone()] + [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]
        finally:
            lib.set_option("gemm_inline_a", 0)
    for a, b in zip(out[1], out[0]):
        assert torch.equal(a, b)


def test_preformatted_table_tiles_give_identical_results():
    """nnue_ft_format_tables + the *_tables entry points (table tiles formatted once, on the side stream) run the same
    kernels on the same tiles as the per-call formatting: bit-identical loss and gradients."""
    from nnue_vision_b200 import nnue as nn_mod
    cfg, model, images, labels = _make("D_staged")

    def run():
        model.zero_grad()
        loss = model.loss(images, labels)
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    l0, g0 = run()
    old = nn_mod.PREFORMAT_TABLES
    nn_mod.PREFORMAT_TABLES = True
    try:
        l1, g1 = run()
    finally:
        nn_mod.PREFORMAT_TABLES = old
    assert torch.equal(l0, l1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
