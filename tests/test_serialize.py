"""Host logic of the drop-in boundary: module surface, state_dict layout and the .nnue writer,
checked against files and tensors produced by the reference itself (tests/golden)."""
import io
import struct

import numpy as np
import pytest
import torch

from util import GOLDEN, GOLDEN_CASES, golden_state, load_golden

from nnue_vision_b200 import nnue, serialize

REFERENCE_PARAM_ORDER = [
    "nnue2score", "visual_threshold", "conv.weight", "input.weight", "input.bias",
    "classifier.classifier.0.weight", "classifier.classifier.0.bias",
    "classifier.classifier.2.weight", "classifier.classifier.2.bias",
    "classifier.classifier.4.weight", "classifier.classifier.4.bias",
]


def build(cfg):
    return nnue.NNUE(nnue.GridFeatureSet(cfg["grid"], cfg["C"]), cfg["L1"], cfg["L2"], cfg["L3"],
                     num_classes=cfg["NC"], input_size=cfg["model_input"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_nnue_file_is_byte_identical_to_reference(name, tmp_path):
    rec = load_golden(name)
    model = build(rec["cfg"])
    model.load_state_dict({k: torch.as_tensor(v) for k, v in golden_state(rec).items()})
    out = tmp_path / "m.nnue"
    serialize.serialize_model(model, out)
    assert out.read_bytes() == (GOLDEN / f"{name}.nnue").read_bytes()
    # serialize clips the live weights like the reference does (nnue.py:528-539)
    clipped = golden_state(rec, clipped=True)
    for k, v in model.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), clipped[k])
    assert not model.training


def test_parameter_order_and_state_dict_keys_match_reference():
    rec = load_golden("default_cfg")
    model = build(rec["cfg"])
    assert [k for k, _ in model.named_parameters()] == REFERENCE_PARAM_ORDER
    assert list(model.state_dict().keys()) == REFERENCE_PARAM_ORDER
    assert set(golden_state(rec)) == set(REFERENCE_PARAM_ORDER)
    for k, v in golden_state(rec).items():
        assert tuple(model.state_dict()[k].shape) == v.shape, k


def test_initialisation_stream_matches_reference():
    """Same construction order => same RNG consumption => torch.manual_seed(42) reproduces the
    reference's initial weights exactly (the golden state was built that way)."""
    rec = load_golden("default_cfg")
    torch.manual_seed(42)
    model = build(rec["cfg"])
    for k, v in golden_state(rec).items():
        np.testing.assert_array_equal(model.state_dict()[k].numpy(), v, err_msg=k)


def test_constructor_defaults_and_attributes():
    m = nnue.NNUE()
    assert (m.l1_size, m.l2_size, m.l3_size, m.num_classes, m.weight_decay, m.input_size) == (1024, 128, 32, 1, 5e-4, 32)
    assert m.feature_set == nnue.GridFeatureSet(10, 8) and m.feature_set.num_features == 800
    assert isinstance(m.loss_params, nnue.LossParams)
    assert m.conv.stride == (3, 3) and m.conv.bias is None and m.conv.out_channels == 8
    assert m.input.num_features == 800 and m.input.output_size == 1024
    assert float(m.nnue2score) == 600.0 and torch.all(m.visual_threshold == 0.1)
    # the reference's tests forbid these (tests/test_model.py:1346-1351)
    assert not hasattr(m, "use_optimizations") and not hasattr(m.input, "_enable_incremental")
    assert nnue.NNUE(nnue.GridFeatureSet(32, 64), input_size=224).conv.stride == (7, 7)
    assert nnue.NNUE(nnue.GridFeatureSet(4, 8), input_size=4).conv.stride == (1, 1)


def test_quantized_model_data_layout():
    rec = load_golden("test_cfg")
    model = build(rec["cfg"])
    q = model.get_quantized_model_data()
    assert set(q) == {"metadata", "conv_layer", "feature_transformer", "classifier"}
    assert set(q["metadata"]) == {"feature_set", "L1", "L2", "L3", "num_classes", "nnue2score", "quantized_one",
                                  "visual_threshold"}
    assert q["metadata"]["quantized_one"] == 127.0
    assert q["conv_layer"]["weight"].dtype == torch.int8 and q["conv_layer"]["bias"].dtype == torch.int32
    assert q["feature_transformer"]["weight"].shape == (256, 64) and len(q["classifier"]["layers"]) == 3


def test_quantiser_rounding_and_clamp():
    lin = torch.nn.Linear(4, 2)
    with torch.no_grad():
        lin.weight.copy_(torch.tensor([[0.0078125, 0.0234375, -0.0078125, 5.0], [-5.0, 0.3, 1.0, -1.0]]))
        lin.bias.copy_(torch.tensor([0.0078125, -100.0]))
    q = serialize.quantize_linear_layer(lin)
    # 0.5 -> 0, 1.5 -> 2, -0.5 -> -0 (round-half-even, serialize.py:234), +-320 clamp to +-127
    assert q["weight"].tolist() == [[0, 2, 0, 127], [-127, 19, 64, -64]]
    assert q["bias"].tolist() == [0, -6400] and q["scale"] == 64.0


def test_header_requires_metadata_keys():
    with pytest.raises(ValueError, match="Missing required NNUE metadata keys"):
        serialize.write_nnue_header(io.BytesIO(), {"L1": 1})


def test_sparse_features_host_layout():
    rec = load_golden("default_cfg")
    model = build(rec["cfg"])
    bits = np.unpackbits(rec["float.bits"], axis=1, bitorder="little")[:, :968].reshape(-1, 8, 11, 11)
    idx, val = model._to_sparse_features(torch.as_tensor(bits).float())
    np.testing.assert_array_equal(idx.numpy().astype(np.int32), rec["float.sparse_idx"])
    np.testing.assert_array_equal(val.numpy(), (idx >= 0).float().numpy())
    idx0, val0 = model._to_sparse_features(torch.zeros(2, 8, 11, 11))
    assert idx0.shape == (2, 1) and (idx0 == -1).all() and (val0 == 0).all()  # K >= 1 (nnue.py:612)


def test_straight_through_binary_matches_reference_formula():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 4, 5, 5, generator=g, requires_grad=True)
    thr = torch.full((1, 4, 1, 1), 0.1, requires_grad=True)
    y = nnue.binary_activation_ste(x, thr)
    go = torch.randn(3, 4, 5, 5, generator=g)
    y.backward(go)
    assert torch.equal(y, (x > thr).float()) and torch.equal(x.grad, go)
    sig = torch.sigmoid(10.0 * (x.detach() - thr.detach()))
    torch.testing.assert_close(thr.grad, -(go * 10.0 * sig * (1 - sig)).sum(dim=(0, 2, 3), keepdim=True))


def test_cpu_forward_raises_instead_of_falling_back():
    from nnue_vision_b200 import _lib
    m = build(load_golden("test_cfg")["cfg"])
    with pytest.raises(_lib.NnueError, match="no CPU fallback"):
        m(torch.zeros(2, 3, 32, 32))
    with pytest.raises(_lib.NnueError):
        m.input(torch.zeros(2, 3, dtype=torch.long), torch.ones(2, 3))


# ---- reader ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reader_parses_reference_written_files(name):
    """read_nnue on the files the reference's serialize.py wrote: header fields and payloads match the
    quantised golden state."""
    from nnue_vision_b200 import serialize
    rec = load_golden(name)
    q = serialize.read_nnue(GOLDEN / f"{name}.nnue")
    cfg, md = rec["cfg"], q["metadata"]
    assert (md["grid_size"], md["num_features_per_square"], md["L1"], md["L2"], md["L3"], md["num_classes"]) == (
        cfg["grid"], cfg["C"], cfg["L1"], cfg["L2"], cfg["L3"], cfg["NC"])
    assert md["num_ls_buckets"] == 1 and md["quantized_one"] == 127.0 and q["trailing_bytes"] == 0
    st = golden_state(rec, clipped=True)
    np.testing.assert_array_equal(q["feature_transformer"]["weight"],
                                  np.clip(np.round(st["input.weight"] * 64.0), -127, 127).astype(np.int16))
    np.testing.assert_array_equal(q["conv_layer"]["weight"],
                                  np.clip(np.round(st["conv.weight"] * 64.0), -127, 127).astype(np.int8))
    s0 = q["layer_stacks"][0]
    assert not s0["l1_weight"][-1].any() and not s0["l2_weight"][:, cfg["L2"]:].any()
    np.testing.assert_array_equal(s0["l1_fact_weight"], np.eye(cfg["L1"], dtype=np.int8) * 127)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_nnue_file_round_trips_through_float_model(name, tmp_path):
    """.nnue -> float model -> .nnue is byte-identical (the quantiser is idempotent on de-quantised values)."""
    from nnue_vision_b200 import serialize
    rec = load_golden(name)
    src = GOLDEN / f"{name}.nnue"
    model = serialize.load_nnue_as_model(src, input_size=rec["cfg"]["model_input"])
    out = tmp_path / "again.nnue"
    serialize.serialize_model(model, out)
    a, b = out.read_bytes(), src.read_bytes()
    # every payload byte survives; the header threshold is the fp32 MEAN of the per-channel values
    # (nnue.py:556-558), which may round in the last bit when re-averaged over C equal entries
    assert a[:36] == b[:36] and a[40:] == b[40:]
    ta, tb = struct.unpack("<f", a[36:40])[0], struct.unpack("<f", b[36:40])[0]
    assert abs(ta - tb) <= 2e-7 * max(abs(tb), 1e-30)


def test_reader_rejects_malformed_files(tmp_path):
    from nnue_vision_b200 import serialize
    good = (GOLDEN / "test_cfg.nnue").read_bytes()
    for name, blob in (("magic", b"XXXX" + good[4:]), ("version", good[:4] + b"\x03\x00\x00\x00" + good[8:]),
                       ("short", good[: len(good) // 2])):
        p = tmp_path / f"{name}.nnue"
        p.write_bytes(blob)
        with pytest.raises(serialize.NnueFormatError):
            serialize.read_nnue(p)
