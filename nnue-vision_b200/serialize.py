"""`.nnue` (format v2) writer and quantiser -- drop-in for the NNUE half of the reference's
`serialize.py`.  Output is byte-identical to /root/reference/serialize.py:500-528 for the same
parameters (tests/test_serialize.py compares against files the reference wrote).

File layout, little endian (serialize.py:30-63, 103-136, 394-491):
  header   "NNUE" u32 version=2 | u32 F, L1, L2, L3 | u32 num_ls_buckets=1 |
           f32 nnue2score, f32 quantized_one=127, f32 visual_threshold (mean over channels)
  conv     u32 type=0 | f32 scale | u32 OC, IC, KH, KW | int8 W[OC,IC,KH,KW] | u32 OC | int32 b[OC]
  FT       f32 scale | u32 F, L1 | int16 W[F,L1] | u32 L1 | int32 b[L1]
  stack    f32 l1, l2, out, l1_fact scales |
           L1 block  u32 L2+1, L1 | int8 [(L2+1),L1] (last row 0) | u32 L2+1 | int32 b
           L1-fact   u32 L1, L1   | int8 127*I                    | u32 L1   | int32 0
           L2 block  u32 L3, 2*L2 | int8 [L3, 2*L2] (cols >= L2 zero) | u32 L3 | int32 b
           output    u32 NC, L3   | int8 [NC, L3]                 | u32 NC   | int32 b
"""
import struct
from pathlib import Path
from typing import Any, Dict

import numpy as np
import torch

QUANT_SCALE = 64.0


def _require_keys(d: Dict[str, Any], required, context: str) -> None:
    missing = [k for k in required if k not in d]
    if missing:
        raise ValueError(f"Missing required {context} keys: {', '.join(missing)}")


def _quantize(weight, bias, n_out, scale):
    """round(w*scale) clamped to +-127 as int8, round(b*scale) as int32; torch.round is
    round-half-to-even (serialize.py:210-239)."""
    weight = weight.detach().cpu().float()
    bias = torch.zeros(n_out) if bias is None else bias.detach().cpu().float()
    weight_q = torch.round(weight * scale).clamp(-127, 127).to(torch.int8)
    bias_q = torch.round(bias * scale).to(torch.int32)
    return {"weight": weight_q, "bias": bias_q, "scale": scale}


def quantize_conv_layer(conv_layer, scale=QUANT_SCALE):
    return _quantize(conv_layer.weight.data, None if conv_layer.bias is None else conv_layer.bias.data,
                     conv_layer.out_channels, scale)


def quantize_linear_layer(linear_layer, scale=QUANT_SCALE):
    bias = linear_layer.bias
    n_out = linear_layer.bias.shape[0] if bias is not None else linear_layer.out_features
    return _quantize(linear_layer.weight.data, None if bias is None else bias.data, n_out, scale)


def _np(t, dtype):
    return (t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)).astype(dtype)


def _u32(f, *vals):
    f.write(struct.pack("<" + "I" * len(vals), *vals))


def _f32(f, *vals):
    f.write(struct.pack("<" + "f" * len(vals), *vals))


def write_nnue_header(f, metadata: Dict[str, Any]) -> None:
    _require_keys(metadata, ("feature_set", "L1", "L2", "L3", "nnue2score", "quantized_one", "visual_threshold"),
                  "NNUE metadata")
    f.write(b"NNUE")
    _u32(f, 2, metadata["feature_set"].num_features, metadata["L1"], metadata["L2"], metadata["L3"], 1)
    _f32(f, metadata["nnue2score"], metadata["quantized_one"], float(metadata["visual_threshold"]))


def write_conv_layer(f, conv_data: Dict[str, Any]) -> None:
    w, b = conv_data["weight"], conv_data["bias"]
    _u32(f, 0)
    _f32(f, conv_data["scale"])
    _u32(f, *w.shape)
    f.write(_np(w, "i1").tobytes())
    _u32(f, b.shape[0])
    f.write(_np(b, "<i4").tobytes())


def write_feature_transformer(f, ft_data: Dict[str, Any]) -> None:
    w, b = ft_data["weight"], ft_data["bias"]
    _f32(f, ft_data["scale"])
    _u32(f, w.shape[0], w.shape[1])
    f.write(_np(w, "<i2").tobytes())
    _u32(f, b.shape[0])
    f.write(_np(b, "<i4").tobytes())


def _block(f, n_out, n_in, w_int8, b_int32):
    _u32(f, n_out, n_in)
    f.write(np.ascontiguousarray(w_int8, "i1").tobytes())
    _u32(f, n_out)
    f.write(np.ascontiguousarray(b_int32, "<i4").tobytes())


def write_layer_stack(f, classifier_data: Dict[str, Any]) -> None:
    l1, l2, l3 = classifier_data["layers"]
    _f32(f, l1["scale"], l2["scale"], l3["scale"], l1["scale"])
    w1, b1 = _np(l1["weight"], "i1"), _np(l1["bias"], "<i4")
    L2, L1 = w1.shape
    w1x = np.zeros((L2 + 1, L1), "i1")
    w1x[:L2] = w1
    b1x = np.zeros(L2 + 1, "<i4")
    b1x[:L2] = b1
    _block(f, L2 + 1, L1, w1x, b1x)
    _block(f, L1, L1, np.eye(L1, dtype="i1") * 127, np.zeros(L1, "<i4"))
    w2, b2 = _np(l2["weight"], "i1"), _np(l2["bias"], "<i4")
    L3 = w2.shape[0]
    w2x = np.zeros((L3, 2 * L2), "i1")
    w2x[:, :L2] = w2
    _block(f, L3, 2 * L2, w2x, b2)
    w3, b3 = _np(l3["weight"], "i1"), _np(l3["bias"], "<i4")
    _block(f, w3.shape[0], L3, w3, b3)


def write_classifier(f, classifier_data: Dict[str, Any]) -> None:
    write_layer_stack(f, classifier_data)


def serialize_model(model, output_path) -> None:
    """Quantise `model` (any object with the NNUE duck type: `_clip_weights`,
    `get_quantized_model_data`) and write `output_path`.  Like the reference, this clips the live
    weights of the model to [-1, 1] in place (nnue.py:528-539)."""
    model.eval()
    model._clip_weights()
    q = model.get_quantized_model_data()
    with open(Path(output_path), "wb") as f:
        write_nnue_header(f, q["metadata"])
        write_conv_layer(f, q["conv_layer"])
        write_feature_transformer(f, q["feature_transformer"])
        write_classifier(f, q["classifier"])


# ---- reader ---------------------------------------------------------------
# The reference only parses `.nnue` files in C++ (NNUEEvaluator::load_model, nnue_engine.cpp:544-657,
# ConvLayer / FeatureTransformer / LayerStack::load_from_stream :11-46, :161-186, :283-380).  The reader
# below applies the same checks in the same order and returns the payloads as numpy arrays, so that a
# file can be inspected, turned back into a float model, or re-written byte for byte.
class NnueFormatError(ValueError):
    pass


def _take(buf, off, fmt):
    n = struct.calcsize(fmt)
    if off + n > len(buf):
        raise NnueFormatError("truncated .nnue file")
    return struct.unpack_from(fmt, buf, off), off + n


def _take_array(buf, off, dtype, count):
    n = np.dtype(dtype).itemsize * count
    if off + n > len(buf):
        raise NnueFormatError("truncated .nnue file")
    return np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy(), off + n


def read_nnue(path) -> Dict[str, Any]:
    """Parse a format-v2 `.nnue` file into
    {"metadata": {...}, "conv_layer": {...}, "feature_transformer": {...}, "layer_stacks": [ {...}, ... ]}
    with integer payloads exactly as stored (conv int8 [OC,IC,KH,KW], FT int16 [F,L1], dense int8)."""
    buf = Path(path).read_bytes()
    if buf[:4] != b"NNUE":
        raise NnueFormatError("bad magic")
    (version, F, L1, L2, L3, n_buckets), off = _take(buf, 4, "<IIIIII")
    if version != 2:
        raise NnueFormatError(f"unsupported version {version}")
    (nnue2score, quantized_one, threshold), off = _take(buf, off, "<fff")
    (layer_type,), off = _take(buf, off, "<I")
    if layer_type != 0:
        raise NnueFormatError("conv layer type must be 0")
    (conv_scale, OC, IC, KH, KW), off = _take(buf, off, "<fIIII")
    conv_w, off = _take_array(buf, off, "i1", OC * IC * KH * KW)
    (nb,), off = _take(buf, off, "<I")
    conv_b, off = _take_array(buf, off, "<i4", nb)
    if OC == 0 or F == 0 or F % OC:
        raise NnueFormatError("invalid feature/channel configuration")
    G = int(np.sqrt(F // OC))
    if G * G * OC != F:
        raise NnueFormatError("invalid feature grid calculation")
    (ft_scale, ftF, ftL1), off = _take(buf, off, "<fII")
    ft_w, off = _take_array(buf, off, "<i2", ftF * ftL1)
    (nb,), off = _take(buf, off, "<I")
    ft_b, off = _take_array(buf, off, "<i4", nb)
    if ftF != F or ftL1 != L1:
        raise NnueFormatError("feature transformer architecture mismatch")
    stacks = []
    for _ in range(n_buckets):
        (s1, s2, so, sf), off = _take(buf, off, "<ffff")
        blocks = []
        for _ in range(4):  # L1 (+1 row), L1-fact, L2 (2*L2 inputs), output
            (n_out, n_in), off = _take(buf, off, "<II")
            w, off = _take_array(buf, off, "i1", n_out * n_in)
            (nb,), off = _take(buf, off, "<I")
            b, off = _take_array(buf, off, "<i4", nb)
            blocks.append((w.reshape(n_out, n_in), b))
        (w1, b1), (wf, bf), (w2, b2), (wo, bo) = blocks
        if w1.shape != (L2 + 1, L1) or w2.shape != (L3, 2 * L2) or wo.shape[1] != L3:
            raise NnueFormatError("layer stack architecture mismatch")
        stacks.append({"l1_scale": s1, "l2_scale": s2, "output_scale": so, "l1_fact_scale": sf,
                       "l1_weight": w1, "l1_bias": b1, "l1_fact_weight": wf, "l1_fact_bias": bf,
                       "l2_weight": w2, "l2_bias": b2, "output_weight": wo, "output_bias": bo})
    return {
        "metadata": {"version": version, "num_features": F, "L1": L1, "L2": L2, "L3": L3, "num_ls_buckets": n_buckets,
                     "nnue2score": nnue2score, "quantized_one": quantized_one, "visual_threshold": threshold,
                     "grid_size": G, "num_features_per_square": OC,
                     "num_classes": int(stacks[0]["output_weight"].shape[0]) if stacks else 0},
        "conv_layer": {"weight": conv_w.reshape(OC, IC, KH, KW), "bias": conv_b, "scale": conv_scale},
        "feature_transformer": {"weight": ft_w.reshape(F, L1), "bias": ft_b, "scale": ft_scale},
        "layer_stacks": stacks,
        "trailing_bytes": len(buf) - off,
    }


def load_nnue_as_model(path, input_size=32, layer_stack_index=0):
    """Float `NNUE` whose parameters are the de-quantised contents of a `.nnue` file (payload / scale).
    Serialising it again reproduces the file byte for byte when the file came from `serialize_model`
    (integers up to 2**24 survive the division by 64 and the multiplication back exactly).  The file
    keeps only the channel mean of the threshold (nnue.py:556-558); it is replicated per channel."""
    from .nnue import NNUE, GridFeatureSet
    q = read_nnue(path)
    md = q["metadata"]
    st = q["layer_stacks"][layer_stack_index]
    model = NNUE(GridFeatureSet(md["grid_size"], md["num_features_per_square"]), md["L1"], md["L2"], md["L3"],
                 num_classes=md["num_classes"], input_size=input_size)
    L2 = md["L2"]
    lin = [m for m in model.classifier.classifier if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        model.conv.weight.copy_(torch.from_numpy(q["conv_layer"]["weight"].astype(np.float32) / q["conv_layer"]["scale"]))
        model.input.weight.copy_(torch.from_numpy(q["feature_transformer"]["weight"].astype(np.float32) / q["feature_transformer"]["scale"]))
        model.input.bias.copy_(torch.from_numpy(q["feature_transformer"]["bias"].astype(np.float32) / q["feature_transformer"]["scale"]))
        lin[0].weight.copy_(torch.from_numpy(st["l1_weight"][:L2].astype(np.float32) / st["l1_scale"]))
        lin[0].bias.copy_(torch.from_numpy(st["l1_bias"][:L2].astype(np.float32) / st["l1_scale"]))
        lin[1].weight.copy_(torch.from_numpy(st["l2_weight"][:, :L2].astype(np.float32) / st["l2_scale"]))
        lin[1].bias.copy_(torch.from_numpy(st["l2_bias"].astype(np.float32) / st["l2_scale"]))
        lin[2].weight.copy_(torch.from_numpy(st["output_weight"].astype(np.float32) / st["output_scale"]))
        lin[2].bias.copy_(torch.from_numpy(st["output_bias"].astype(np.float32) / st["output_scale"]))
        model.nnue2score.fill_(md["nnue2score"])
        model.visual_threshold.fill_(md["visual_threshold"])
    return model
