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
