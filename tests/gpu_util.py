"""Helpers shared by the -m gpu parity tests (test infrastructure only)."""
import ctypes

import numpy as np
import torch

from oracle import float_oracle as fo
from util import golden_state, load_golden

RTOL = 1e-5  # float-path bar: |got - ref| <= RTOL*|ref| + RTOL*max|ref| per tensor (BASELINE north_star)


def assert_close(got, ref, what, rtol=RTOL):
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, np.float64)
    ref = np.asarray(ref.detach().cpu() if torch.is_tensor(ref) else ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    scale = np.abs(ref).max() if ref.size else 0.0
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + rtol * scale
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} outside tolerance; max err {err.max():.3e}, "
                           f"max|ref| {scale:.3e}, first bad index {np.argwhere(bad)[0].tolist()}")


def build_model(cfg, state=None, device="cuda"):
    from nnue_vision_b200 import nnue
    model = nnue.NNUE(
        feature_set=nnue.GridFeatureSet(grid_size=cfg["grid"], num_features_per_square=cfg["C"]),
        l1_size=cfg["L1"], l2_size=cfg["L2"], l3_size=cfg["L3"], num_classes=cfg["NC"],
        input_size=cfg["model_input"])
    if state is not None:
        model.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()})
    return model.to(device)


def model_state_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def golden_model(name, device="cuda"):
    rec = load_golden(name)
    return rec, build_model(rec["cfg"], golden_state(rec), device)


def unpack_bits(model_shape, bits_s):
    """int32 [B, NW] internal bitmask -> bool [B, C, Gh*Gw] in CHW order."""
    B = bits_s.shape[0]
    C, CW, cells = model_shape.C, model_shape.CW, model_shape.Gh * model_shape.Gw
    words = bits_s.detach().cpu().numpy().astype(np.uint32).reshape(B, C, CW)
    bits = ((words[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(B, C, CW * 32)
    assert not bits[:, :, cells:].any(), "padding bits must be zero"
    return bits[:, :, :cells]


def ambiguous_samples(conv_out64, thr, eps=1e-5):
    """Samples that have a conv activation within eps of its threshold: a last-bit difference in the
    fp32 conv sum may legitimately flip that bit, so bit-for-bit checks skip those samples."""
    d = np.abs(np.asarray(conv_out64, np.float64) - np.asarray(thr, np.float64).reshape(1, -1, 1, 1))
    return (d < eps).reshape(d.shape[0], -1).any(axis=1)


def oracle_step(model, images, labels, dtype=torch.float64):
    cfg_stride = model.conv.stride[0]
    return fo.step(model_state_numpy(model), images.detach().cpu().numpy(), labels.detach().cpu().numpy(),
                   cfg_stride, dtype=dtype)


def shape_of(model, images):
    from nnue_vision_b200 import _lib
    B, _, H, W = images.shape
    fs = model.feature_set
    return _lib.make_shape(B, H, W, fs.num_features_per_square, fs.grid_size, model.l1_size, model.l2_size,
                           model.l3_size, model.num_classes, model.conv.stride[0])


def deambiguate(model, images, eps=1e-4, rounds=8):
    """Nudges input pixels until no conv activation lies within `eps` of its threshold, so that the fp32 kernels and the
    fp64 oracle must agree on every bit of the active set.  With 65536 positions per sample (SURVEY config I) nearly
    every random sample has SOME activation within 1e-5 of its threshold and dropping such samples would leave none."""
    import torch.nn.functional as F
    w = model.conv.weight.detach().double().cpu()
    thr = model.visual_threshold.detach().double().cpu().view(1, -1, 1, 1)
    s = model.conv.stride[0]
    img = images.detach().double().cpu().clone()
    for _ in range(rounds):
        x = F.conv2d(img, w, None, stride=s, padding=1)
        near = ((x - thr).abs() < eps).any(dim=1)  # [B, Gh, Gw]
        idx = near.nonzero()
        if idx.numel() == 0:
            break
        # the centre tap of cell (oy, ox) is pixel (oy*s, ox*s), always inside the image
        img[idx[:, 0], 0, idx[:, 1] * s, idx[:, 2] * s] += 1e-2
    else:
        raise AssertionError("could not move the activations away from their thresholds")
    return img.float().to(images.device)
