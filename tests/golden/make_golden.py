#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference and oracle/_ref):

    make -C oracle && python tests/golden/make_golden.py

What it records, per case (one .npz + one .nnue each):
  * the reference model's state_dict (so tests do not depend on RNG streams),
  * float path: images, labels, logits, mean-CE loss and every parameter
    gradient from the reference's own `nnue.NNUE` + autograd
    (/root/reference/nnue.py:447-738, train.py:250-254), CPU fp32,
  * the `.nnue` bytes written by the reference's `serialize.serialize_model`
    (/root/reference/serialize.py:500-528) for that state_dict,
  * int path: logits + density from the reference C++ engine
    (engine/src/nnue_engine.cpp:704-734 via oracle/_ref/libnnue_ref.so) on
    images handed over exactly as evaluate.py:154-168 does (raw CHW bytes).

Nothing here is imported by the product; the fixtures are small and committed.
"""
import ctypes
import io
import os
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("NNUE_REFERENCE_ROOT", "/root/reference"))
sys.path.insert(0, str(REF))

import nnue as ref_nnue  # noqa: E402  (the reference)
import serialize as ref_serialize  # noqa: E402

torch.set_num_threads(1)

CASES = {
    # name: (grid, C, L1, L2, L3, NC, model_input_size, image_size, B_float, B_int, stress)
    "test_cfg": (8, 4, 64, 4, 8, 10, 32, 32, 16, 8, False),  # config/train_nnue_test.py:9-23
    "default_cfg": (10, 8, 64, 32, 8, 10, 32, 32, 12, 8, False),  # config/train_nnue_default.py:16-35
    "default_stress": (10, 8, 64, 32, 8, 10, 32, 32, 4, 8, True),  # clamp / int16-wrap exercise
    "parity_small": (4, 8, 32, 4, 4, 10, 32, 32, 8, 8, False),  # tests/test_compiled_parity.py:47-56
    "big_image": (4, 8, 32, 4, 4, 10, 32, 96, 3, 2, False),  # 96x96 into a 32-configured model
    "many_classes": (6, 16, 128, 16, 32, 1000, 64, 64, 2, 2, False),  # 1000-class head, L1=128
}


def build_reference_model(name, spec):
    grid, C, L1, L2, L3, NC, msize, isize, Bf, Bi, stress = spec
    torch.manual_seed(42)
    model = ref_nnue.NNUE(
        feature_set=ref_nnue.GridFeatureSet(grid_size=grid, num_features_per_square=C),
        l1_size=L1, l2_size=L2, l3_size=L3, num_classes=NC, input_size=msize,
    )
    if stress:
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            model.input.weight.mul_(3.0)
            model.input.bias.copy_(torch.randn(L1, generator=g) * 0.5)
            model.conv.weight.mul_(4.0)
            for m in model.classifier.classifier:
                if isinstance(m, torch.nn.Linear):
                    m.weight.mul_(2.5)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g))
    return model


def float_golden(model, images, labels):
    model.train()
    model.zero_grad()
    logits = model(images)
    loss = F.cross_entropy(logits, labels.long())
    loss.backward()
    out = {"logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
    for k, p in model.named_parameters():
        out["grad." + k] = None if p.grad is None else p.grad.detach().numpy().copy()
    assert out["grad.nnue2score"] is None  # tests/test_model.py:179-182
    del out["grad.nnue2score"]
    # intermediates that make a failing parity test debuggable
    with torch.no_grad():
        x = model.conv(images)
        bits = (x > model.visual_threshold.view(1, -1, 1, 1))
        out["conv_out"] = x.numpy()
        out["bits"] = np.packbits(bits.reshape(bits.shape[0], -1).numpy(), axis=1, bitorder="little")
        idx, val = model._to_sparse_features(bits.float())
        out["ft_out"] = model.input(idx, val).numpy()
        out["sparse_idx"] = idx.numpy().astype(np.int32)
    return out


def int_golden(lib, model_path, images_chw):
    B, _, H, W = images_chw.shape
    h = lib.ref_load(str(model_path).encode())
    assert h, "reference engine refused the .nnue"
    nc = lib.ref_num_classes(h, H, W)
    raw = np.ascontiguousarray(images_chw.numpy())  # CHW bytes, read as HWC by the engine
    logits = np.zeros((B, nc), np.float32)
    dens = np.zeros((B,), np.float32)
    rc = lib.ref_eval_batch(h, raw.ctypes.data_as(ctypes.c_void_p), B, H, W,
                            logits.ctypes.data_as(ctypes.c_void_p), dens.ctypes.data_as(ctypes.c_void_p), 1)
    assert rc == nc, rc
    lib.ref_free(h)
    return logits, dens


def main():
    lib = ctypes.CDLL(str(ROOT / "oracle" / "_ref" / "libnnue_ref.so"))
    lib.ref_load.restype = ctypes.c_void_p
    lib.ref_load.argtypes = [ctypes.c_char_p]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    lib.ref_num_classes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.ref_eval_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for name, spec in CASES.items():
        grid, C, L1, L2, L3, NC, msize, isize, Bf, Bi, stress = spec
        model = build_reference_model(name, spec)
        state = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        g = torch.Generator().manual_seed(1234)
        images = torch.randn(Bf, 3, isize, isize, generator=g)
        labels = torch.randint(0, NC, (Bf,), generator=g)
        rec = {"spec": np.array(spec[:10], dtype=np.int64), "images": images.numpy(), "labels": labels.numpy()}
        rec.update({"state." + k: v for k, v in state.items()})
        rec.update({"float." + k: v for k, v in float_golden(model, images, labels).items()})
        # serialize AFTER the float golden: serialize_model clips weights in place (nnue.py:528-539)
        nnue_path = HERE / f"{name}.nnue"
        with redirect_stdout(io.StringIO()):
            ref_serialize.serialize_model(model, nnue_path)
        rec.update({"clipped_state." + k: v.detach().numpy().copy() for k, v in model.state_dict().items()
                    if not np.array_equal(v.detach().numpy(), state[k])})
        int_images = images[:Bi].contiguous()
        logits, dens = int_golden(lib, nnue_path, int_images)
        rec["int.n_images"] = np.int64(Bi)
        rec["int.logits"] = logits
        rec["int.density"] = dens
        np.savez_compressed(HERE / f"{name}.npz", **rec)
        print(f"{name}: .nnue {nnue_path.stat().st_size} B, npz {(HERE / (name + '.npz')).stat().st_size} B, "
              f"loss {rec['float.loss']:.6f}, density {dens.mean():.3f}, logits[0,:3] {logits[0, :3]}")


if __name__ == "__main__":
    main()
