"""Shared helpers for the test-suite (test infrastructure only)."""
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_CASES = ["test_cfg", "default_cfg", "default_stress", "parity_small", "big_image", "many_classes"]


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    rec = {k: z[k] for k in z.files}
    grid, C, L1, L2, L3, NC, msize, isize, Bf, Bi = (int(v) for v in rec["spec"])
    rec["cfg"] = dict(grid=grid, C=C, L1=L1, L2=L2, L3=L3, NC=NC, model_input=msize, image=isize)
    return rec


def golden_state(rec, clipped=False):
    """state_dict (numpy) of the golden model; `clipped` applies the serialize-time clip."""
    st = {k[len("state."):]: v for k, v in rec.items() if k.startswith("state.")}
    if clipped:
        for k, v in rec.items():
            if k.startswith("clipped_state."):
                st[k[len("clipped_state."):]] = v
    return st


def write_random_nnue(path, rng, G, C, L1, L2, L3, NC, wild=False, threshold=None, n_buckets=1):
    """Write a syntactically valid .nnue v2 file with random integer payloads.

    Layout per the reference writer (serialize.py:30-63, 103-136, 394-491).  With
    wild=True the payloads use the full ranges the FORMAT allows (int16 FT rows,
    large biases) so that int16 wraparound and every clamp are exercised, which
    serialize.py's own quantiser (|w| <= 127) never produces.
    """
    F = G * G * C
    thr = float(rng.uniform(-3, 6)) if threshold is None else float(threshold)
    with open(path, "wb") as f:
        f.write(b"NNUE")
        f.write(struct.pack("<IIIIII", 2, F, L1, L2, L3, n_buckets))
        f.write(struct.pack("<fff", 600.0, 127.0, thr))
        f.write(struct.pack("<IfIIII", 0, 64.0, C, 3, 3, 3))
        f.write(rng.integers(-127, 128, size=C * 27, dtype=np.int8).tobytes())
        f.write(struct.pack("<I", C))
        f.write(rng.integers(-3000 if wild else 0, 3001 if wild else 1, size=C).astype("<i4").tobytes())
        f.write(struct.pack("<fII", 64.0, F, L1))
        lim = 32767 if wild else 127
        f.write(rng.integers(-lim, lim + 1, size=F * L1).astype("<i2").tobytes())
        f.write(struct.pack("<I", L1))
        blim = 70000 if wild else 64
        f.write(rng.integers(-blim, blim + 1, size=L1).astype("<i4").tobytes())
        for _ in range(n_buckets):
            f.write(struct.pack("<ffff", 64.0, 64.0, 64.0, 64.0))
            f.write(struct.pack("<II", L2 + 1, L1))
            f.write(rng.integers(-127, 128, size=(L2 + 1) * L1, dtype=np.int8).tobytes())
            f.write(struct.pack("<I", L2 + 1))
            f.write(rng.integers(-20000, 20001, size=L2 + 1).astype("<i4").tobytes())
            f.write(struct.pack("<II", L1, L1))
            f.write((np.eye(L1, dtype=np.int8) * 127).tobytes())
            f.write(struct.pack("<I", L1))
            f.write(np.zeros(L1, "<i4").tobytes())
            f.write(struct.pack("<II", L3, 2 * L2))
            f.write(rng.integers(-127, 128, size=L3 * 2 * L2, dtype=np.int8).tobytes())
            f.write(struct.pack("<I", L3))
            f.write(rng.integers(-9000, 9001, size=L3).astype("<i4").tobytes())
            f.write(struct.pack("<II", NC, L3))
            f.write(rng.integers(-127, 128, size=NC * L3, dtype=np.int8).tobytes())
            f.write(struct.pack("<I", NC))
            f.write(rng.integers(-9000, 9001, size=NC).astype("<i4").tobytes())
    return thr


# (G, C, L1, L2, L3, NC, H) -- architectures the randomized int-path tests sweep
INT_ARCHS = [
    (8, 4, 64, 4, 8, 10, 32),
    (10, 8, 64, 32, 8, 10, 32),
    (4, 8, 32, 4, 4, 10, 96),
    (6, 16, 128, 16, 32, 1000, 64),
    (5, 3, 30, 5, 7, 3, 17),      # odd everything: L1 not a multiple of 4, odd image, C=3
    (3, 70, 16, 4, 4, 5, 20),     # more than 64 channels: channels >= 64 never fire
    (1, 8, 16, 4, 4, 2, 9),       # single-cell grid: stride collapses to H
    (16, 32, 256, 16, 32, 1000, 224),  # engine test shape (tests/test_nnue_engine.cpp:12-16)
]


def load_incremental_golden(name):
    """Feature lists and the reference engine's evaluate_incremental scores for one golden model
    (tests/golden/make_incremental_golden.py)."""
    z = np.load(GOLDEN / "incremental.npz")

    def unpack(off, idx):
        return [idx[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]

    return {"fresh": unpack(z[f"{name}.fresh_off"], z[f"{name}.fresh_idx"]), "fresh_score": z[f"{name}.fresh_score"],
            "walk": unpack(z[f"{name}.walk_off"], z[f"{name}.walk_idx"]), "walk_score": z[f"{name}.walk_score"]}
