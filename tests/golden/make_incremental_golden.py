"""Golden vectors for the engine's incremental interface (nnue_engine.cpp:739-821), produced by the REFERENCE ENGINE
itself (oracle/_ref/libnnue_ref.so, compiled from /root/reference by oracle/Makefile) on the committed golden models.

    python tests/golden/make_incremental_golden.py      # writes tests/golden/incremental.npz

Per model: 12 independent feature lists scored after mark_dirty() (full refresh), and a 6-step walk scored
incrementally (the engine diffs every call against its previous one).  Lists are stored as offsets + flat indices."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import int_oracle  # noqa: E402
from util import GOLDEN, GOLDEN_CASES  # noqa: E402


def pack(lists):
    off = np.zeros(len(lists) + 1, np.int32)
    np.cumsum([len(l) for l in lists], out=off[1:])
    flat = np.array([f for l in lists for f in l], np.int32)
    return off, flat


def main():
    assert int_oracle.RefEngine.available(), "build oracle/_ref first (make -C oracle)"
    out = {}
    for name in GOLDEN_CASES:
        rng = np.random.default_rng(abs(hash(name)) % (2**32))
        rng = np.random.default_rng(sum(map(ord, name)))  # stable across interpreter runs
        ref = int_oracle.RefEngine(GOLDEN / f"{name}.nnue")
        F = ref.F
        fresh = [rng.choice(F, size=int(rng.integers(0, min(F, 120))), replace=False).tolist() for _ in range(12)]
        fresh[0] = []
        fresh[1] = [0, F - 1, F + 3, -2]          # out-of-range indices are ignored by the engine
        scores = []
        for feats in fresh:
            ref.mark_dirty()
            scores.append(ref.eval_incremental(feats))
        walk, wscores = [], []
        feats = rng.choice(F, size=min(F, 50), replace=False).tolist()
        ref.mark_dirty()
        for _ in range(6):
            walk.append(list(feats))
            wscores.append(ref.eval_incremental(feats))
            drop = set(feats[:5])
            feats = [f for f in feats if f not in drop] + [int(f) for f in rng.choice(F, size=6, replace=False) if f not in feats]
        out[f"{name}.fresh_off"], out[f"{name}.fresh_idx"] = pack(fresh)
        out[f"{name}.fresh_score"] = np.array(scores, np.float32)
        out[f"{name}.walk_off"], out[f"{name}.walk_idx"] = pack(walk)
        out[f"{name}.walk_score"] = np.array(wscores, np.float32)
    np.savez_compressed(GOLDEN / "incremental.npz", **out)
    print("wrote", GOLDEN / "incremental.npz", {k: v.shape for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
