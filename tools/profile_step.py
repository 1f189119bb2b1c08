#!/usr/bin/env python3
"""Short driver for ncu: a few training steps of the bench workload (no timing, no CPU legs).

    python tools/profile_step.py [--workload default_cifar_b16384] [--batch B] [--steps 4] [--int]
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="default_cifar_b16384")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--int", action="store_true", help="also run the integer-inference kernel")
    a = ap.parse_args()
    from nnue_vision_b200 import train
    w = dict(bench.WORKLOADS[a.workload])
    if a.batch:
        w["batch"] = a.batch
    dev = torch.device("cuda", 0)
    model = bench.build_model(w, dev)
    dp = train.DataParallelStep(model)
    images, labels = bench.synthetic_batch(w, w["batch"], seed=1, device=dev)
    for _ in range(a.steps):
        loss = dp.step(images, labels)
    torch.cuda.synchronize()
    print("loss", float(loss))
    if a.int:
        import tempfile
        from nnue_vision_b200 import engine, serialize
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / "m.nnue"
            serialize.serialize_model(bench.build_model(w, "cpu"), p)
            ev = engine.NNUEEvaluator(p)
            imgs = torch.randn(w["batch"], w["image"], w["image"], 3, device=dev)
            for _ in range(a.steps):
                ev.evaluate_logits(imgs)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
