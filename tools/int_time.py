"""Integer inference timed per batch size and kernel choice (config-D model): eager call loop and the same calls
replayed as one CUDA graph (device time per call without host launch cost).

    python tools/int_time.py [B ...]          # options A/B: q_conv_fixed, q_cta_max_batch
"""
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from nnue_vision_b200 import _lib, engine, serialize


def timed(ev, imgs, reps=50):
    B = imgs.shape[0]
    out = (torch.empty(B, ev.num_classes, device="cuda"), torch.empty(B, device="cuda"))
    for _ in range(3):
        ev.evaluate_logits(imgs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ev.evaluate_logits(imgs, out=out)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            ev.evaluate_logits(imgs, out=out)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return eager, e0.elapsed_time(e1) / reps, float(out[1].mean())


def main():
    w = dict(bench.WORKLOADS["default_cifar_b16384"])
    profile = "--profile" in sys.argv  # one batch, three eager calls, nothing else (for ncu)
    if profile:
        sys.argv.remove("--profile")
    batches = [int(x) for x in sys.argv[1:]] or [1, 16, 256, 592, 1024, 2048, 4096, 16384, 65536]
    res = {}
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "m.nnue"
        serialize.serialize_model(bench.build_model(w, "cpu"), p)
        ev = engine.NNUEEvaluator(p)
        big = torch.randn(max(batches), 32, 32, 3, generator=torch.Generator().manual_seed(3)).cuda()
        if profile:
            for _ in range(3):
                ev.evaluate_logits(big[:batches[0]])
            torch.cuda.synchronize()
            return
        for B in batches:
            imgs = big[:B]
            variants = {"default": {}}
            if B >= 2048:
                variants["general_conv"] = {"q_conv_fixed": 0}
                variants["stack_other"] = {"q_stack_fused": 0 if B >= 8192 else 1}
            else:
                variants["warp_per_sample"] = {"q_cta_max_batch": 0}
            for name, opts in variants.items():
                for k, v in opts.items():
                    _lib.set_option(k, v)
                try:
                    eager, replay, dens = timed(ev, imgs)
                finally:
                    _lib.set_option("q_conv_fixed", 1)
                    _lib.set_option("q_stack_fused", 8192)
                    _lib.set_option("q_cta_max_batch", 592)
                res[f"{B}/{name}"] = {"eager_us": 1e3 * eager, "graph_us": 1e3 * replay, "samples_per_s": B / (replay * 1e-3),
                                      "density": dens}
                print(f"B={B:6d} {name:16s} eager {1e3 * eager:9.2f} us  graph {1e3 * replay:9.2f} us  "
                      f"{B / (replay * 1e-3) / 1e6:9.3f} M samples/s  density {dens:.4f}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
