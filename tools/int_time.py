import sys, tempfile, time
from pathlib import Path
sys.path.insert(0, '/root/repo')
import torch, bench
from nnue_vision_b200 import engine, serialize
w = dict(bench.WORKLOADS["default_cifar_b16384"])
with tempfile.TemporaryDirectory() as td:
    p = Path(td) / "m.nnue"
    serialize.serialize_model(bench.build_model(w, "cpu"), p)
    ev = engine.NNUEEvaluator(p)
    for B in ([int(x) for x in sys.argv[1:]] or (1, 256, 4096, 16384, 65536)):
        imgs = torch.randn(B, 32, 32, 3, generator=torch.Generator().manual_seed(3)).cuda()
        for _ in range(3): ev.evaluate_logits(imgs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): l, d = ev.evaluate_logits(imgs)
        e1.record(); torch.cuda.synchronize()
        print(B, e0.elapsed_time(e1) / 20, "ms", float(d.mean()))
