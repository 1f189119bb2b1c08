"""Experiment: the config-D training step as K micro-batches on K streams inside one CUDA graph (the HBM-bound
extraction / conv-gradient kernels of one micro-batch overlap the latency-bound contractions of the others).

    python tools/microbatch_try.py [K ...]

Prints the replayed step time per K; K = 1 is the step as DataParallelStep runs it."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from nnue_vision_b200 import _lib, nnue as _nnue


def main():
    ks = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
    w = dict(bench.WORKLOADS["default_cifar_b16384"])
    dev = torch.device("cuda", 0)
    model = bench.build_model(w, dev)
    B = w["batch"]
    g = torch.Generator().manual_seed(1)
    sets = [torch.randn(B, 3, 32, 32, generator=g).to(dev) for _ in range(3)]  # 3 x 201 MB cycled: larger than L2
    labels = torch.randint(0, 10, (B,), generator=g).to(dev)
    params = tuple(p.detach().contiguous() for p in model._hot_params())
    fs = model.feature_set
    res = {}
    for K in ks:
        mb = B // K
        shape = _lib.make_shape(mb, 32, 32, fs.num_features_per_square, fs.grid_size, model.l1_size, model.l2_size,
                                model.l3_size, model.num_classes, model.conv.stride[0])
        streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
        grads = [tuple(torch.empty_like(p) for p in params) for _ in range(K)]
        losses = [torch.empty(1, device=dev) for _ in range(K)]
        total = tuple(torch.empty_like(p) for p in params)

        def step(images):
            cur = torch.cuda.current_stream()
            for k in range(K):
                s = streams[k] if K > 1 else cur
                if K > 1:
                    s.wait_stream(cur)
                with torch.cuda.stream(s):
                    _nnue._run_train_step(shape, images[k * mb:(k + 1) * mb], labels[k * mb:(k + 1) * mb], params, 1.0 / B,
                                          grads=grads[k], loss_out=losses[k])
            if K > 1:
                for s in streams:
                    cur.wait_stream(s)
                for i, t in enumerate(total):  # (an experiment: the product would fold these in one kernel)
                    torch.sum(torch.stack([gk[i] for gk in grads]), dim=0, out=t)

        graphs = []
        for images in sets:
            step(images)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                step(images)
            graphs.append(gr)
        for gr in graphs:
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(5):
            e0.record()
            for i in range(30):
                graphs[i % 3].replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 30)
        times.sort()
        loss = float(sum(l.item() for l in losses))
        res[K] = {"ms_per_step": times[2], "samples_per_s": B / (times[2] * 1e-3), "loss": loss}
        print(f"K={K}: {times[2] * 1e3:8.2f} us per step  {B / (times[2] * 1e-3) / 1e6:7.2f} M samples/s  loss {loss:.6f}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
