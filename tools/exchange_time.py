#!/usr/bin/env python3
"""Latency of the library's one-shot all-reduce (csrc/allreduce.cu) per slice size, N ranks of one box:
100 back-to-back exchanges replayed as one CUDA graph (no host launch cost), CUDA events, max over ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/exchange_time.py
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


class _Buf:
    def __init__(self, n, split, device):
        self.flat = torch.zeros(n, device=device)
        self.split = split

    def numel(self):
        return self.flat.numel()


def main():
    from nnue_vision_b200 import train
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    out = {}
    for n in (256, 4096, 55296, 1 << 20):
        buf = _Buf(n, 0, dev)  # split 0: the whole buffer is the "early" slice
        x = train.OneShotExchange(buf, world, None, dev)
        reps = 100
        for _ in range(3):
            x.full()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                x.full()
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # NCCL for comparison (eager launches: includes its host cost)
        y = torch.zeros(n, device=dev)
        for _ in range(5):
            dist.all_reduce(y)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(reps):
            dist.all_reduce(y)
        e1.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        out[str(n)] = {"oneshot_us": float(t), "nccl_us": float(t2)}
    if rank == 0:
        print(json.dumps({"world": world, "floats": out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
